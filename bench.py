#!/usr/bin/env python
"""bench.py -- the hot path measured on B200s (contract: the driver's `python bench.py --gpus N --steps K --warmup W`).

Headline workload (BASELINE.json configs[1], the largest single-GPU configuration; weak scaling for N > 1):
  a (32 + log2 N)-qubit state vector, 64 GiB shard per GPU; one STEP = one sweep over ALL target positions,
  each position getting a oneTargGate and a manyCtrlOneTargGate with 1-3 controls (64 + 2 log2 N gates).
  Prefix targets (the top log2 N qubits) exercise the NVLink pairwise exchange.
Metric: "34-qubit-equivalent SV gates/s" = amplitude-updates per second / 2^34 (BASELINE's "34-qubit SV gates/s" made additive
  over GPUs under weak scaling -- a 34-qubit state does not fit one GPU); `gates_per_s_actual` is the plain rate at the actual
  size and `amp_updates_per_s` the un-normalised throughput.
Timing: CUDA events on the library's compute stream, barrier + device sync on both sides, max over ranks.

The same JSON line also carries
  "parity"   : BEFORE any timing, a (24 + log2 N)-qubit sweep + config-3 mini-circuit and a 12-qubit noisy density-matrix layer
               run through the SAME transports the timed region uses (fused remote-load kernels over NVLink, then again on the
               staged NCCL path) and are compared with the C oracle amplitude by amplitude on the device; a failure exits 3.
  "selfcheck": at the FULL size (33-35 qubits are beyond any oracle): the sweep, then its inverse, compared with the
               regenerated initial state over all amplitudes.
  "configs"  : BASELINE configs 3, 4, 5 (random 5-target dense gates / gadgets; noisy density-matrix layer at 14 and 16 qubits;
               expecPauliString over 256 strings + partialTrace), per-op device times and fractions of the roofline
               max(HBM bytes / measured HBM peak, NVLink bytes per direction / NVLink rate MEASURED in this run, flop / FP64 peak).
`--impl reference` times the reference's own CPU implementation (oracle/_ref/ref_driver: unmodified reference +
setBit patch + fork/shm MPI stand-in) on the box's host cores on a bounded sample of the same sweep.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

PKG = "distributed-full-state-algorithms_b200"
REF_SAMPLE_QUBITS = 30          # bounded CPU sample of the sweep (16 GiB state + 16 GiB buffer)
SEED = 20261017
FP64_PEAK_TFLOPS = 36.6         # measured on B200: profiles/r01_fp64_peak_b200.jsonl (DFMA 36.6, DMMA 37.0)
NVLINK_FALLBACK_GBS = 770.0     # per direction, B200_PROFILING.md; replaced by the in-run measurement when N > 1
PARITY_TOL = 1e-12
# ops that address amplitudes by index bit and therefore restore a lazily relabelled layout first (host/layout.hpp)
INDEX_ADDRESSED_OPS = ("dm_oneQubitDephasing", "dm_twoQubitDephasing", "dm_oneQubitDepolarising", "dm_twoQubitDepolarising", "dm_damping",
                       "dm_expecPauliString", "dm_partialTrace")


# ------------------------------------------------------------------------------------------------ workloads

def haar(rng, d):
    q, r = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def haar_2x2(rng):
    return haar(rng, 2)


def make_sweep(num_qubits, seed=SEED):
    """One step: for every target position a oneTargGate and a manyCtrlOneTargGate (1-3 random controls)."""
    rng = np.random.default_rng(seed)
    ops = []
    for t in range(num_qubits):
        ops.append(("sv_oneTargGate", t, haar_2x2(rng)))
        nc = 1 + t % 3
        ctrls = [int(c) for c in rng.permutation([q for q in range(num_qubits) if q != t])[:nc]]
        ops.append(("sv_manyCtrlOneTargGate", ctrls, t, haar_2x2(rng)))
    return ops


def inverse_ops(ops):
    """The op list that undoes `ops` (unitary gates only)."""
    out = []
    for op in reversed(ops):
        name = op[0]
        if name == "sv_oneTargGate":
            out.append((name, op[1], op[2].conj().T))
        elif name == "sv_manyCtrlOneTargGate":
            out.append((name, op[1], op[2], op[3].conj().T))
        elif name == "sv_manyTargGate":
            out.append((name, op[1], op[2].conj().T))
        elif name == "sv_pauliGadget":
            out.append((name, op[1], op[2], -op[3]))
        elif name == "sv_phaseGadget":
            out.append((name, op[1], -op[2]))
        elif name in ("sv_swapGate", "sv_pauliTensor"):
            out.append(op)
        else:
            raise ValueError(name)
    return out


def config_workload(name, world, seed=7, dm_qubits=None):
    """BASELINE configs 3-5."""
    rng = np.random.default_rng(seed)
    if name == "circuit":
        nq = {1: 32, 2: 33, 4: 34, 8: 34}[world]
        ops = []
        for _ in range(8):
            ops.append(("sv_manyTargGate", [int(x) for x in rng.permutation(nq)[:5]], haar(rng, 32)))
            nt = int(rng.integers(3, 7))
            paulis = [int(x) for x in rng.integers(1, 4, size=nt)]
            if all(p == 3 for p in paulis):
                paulis[0] = 1
            ops.append(("sv_pauliGadget", [int(x) for x in rng.permutation(nq)[:nt]], paulis, float(rng.uniform(-np.pi, np.pi))))
            ops.append(("sv_phaseGadget", [int(x) for x in rng.permutation(nq)[:int(rng.integers(1, 8))]], float(rng.uniform(-np.pi, np.pi))))
        return "sv", nq, ops, "config 3: random circuit of manyTargGate(5 targets) / pauliGadget / phaseGadget, random targets"
    if name == "dm":
        N = dm_qubits or (14 if world == 1 else 16)
        ops = []
        for q in range(N):
            ops.append(("dm_manyTargGate", [q, (q + 1) % N], haar(rng, 4)))
            ops.append(("dm_oneQubitDepolarising", q, float(rng.uniform(0, 0.5))))
            ops.append(("dm_twoQubitDephasing", q, (q + 1) % N, float(rng.uniform(0, 0.5))))
            ops.append(("dm_damping", q, float(rng.uniform(0, 0.5))))
        return "dm", N, ops, "config 4: noisy layer (manyTargGate t=2, oneQubitDepolarising, twoQubitDephasing, damping) on every qubit"
    if name == "expec":
        N = dm_qubits or (14 if world == 1 else 16)
        coeffs = rng.uniform(-10, 10, 256)
        paulis = rng.integers(0, 4, size=(256, N))
        ops = [("dm_expecPauliString", coeffs, paulis)] * 4
        ops.append(("dm_partialTrace", [0, 3, 5, 8]))                       # all-suffix: local gather-sum
        ops.append(("dm_partialTrace", [N - 4, N - 3, N - 2, N - 1]))      # top qubits: bra bits are rank bits -> relocation
        return "dm", N, ops, "config 5: expecPauliString over 256 random Pauli strings (x4) + partialTrace of 4 qubits (local / relocating)"
    raise ValueError(name)


def op_cost(op, kind, nq, k, where=None, lazy=False):
    """[(hbm_bytes, nvlink_bytes_per_direction, flops), ...] per rank for the phases of one API call -- the ALGORITHMIC work
    (SURVEY 8d), not what a particular implementation moves. nq = n (sv) or N (dm); A = amplitudes per rank.
    `where` = the state's current qubit layout (index bit of every logical qubit, host/layout.hpp): whether a gate needs an
    exchange depends on where its qubits sit NOW. lazy: a dense gate relocates its prefix targets once and leaves them
    (the undo is a separate, deferred step: "layout restore"); otherwise twice, as the reference does."""
    name = op[0]
    bits = nq if kind == "sv" else 2 * nq
    L = bits - k
    A = float(1 << L)
    if name == "layout_restore":
        return restore_cost(op[1], kind, nq, k)
    if where is not None and name.startswith("sv_"):
        m = lambda qs: [where[q] for q in qs]           # noqa: E731
        if name == "sv_oneTargGate":
            op = (name, where[op[1]]) + tuple(op[2:])
        elif name == "sv_manyCtrlOneTargGate":
            op = (name, m(op[1]), where[op[2]]) + tuple(op[3:])
        elif name == "sv_swapGate":
            op = (name, where[op[1]], where[op[2]])
        else:
            op = (name, m(op[1])) + tuple(op[2:])
    elif where is not None and name in ("dm_manyTargGate",):
        pass                                               # handled below through many_targ (ket and bra bits mapped there)
    mapq = (lambda q: where[q]) if where is not None else (lambda q: q)

    def relocation(npre):
        # npre (suffix, prefix) qubit pairs swapped in one step: every rank keeps 2^-npre of its shard and pulls the rest
        # from the other members of its group; one read + one write of the shard in HBM
        return (32 * A, (1.0 - 0.5 ** npre) * 16 * A) if npre else (0.0, 0.0)

    def many_targ(targets):
        npre = sum(1 for t in targets if t >= L)
        r = relocation(npre)
        # FP64 work in the cheapest known form of the complex product (3M: 6 * 2^t flop per amplitude; the 4M form is 8 * 2^t).
        # The reference relocates before AND after the gate (distributed_statevector.hpp:213-223): two relocations; with the
        # lazy layout one (the undo is deferred and accounted for as "layout_restore").
        return [(32 * A, 0.0, 6.0 * (1 << len(targets)) * A)] + [(r[0], r[1], 0.0)] * ((1 if lazy else 2) if npre else 0)

    if name == "sv_manyTargGate":
        return many_targ(op[1])
    if name == "sv_oneTargGate":
        return [(32 * A, 0, 0)] if op[1] < L else [(48 * A, 16 * A, 0)]
    if name == "sv_manyCtrlOneTargGate":
        ctrls, t = op[1], op[2]
        m = A / (1 << len([c for c in ctrls if c < L]))
        # ranks failing a prefix control do nothing; the work of a participating rank is counted
        return [(32 * m, 0, 0)] if t < L else [(48 * m, 16 * m, 0)]
    if name in ("sv_pauliGadget", "sv_pauliTensor"):
        prefix_xy = any(t >= L and p in (1, 2) for t, p in zip(op[1], op[2]))
        return [(48 * A, 16 * A, 0)] if prefix_xy else [(32 * A, 0, 0)]
    if name == "sv_phaseGadget":
        return [(32 * A, 0, 0)]
    if name == "sv_swapGate":
        a, b = sorted(op[1:3])
        if b < L:
            return [(16 * A, 0, 0)]
        if a >= L:
            return [] if lazy else [(32 * A, 16 * A, 0)]         # both on rank bits: a relabelling under the lazy layout
        return [(24 * A, 8 * A, 0)]
    if name == "dm_manyTargGate":
        t = len(op[1])
        if t <= 2:
            # U (x) conj(U) on the 2t bits {targets, targets + N} is ONE pass over the shard
            return many_targ([mapq(q) for q in op[1]] + [mapq(q + nq) for q in op[1]])
        return many_targ([mapq(q) for q in op[1]]) + many_targ([mapq(q + nq) for q in op[1]])
    thr = nq - k
    if name == "dm_oneQubitDepolarising":
        return [(32 * A, 0, 0)] if op[1] < thr else [(48 * A, 8 * A, 0)]
    if name == "dm_twoQubitDephasing":
        return [(28 * A, 0, 0)]            # every amplitude read, the 3/4 that change written
    if name == "dm_oneQubitDephasing":
        return [(16 * A, 0, 0)]
    if name == "dm_twoQubitDepolarising":
        a, b = sorted(op[1:3])
        return [(32 * A, 0, 0)] if b < thr else ([(32 * A + 4 * A, 2 * A, 0)] if a < thr else [(32 * A + 8 * A, 8 * A, 0)])
    if name == "dm_damping":
        return [(32 * A, 0, 0)] if op[1] < thr else [(40 * A, 0, 0, 8 * A)]      # population flows one way only
    if name == "dm_expecPauliString":
        T = len(op[1])
        return [(min(16 * A, 32.0 * T * (1 << nq) / (1 << k)), 0, 0)]
    if name == "dm_partialTrace":
        t = len(op[1])
        npre = sum(1 for q in op[1] if q + nq >= L)
        r = relocation(npre)
        return [(16 * A / (1 << t) + 16 * A / (1 << (2 * t)), 0, 0)] + ([(r[0], r[1], 0.0)] if npre else [])
    raise ValueError(name)


def bound_ms(cost, peaks):
    """sum over the phases of one call of max(HBM time, NVLink time, FP64 time). A phase is (hbm bytes, NVLink bytes per direction
    with both directions busy, flop[, NVLink bytes of a ONE-WAY transfer]); the two NVLink rates are measured separately."""
    one_way = peaks.get("nvlink_one_way_GBs", peaks["nvlink_GBs_per_dir"])
    return sum(max(c[0] / (peaks["hbm_GBs"] * 1e9), c[1] / (peaks["nvlink_GBs_per_dir"] * 1e9), c[2] / (peaks["fp64_TFLOPs"] * 1e12),
                   (c[3] if len(c) > 3 else 0.0) / (one_way * 1e9)) for c in cost) * 1e3


def restore_cost(steps, kind, nq, k):
    """cost of StateVector::restoreLayout for the plan `steps` = [(0 = relocation pair | 1 = index-bit swap, a, b), ...]"""
    L = (nq if kind == "sv" else 2 * nq) - k
    A = float(1 << L)
    out = []
    npairs = sum(1 for s in steps if s[0] == 0)
    if npairs:
        out.append((32 * A, (1.0 - 0.5 ** npairs) * 16 * A, 0.0))
    for s in steps:
        if s[0] == 1:
            a, b = sorted(s[1:3])
            out.append((16 * A, 0.0, 0.0) if b < L else ((24 * A, 8 * A, 0.0) if a < L else (32 * A, 16 * A, 0.0)))
    return out


# ------------------------------------------------------------------------------------------------ helpers

class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""

    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val == "Active":
                    reasons.add(name)
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes (read+write) per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def parse_cpulist(text):
    """'0-31,64-95' -> {0, ..., 31, 64, ..., 95} (the format of /sys/devices/system/node/nodeN/cpulist)"""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU runs: put this rank's host threads -- and with them the pinned buffer of the end-to-end leg, which the kernel
    places on the node of the thread that allocates it -- on the NUMA node its GPU hangs off (what `numactl` does for a launcher
    that knows the topology; torchrun does not). Without it a rank may sit on the other socket and every host<->device copy
    crosses the socket interconnect, which several ranks then share. Best effort, never raises; the outcome goes into the line."""
    try:
        sel = str(local_rank)
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if visible:
            ids = [x.strip() for x in visible.split(",") if x.strip()]
            if local_rank < len(ids):
                sel = ids[local_rank]                       # nvidia-smi -i takes physical indices or UUIDs
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", sel], capture_output=True, text=True, timeout=60).stdout
        bus = out.strip().splitlines()[0].strip().lower()     # "00000000:1b:00.0": sysfs uses a 4-digit domain
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return {"bound": False, "why": "the platform reports no NUMA node for %s" % bus}
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = parse_cpulist(f.read())
        target = cpus & os.sched_getaffinity(0)
        if not target:
            return {"bound": False, "why": "no allowed CPU on node %d" % node}
        os.sched_setaffinity(0, target)
        return {"bound": True, "node": node, "cpus": len(target), "gpu": bus}
    except Exception as e:                                   # noqa: BLE001 -- placement is an optimisation, not a requirement
        return {"bound": False, "why": repr(e)[:200]}


def gates_equiv(num_gates, num_qubits, seconds):
    return num_gates * (2.0 ** (num_qubits - 34)) / seconds


# ------------------------------------------------------------------------------------------------ reference arm

def run_reference(args, world, rank):
    """The reference's own CPU code path (real reference build when present, else unavailable) on a bounded sample: the SAME
    sweep on a 30-qubit state (smaller only if host RAM or the time budget forces it -- never a single pass scaled up),
    np = N ranks x (cores / N) OpenMP threads, reported in the product arm's metric (per-amplitude normalised)."""
    if rank != 0:
        return
    from oracle import refrun
    nodes = max(1, args.gpus)
    k = nodes.bit_length() - 1
    config = {"workload": "oneTargGate + manyCtrlOneTargGate sweep over all target positions (BASELINE configs[1])",
              "qubits": 32 + k, "gates_per_step": 2 * (32 + k), "parallelism": "%d-way state sharding" % nodes, "l2": "inputs >> L2"}
    line = {"impl": "reference", "metric": "34-qubit-equivalent SV gates/s", "unit": "gates/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)",
            "data": "synthetic", "config": config}
    if not refrun.available():
        line["unavailable"] = "oracle/_ref/ref_driver not built (needs /root/reference at build time)"
        print(json.dumps(line), flush=True)
        return
    cores = os.cpu_count() or 1
    threads = max(1, cores // nodes)
    budget_s = 200.0
    ram = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
    nq = REF_SAMPLE_QUBITS
    while (16 << nq) * 2 > 0.5 * ram:
        nq -= 1
    # a short probe (same sweep, 24 qubits) gives the per-amplitude cost; the sample is the largest size <= 30 qubits at which
    # one untimed + at least two timed passes fit the budget
    probe_q = 24
    probe = refrun.run("sv", probe_q, make_sweep(probe_q), num_nodes=nodes, init_seed=SEED, threads=threads, want_state=False, timed=True, timeout=600)
    per_gate_amp = probe["seconds"] / (2 * probe_q * (1 << probe_q))
    timed_passes = max(2, min(args.steps, 3))
    while nq > 26 and (1 + timed_passes) * per_gate_amp * 2 * nq * (1 << nq) * 1.15 > budget_s:
        nq -= 1
    ops = make_sweep(nq)
    times = []
    for i in range(1 + timed_passes):
        r = refrun.run("sv", nq, ops, num_nodes=nodes, init_seed=SEED, threads=threads, want_state=False, timed=True, timeout=1800)
        if i >= 1:
            times.append(r["seconds"])
    sec = float(np.mean(times))
    val = gates_equiv(len(ops), nq, sec)
    sample = "%d-qubit sweep (%d gates), %d rank(s) x %d OpenMP threads, 1 untimed + %d timed passes (%s s); %s" % (
        nq, len(ops), nodes, threads, len(times), ", ".join("%.2f" % t for t in times), os.path.basename(refrun.driver_path()))
    line.update({"value": val, "ms_per_step": sec * 1e3, "sample_qubits": nq, "gates_per_s_actual_at_sample": len(ops) / sec,
                 "amp_updates_per_s": len(ops) * float(1 << nq) / sec,
                 "cpu_baseline": {"value": val, "unit": "gates/s", "cores": threads * nodes, "kind": "reference", "sample": sample},
                 "e2e": {"value": val, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0, "steps_timed": len(times)})
    print(json.dumps(line), flush=True)


def cpu_baseline_sample():
    """Rank 0, N=1: the real reference (kind=reference) or the C port (kind=port) on a bounded sample of the sweep (28 qubits:
    about 10 s of CPU work on 16 cores; the reference ARM uses the 30-qubit sample)."""
    from oracle import refrun
    cores = os.cpu_count() or 1
    nq = 28
    ops = make_sweep(nq)
    if refrun.available():
        r = refrun.run("sv", nq, ops, num_nodes=1, init_seed=SEED, threads=cores, want_state=False, timed=True, timeout=1800)
        return {"value": gates_equiv(len(ops), nq, r["seconds"]), "unit": "gates/s", "cores": cores, "kind": "reference",
                "sample": "%d-qubit sweep, %d gates in %.2f s, 1 rank x %d OpenMP threads (unmodified reference + setBit patch)" % (nq, len(ops), r["seconds"], cores)}
    from oracle import capi
    import cases
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    st = capi.OracleState("sv", nq, 1)
    st.init_hash(SEED)
    t0 = time.time()
    for op in ops:
        cases.apply(st, op)
    sec = time.time() - t0
    return {"value": gates_equiv(len(ops), nq, sec), "unit": "gates/s", "cores": cores, "kind": "port",
            "sample": "%d-qubit sweep, %d gates in %.2f s, C restatement with OpenMP" % (nq, len(ops), sec)}


# ------------------------------------------------------------------------------------------------ product arm

class Job:
    """One rank of the product run: library handles + the torch.distributed plumbing (barrier, max over ranks)."""

    def __init__(self, world, rank, local_rank):
        self.world, self.rank, self.local_rank = world, rank, local_rank
        self.k = world.bit_length() - 1
        self.torch = self.dist = None
        # host placement (multi-GPU only: at one rank the same process later times the CPU reference on all cores)
        self.numa = bind_to_gpu_numa_node(local_rank) if world > 1 else {"bound": False, "why": "single rank"}
        if world > 1:
            import torch                                  # load torch's NCCL before ours; torch.distributed = plumbing only
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
        self.dfsa = importlib.import_module(PKG)
        self.lib = self.dfsa.device_lib()
        self.check = self.dfsa.api.check
        self.dfsa.comm_init()                             # RANK/WORLD_SIZE/LOCAL_RANK from the launcher
        assert self.dfsa.comm_size() == world and self.dfsa.comm_rank() == rank
        if world > 1:
            self.torch.cuda.set_device(local_rank)
            self.dist.init_process_group("nccl", device_id=self.torch.device("cuda", local_rank))
        self.lib.dfsa_comm_fused_active.restype = C.c_int

    def event(self):
        e = C.c_void_p()
        self.check(self.lib.dfsa_event_create(C.byref(e)))
        return e

    def record(self, e):
        self.check(self.lib.dfsa_event_record(e))

    def elapsed(self, e0, e1):
        ms = C.c_double()
        self.check(self.lib.dfsa_event_elapsed_ms(e0, e1, C.byref(ms)))
        return ms.value

    def barrier(self):
        self.dfsa.comm_synch()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(self, v):
        return -self.max_over_ranks(-float(v))

    def broadcast_obj(self, obj):
        if self.world == 1:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()
        self.dfsa.comm_end()


def measure_nvlink(job, st):
    """GB/s per direction with every pair of ranks (rank ^ 1) pulling each other's shard at once: remote loads from a kernel
    (what the fused exchange kernels do) and the copy engine. None at N = 1."""
    if job.world == 1 or job.lib.dfsa_comm_fused_active() == 0:
        return None
    out = {}
    shard_bytes = 16.0 * st.num_amps_per_node
    for mode, label in ((0, "kernel_remote_loads"), (1, "copy_engine"), (2, "kernel_remote_loads_one_way")):
        best = None
        for rep in range(3):
            ms = C.c_double()
            job.check(job.lib.dfsa_xk_measure_link(st.handle, job.rank ^ 1, mode, C.byref(ms)))
            t = job.max_over_ranks(ms.value)
            if rep > 0:
                best = t if best is None else min(best, t)
        out[label + "_GBs_per_dir"] = shard_bytes / (best * 1e-3) / 1e9
    out["how"] = "all %d pairs (rank ^ 1) pull each other's %d GiB shard simultaneously, both directions; best of 2 after warm-up, max over ranks" % (job.world // 2, int(shard_bytes) >> 30)
    return out


def parity_selfcheck(job):
    """Small-size parity on the transports the timed region uses, against the C oracle (virtual ranks), compared on the device."""
    import cases
    from oracle import capi
    k, world, rank = job.k, job.world, job.rank
    rng = np.random.default_rng(SEED + 1)
    nq = 24 + k
    L = nq - k
    sv_ops = make_sweep(nq, seed=SEED + 2)
    top = list(range(nq - 1, nq - 1 - max(k, 1), -1))                 # the prefix qubits (or just the top qubit at N = 1)
    for rep in range(2):
        t5 = top[: min(len(top), 1 + rep)] + [int(x) for x in rng.permutation(L)[: 5 - min(len(top), 1 + rep)]]
        sv_ops.append(("sv_manyTargGate", [int(x) for x in rng.permutation(t5)], haar(rng, 32)))
        sv_ops.append(("sv_pauliGadget", [top[0], 3, 11, top[-1]] if len(top) > 1 else [top[0], 3, 11], [1, 3, 2, 2][: (4 if len(top) > 1 else 3)], float(rng.uniform(-3, 3))))
        sv_ops.append(("sv_phaseGadget", [top[0], 0, 7], float(rng.uniform(-3, 3))))
        sv_ops.append(("sv_swapGate", top[0], 5 + rep))
        if k >= 2:
            sv_ops.append(("sv_swapGate", top[0], top[1]))
        sv_ops.append(("sv_pauliTensor", [top[-1], 2], [2, 1]))
    N = 12
    dm_ops = []
    for q in (0, N - 1, N - 2, 5):
        dm_ops.append(("dm_manyTargGate", [q, (q + 1) % N], haar(rng, 4)))
        dm_ops.append(("dm_oneQubitDepolarising", q, float(rng.uniform(0, 0.5))))
        dm_ops.append(("dm_twoQubitDephasing", q, (q + 1) % N, float(rng.uniform(0, 0.5))))
        dm_ops.append(("dm_damping", q, float(rng.uniform(0, 0.5))))
    dm_ops.append(("dm_twoQubitDepolarising", N - 1, 3, 0.21))
    dm_ops.append(("dm_twoQubitDepolarising", N - 1, N - 2, 0.17))
    dm_ops.append(("dm_oneQubitDephasing", N - 1, 0.1))
    suites = [("sv", nq, sv_ops, "%d-qubit sweep (%d gates) + config-3 mini-circuit" % (nq, 2 * nq)), ("dm", N, dm_ops, "12-qubit noisy density-matrix layer (config 4's ops)")]

    # the oracle's results: rank 0 computes, the others map the file
    path = "/dev/shm/dfsa_parity_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getuid())
    t0 = time.time()
    if rank == 0:
        os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
        for kind, n, ops, _ in suites:
            o = capi.OracleState(kind, n, world)
            o.init_hash(SEED)
            for op in ops:
                cases.apply(o, op)
            np.save(path + "_" + kind + ".npy", o.get_amps())
            del o
    oracle_s = time.time() - t0
    job.barrier()
    if world > 1:
        job.dist.barrier()
    results, worst, ok = [], 0.0, True
    modes = [("fused", 1), ("staged", 0)] if world > 1 else [("single", -1)]
    for kind, n, ops, what in suites:
        want = np.load(path + "_" + kind + ".npy", mmap_mode="r")
        st = job.dfsa.DeviceState(kind, n)
        ref = job.dfsa.DeviceState(kind, n)
        A = st.num_amps_per_node
        ref.set_local_amps(np.ascontiguousarray(want[rank * A:(rank + 1) * A]))
        for label, mode in modes:
            if world > 1:
                job.check(job.lib.dfsa_comm_set_fused(mode))
            active = job.lib.dfsa_comm_fused_active()
            st.init_hash(SEED)
            for op in ops:
                cases.apply(st, op)
            d, ne, mr = st.compare(ref)
            good = bool(d <= PARITY_TOL * max(1.0, mr))
            ok = ok and good
            worst = max(worst, d if d == d else float("inf"))
            results.append({"what": what, "exchange": label, "fused_active": active, "max_abs_diff": d, "max_abs_ref": mr, "ok": good})
        st.close()
        ref.close()
    if world > 1:
        job.check(job.lib.dfsa_comm_set_fused(-1))
        job.dist.barrier()
    if rank == 0:
        for kind, _, _, _ in suites:
            try:
                os.unlink(path + "_" + kind + ".npy")
            except OSError:
                pass
    return {"ok": ok, "max_abs_diff": worst, "tol": "%g * max(1, max|ref|)" % PARITY_TOL, "transport": job.lib.dfsa_comm_transport().decode(),
            "checker": "oracle/dfsa_oracle.c with %d virtual rank(s) (%.1f s on rank 0), compared on the device (dfsa_state_compare)" % (world, oracle_s),
            "cases": results}


def relocation_cost(num_pairs, shard_amps):
    """One relocation step of `num_pairs` disjoint (suffix bit, rank bit) pairs (dfsa_xk_relocate): an out-of-place pass over the shard
    that gathers from the 2^k shards of the rank's group -- (HBM bytes, NVLink bytes per direction, flop)."""
    return (32.0 * shard_amps, (1.0 - 0.5 ** num_pairs) * 16.0 * shard_amps, 0.0)


def fused_plan_summary(steps, L, count_passes):
    """Pure part of the fused-step accounting: from the flush plan (api.plan_flush format) the relocation steps (pairs each), the
    number of passes over HBM (count_passes(gates) per run of gates) and a short description."""
    relocation_pairs, passes, summary = [], 0, []
    for kind, body in steps:
        if kind == "relocate":
            relocation_pairs.append(len(body))
            summary.append("relocate %d pair(s) %s" % (len(body), body))
        else:
            p = count_passes([(t, [c for c in range(L) if (mask >> c) & 1]) for t, mask in body])
            passes += p
            summary.append("%d gates in %d pass(es)" % (len(body), p))
    return relocation_pairs, passes, summary


def fused_accounting(plans, ops, lib, L, check):
    """The flush plans the host layer reported for the timed steps (api.plan_pending_flush, taken just before every flush) -> per-step
    relocation steps and passes over HBM. Never raises: a failure of the accounting must not cost the measurement (it is reported
    in the line instead)."""
    class G(C.Structure):
        _fields_ = [("matrix", C.c_double * 8), ("ctrlMask", C.c_uint64), ("target", C.c_uint32), ("reserved", C.c_uint32)]

    def count_passes(segment):
        """passes over HBM the library makes for a run of (index-bit target, index-bit suffix controls) gates"""
        if not segment:
            return 0
        arr = (G * len(segment))()
        for i, (t, ctrls) in enumerate(segment):
            arr[i].target = t
            arr[i].ctrlMask = sum(1 << c for c in ctrls if c < L)
        nb = C.c_uint()
        scratch = [(C.c_uint32 * max(1, 11 * len(segment)))() for _ in range(5)]
        check(lib.dfsa_plan_gateSequence(arr, len(segment), L, scratch[0], scratch[1], scratch[2], scratch[3], scratch[4], C.byref(nb)))
        return nb.value

    shard_amps = float(1 << L)
    pairs = sum((shard_amps / 2.0) / (1 << len([] if op[0] == "sv_oneTargGate" else op[1])) for op in ops)     # averaged over ranks for rank-bit controls
    out = {"relocation_pairs": [], "passes": [], "pairs": pairs, "summary": [], "error": None}
    for plan in plans or []:
        if isinstance(plan, str):
            out["error"] = plan
            continue
        try:
            rel, passes, summary = fused_plan_summary(plan, L, count_passes)
        except Exception as e:                               # noqa: BLE001 -- reported, not fatal
            out["error"] = "fused_plan_summary failed: %r" % (e,)
            continue
        out["relocation_pairs"].append(rel)
        out["passes"].append(passes)
        out["summary"].append(summary)
    return out


def run_config(job, name, peaks, reps, dm_qubits=None):
    """One BASELINE config: per-op device times (max over ranks) and roofline fractions."""
    import cases
    kind, nq, ops, desc = config_workload(name, job.world, dm_qubits=dm_qubits)
    k = job.k
    st = job.dfsa.DeviceState(kind, nq)
    st.init_hash(SEED)
    lazy = bool(job.dfsa.host_lib().dfsa_host_lazyLayoutEnabled())
    per_type, total = {}, []
    for rep in range(reps + 1):                              # rep 0 = warm-up
        evs = [(job.event(), job.event()) for _ in range(len(ops) + 1)]
        costs = []
        job.barrier()
        def restore_steps():
            """what restoreLayout would do now (host/layout.hpp planRestore)"""
            where = st.layout()
            n = len(where)
            out = (C.c_uint * (9 * n))()
            hl = job.dfsa.host_lib()
            hl.dfsa_host_plan_restoreLayout.restype = C.c_uint
            cnt = hl.dfsa_host_plan_restoreLayout((C.c_uint * n)(*where), n, (nq if kind == "sv" else 2 * nq) - k, out)
            return [(out[3 * i], out[3 * i + 1], out[3 * i + 2]) for i in range(cnt)]

        for (e0, e1), op in zip(evs, ops):
            cost = op_cost(op, kind, nq, k, where=st.layout(), lazy=lazy)            # what the gate needs given where its qubits sit now
            if op[0] in INDEX_ADDRESSED_OPS:
                cost = restore_cost(restore_steps(), kind, nq, k) + cost             # these ops put a lazily relabelled layout back first
            costs.append(cost)
            job.record(e0)
            r = cases.apply(st, op)
            job.record(e1)
            if op[0] == "dm_partialTrace":
                r.close()
                st.init_hash(SEED)                           # partialTrace mutates its input: same contents for the next op (the shard and
                                                             # its peer mappings stay: a fresh allocation would be re-mapped inside the next timed op)
        # the deferred part of the pass: put the relocated qubits back (nothing to do unless a dense gate left some displaced)
        steps = restore_steps()
        displaced = len(steps)
        restore_op = ("layout_restore", steps)
        costs.append(op_cost(restore_op, kind, nq, k))
        job.record(evs[-1][0])
        st.restore_layout()
        job.record(evs[-1][1])
        job.barrier()
        if rep == 0:
            continue
        tot = 0.0
        for (e0, e1), op, cost in zip(evs, list(ops) + [restore_op], costs):
            v = job.max_over_ranks(job.elapsed(e0, e1))
            tot += v
            label = op[0]
            if label == "layout_restore":
                if not displaced:
                    continue
                label = "layout restore (deferred undo of relocations)"
            if op[0] == "dm_partialTrace":
                label += " (relocating)" if max(op[1]) + nq >= (2 * nq - k) else " (local)"
            elif any(c[1] > 0 or (len(c) > 3 and c[3] > 0) for c in cost) and op[0] != "layout_restore":      # NVLink bytes, both ways or one way
                label += " [exchange]"
            d = per_type.setdefault(label, {"n": 0, "ms": 0.0, "bound_ms": 0.0})
            d["n"] += 1
            d["ms"] += v
            d["bound_ms"] += bound_ms(cost, peaks)
        total.append(tot)
    st.close()
    step_ms = float(np.mean(total))
    bound_total = sum(d["bound_ms"] for d in per_type.values()) / reps
    return {"what": desc, "qubits": nq, "kind": kind, "ops_per_pass": len(ops), "passes_timed": reps, "lazy_layout": lazy, "ms_per_pass": step_ms,
            "gates_per_s": len(ops) / (step_ms * 1e-3), "roofline_ms_per_pass": bound_total, "roofline_frac": bound_total / step_ms,
            "per_op": {lab: {"count_per_pass": d["n"] // reps, "mean_ms": d["ms"] / d["n"], "roofline_frac": d["bound_ms"] / d["ms"]} for lab, d in sorted(per_type.items())}}


def run_product(args, world, rank, local_rank):
    import cases
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version there) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    job = Job(world, rank, local_rank)
    lib, check, dfsa = job.lib, job.check, job.dfsa
    k = job.k

    # ---- parity first (small states, real transports)
    parity = None if args.skip_parity else parity_selfcheck(job)

    nq = args.qubits if args.qubits else 32 + k
    ops = make_sweep(nq)
    L = nq - k
    shard_amps = 1 << L
    st = dfsa.DeviceState("sv", nq)
    st.init_hash(SEED)

    hbm_peak, peak_src = measured_peak()
    nvlink = measure_nvlink(job, st)
    peaks = {"hbm_GBs": hbm_peak, "fp64_TFLOPs": FP64_PEAK_TFLOPS,
             "nvlink_GBs_per_dir": nvlink["kernel_remote_loads_GBs_per_dir"] if nvlink else NVLINK_FALLBACK_GBS,
             "nvlink_one_way_GBs": nvlink["kernel_remote_loads_one_way_GBs_per_dir"] if nvlink else NVLINK_FALLBACK_GBS,
             "nvlink_source": "measured in this run (kernel remote loads, all pairs at once)" if nvlink else "not used at 1 GPU (fallback %g)" % NVLINK_FALLBACK_GBS,
             "hbm_source": peak_src, "fp64_source": "profiles/r01_fp64_peak_b200.jsonl (DFMA 36.6, DMMA 37.0 TFLOP/s)"}

    def run_step(per_gate=None, the_ops=ops):
        for i, op in enumerate(the_ops):
            if per_gate is not None:
                job.record(per_gate[i][0])
            cases.apply(st, op)
            if per_gate is not None:
                job.record(per_gate[i][1])

    # ---- full-size self-consistency (also the first warm-up passes): the sweep, then its inverse, against the initial state
    selfcheck = None
    if not args.skip_parity:
        inv = inverse_ops(ops)
        run_step()
        run_step(the_ops=inv)
        d, ne, mr = st.compare_hash(SEED)
        selfcheck = {"what": "%d-qubit sweep then its inverse (%d gates) vs the regenerated initial state, all 2^%d amplitudes, on the device" % (nq, 2 * len(ops), nq),
                     "max_abs_diff": d, "max_abs_ref": mr, "ok": bool(d <= 1e-12)}
        st.init_hash(SEED)
    for _ in range(max(args.warmup, 3)):
        run_step()
    job.barrier()

    # ---- timed region (the product's default behaviour): K steps, device-timed. With gate fusion on (default) the library defers
    #      the one-target gates of a step and launches them as a few shared passes over HBM when the step is flushed.
    fusion = dfsa.gate_fusion_enabled()

    def timed_steps(steps, per_gate=None, plans=None):
        e0, e1 = job.event(), job.event()
        launches0 = lib.dfsa_launch_count()
        job.barrier()
        job.record(e0)
        for s_ in range(steps):
            run_step(per_gate[s_] if per_gate else None)
            if plans is not None:
                # what this flush is about to do (host-only, microseconds): relocation steps and runs of gates, for the roofline
                try:
                    plans.append(st.plan_pending_flush()[0] if st.pending_gates() == len(ops) else
                                 "%d of %d gates pending before the flush: plan not taken" % (st.pending_gates(), len(ops)))
                except Exception as e:                       # noqa: BLE001 -- the accounting must not cost the measurement
                    plans.append("plan_pending_flush failed: %r" % (e,))
            st.flush()                                       # a step ends with its gates launched (no fusing across steps)
        job.record(e1)
        job.barrier()
        return job.max_over_ranks(job.elapsed(e0, e1)), int(lib.dfsa_launch_count() - launches0)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    flush_plans = [] if fusion else None
    total_ms, launches = timed_steps(args.steps, plans=flush_plans)
    clocks = sampler.stop() if rank == 0 else None
    step_ms = total_ms / args.steps
    value = gates_equiv(len(ops), nq, step_ms * 1e-3)

    # ---- the same sweep gate by gate (fusion off): one kernel per gate, per-gate events -> the per-gate roofline numbers.
    #      Index order first: the default mode leaves rank-bit qubits swapped into the shard, and in this mode the gates on the top
    #      log2 N qubits are meant to be the exchange gates.
    st.restore_layout()
    dfsa.set_gate_fusion(False)
    pg_steps = max(2, min(args.steps, 5))
    run_step()
    per_gate = [[(job.event(), job.event()) for _ in ops] for _ in range(pg_steps)]
    pg_total_ms, pg_launches = timed_steps(pg_steps, per_gate)
    pg_step_ms = pg_total_ms / pg_steps
    # dominant per-gate kernel: local oneTargGate (ctrlOneTarg kernel with no controls), 32*A bytes per launch
    one_ms = [job.elapsed(per_gate[s][i][0], per_gate[s][i][1]) for s in range(pg_steps) for i, op in enumerate(ops) if op[0] == "sv_oneTargGate" and op[1] < L]
    one_avg_ms = float(np.mean(one_ms))
    gate_ms = [job.max_over_ranks(float(np.mean([job.elapsed(per_gate[s][i][0], per_gate[s][i][1]) for s in range(pg_steps)]))) for i in range(len(ops))] if (args.per_gate or world > 1) else None
    exchange_summary = None
    if gate_ms is not None:
        ex = [(i, op) for i, op in enumerate(ops) if any(c[1] > 0 for c in op_cost(op, "sv", nq, k))]
        if ex:
            t = sum(gate_ms[i] for i, _ in ex)
            b = sum(bound_ms(op_cost(op, "sv", nq, k), peaks) for _, op in ex)
            nvb = sum(sum(c[1] for c in op_cost(op, "sv", nq, k)) for _, op in ex)
            exchange_summary = {"gates": len(ex), "ms": t, "bound_ms": b, "frac_of_measured_nvlink_line": b / t, "achieved_GBs_per_dir": nvb / (t * 1e-3) / 1e9,
                                "note": "per-gate device time, max over ranks: includes waiting for a partner that arrives late (ranks gated out by a prefix control run ahead)"}
            # the exchange kernels alone (one untimed pass): device time between the partners' READY and this rank's DONE
            kt = 0.0
            for _, op in ex:
                cases.apply(st, op)
                ms = C.c_double()
                check(lib.dfsa_comm_last_exchange_ms(C.byref(ms)))
                kt += job.max_over_ranks(ms.value)
            exchange_summary["kernels_only"] = {"ms": kt, "frac_of_measured_nvlink_line": b / kt, "achieved_GBs_per_dir": nvb / (kt * 1e-3) / 1e9}
        if args.per_gate and rank == 0:
            for i, op in enumerate(ops):
                what = "%s t=%d%s" % (op[0][3:], op[1] if op[0] == "sv_oneTargGate" else op[2], "" if op[0] == "sv_oneTargGate" else " ctrls=%s" % (op[1],))
                b = bound_ms(op_cost(op, "sv", nq, k), peaks)
                sys.stderr.write("gate %2d %-48s %9.3f ms  (roofline %8.3f ms, %5.1f%%)\n" % (i, what, gate_ms[i], b, 100 * b / gate_ms[i]))
    pg_achieved = 32.0 * shard_amps / (one_avg_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    bound_step = sum(bound_ms(op_cost(op, "sv", nq, k), peaks) for op in ops)
    per_gate_mode = {
        "what": "DFSA_FUSE_GATES=0: every gate is its own kernel and its own sweep over the shard, as in the reference",
        "value": gates_equiv(len(ops), nq, pg_step_ms * 1e-3), "ms_per_step": pg_step_ms, "steps": pg_steps, "gates_per_s_actual": len(ops) / (pg_step_ms * 1e-3),
        "gpu_launches": pg_launches,
        "step_roofline": {"bound_ms": bound_step, "frac": bound_step / pg_step_ms,
                          "how": "sum over gates of max(HBM bytes / HBM peak, NVLink bytes per direction / measured NVLink rate): the bound of any one-sweep-per-gate implementation"},
        "roofline": {"bound": "hbm", "kernel": "streamKernel<ctrlOneTarg> (local oneTargGate)", "achieved": pg_achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": pg_achieved / hbm_peak, "algorithmic_bytes_per_launch": 32 * shard_amps, "avg_launch_ms": one_avg_ms, "launches_timed": len(one_ms),
                     "traffic": (traffic["dram_over_algorithmic"] * 32 * shard_amps) if traffic else None,
                     "traffic_note": ("NOT measured in this run: " + traffic.get("note", "")) if traffic else None}}
    dfsa.set_gate_fusion(fusion)

    # roofline of the dominant kernel of the timed region
    if fusion:
        # fused passes: what must cross HBM for a pass is one read and one write of the shard, however many gates it carries;
        # the FP64 work of its gates (16 FMA per touched pair) is the other bound.
        # In this mode gates on rank-bit qubits are queued like the others; when the step is flushed the host layer brings those
        # qubits into the shard (one relocation step for all it can, host/layout.hpp planFlush) and the layout remembers it. Which
        # gates that hits depends on the layout the previous step left behind, so the host layer was asked for the plan of every
        # timed flush (exactly what flush() then did); the bound is computed from those plans, averaged over the timed steps.
        fused_plan = fused_accounting(flush_plans, ops, lib, L, check)
        n = len(ops)
        planned = max(1, len(fused_plan["passes"]))
        relocation_pairs = [m for step in fused_plan["relocation_pairs"] for m in step]
        passes = max(1.0, sum(fused_plan["passes"]) / planned)                 # per step
        exch_ms = sum(bound_ms([relocation_cost(m, shard_amps)], peaks) for m in relocation_pairs) / planned
        local_ms = max(step_ms - exch_ms, 1e-9)            # relocations at their bound: a lower bound on what the passes took
        achieved = 32.0 * shard_amps * passes / (local_ms * 1e-3) / 1e9
        fp64_ms = fused_plan["pairs"] * 32.0 / (FP64_PEAK_TFLOPS * 1e12) * 1e3
        roofline = {"bound": "hbm", "kernel": "fusedGateTileKernel (all one-target gates of a pass applied to 32 KiB tiles in shared memory)", "achieved": achieved,
                    "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src, "algorithmic_bytes_per_launch": 32 * shard_amps,
                    "avg_launch_ms": local_ms / passes, "passes_per_step": passes, "gates_in_passes": n, "relocations_per_step": len(relocation_pairs) / planned,
                    "passes_by_step": fused_plan["passes"], "relocation_pairs_by_step": fused_plan["relocation_pairs"],
                    "flush_plan_first_timed_step": fused_plan["summary"][0] if fused_plan["summary"] else None,
                    "fp64_bound_ms_per_step": fp64_ms, "hbm_bound_ms_per_step": 32.0 * shard_amps * passes / (hbm_peak * 1e9) * 1e3,
                    "note": "algorithmic bytes of a fused pass = one read + one write of the shard (32*A), whatever the number of gates it carries; "
                            "time per launch = (step time - relocations at their roofline) / passes, measured with CUDA events around whole steps",
                    "traffic": (traffic["fused"]["dram_over_algorithmic"] * 32 * shard_amps) if traffic and "fused" in traffic else None,
                    "traffic_note": ("NOT measured in this run: " + traffic["fused"].get("note", "")) if traffic and "fused" in traffic else None}
        if fused_plan.get("error"):
            roofline["accounting_error"] = fused_plan["error"]
        step_roofline = {"bound_ms": max(roofline["hbm_bound_ms_per_step"], fp64_ms) + exch_ms, "frac": (max(roofline["hbm_bound_ms_per_step"], fp64_ms) + exch_ms) / step_ms,
                         "how": "fused passes: max(passes x 32*A / HBM peak, FP64 work of all gates / FP64 peak) + one relocation step per group of rank-bit qubits brought into the shard "
                                "(k pairs: (1 - 2^-k) * 16*A bytes per direction at the measured NVLink rate, 32*A bytes of HBM)",
                         "speedup_over_per_gate_roofline": bound_step / step_ms}
    else:
        roofline = dict(per_gate_mode["roofline"], peak_source=peak_src)
        step_roofline = per_gate_mode["step_roofline"]

    # ---- BASELINE config 3 (state-vector circuit; its own state -- the sweep state is needed again for e2e at the same size)
    configs = {}
    if not args.skip_configs:
        st.close()
        st = None
        configs["config3_circuit"] = run_config(job, "circuit", peaks, reps=2)
        st = dfsa.DeviceState("sv", nq)
        st.init_hash(SEED)

    # ---- e2e: the same step through the host API with HOST buffers: pinned-host -> HBM upload of the shard,
    #      the sweep, HBM -> pinned-host download of the result, every step
    shard_bytes = 16 * shard_amps
    e2e = None
    host = C.c_void_p()
    # every rank pins a shard-sized host buffer: only when the box has the RAM for it (a box driven out of memory is lost)
    mem_avail = 0
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable"):
                    mem_avail = int(ln.split()[1]) * 1024
    except OSError:
        pass
    ram_ok = job.min_over_ranks(1.0 if shard_bytes * world <= 0.6 * mem_avail else 0.0) > 0.5
    if not ram_ok:
        e2e = {"value": None, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "skipped": "host RAM too small to pin %d x %d GiB (MemAvailable %d GiB)" % (world, shard_bytes >> 30, mem_avail >> 30)}
    if ram_ok and lib.dfsa_host_alloc_pinned(C.c_uint64(shard_bytes), C.byref(host)) == 0:
        hp = C.cast(host, C.POINTER(C.c_double))
        st.download_shard_to(hp)                             # fill the host buffer (untimed)
        e2e_steps = max(1, min(args.steps, 2))
        job.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            st.upload_shard_from(hp)
            run_step()
            st.download_shard_to(hp)                         # launches the deferred gates, restores index order, copies out
        job.barrier()
        e2e_s = job.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e = {"value": gates_equiv(len(ops), nq, e2e_s), "unit": "gates/s", "h2d_bytes_per_step": shard_bytes * world,
               "d2h_bytes_per_step": shard_bytes * world, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "what": "per step: upload of every rank's shard from pinned host memory, the sweep through the host C++ API, download of the result"}
        lib.dfsa_host_free_pinned(host)
    norm2 = st.norm2()
    st.close()

    # ---- BASELINE configs 4 and 5 (density matrices; the 64 GiB state-vector shard is released first)
    if not args.skip_configs:
        for N in ((14, 16) if world == 1 else (16,)):
            configs["config4_noisy_dm_%dq" % N] = run_config(job, "dm", peaks, reps=2, dm_qubits=N)
        for N in ((14, 16) if world == 1 else (16,)):
            configs["config5_expec_ptrace_%dq" % N] = run_config(job, "expec", peaks, reps=2, dm_qubits=N)

    if rank == 0:
        line = {
            "metric": "34-qubit-equivalent SV gates/s", "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {"workload": "oneTargGate + manyCtrlOneTargGate sweep over all target positions (BASELINE configs[1])",
                       "qubits": nq, "gates_per_step": len(ops), "shard_GiB_per_gpu": shard_bytes / 2 ** 30,
                       "parallelism": "%d-way state sharding (top %d qubits = rank)" % (world, k), "transport": lib.dfsa_comm_transport().decode(),
                       "exchange": {0: "staged", 1: "fused, host-synchronised", 2: "fused, stream-ordered"}[lib.dfsa_comm_fused_active()] if world > 1 else "none",
                       "gate_fusion": "one-target gates deferred and applied in shared passes over HBM (DFSA_FUSE_GATES=0 turns it off)" if fusion else "off",
                       "l2": "inputs >> L2 (every gate streams the whole %d GiB shard)" % (shard_bytes >> 30),
                       "host_numa_binding": getattr(job, "numa", None)},
            "gates_per_s_actual": len(ops) / (step_ms * 1e-3),
            "amp_updates_per_s": len(ops) * float(1 << nq) / (step_ms * 1e-3),
            "gate_fusion": fusion, "step_roofline": step_roofline, "roofline": roofline, "per_gate_mode": per_gate_mode,
            "peaks": peaks, "nvlink": nvlink, "exchange_gates": exchange_summary,
            "parity": parity, "selfcheck": selfcheck, "configs": configs,
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "norm2_after": norm2,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    job.close()
    bad = (parity is not None and not parity["ok"]) or (selfcheck is not None and not selfcheck["ok"])
    if bad:
        sys.stderr.write("bench.py: PARITY FAILURE (see the parity / selfcheck blocks of the JSON line)\n")
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="override the state size (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-parity", action="store_true", help="no parity / self-consistency blocks (debugging only)")
    ap.add_argument("--skip-configs", action="store_true", help="no BASELINE configs 3-5 block (debugging only)")
    ap.add_argument("--per-gate", action="store_true", help="print the mean device time of every gate of the step to stderr")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, world, rank)
        return
    if world != args.gpus and not (world == 1 and args.gpus == 1):
        raise SystemExit("launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world))
    run_product(args, world, rank, local_rank)


if __name__ == "__main__":
    main()
