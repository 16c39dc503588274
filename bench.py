#!/usr/bin/env python
"""bench.py -- the hot path measured on B200s (contract: the driver's `python bench.py --gpus N --steps K --warmup W`).

Workload (BASELINE.json configs[1], the largest single-GPU configuration; weak scaling for N > 1):
  a (32 + log2 N)-qubit state vector, 64 GiB shard per GPU; one STEP = one sweep over ALL target positions,
  each position getting a oneTargGate and a manyCtrlOneTargGate with 1-3 controls (64 + 2 log2 N gates).
  Prefix targets (the top log2 N qubits) exercise the NVLink pairwise exchange.
Metric: "34-qubit-equivalent SV gates/s" = amplitude-updates per second / 2^34 (so that the number is a whole-job
  throughput that adds up over GPUs under weak scaling; `gates_per_s_actual` is the plain rate at the actual size).
Timing: CUDA events on the library's compute stream, barrier + device sync on both sides, max over ranks.
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
REF_SAMPLE_QUBITS = 28          # bounded CPU sample of the sweep (4 GiB state)
SEED = 20261017


# ------------------------------------------------------------------------------------------------ workload

def haar_2x2(rng):
    q, r = np.linalg.qr(rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


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


def op_algorithmic_bytes(op, num_qubits, log_ranks):
    """per-rank HBM bytes of one gate (A = amps per rank); (hbm_bytes, nvlink_bytes_per_direction)."""
    L = num_qubits - log_ranks
    A = 1 << L
    if op[0] == "sv_oneTargGate":
        return (32 * A, 0) if op[1] < L else (48 * A, 16 * A)
    ctrls, t = op[1], op[2]
    suffix = [c for c in ctrls if c < L]
    # ranks failing a prefix control do nothing; count the work of a participating rank
    if t < L:
        return (32 * A >> len(suffix), 0)
    if not suffix:
        return (48 * A, 16 * A)
    m = A >> len(suffix)
    return (16 * m + 16 * m + 48 * m, 16 * m)       # pack (r+w), send/recv staging, combine


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


def gates_equiv(num_gates, num_qubits, seconds):
    return num_gates * (2.0 ** (num_qubits - 34)) / seconds


# ------------------------------------------------------------------------------------------------ reference arm

def run_reference(args, world, rank):
    """The reference's own CPU code path (real reference build when present, else unavailable) on a bounded sample."""
    if rank != 0:
        return
    from oracle import refrun
    nodes = max(1, args.gpus)
    k = nodes.bit_length() - 1
    nq = REF_SAMPLE_QUBITS + k
    while (16 << nq) * 2 > 0.5 * os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES"):
        nq -= 1
    config = {"workload": "oneTargGate + manyCtrlOneTargGate sweep over all target positions (BASELINE configs[1])",
              "qubits": 32 + k, "gates_per_step": 2 * (32 + k), "parallelism": "%d-way state sharding" % nodes, "l2": "inputs >> L2"}
    line = {"impl": "reference", "metric": "34-qubit-equivalent SV gates/s", "unit": "gates/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)",
            "data": "synthetic", "config": config}
    if not refrun.available():
        line["unavailable"] = "oracle/_ref/ref_driver not built (needs /root/reference at build time)"
        print(json.dumps(line), flush=True)
        return
    ops = make_sweep(nq)
    cores = os.cpu_count() or 1
    threads = max(1, cores // nodes)
    times = []
    for i in range(args.warmup + args.steps):
        if i >= 1 and (time.time() - t_start) > 240:      # keep the whole arm within a few minutes
            break
        if i == 0:
            t_start = time.time()
        r = refrun.run("sv", nq, ops, num_nodes=nodes, init_seed=SEED, threads=threads, want_state=False, timed=True, timeout=1800)
        if i >= min(args.warmup, 1):
            times.append(r["seconds"])
    sec = float(np.mean(times))
    val = gates_equiv(len(ops), nq, sec)
    sample = "%d-qubit sweep (%d gates), %d rank(s) x %d OpenMP threads, %d timed pass(es); %s" % (
        nq, len(ops), nodes, threads, len(times), os.path.basename(refrun.driver_path()))
    line.update({"value": val, "ms_per_step": sec * 1e3, "sample_qubits": nq, "gates_per_s_actual_at_sample": len(ops) / sec,
                 "cpu_baseline": {"value": val, "unit": "gates/s", "cores": threads * nodes, "kind": "reference", "sample": sample},
                 "e2e": {"value": val, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0, "steps_timed": len(times)})
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(budget_s=25.0):
    """Rank 0, N=1: the real reference (kind=reference) or the C port (kind=port) on a bounded sample of the sweep."""
    from oracle import refrun
    cores = os.cpu_count() or 1
    nq = REF_SAMPLE_QUBITS
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

def run_product(args, world, rank, local_rank):
    import cases
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version there) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch                                  # load torch's NCCL before ours; torch.distributed = plumbing only
        import torch.distributed as dist
    dfsa = importlib.import_module(PKG)
    lib = dfsa.device_lib()
    check = dfsa.api.check
    dfsa.comm_init()                                  # RANK/WORLD_SIZE/LOCAL_RANK from the launcher
    assert dfsa.comm_size() == world and dfsa.comm_rank() == rank
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    k = world.bit_length() - 1
    nq = args.qubits if args.qubits else 32 + k
    ops = make_sweep(nq)
    L = nq - k
    shard_amps = 1 << L
    st = dfsa.DeviceState("sv", nq)
    st.init_hash(SEED)

    def event():
        e = C.c_void_p()
        check(lib.dfsa_event_create(C.byref(e)))
        return e

    def barrier():
        dfsa.comm_synch()

    def run_step(per_gate=None):
        for i, op in enumerate(ops):
            if per_gate is not None:
                check(lib.dfsa_event_record(per_gate[i][0]))
            cases.apply(st, op)
            if per_gate is not None:
                check(lib.dfsa_event_record(per_gate[i][1]))

    for _ in range(max(args.warmup, 3)):
        run_step()
    barrier()

    # ---- timed region: K steps, device-timed, per-gate events inside for the roofline of the dominant kernel
    per_gate = [[(event(), event()) for _ in ops] for _ in range(args.steps)]
    e0, e1 = event(), event()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.dfsa_launch_count()
    barrier()
    check(lib.dfsa_event_record(e0))
    for s in range(args.steps):
        run_step(per_gate[s])
    check(lib.dfsa_event_record(e1))
    barrier()
    launches = int(lib.dfsa_launch_count() - launches0)
    clocks = sampler.stop() if rank == 0 else None
    ms = C.c_double()
    check(lib.dfsa_event_elapsed_ms(e0, e1, C.byref(ms)))
    total_ms = ms.value
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # dominant kernel: local oneTargGate (ctrlOneTarg kernel with no controls), 32*A bytes per launch
    one_ms = []
    for s in range(args.steps):
        for i, op in enumerate(ops):
            if op[0] == "sv_oneTargGate" and op[1] < L:
                check(lib.dfsa_event_elapsed_ms(per_gate[s][i][0], per_gate[s][i][1], C.byref(ms)))
                one_ms.append(ms.value)
    one_avg_ms = float(np.mean(one_ms))
    if args.per_gate and rank == 0:
        for i, op in enumerate(ops):
            acc = 0.0
            for s_ in range(args.steps):
                check(lib.dfsa_event_elapsed_ms(per_gate[s_][i][0], per_gate[s_][i][1], C.byref(ms)))
                acc += ms.value
            hb, nb = op_algorithmic_bytes(op, nq, k)
            what = "%s t=%d%s" % (op[0][3:], op[1] if op[0] == "sv_oneTargGate" else op[2], "" if op[0] == "sv_oneTargGate" else " ctrls=%s" % (op[1],))
            bound = max(hb / (measured_peak()[0] * 1e9), nb / 770e9) * 1e3
            sys.stderr.write("gate %2d %-48s %9.3f ms  (roofline %8.3f ms, %5.1f%%)\n" % (i, what, acc / args.steps, bound, 100 * bound / (acc / args.steps)))
    peak, peak_src = measured_peak()
    achieved = 32.0 * shard_amps / (one_avg_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    # whole-step roofline: sum over gates of max(HBM bytes / HBM peak, NVLink bytes / 770 GB/s)
    bound_ms = 0.0
    for op in ops:
        hb, nb = op_algorithmic_bytes(op, nq, k)
        bound_ms += max(hb / (peak * 1e9), nb / 770e9) * 1e3
    step_ms = total_ms / args.steps
    value = gates_equiv(len(ops), nq, step_ms * 1e-3)

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
    ram_ok = shard_bytes * world <= 0.6 * mem_avail
    if world > 1:
        t = torch.tensor([1.0 if ram_ok else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ram_ok = bool(t.item() > 0.5)
    if not ram_ok:
        e2e = {"value": None, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "skipped": "host RAM too small to pin %d x %d GiB (MemAvailable %d GiB)" % (world, shard_bytes >> 30, mem_avail >> 30)}
    if ram_ok and lib.dfsa_host_alloc_pinned(C.c_uint64(shard_bytes), C.byref(host)) == 0:
        hp = C.cast(host, C.POINTER(C.c_double))
        check(lib.dfsa_state_download(st.handle, 0, C.c_uint64(0), C.c_uint64(shard_amps), hp))     # fill the host buffer (untimed)
        e2e_steps = max(1, min(args.steps, 2))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            check(lib.dfsa_state_upload(st.handle, 0, C.c_uint64(0), C.c_uint64(shard_amps), hp))
            run_step()
            check(lib.dfsa_state_download(st.handle, 0, C.c_uint64(0), C.c_uint64(shard_amps), hp))
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": gates_equiv(len(ops), nq, e2e_s), "unit": "gates/s", "h2d_bytes_per_step": shard_bytes * world,
               "d2h_bytes_per_step": shard_bytes * world, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "what": "per step: upload of every rank's shard from pinned host memory, the sweep through the host C++ API, download of the result"}
        lib.dfsa_host_free_pinned(host)
    norm2 = st.norm2()

    if rank == 0:
        line = {
            "metric": "34-qubit-equivalent SV gates/s", "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {"workload": "oneTargGate + manyCtrlOneTargGate sweep over all target positions (BASELINE configs[1])",
                       "qubits": nq, "gates_per_step": len(ops), "shard_GiB_per_gpu": shard_bytes / 2 ** 30,
                       "parallelism": "%d-way state sharding (top %d qubits = rank)" % (world, k), "transport": lib.dfsa_comm_transport().decode(),
                       "l2": "inputs >> L2 (every gate streams the whole %d GiB shard)" % (shard_bytes >> 30)},
            "gates_per_s_actual": len(ops) / (step_ms * 1e-3),
            "step_roofline": {"bound_ms": bound_ms, "frac": bound_ms / step_ms, "how": "sum over gates of max(HBM bytes/peak, NVLink bytes per direction/770 GB/s)"},
            "roofline": {"bound": "hbm", "kernel": "streamKernel<ctrlOneTarg> (local oneTargGate)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "algorithmic_bytes_per_launch": 32 * shard_amps,
                         "avg_launch_ms": one_avg_ms, "launches_timed": len(one_ms),
                         "traffic": (traffic["dram_over_algorithmic"] * 32 * shard_amps) if traffic else None,
                         "traffic_note": (traffic or {}).get("note")},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "norm2_after": norm2,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    st.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    dfsa.comm_end()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="override the state size (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
