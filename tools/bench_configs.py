#!/usr/bin/env python
"""Measures BASELINE.json configs 3-5 (bench.py itself covers configs[1], the contract workload):

  circuit  config 3: 34-qubit-class state-vector random circuit (manyTargGate with 5 targets / pauliGadget / phaseGadget,
           random targets, so prefix targets force NVLink exchange and relocation). 32/33/34/34 qubits at 1/2/4/8 GPUs.
  dm       config 4: noisy density-matrix layer (manyTargGate t=2 + oneQubitDepolarising + twoQubitDephasing + damping on
           every qubit). N = 14 at 1 GPU, 16 at 2/4/8.
  expec    config 5: expecPauliString over 256 random Pauli strings + partialTrace of 4 qubits (local and relocating case).

One JSON line per workload on rank 0: total device time (CUDA events, barrier both sides, max over ranks), gates/s, the
per-op-type mean times and each op type's fraction of its roofline = max(HBM bytes / measured HBM peak,
NVLink bytes per direction / 770 GB/s, flop / measured FP64 peak).
Launch: python tools/bench_configs.py [--only circuit,dm,expec]    or under torchrun --nproc-per-node N.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

import bench  # noqa: E402

FP64_PEAK_TFLOPS = 36.6          # measured on B200: profiles/r01_fp64_peak_b200.jsonl (DFMA 36.6, DMMA 37.0)
NVLINK_GBS = 770.0               # per direction, B200_PROFILING.md (measured peer copy)


def haar(rng, d):
    q, r = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def op_cost(op, kind, nq, k):
    """(hbm_bytes, nvlink_bytes_per_direction, flops) per rank, SURVEY 8(d). nq = n (sv) or N (dm)."""
    name = op[0]
    bits = nq if kind == "sv" else 2 * nq
    L = bits - k
    A = float(1 << L)
    swap = (16 * A, 8 * A)       # one suffix<->prefix swap: HBM bytes, NVLink bytes per direction

    def many_targ(targets):
        npre = sum(1 for t in targets if t >= L)
        # FP64 work counted in the cheapest known form of the complex product (3M: three real multiply-adds per complex
        # one, 6 * 2^t flop per amplitude) -- the count the t = 4, 5 tensor-core kernels issue; the 4M form is 8 * 2^t
        return 32 * A + 2 * npre * swap[0], 2 * npre * swap[1], 6.0 * (1 << len(targets)) * A

    if name == "sv_manyTargGate":
        return many_targ(op[1])
    if name in ("sv_pauliGadget", "sv_pauliTensor"):
        prefix_xy = any(t >= L and p in (1, 2) for t, p in zip(op[1], op[2]))
        return (48 * A, 16 * A, 0) if prefix_xy else (32 * A, 0, 0)
    if name == "sv_phaseGadget":
        return 32 * A, 0, 0
    if name == "dm_manyTargGate":
        a = many_targ(op[1])
        b = many_targ([t + nq for t in op[1]])
        return a[0] + b[0], a[1] + b[1], a[2] + b[2]
    thr = nq - k
    if name == "dm_oneQubitDepolarising":
        return (32 * A, 0, 0) if op[1] < thr else (48 * A, 8 * A, 0)
    if name == "dm_twoQubitDephasing":
        return 32 * A, 0, 0
    if name == "dm_oneQubitDephasing":
        return 16 * A, 0, 0
    if name == "dm_damping":
        return (32 * A, 0, 0) if op[1] < thr else (40 * A, 8 * A, 0)
    if name == "dm_expecPauliString":
        T = len(op[1])
        return min(16 * A, 32.0 * T * (1 << nq) / (1 << k)), 0, 0
    if name == "dm_partialTrace":
        t = len(op[1])
        npre = sum(1 for q in op[1] if q + nq >= L)
        return 16 * A / (1 << t) + 16 * A / (1 << (2 * t)) + npre * swap[0], npre * swap[1], 0
    raise ValueError(name)


def bound_ms(cost, hbm_peak):
    return max(cost[0] / (hbm_peak * 1e9), cost[1] / (NVLINK_GBS * 1e9), cost[2] / (FP64_PEAK_TFLOPS * 1e12)) * 1e3


def workload(name, world, seed=7):
    rng = np.random.default_rng(seed)
    k = world.bit_length() - 1
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
        N = 14 if world == 1 else 16
        ops = []
        for q in range(N):
            ops.append(("dm_manyTargGate", [q, (q + 1) % N], haar(rng, 4)))
            ops.append(("dm_oneQubitDepolarising", q, float(rng.uniform(0, 0.5))))
            ops.append(("dm_twoQubitDephasing", q, (q + 1) % N, float(rng.uniform(0, 0.5))))
            ops.append(("dm_damping", q, float(rng.uniform(0, 0.5))))
        return "dm", N, ops, "config 4: noisy layer (manyTargGate t=2, oneQubitDepolarising, twoQubitDephasing, damping) on every qubit"
    if name == "expec":
        N = 14 if world == 1 else 16
        coeffs = rng.uniform(-10, 10, 256)
        paulis = rng.integers(0, 4, size=(256, N))
        ops = [("dm_expecPauliString", coeffs, paulis)] * 4
        ops.append(("dm_partialTrace", [0, 3, 5, 8]))                       # all-suffix: local gather-sum
        ops.append(("dm_partialTrace", [N - 4, N - 3, N - 2, N - 1]))      # top qubits: bra bits are rank bits -> relocation
        return "dm", N, ops, "config 5: expecPauliString over 256 random Pauli strings (x4) + partialTrace of 4 qubits (local / relocating)"
    raise ValueError(name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="circuit,dm,expec")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import cases
    if world > 1:
        import torch
        import torch.distributed as dist
    dfsa = importlib.import_module(bench.PKG)
    lib = dfsa.device_lib()
    check = dfsa.api.check
    dfsa.comm_init()
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    k = world.bit_length() - 1
    hbm_peak, _ = bench.measured_peak()

    def event():
        e = C.c_void_p()
        check(lib.dfsa_event_create(C.byref(e)))
        return e

    for name in args.only.split(","):
        kind, nq, ops, desc = workload(name, world)
        st = dfsa.DeviceState(kind, nq)
        st.init_hash(bench.SEED)
        per_type = {}
        total = []
        for rep in range(args.reps + 1):                    # rep 0 = warm-up
            evs = [(event(), event()) for _ in ops]
            dfsa.comm_synch()
            for (e0, e1), op in zip(evs, ops):
                check(lib.dfsa_event_record(e0))
                r = cases.apply(st, op)
                check(lib.dfsa_event_record(e1))
                if op[0] == "dm_partialTrace":
                    r.close()
                    st.close()                               # partialTrace mutates its input: start from a fresh state
                    st = dfsa.DeviceState(kind, nq)
                    st.init_hash(bench.SEED)
            dfsa.comm_synch()
            if rep == 0:
                continue
            ms = C.c_double()
            tot = 0.0
            for (e0, e1), op in zip(evs, ops):
                check(lib.dfsa_event_elapsed_ms(e0, e1, C.byref(ms)))
                v = ms.value
                if world > 1:
                    t = torch.tensor([v], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    v = float(t.item())
                tot += v
                label = op[0]
                if op[0] == "dm_partialTrace":
                    label += " (relocating)" if max(op[1]) + nq >= (2 * nq - k) else " (local)"
                cost = op_cost(op, kind, nq, k)
                b = bound_ms(cost, hbm_peak)
                d = per_type.setdefault(label, {"n": 0, "ms": 0.0, "bound_ms": 0.0, "nvlink_gates": 0})
                d["n"] += 1
                d["ms"] += v
                d["bound_ms"] += b
                d["nvlink_gates"] += 1 if cost[1] > 0 else 0
            total.append(tot)
        if rank == 0:
            step_ms = float(np.mean(total))
            bound_total = sum(d["bound_ms"] for d in per_type.values()) / args.reps
            line = {"workload": name, "what": desc, "n_gpus": world, "qubits": nq, "kind": kind, "ops_per_pass": len(ops),
                    "ms_per_pass": step_ms, "gates_per_s": len(ops) / (step_ms * 1e-3), "roofline_ms_per_pass": bound_total,
                    "roofline_frac": bound_total / step_ms, "transport": lib.dfsa_comm_transport().decode(),
                    "per_op": {lab: {"count_per_pass": d["n"] // args.reps, "mean_ms": d["ms"] / d["n"], "roofline_frac": d["bound_ms"] / d["ms"],
                                     "with_exchange_per_pass": d["nvlink_gates"] // args.reps} for lab, d in per_type.items()},
                    "peaks": {"hbm_GBs": hbm_peak, "nvlink_GBs_per_dir": NVLINK_GBS, "fp64_TFLOPs": FP64_PEAK_TFLOPS}}
            print(json.dumps(line), flush=True)
        st.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    dfsa.comm_end()


if __name__ == "__main__":
    main()
