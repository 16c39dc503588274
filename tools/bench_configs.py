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

# workloads, algorithmic costs and the timing loop live in bench.py (which runs them inside the contract bench line too)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="circuit,dm,expec")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--dm-qubits", type=int, default=0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    job = bench.Job(world, rank, local_rank)
    hbm_peak, src = bench.measured_peak()
    nvlink = None
    if world > 1:
        probe = job.dfsa.DeviceState("sv", 30 + job.k)
        nvlink = bench.measure_nvlink(job, probe)
        probe.close()
    peaks = {"hbm_GBs": hbm_peak, "fp64_TFLOPs": bench.FP64_PEAK_TFLOPS,
             "nvlink_GBs_per_dir": nvlink["kernel_remote_loads_GBs_per_dir"] if nvlink else bench.NVLINK_FALLBACK_GBS}
    for name in args.only.split(","):
        line = bench.run_config(job, name, peaks, reps=args.reps, dm_qubits=args.dm_qubits or None)
        if rank == 0:
            line.update({"workload": name, "n_gpus": world, "transport": job.lib.dfsa_comm_transport().decode(), "peaks": peaks, "nvlink": nvlink})
            print(json.dumps(line), flush=True)
    job.close()


if __name__ == "__main__":
    main()
