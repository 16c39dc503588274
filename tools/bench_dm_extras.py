#!/usr/bin/env python
"""Timings VERDICT r1 asked for that no BASELINE config contains (1 GPU, 14-qubit density matrix = 2^28 amplitudes):
krausMap on 3, 4 and 5 qubits (6 / 8 / 10 effective targets: tensor-core tile kernel / GEMM with the device-built superoperator),
local twoQubitDepolarising (one 32*A pass), partialTrace of 4 qubits at three placements of the traced bits.
One JSON line per measurement (CUDA events on the compute stream, best of 3 after a warm-up)."""
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


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 14
    job = bench.Job(1, 0, 0)
    hbm, _ = bench.measured_peak()
    rho = job.dfsa.DeviceState("dm", N)
    rho.init_hash(3)
    A = float(rho.num_amps_per_node)
    rng = np.random.default_rng(0)

    def timed(fn, reps=3):
        best = None
        for r in range(reps + 1):
            e0, e1 = job.event(), job.event()
            job.barrier()
            job.record(e0)
            out = fn()
            job.record(e1)
            job.barrier()
            t = job.elapsed(e0, e1)
            if out is not None:
                out.close()
            if r > 0:
                best = t if best is None else min(best, t)
        return best

    def emit(what, ms, hbm_bytes, flops=0.0):
        bound = max(hbm_bytes / (hbm * 1e9), flops / (bench.FP64_PEAK_TFLOPS * 1e12)) * 1e3
        print(json.dumps({"what": what, "dm_qubits": N, "ms": round(ms, 3), "bound_ms": round(bound, 3), "roofline_frac": round(bound / ms, 3),
                          "GBps": round(hbm_bytes / ms / 1e6, 1), "TFLOPs_3M": round(flops / ms / 1e9, 2)}), flush=True)

    for t in (3, 4, 5):
        ops = [bench.haar(rng, 1 << t) / np.sqrt(3) for _ in range(3)]
        targets = [int(x) for x in rng.permutation(N)[:t]]
        ms = timed(lambda: rho.dm_krausMap(targets, ops))
        emit("krausMap on %d qubits %r (3 Kraus operators): %dx%d superoperator" % (t, targets, 4 ** t, 4 ** t), ms, 32 * A, 6.0 * (4 ** t) * A)
    emit("twoQubitDepolarising local (5, 9)", timed(lambda: rho.dm_twoQubitDepolarising(5, 9, 0.1)), 32 * A)
    emit("twoQubitDepolarising local (0, 1)", timed(lambda: rho.dm_twoQubitDepolarising(0, 1, 0.1)), 32 * A)
    for targets in ([0, 3, 5, 8], [3, 5, 8, 10], [N - 6, N - 5, N - 4, N - 3]):
        emit("partialTrace %r" % (targets,), timed(lambda: rho.dm_partialTrace(targets)), 16 * A / 16 + 16 * A / 256)
    coeffs = rng.uniform(-10, 10, 256)
    paulis = rng.integers(0, 4, size=(256, N))
    emit("expecPauliString T=256", timed(lambda: rho.dm_expecPauliString(coeffs, paulis) and None), min(16 * A, 32.0 * 256 * (1 << N)))
    rho.close()
    job.close()


if __name__ == "__main__":
    main()
