"""Kernel-level timing sweep on one GPU (CUDA events on the library's compute stream).
Usage: python tools/microbench.py [numQubits]   -> JSON lines, one per kernel configuration."""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dfsa = importlib.import_module("distributed-full-state-algorithms_b200")
import torch  # noqa: E402  (events on an external stream)


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dfsa.comm_init()
    dfsa.set_gate_fusion(False)      # time every gate as its own kernel
    st = dfsa.DeviceState("sv", nq)
    st.init_hash(1)
    stream = torch.cuda.ExternalStream(dfsa.device_lib().dfsa_stream_compute())
    A = 1 << nq
    rng = np.random.default_rng(0)
    q, _ = np.linalg.qr(rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2)))
    q32, _ = np.linalg.qr(rng.standard_normal((32, 32)) + 1j * rng.standard_normal((32, 32)))

    def timeit(label, fn, bytes_per_call, flops=0, reps=5):
        fn(); fn()
        dfsa.api.check(dfsa.device_lib().dfsa_device_sync())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rec = {"kernel": label, "qubits": nq, "ms": round(ms, 4), "GBps": round(bytes_per_call / ms / 1e6, 1)}
        if flops:
            rec["TFLOPs"] = round(flops / ms / 1e9, 2)
        print(json.dumps(rec), flush=True)

    for t in [0, 1, 2, 3, 4, 5, 8, 12, nq // 2, nq - 2, nq - 1]:
        timeit("oneTarg t=%d" % t, lambda: st.sv_oneTargGate(t, q), 32 * A)
    for ctrls, t in [([1], 0), ([0], 5), ([nq - 1], 3), ([2, 9], 20), ([0, 1, 2], 10)]:
        timeit("ctrl%s t=%d" % (ctrls, t), lambda: st.sv_manyCtrlOneTargGate(ctrls, t, q), 32 * A / 2 ** len(ctrls))
    for a, b in [(0, 1), (0, nq - 1), (5, 17), (nq - 2, nq - 1)]:
        timeit("swap %d,%d" % (a, b), lambda: st.sv_swapGate(a, b), 16 * A)
    timeit("phaseGadget", lambda: st.sv_phaseGadget([0, 7, nq - 1], 0.3), 32 * A)
    timeit("pauliGadget XYZ", lambda: st.sv_pauliGadget([1, 9, nq - 1], [1, 2, 3], 0.3), 32 * A)
    timeit("pauliTensor XYZ", lambda: st.sv_pauliTensor([1, 9, nq - 1], [1, 2, 3]), 32 * A)
    for targs in ([0, 1, 2, 3, 4], [3, 0, 17, nq - 1, 9], [nq - 5, nq - 4, nq - 3, nq - 2, nq - 1]):
        timeit("manyTarg5 %s" % targs, lambda: st.sv_manyTargGate(targs, q32), 32 * A, flops=8 * 32 * A, reps=3)
    for nt in (1, 2, 3, 4, 6):
        d = 1 << nt
        g, _ = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
        targs = [int(x) for x in rng.permutation(nq)[:nt]]
        timeit("manyTarg%d %s" % (nt, targs), lambda: st.sv_manyTargGate(targs, g), 32 * A, flops=8 * d * A, reps=3)
    st.close()
    if nq % 2 == 0:
        N = nq // 2
        rho = dfsa.DeviceState("dm", N)
        rho.init_hash(2)
        timeit("dm oneQubitDephasing", lambda: rho.dm_oneQubitDephasing(3, 0.1), 16 * A)
        timeit("dm twoQubitDephasing", lambda: rho.dm_twoQubitDephasing(1, 4, 0.1), 32 * A)
        timeit("dm oneQubitDepolarising", lambda: rho.dm_oneQubitDepolarising(2, 0.1), 32 * A)
        timeit("dm twoQubitDepolarising", lambda: rho.dm_twoQubitDepolarising(2, 5, 0.1), 40 * A)
        timeit("dm damping", lambda: rho.dm_damping(N - 1, 0.1), 32 * A)
        g4, _ = np.linalg.qr(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))
        timeit("dm manyTarg2", lambda: rho.dm_manyTargGate([1, 6], g4), 64 * A)
        coeffs = rng.uniform(-10, 10, 256)
        paulis = rng.integers(0, 4, size=(256, N))
        timeit("dm expecPauliString T=256", lambda: rho.dm_expecPauliString(coeffs, paulis), 256 * (1 << N) * 32, reps=3)
        rho.close()


if __name__ == "__main__":
    main()
