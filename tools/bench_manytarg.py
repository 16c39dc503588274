"""manyTargGate kernel timings on one GPU: t = 2..6 at target placements that exercise each tile layout.
Usage: python tools/bench_manytarg.py [numQubits]   -> JSON lines (CUDA events on the library's compute stream; each
configuration is timed in ROUNDS interleaved rounds, median and best reported: FP64-heavy kernels move the clocks)."""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dfsa = importlib.import_module("distributed-full-state-algorithms_b200")

HBM_GBS, FP64_TF, ROUNDS, REPS = 6448.4, 36.6, 3, 3


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dfsa.comm_init()
    lib = dfsa.device_lib()
    st = dfsa.DeviceState("sv", nq)
    st.init_hash(1)
    A = 1 << nq
    rng = np.random.default_rng(0)

    def event():
        e = C.c_void_p()
        dfsa.api.check(lib.dfsa_event_create(C.byref(e)))
        return e

    def time_once(fn):
        e0, e1 = event(), event()
        dfsa.api.check(lib.dfsa_event_record(e0))
        for _ in range(REPS):
            fn()
        dfsa.api.check(lib.dfsa_event_record(e1))
        ms = C.c_double()
        dfsa.api.check(lib.dfsa_event_elapsed_ms(e0, e1, C.byref(ms)))
        return ms.value / REPS

    only = [int(x) for x in os.environ.get("MT_ONLY", "2,3,4,5,6").split(",")]
    for nt in only:
        d = 1 << nt
        g, _ = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
        placements = {"low": list(range(nt)), "top": list(range(nq - nt, nq)), "mid": [6 + 2 * i for i in range(nt)][::-1],
                      "mixed": [3, 0, 17, nq - 1, 9, 12][:nt]}
        fns = {p: (lambda targs=targs: st.sv_manyTargGate(targs, g)) for p, targs in placements.items()}
        for fn in fns.values():
            fn()
        dfsa.comm_synch()
        times = {p: [] for p in fns}
        for _ in range(ROUNDS):
            for p, fn in fns.items():
                times[p].append(time_once(fn))
        # the roofline counts FP64 work in the 3M form (6 * 2^t flop per amplitude), the cheapest known complex product
        hbm_ms = 32 * A / HBM_GBS / 1e6
        fp_ms = 6 * d * A / FP64_TF / 1e9
        for p, ts in times.items():
            t = float(np.median(ts))
            print(json.dumps({"kernel": "manyTarg t=%d %s" % (nt, p), "targets": placements[p], "qubits": nq, "ms": round(t, 3), "ms_best": round(min(ts), 3),
                              "GBps": round(32 * A / t / 1e6, 1), "TFLOPs_3M_count": round(6 * d * A / t / 1e9, 2),
                              "TFLOPs_4M_equiv": round(8 * d * A / t / 1e9, 2),
                              "bound_ms": round(max(hbm_ms, fp_ms), 3), "roofline_frac": round(max(hbm_ms, fp_ms) / t, 3)}), flush=True)
    st.close()


if __name__ == "__main__":
    main()
