"""manyTargGate kernel timings on one GPU: t = 2..6, target placements that select each kernel form (bulk rows / gather),
and for t = 5 both the warp-pair 3M kernel and the one-warp 4M kernel (DFSA_MANYTARG5=warp).
Usage: python tools/bench_manytarg.py [numQubits]   -> JSON lines (CUDA events on the library's compute stream)."""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dfsa = importlib.import_module("distributed-full-state-algorithms_b200")

HBM_GBS, FP64_TF = 6448.4, 36.6


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

    def timeit(label, fn, flops_per_amp_4m, flops_per_amp_issued, reps=4):
        fn(); fn()
        dfsa.comm_synch()
        e0, e1 = event(), event()
        dfsa.api.check(lib.dfsa_event_record(e0))
        for _ in range(reps):
            fn()
        dfsa.api.check(lib.dfsa_event_record(e1))
        ms = C.c_double()
        dfsa.api.check(lib.dfsa_event_elapsed_ms(e0, e1, C.byref(ms)))
        t = ms.value / reps
        hbm_ms = 32 * A / HBM_GBS / 1e6
        fp_ms = flops_per_amp_issued * A / FP64_TF / 1e9
        print(json.dumps({"kernel": label, "qubits": nq, "ms": round(t, 3), "GBps": round(32 * A / t / 1e6, 1),
                          "TFLOPs_4M_equiv": round(flops_per_amp_4m * A / t / 1e9, 2),
                          "TFLOPs_3M_count": round(flops_per_amp_issued * A / t / 1e9, 2),
                          "bound_ms": round(max(hbm_ms, fp_ms), 3), "roofline_frac": round(max(hbm_ms, fp_ms) / t, 3)}), flush=True)

    only = [int(x) for x in os.environ.get("MT_ONLY", "2,3,4,5,6").split(",")]
    for nt in only:
        d = 1 << nt
        g, _ = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
        placements = {"low": list(range(nt)), "top": list(range(nq - nt, nq)), "mid": [6 + 2 * i for i in range(nt)][::-1],
                      "mixed": [3, 0, 17, nq - 1, 9, 12][:nt]}
        for pname, targs in placements.items():
            variants = [("", None)]
            if nt == 5:
                variants = [("spec", {"DFSA_MANYTARG5": "spec"}), ("spec-nounroll", {"DFSA_MANYTARG5": "spec", "DFSA_SPEC5_UNROLL": "0"}),
                            ("spec-cg", {"DFSA_MANYTARG5": "spec", "DFSA_SPEC5_CA": "0"}), ("spec-ca", {"DFSA_MANYTARG5": "spec", "DFSA_SPEC5_CA": "1"}),
                            ("pair", {"DFSA_MANYTARG5": "pair"}), ("warp4m", {"DFSA_MANYTARG5": "warp"})]
            for vname, env in variants:
                for k in ("DFSA_MANYTARG5", "DFSA_SPEC5_NIN", "DFSA_SPEC5_UNROLL", "DFSA_SPEC5_CA"):
                    os.environ.pop(k, None)
                os.environ.update(env or {})
                # the roofline counts FP64 work in the 3M form (6 * 2^t flop per amplitude) whichever form the kernel issues
                timeit("manyTarg t=%d %s %s" % (nt, pname, vname), lambda: st.sv_manyTargGate(targs, g), 8 * d, 6 * d)
    os.environ.pop("DFSA_MANYTARG5", None)
    st.close()


if __name__ == "__main__":
    main()
