"""ncu target: manyTargGate launches on one state vector.
Usage: python tools/prof_manytarg.py [numQubits] [t,t,...] [placement: random|low|mid|top|mixed]
"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dfsa = importlib.import_module("distributed-full-state-algorithms_b200")
dfsa.comm_init()
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 26
placement = sys.argv[3] if len(sys.argv) > 3 else "random"
st = dfsa.DeviceState("sv", nq)
st.init_hash(1)
rng = np.random.default_rng(0)
for nt in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "4,5,6".split(","))]:
    d = 1 << nt
    g, _ = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    targs = {"random": [int(x) for x in rng.permutation(nq)[:nt]], "low": list(range(nt)), "top": list(range(nq - nt, nq)),
             "mid": [6 + 2 * i for i in range(nt)][::-1], "mixed": [3, 0, 17, nq - 1, 9, 12][:nt],
             "mixed8": [3, 0, 17, nq - 1, 9, 12, 5, 20, 22, 14, 7][:nt]}[placement]
    for _ in range(2):
        st.sv_manyTargGate(targs, g)
dfsa.comm_synch()
