"""ncu target: one manyTargGate launch per t on a 26-qubit state."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dfsa = importlib.import_module("distributed-full-state-algorithms_b200")
dfsa.comm_init()
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 26
st = dfsa.DeviceState("sv", nq)
st.init_hash(1)
rng = np.random.default_rng(0)
for nt in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "4,5,6".split(","))]:
    d = 1 << nt
    g, _ = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    targs = [int(x) for x in rng.permutation(nq)[:nt]]
    for _ in range(2):
        st.sv_manyTargGate(targs, g)
dfsa.comm_synch()
