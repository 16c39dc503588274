import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import capi
dfsa = importlib.import_module("distributed-full-state-algorithms_b200")
dfsa.comm_init()
rng = np.random.default_rng(1)
for nq in (18, 11, 7):
    for t in range(1, min(nq, 9) + 1):
        for trial in range(3):
            targs = [int(x) for x in rng.permutation(nq)[:t]]
            g = rng.standard_normal((1 << t, 1 << t)) + 1j * rng.standard_normal((1 << t, 1 << t))
            amps = rng.standard_normal(1 << nq) + 1j * rng.standard_normal(1 << nq)
            print("nq=%d t=%d targs=%s ..." % (nq, t, targs), end="", flush=True)
            t0 = time.time()
            st = dfsa.DeviceState("sv", nq); st.set_amps(amps); st.sv_manyTargGate(targs, g); got = st.get_amps(); st.close()
            o = capi.OracleState("sv", nq, 1); o.set_amps(amps); o.sv_manyTargGate(targs, g)
            d = np.abs(got - o.get_amps()).max() / max(1, np.abs(o.get_amps()).max())
            print(" rel=%.2e  %.2fs" % (d, time.time() - t0), flush=True)
# explicit high-target cases (bulk-copy path: all targets >= bit 5)
for nq, t in ((18, 3), (18, 4), (18, 5), (13, 5), (12, 4)):
    targs = [int(x) for x in (5 + rng.permutation(nq - 5)[:t])]
    g = rng.standard_normal((1 << t, 1 << t)) + 1j * rng.standard_normal((1 << t, 1 << t))
    amps = rng.standard_normal(1 << nq) + 1j * rng.standard_normal(1 << nq)
    st = dfsa.DeviceState("sv", nq); st.set_amps(amps); st.sv_manyTargGate(targs, g); st.sv_manyTargGate(targs, g); got = st.get_amps(); st.close()
    o = capi.OracleState("sv", nq, 1); o.set_amps(amps); o.sv_manyTargGate(targs, g); o.sv_manyTargGate(targs, g)
    print("HIGH nq=%d t=%d targs=%s rel=%.2e" % (nq, t, targs, np.abs(got - o.get_amps()).max() / max(1, np.abs(o.get_amps()).max())), flush=True)
