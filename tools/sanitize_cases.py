"""Small single-rank pass over every kernel family, meant to run under compute-sanitizer (memcheck / racecheck)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
dfsa = importlib.import_module("distributed-full-state-algorithms_b200")
dfsa.comm_init()
rng = np.random.default_rng(3)
nq = 12
st = dfsa.DeviceState("sv", nq); st.set_amps(cases.random_state(rng, nq))
for t, targs in ((1, [3]), (2, [0, 7]), (3, [6, 9, 11]), (3, [0, 2, 5]), (4, [5, 6, 8, 10]), (4, [1, 4, 7, 9]), (5, [6, 7, 8, 9, 10]), (5, [0, 3, 5, 8, 11]), (6, [0, 1, 4, 6, 9, 11]), (7, [0, 1, 2, 5, 7, 9, 10])):
    st.sv_manyTargGate(targs, cases.random_matrix(rng, 1 << t) / (1 << t))
for name in cases.SV_OPS:
    for _ in range(2):
        cases.apply(st, cases.make_op(rng, name, nq, 0, max_targets=4))
st.get_amps(); st.close()
N = 5
rho = dfsa.DeviceState("dm", N); rho.set_amps(cases.random_state(rng, 2 * N))
for name in cases.DM_OPS:
    op = cases.make_op(rng, name, N, 0, max_targets=2)
    r = cases.apply(rho, op)
    if name == "dm_partialTrace":
        r.get_amps(); r.close()
rho.get_amps(); rho.close()
dfsa.comm_end()
print("sanitize pass complete")
