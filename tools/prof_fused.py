"""ncu target: the bench sweep (oneTargGate + manyCtrlOneTargGate on every target) through the fused gate queue.
Usage: python tools/prof_fused.py [numQubits]      (ncu -k regex:fusedGateTile ...)"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, cases
dfsa = importlib.import_module("distributed-full-state-algorithms_b200")
dfsa.comm_init()
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 28
st = dfsa.DeviceState("sv", nq)
st.init_hash(1)
for rep in range(2):
    for op in bench.make_sweep(nq):
        cases.apply(st, op)
    st.flush()
dfsa.comm_synch()
