import importlib, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dfsa = importlib.import_module("distributed-full-state-algorithms_b200")
import ctypes as C
dfsa.comm_init()
dfsa.set_gate_fusion(False)      # time every gate as its own kernel
lib = dfsa.device_lib(); check = dfsa.api.check
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 31
st = dfsa.DeviceState("sv", nq); st.init_hash(1)
q, _ = np.linalg.qr(np.random.default_rng(0).standard_normal((2, 2)) + 0j)
def ev():
    e = C.c_void_p(); check(lib.dfsa_event_create(C.byref(e))); return e
res = []
for t in (1, 5, 12, 20, nq - 1):
    st.sv_oneTargGate(t, q); st.sv_oneTargGate(t, q)
    e0, e1 = ev(), ev()
    check(lib.dfsa_event_record(e0))
    for _ in range(5): st.sv_oneTargGate(t, q)
    check(lib.dfsa_event_record(e1))
    ms = C.c_double(); check(lib.dfsa_event_elapsed_ms(e0, e1, C.byref(ms)))
    res.append("t=%d:%.0f" % (t, 32 * (1 << nq) / (ms.value / 5) / 1e6))
print(os.environ.get("DFSA_STREAM_BPS"), os.environ.get("DFSA_PAIR_UNROLL"), " ".join(res), flush=True)
