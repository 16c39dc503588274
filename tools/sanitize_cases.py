"""A small pass over every kernel family, meant to run under compute-sanitizer (memcheck / racecheck / synccheck):

  compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
  DFSA_NP=4 compute-sanitizer --tool memcheck --target-processes all python tools/sanitize_cases.py      (4 ranks on one GPU)

Single rank: every dense-gate kernel (pair / quad stream, the warp-specialised tensor-core kernel for t = 3..6 at two
placements each, the GEMM kernel for t = 7, 8, the generic kernel on a tiny shard), every state-vector and density-matrix
op, expecPauliString (gather and scan), partialTrace, krausMap with the device-built superoperator, the comparator.
With DFSA_NP = P > 1 (comm_init forks the ranks) additionally every fused remote-load kernel: prefix oneTarg / controlled
sub-cube / Pauli, suffix<->prefix swap, single-shot relocation (2 and, at 8 ranks, 3 pairs), lazy layout restore, prefix
oneQubitDepolarising / damping / twoQubitDepolarising pair and quad, relocating partialTrace, the shared-page expectation value."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
dfsa = importlib.import_module("distributed-full-state-algorithms_b200")
dfsa.comm_init()
P = dfsa.comm_size()
k = P.bit_length() - 1
rng = np.random.default_rng(3)
nq = 12
st = dfsa.DeviceState("sv", nq); st.set_amps(cases.random_state(rng, nq))
L = nq - k
dense = [(1, [3]), (2, [0, 7]), (3, [6, 5, 8]), (3, [0, 2, 5]), (4, [5, 6, 8, 3]), (4, [1, 4, 7, 0]), (5, [6, 7, 8, 2, 4]), (5, [0, 3, 5, 8, 1]),
         (6, [0, 1, 4, 6, 8, 3]), (7, [0, 1, 2, 5, 7, 8, 3]), (8, [0, 1, 2, 3, 4, 5, 6, 7])]
for t, targs in dense:
    if t <= L:
        st.sv_manyTargGate([q for q in targs], cases.random_matrix(rng, 1 << t) / (1 << t))
for name in cases.SV_OPS:
    for _ in range(2):
        cases.apply(st, cases.make_op(rng, name, nq, k, max_targets=4))
if P > 1:
    top = nq - 1
    g = cases.random_matrix(rng, 2) / 1.5
    # one-target gates on a rank-bit qubit: with the queue off through the fused combine and the fused sub-cube combine + unpack (the
    # layout is still the identity here), then (default) queued and swapped into the shard by the flush
    st.restore_layout()
    dfsa.set_gate_fusion(False)
    st.sv_oneTargGate(top, g)
    st.sv_manyCtrlOneTargGate([1, 4], top, g)
    dfsa.set_gate_fusion(True)
    st.sv_oneTargGate(top, g); st.sv_manyCtrlOneTargGate([1, 4], top, g); st.sv_oneTargGate(top - 1 if k >= 2 else 0, g); st.flush()
    st.sv_pauliGadget([top, 0, 3], [2, 3, 1], 0.4)                       # fused Pauli combine
    st.sv_pauliTensor([top, 2], [1, 2])
    st.sv_swapGate(top, 2); st.sv_swapGate(top, L - 1)                   # fused suffix<->prefix swap
    if k >= 2:
        st.sv_swapGate(top, top - 1)                                     # rank relabelling (lazy) / full-shard exchange
        st.sv_manyTargGate([top, top - 1, 0], cases.random_matrix(rng, 8) / 8)          # single-shot relocation, 2 pairs
    if k >= 3:
        st.sv_manyTargGate([top, top - 1, top - 2, 5], cases.random_matrix(rng, 16) / 16)   # 3 pairs
    st.sv_manyTargGate([top, 1, 6], cases.random_matrix(rng, 8) / 8)     # leaves the layout displaced
    st.restore_layout()
st.get_amps()
other = dfsa.DeviceState("sv", nq); other.copy_from(st); assert st.compare(other)[:2] == (0.0, 0); other.close()
st.close()
tiny = dfsa.DeviceState("sv", 5 + k); tiny.init_hash(1); tiny.sv_manyTargGate([0, 2, 4], cases.random_matrix(rng, 8) / 8); tiny.get_amps(); tiny.close()   # generic kernel
N = 6
rho = dfsa.DeviceState("dm", N); rho.set_amps(cases.random_state(rng, 2 * N))
for name in cases.DM_OPS:
    op = cases.make_op(rng, name, N, k, max_targets=2)
    r = cases.apply(rho, op)
    if name == "dm_partialTrace":
        r.get_amps(); r.close()
rho.dm_krausMap([0, 2, 3, 1], [cases.random_matrix(rng, 16) / 16 for _ in range(2)])     # 8 effective targets: device superoperator + GEMM
os.environ["DFSA_EXPEC_FORCE_SCAN"] = "1"
rho.dm_expecPauliString(rng.uniform(-1, 1, 5), rng.integers(0, 4, size=(5, N)))
del os.environ["DFSA_EXPEC_FORCE_SCAN"]
if P > 1:
    rho.dm_oneQubitDepolarising(N - 1, 0.1); rho.dm_damping(N - 1, 0.2); rho.dm_oneQubitDephasing(N - 1, 0.1)
    rho.dm_twoQubitDepolarising(N - 1, 1, 0.2)                           # pair
    rho.dm_twoQubitDepolarising(N - 1, 0, 0.2, True)
    if k >= 2:
        rho.dm_twoQubitDepolarising(N - 1, N - 2, 0.2)                   # quad
        rho.dm_twoQubitDepolarising(N - 2, N - 1, 0.2, True)
    rho.dm_expecPauliString(rng.uniform(-1, 1, 9), rng.integers(0, 4, size=(9, N)))
    r = rho.dm_partialTrace([N - 1, 0]); r.get_amps(); r.close()         # relocating trace
rho.get_amps(); rho.close()
dfsa.comm_end()
print("sanitize pass complete (%d rank(s))" % P)
