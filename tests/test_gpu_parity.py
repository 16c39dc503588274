"""GPU parity tests (-m gpu): the CUDA path, driven through the host API / C-ABI, against
(1) the committed golden outputs of the real reference (tests/golden/), (2) the C oracle on seeded random
inputs at larger sizes, (3) size-independent properties at sizes the oracle cannot reach.
Bar (BASELINE.json north_star): permutation / sign-only ops bit-exact, everything else max|delta amp| <= 1e-12
(relative to the largest reference amplitude), expectation values relative error <= 1e-12."""
import zlib

import numpy as np
import pytest

import cases
import compare
import golden_io
import product
from oracle import capi, dense

pytestmark = pytest.mark.gpu

GOLDEN = golden_io.load("sv") + golden_io.load("dm")
GOLDEN_1 = [c for c in GOLDEN if c["nodes"] == 1]


@pytest.fixture(scope="module")
def dfsa():
    m = product.pkg()
    m.comm_init()
    assert m.comm_size() == 1
    return m


def check_against(case_op, kind, got_state, got_result, want_amps=None, want_val=None, want_mut=None, mutated=None):
    name = case_op[0]
    if name == "dm_expecPauliString":
        compare.assert_value_close(got_result, want_val, what=name)
    elif name == "dm_partialTrace":
        compare.assert_exact(got_result.get_amps(), want_amps, what=name)
        if want_mut is not None:
            compare.assert_exact(got_state.get_amps(), want_mut, what=name + " (mutated input)")
    elif name in cases.EXACT_OPS:
        compare.assert_exact(got_state.get_amps(), want_amps, what=name)
    else:
        compare.assert_close(got_state.get_amps(), want_amps, what=name)


@pytest.mark.parametrize("case", GOLDEN_1, ids=golden_io.case_id)
def test_single_rank_matches_reference_golden(dfsa, case):
    st = dfsa.DeviceState(case["kind"], case["nq"])
    st.set_amps(case["amps"])
    res = cases.apply(st, case["op"])
    check_against(case["op"], case["kind"], st, res, want_amps=case.get("out"), want_val=(case["val"][0] if "val" in case else None), want_mut=case.get("mut"))


def _oracle_run(kind, nq, nodes, amps, op):
    o = capi.OracleState(kind, nq, nodes)
    o.set_amps(amps)
    r = cases.apply(o, op)
    return o, r


@pytest.mark.parametrize("name", cases.SV_OPS)
@pytest.mark.parametrize("nq", [3, 11, 18])
def test_sv_ops_match_oracle_at_larger_sizes(dfsa, name, nq):
    rng = np.random.default_rng(zlib.crc32(("%s-%d" % (name, nq)).encode()))
    for trial in range(4 if nq < 18 else 2):
        op = cases.make_op(rng, name, nq, 0, max_targets=9)
        amps = cases.random_state(rng, nq)
        st = dfsa.DeviceState("sv", nq)
        st.set_amps(amps)
        cases.apply(st, op)
        o, _ = _oracle_run("sv", nq, 1, amps, op)
        if name in cases.EXACT_OPS:
            compare.assert_exact(st.get_amps(), o.get_amps(), what=name)
        else:
            compare.assert_close(st.get_amps(), o.get_amps(), what="%s %r" % (name, op[1]))


@pytest.mark.parametrize("name", cases.DM_OPS)
@pytest.mark.parametrize("nq", [2, 6, 9])
def test_dm_ops_match_oracle_at_larger_sizes(dfsa, name, nq):
    rng = np.random.default_rng(zlib.crc32(("%s-%d" % (name, nq)).encode()))
    for trial in range(3 if nq < 9 else 1):
        op = cases.make_op(rng, name, nq, 0, max_targets=(3 if name == "dm_krausMap" else 5))
        if name == "dm_partialTrace" and nq == 2:
            op = (name, op[1][:1])
        amps = cases.random_state(rng, 2 * nq)
        st = dfsa.DeviceState("dm", nq)
        st.set_amps(amps)
        res = cases.apply(st, op)
        o, ores = _oracle_run("dm", nq, 1, amps, op)
        if name == "dm_expecPauliString":
            compare.assert_value_close(res, ores, what=name)
        elif name == "dm_partialTrace":
            compare.assert_exact(res.get_amps(), ores.get_amps(), what=name)
        elif name in cases.EXACT_OPS:
            compare.assert_exact(st.get_amps(), o.get_amps(), what=name)
        else:
            compare.assert_close(st.get_amps(), o.get_amps(), what="%s %r" % (name, op[1:3]))


def test_every_target_position_one_and_ctrl_gate(dfsa):
    """BASELINE config 2 in miniature: oneTargGate / manyCtrlOneTargGate over all target positions."""
    nq = 14
    rng = np.random.default_rng(5)
    amps = cases.random_state(rng, nq)
    st = dfsa.DeviceState("sv", nq)
    st.set_amps(amps)
    o = capi.OracleState("sv", nq, 1)
    o.set_amps(amps)
    for t in range(nq):
        g = cases.random_matrix(rng, 2) / 1.5
        ops = [("sv_oneTargGate", t, g),
               ("sv_manyCtrlOneTargGate", [int(c) for c in rng.permutation([q for q in range(nq) if q != t])[: 1 + t % 3]], t, g)]
        for op in ops:
            cases.apply(st, op)
            cases.apply(o, op)
    compare.assert_close(st.get_amps(), o.get_amps(), tol=1e-11, what="target sweep")   # 28 chained non-unitary gates


@pytest.mark.parametrize("nt", [1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_many_targ_gate_every_kernel_and_placement(dfsa, nt):
    """manyTargGate (local_statevector.hpp:72-99) picks its kernel by target count (pair stream, quad stream, tensor-core
    tiles for t = 3..6, the tiled tensor-core GEMM with the gate streamed from L2 for t = 7..11, generic on shards smaller than a tile), and the tensor-core kernel's
    tile layout and shared-memory swizzle depend on where the targets sit. Every kernel, targets low / high / scattered /
    just above the free bits, in caller order (not sorted), against the oracle."""
    rng = np.random.default_rng(100 + nt)
    for nq in (nt, nt + 2, 13, 16):
        if nq < nt:
            continue
        placements = [list(range(nt)), list(range(nq - nt, nq))[::-1], [int(x) for x in rng.permutation(nq)[:nt]]]
        if nq >= nt + 6:
            placements.append([5 + i for i in range(nt)][::-1])                 # just above the free bits
            placements.append([0] + [nq - 1 - i for i in range(nt - 1)])        # one low target, the rest on top
        for targets in placements:
            amps = cases.random_state(rng, nq)
            gate = cases.random_matrix(rng, 1 << nt) / (1 << nt) ** 0.5
            st = dfsa.DeviceState("sv", nq)
            st.set_amps(amps)
            st.sv_manyTargGate(targets, gate)
            o, _ = _oracle_run("sv", nq, 1, amps, ("sv_manyTargGate", targets, gate))
            compare.assert_close(st.get_amps(), o.get_amps(), what="manyTargGate nq=%d targets=%r" % (nq, targets))
            st.close()


@pytest.mark.parametrize("nq,nt", [(7, 4), (8, 4), (8, 5), (6, 3)])
def test_kraus_map_on_4_and_5_qubits_builds_its_superoperator_on_the_device(dfsa, nq, nt):
    """krausMap on 4 / 5 qubits = a dense gate on 8 / 10 index bits with a 256^2 / 1024^2 superoperator
    (distributed_densitymatrix.hpp:79-89, misc.hpp:58-81): built on the device, applied by the GEMM kernel."""
    rng = np.random.default_rng(nq * 10 + nt)
    targets = [int(x) for x in rng.permutation(nq)[:nt]]
    ops = [cases.random_matrix(rng, 1 << nt) / (1 << nt) for _ in range(3)]
    amps = cases.random_state(rng, 2 * nq)
    st = dfsa.DeviceState("dm", nq)
    st.set_amps(amps)
    st.dm_krausMap(targets, ops)
    o, _ = _oracle_run("dm", nq, 1, amps, ("dm_krausMap", targets, ops))
    compare.assert_close(st.get_amps(), o.get_amps(), what="krausMap nq=%d targets=%r" % (nq, targets))
    st.close()


def test_all_z_pauli_string_applies_the_operator(dfsa):
    """Documented divergence: the reference silently skips X/Y-free Pauli strings; this build applies them."""
    rng = np.random.default_rng(9)
    amps = cases.random_state(rng, 7)
    for op in (("sv_pauliTensor", [0, 3, 6], [3, 3, 3]), ("sv_pauliGadget", [1, 5], [3, 3], 0.81)):
        st = dfsa.DeviceState("sv", 7)
        st.set_amps(amps)
        cases.apply(st, op)
        compare.assert_close(st.get_amps(), dense.apply_op("sv", 7, amps, op), what=op[0] + " all-Z")


def test_expec_gather_and_scan_agree(dfsa, monkeypatch):
    rng = np.random.default_rng(11)
    nq = 6
    amps = cases.random_state(rng, 2 * nq)
    op = cases.make_op(rng, "dm_expecPauliString", nq, 0)
    vals = []
    for force in ("DFSA_EXPEC_FORCE_GATHER", "DFSA_EXPEC_FORCE_SCAN"):
        monkeypatch.setenv(force, "1")
        st = dfsa.DeviceState("dm", nq)
        st.set_amps(amps)
        vals.append(cases.apply(st, op))
        monkeypatch.delenv(force)
    o, want = _oracle_run("dm", nq, 1, amps, op)
    compare.assert_value_close(vals[0], want)
    compare.assert_value_close(vals[1], want)


def test_corrected_two_qubit_depolarising_is_the_channel(dfsa):
    rng = np.random.default_rng(13)
    nq = 4
    amps = cases.random_state(rng, 2 * nq)
    st = dfsa.DeviceState("dm", nq)
    st.set_amps(amps)
    st.dm_twoQubitDepolarising(1, 3, 0.31, corrected=True)
    compare.assert_close(st.get_amps(), dense.apply_op("dm", nq, amps, ("dm_twoQubitDepolarising", 1, 3, 0.31)))


# ---------------------------------------------------------------- properties at sizes beyond the oracle

def test_large_state_properties(dfsa):
    """28-qubit state vector (4 GiB): hash init reproducible on the host, unitary layer preserves the norm,
    U then U^dagger restores the state bit-for-bit-close, swap twice is the identity exactly."""
    nq = 28
    st = dfsa.DeviceState("sv", nq)
    st.init_hash(2024)
    n0 = st.norm2()
    # spot-check the device hash against the numpy restatement
    k = np.arange(2 * 4096, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(2024) + (k + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    vals = (z >> np.uint64(11)).astype(np.float64) / 9007199254740992.0 - 0.5
    head = np.empty(4096, dtype=np.complex128)
    import ctypes as C
    product.pkg().api.check(dfsa.device_lib().dfsa_state_download(st.handle, 0, C.c_uint64(0), C.c_uint64(4096), head.ctypes.data_as(C.POINTER(C.c_double))))
    compare.assert_exact(head, vals[0::2] + 1j * vals[1::2], what="hash init")

    rng = np.random.default_rng(3)
    def unitary(d):
        q, r = np.linalg.qr(cases.random_matrix(rng, d))
        return q
    layer = []
    for t in (0, 1, 5, 13, 27):
        layer.append(("sv_oneTargGate", t, unitary(2)))
    layer.append(("sv_manyCtrlOneTargGate", [2, 20], 9, unitary(2)))
    layer.append(("sv_manyTargGate", [3, 0, 17, 26, 9], unitary(32)))
    layer.append(("sv_pauliGadget", [1, 8, 22], [1, 2, 3], 0.4))
    layer.append(("sv_phaseGadget", [0, 27, 14], -1.1))
    for op in layer:
        cases.apply(st, op)
    assert abs(st.norm2() / n0 - 1) < 1e-12
    for op in reversed(layer):
        name = op[0]
        if name in ("sv_oneTargGate",):
            cases.apply(st, (name, op[1], op[2].conj().T))
        elif name == "sv_manyCtrlOneTargGate":
            cases.apply(st, (name, op[1], op[2], op[3].conj().T))
        elif name == "sv_manyTargGate":
            cases.apply(st, (name, op[1], op[2].conj().T))
        elif name == "sv_pauliGadget":
            cases.apply(st, (name, op[1], op[2], -op[3]))
        else:
            cases.apply(st, (name, op[1], -op[2]))
    product.pkg().api.check(dfsa.device_lib().dfsa_state_download(st.handle, 0, C.c_uint64(0), C.c_uint64(4096), head.ctypes.data_as(C.POINTER(C.c_double))))
    compare.assert_close(head, vals[0::2] + 1j * vals[1::2], tol=1e-13, what="U then U^dagger")
    d_all, _, _ = st.compare_hash(2024)                       # ... and over ALL 2^28 amplitudes, on the device
    assert d_all <= 1e-13, "U then U^dagger, full state: %.3e" % d_all
    st.sv_swapGate(3, 25)
    st.sv_swapGate(25, 3)
    tail = np.empty(4096, dtype=np.complex128)
    product.pkg().api.check(dfsa.device_lib().dfsa_state_download(st.handle, 0, C.c_uint64(0), C.c_uint64(4096), tail.ctypes.data_as(C.POINTER(C.c_double))))
    compare.assert_exact(tail, head, what="swap twice")
    assert st.compare_hash(2024)[0] == d_all                  # swap twice moved nothing, anywhere


# ---------------------------------------------------------------- multi-rank (NCCL with >= P GPUs, else IPC on one GPU)

def _golden_multirank(nodes):
    return [c for c in GOLDEN if c["nodes"] == nodes]


# the three ways an exchange can run: fused into the consuming kernel with stream-ordered signalling between the ranks
# (default), fused with host synchronisation (no stream memory ops), staged pack / exchange / combine like the reference
EXCHANGE_MODES = {"fused": None, "fused-hostsync": {"DFSA_STREAM_SIGNALS": "0"}, "staged": {"DFSA_FUSED_EXCHANGE": "0"}}


@pytest.mark.parametrize("mode", sorted(EXCHANGE_MODES))
@pytest.mark.parametrize("nodes", [2, 4, 8])
def test_multi_rank_matches_reference_golden(nodes, mode):
    todo = _golden_multirank(nodes)
    jobs = [dict(kind=c["kind"], nq=c["nq"], op=c["op"], amps=c["amps"]) for c in todo]
    results = product.run_cases_multirank(jobs, nodes, extra_env=EXCHANGE_MODES[mode])
    for c, r in zip(todo, results):
        name = c["op"][0]
        if name == "dm_expecPauliString":
            compare.assert_value_close(r["values"][0], c["val"][0], what=golden_io.case_id(c))
        elif name == "dm_partialTrace":
            compare.assert_exact(r["amps"], c["out"], what=golden_io.case_id(c))
            compare.assert_exact(r["mutated"], c["mut"], what=golden_io.case_id(c) + " mutated")
        elif name in cases.EXACT_OPS:
            compare.assert_exact(r["amps"], c["out"], what=golden_io.case_id(c))
        else:
            compare.assert_close(r["amps"], c["out"], what=golden_io.case_id(c))


@pytest.mark.parametrize("mode", sorted(EXCHANGE_MODES))
@pytest.mark.parametrize("nodes", [2, 4, 8])
def test_multi_rank_circuit_matches_oracle(nodes, mode):
    """Chained ops on one state (exercises buffer reuse between exchanges, and -- in the stream-ordered mode -- gates
    queued behind each other with no host synchronisation in between) at a larger size."""
    rng = np.random.default_rng(40 + nodes)
    k = nodes.bit_length() - 1
    jobs, wants = [], []
    for kind, nq, names in (("sv", 12, cases.SV_OPS), ("dm", 6, [n for n in cases.DM_OPS if n not in ("dm_partialTrace", "dm_expecPauliString")])):
        ops = []
        for i in range(3 * len(names)):
            op = cases.make_op(rng, names[i % len(names)], nq, k, max_targets=3)
            ops.append(tuple((a / np.linalg.norm(a, 2) if isinstance(a, np.ndarray) and a.ndim == 2 and a.dtype == np.complex128 else a) for a in op))
        amps = cases.random_state(rng, nq if kind == "sv" else 2 * nq)
        o = capi.OracleState(kind, nq, nodes)
        o.set_amps(amps)
        for op in ops:
            cases.apply(o, op)
        jobs.append(dict(kind=kind, nq=nq, ops=ops, amps=amps))
        wants.append(o.get_amps())
    results = product.run_cases_multirank(jobs, nodes, extra_env=EXCHANGE_MODES[mode])
    for r, w in zip(results, wants):
        compare.assert_close(r["amps"], w, tol=1e-11, what="circuit np=%d (%s, %s)" % (nodes, r["transport"], mode))


@pytest.mark.parametrize("mode", sorted(EXCHANGE_MODES))
@pytest.mark.parametrize("nodes", [2, 4, 8])
def test_corrected_two_qubit_depolarising_on_prefix_qubits(nodes, mode):
    """corrected=True on the pair and quad branches (one / both bra bits in the rank index) is the true channel
    (1-16p/15) rho + (4p/15) I (x) Tr_2 rho: against the dense Kraus map."""
    rng = np.random.default_rng(700 + nodes)
    k = nodes.bit_length() - 1
    nq = 5
    pairs = [(0, nq - 1), (nq - 1, 2)] + ([(nq - 1, nq - 2), (nq - k, nq - 1)] if k >= 2 else [])
    jobs, wants = [], []
    for (a, b) in pairs:
        amps = cases.random_state(rng, 2 * nq)
        p = float(rng.uniform(0, 0.75))
        jobs.append(dict(kind="dm", nq=nq, op=("dm_twoQubitDepolarising", a, b, p, True), amps=amps))
        wants.append(dense.apply_op("dm", nq, amps, ("dm_twoQubitDepolarising", a, b, p)))
    results = product.run_cases_multirank(jobs, nodes, extra_env=EXCHANGE_MODES[mode])
    for job, r, w in zip(jobs, results, wants):
        compare.assert_close(r["amps"], w, what="corrected depol2 %r np=%d (%s)" % (job["op"][1:3], nodes, mode))


@pytest.mark.parametrize("nodes", [4, 8])
def test_relocation_of_several_prefix_targets_matches_oracle(nodes):
    """manyTargGate with 2, 3 (and at 8 ranks all) rank bits among its targets: the single-shot relocation gathers from the
    2^k shards of a rank's group in one pass (distributed_statevector.hpp:193-223 does k swaps); also through the staged
    transport (DFSA_FUSED_EXCHANGE=0), which falls back to the sequence of swaps."""
    rng = np.random.default_rng(900 + nodes)
    k = nodes.bit_length() - 1
    nq = 12
    L = nq - k
    jobs, wants = [], []
    for npre in range(2, k + 1):
        for trial in range(2):
            prefix = [int(x) for x in rng.permutation(np.arange(L, nq))[:npre]]
            suffix = [int(x) for x in rng.permutation(L)[: int(rng.integers(0, 3))]]
            targets = [int(x) for x in rng.permutation(prefix + suffix)]
            gate = cases.random_matrix(rng, 1 << len(targets)) / (1 << len(targets)) ** 0.5
            amps = cases.random_state(rng, nq)
            op = ("sv_manyTargGate", targets, gate)
            o, _ = _oracle_run("sv", nq, nodes, amps, op)
            jobs.append(dict(kind="sv", nq=nq, op=op, amps=amps))
            wants.append(o.get_amps())
    for env in (None, {"DFSA_FUSED_EXCHANGE": "0"}):
        results = product.run_cases_multirank(jobs, nodes, extra_env=env)
        for job, r, w in zip(jobs, results, wants):
            compare.assert_close(r["amps"], w, what="relocation np=%d targets=%r (%s)" % (nodes, job["op"][1], r["transport"]))


@pytest.mark.parametrize("nodes", [2, 4, 8])
def test_lazy_layout_defers_relocation_and_matches_oracle(nodes):
    """host/layout.hpp: manyTargGate leaves its relocated targets on the suffix bits they landed on, prefix<->prefix swaps
    only relabel ranks, every later gate is translated through the layout, and the read-back restores index order. Same
    results as the oracle (and as DFSA_LAZY_LAYOUT=0, where every relocation is undone at once); with laziness on, the layout
    just before the read-back must be non-trivial and just after it the identity."""
    rng = np.random.default_rng(1200 + nodes)
    k = nodes.bit_length() - 1
    nq = 12
    L = nq - k
    ops = []
    for rep in range(3):
        pre = [int(x) for x in rng.permutation(np.arange(L, nq))[: 1 + rep % k if k > 1 else 1]]
        suf = [int(x) for x in rng.permutation(L)[:2]]
        targets = [int(x) for x in rng.permutation(pre + suf)]
        ops.append(("sv_manyTargGate", targets, cases.random_matrix(rng, 1 << len(targets)) / (1 << len(targets)) ** 0.5))
        ops.append(("sv_oneTargGate", int(rng.integers(0, nq)), cases.random_matrix(rng, 2) / 1.5))
        ops.append(("sv_manyCtrlOneTargGate", [nq - 1, 2], int(rng.integers(3, nq - 1)), cases.random_matrix(rng, 2) / 1.5))
        ops.append(("sv_pauliGadget", [nq - 1, 0, 5], [1, 3, 2], float(rng.uniform(-3, 3))))
        ops.append(("sv_phaseGadget", [nq - 1, 1], float(rng.uniform(-3, 3))))
        ops.append(("sv_swapGate", nq - 1, int(rng.integers(0, L))))
        if k >= 2:
            ops.append(("sv_swapGate", nq - 1, nq - 2))
        ops.append(("sv_pauliTensor", [nq - 2, 3], [2, 1]))
    ops.append(("sv_manyTargGate", [nq - 1, 4, 0], cases.random_matrix(rng, 8) / 8 ** 0.5))      # ends with displaced qubits
    amps = cases.random_state(rng, nq)
    o = capi.OracleState("sv", nq, nodes)
    o.set_amps(amps)
    for op in ops:
        cases.apply(o, op)
    want = o.get_amps()
    # a density matrix: unitaries ride the lazy layout, the channels force a restore in between
    N = 6
    dm_ops = [("dm_manyTargGate", [N - 1, 0, 2], cases.random_matrix(rng, 8) / 8 ** 0.5), ("dm_pauliGadget", [N - 1, 1], [2, 3], 0.3),
              ("dm_damping", N - 1, 0.2), ("dm_manyTargGate", [N - 1, N - 2], cases.random_matrix(rng, 4) / 2), ("dm_swapGate", N - 1, N - 2),
              ("dm_oneQubitDepolarising", 1, 0.1), ("dm_krausMap", [N - 1], [cases.random_matrix(rng, 2) / 2 for _ in range(3)]),
              ("dm_manyTargGate", [N - 1, 3, 1], cases.random_matrix(rng, 8) / 8 ** 0.5)]
    dm_amps = cases.random_state(rng, 2 * N)
    od = capi.OracleState("dm", N, nodes)
    od.set_amps(dm_amps)
    for op in dm_ops:
        cases.apply(od, op)
    jobs = [dict(kind="sv", nq=nq, ops=ops, amps=amps), dict(kind="dm", nq=N, ops=dm_ops, amps=dm_amps)]
    for lazy in ("1", "0"):
        res = product.run_cases_multirank(jobs, nodes, extra_env={"DFSA_LAZY_LAYOUT": lazy})
        compare.assert_close(res[0]["amps"], want, tol=1e-11, what="lazy layout=%s np=%d" % (lazy, nodes))
        compare.assert_close(res[1]["amps"], od.get_amps(), tol=1e-11, what="lazy layout=%s np=%d (dm)" % (lazy, nodes))
        for r, bits in ((res[0], nq), (res[1], 2 * N)):
            assert r["layout_after_readback"] == list(range(bits))
            assert (r["layout_before_readback"] != list(range(bits))) == (lazy == "1"), (lazy, r["layout_before_readback"])


def test_chunk_pipelined_exchange_matches_oracle():
    """Forces the chunked exchange+combine pipeline (16 chunks even on small shards) on the NCCL transport; on a
    single-GPU box the ranks share the device, the IPC transport is used and the same results must come out."""
    rng = np.random.default_rng(77)
    nodes, nq = 2, 12
    ops = []
    for i in range(12):
        ops.append(("sv_oneTargGate", nq - 1, cases.random_matrix(rng, 2) / 1.5))
        ops.append(("sv_pauliGadget", [nq - 1, 3, 7], [1 + i % 2, 3, 1], float(rng.uniform(-3, 3))))
        ops.append(("sv_pauliTensor", [nq - 1, 1], [2, 1]))
        ops.append(("sv_swapGate", nq - 1, nq - 2))
        ops.append(("sv_manyCtrlOneTargGate", [nq - 2], nq - 1, cases.random_matrix(rng, 2) / 1.5))
    amps = cases.random_state(rng, nq)
    o = capi.OracleState("sv", nq, nodes)
    o.set_amps(amps)
    for op in ops:
        cases.apply(o, op)
    res = product.run_cases_multirank([dict(kind="sv", nq=nq, ops=ops, amps=amps)], nodes,
                                      extra_env={"DFSA_XCHG_MIN_CHUNK_LOG2": "4", "DFSA_XCHG_CHUNKS": "16"})
    compare.assert_close(res[0]["amps"], o.get_amps(), tol=1e-11, what="chunked exchange (%s)" % res[0]["transport"])
