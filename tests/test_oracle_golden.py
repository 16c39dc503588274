"""CPU tests (-m "not gpu"): pin the C restatement (oracle/dfsa_oracle.c) against
(1) the committed outputs of the real patched reference at 1/2/4/8 ranks (tests/golden/),
(2) an independent dense ground truth (oracle/dense.py),
(3) the live reference build oracle/_ref/ref_driver when it is present.
"""
import numpy as np
import pytest

import cases
import compare
import golden_io
from oracle import capi, dense, refrun

ALL_GOLDEN = golden_io.load("sv") + golden_io.load("dm")


def run_oracle(case):
    st = capi.OracleState(case["kind"], case["nq"], case["nodes"])
    st.set_amps(case["amps"])
    res = cases.apply(st, case["op"])
    return st, res


@pytest.mark.parametrize("case", ALL_GOLDEN, ids=golden_io.case_id)
def test_oracle_matches_reference_golden(case):
    st, res = run_oracle(case)
    name = case["op"][0]
    if name == "dm_expecPauliString":
        compare.assert_value_close(res, case["val"][0], what=name)
    elif name == "dm_partialTrace":
        compare.assert_exact(res.get_amps(), case["out"], what=name)           # pure sums in the same order
        compare.assert_exact(st.get_amps(), case["mut"], what=name + " (mutated input)")
    elif name in cases.EXACT_OPS:
        compare.assert_exact(st.get_amps(), case["out"], what=name)
    else:
        compare.assert_close(st.get_amps(), case["out"], what=name)


# twoQubitDepolarising is excluded: the reference's formulas are not the channel;
# parity with the reference is what the golden test above pins.
DENSE_CASES = [c for c in ALL_GOLDEN if c["op"][0] != "dm_twoQubitDepolarising" and c["id"] % 3 == 0]


@pytest.mark.parametrize("case", DENSE_CASES, ids=golden_io.case_id)
def test_oracle_matches_dense_truth(case):
    st, res = run_oracle(case)
    name = case["op"][0]
    truth = dense.apply_op(case["kind"], case["nq"], case["amps"], case["op"])
    if name == "dm_expecPauliString":
        compare.assert_value_close(res, truth, what=name)
    elif name == "dm_partialTrace":
        compare.assert_close(res.get_amps(), truth, what=name)
    else:
        compare.assert_close(st.get_amps(), truth, tol=1e-11 if name == "dm_krausMap" else compare.TOL, what=name)


def test_all_z_pauli_quirk_and_fix():
    """The reference is a no-op for Pauli strings without X/Y when no rank exchange happens
    (local loop has 2^0/2 = 0 inner iterations, src/local_statevector.hpp:108). The oracle reproduces that
    by default and applies the true diagonal operator with quirks off."""
    rng = np.random.default_rng(7)
    amps = cases.random_state(rng, 6)
    op = ("sv_pauliGadget", [1, 4], [3, 3], 0.37)
    st = capi.OracleState("sv", 6, 2)
    st.set_amps(amps)
    cases.apply(st, op)
    compare.assert_exact(st.get_amps(), amps, what="quirk no-op")
    capi.lib().orc_set_quirks(0)
    try:
        st = capi.OracleState("sv", 6, 2)
        st.set_amps(amps)
        cases.apply(st, op)
        compare.assert_close(st.get_amps(), dense.apply_op("sv", 6, amps, op))
    finally:
        capi.lib().orc_set_quirks(1)


def test_hash_state_is_reproducible_in_numpy():
    """The synthetic bench state (splitmix64 of the global index) restated in numpy."""
    st = capi.OracleState("sv", 8, 4)
    st.init_hash(12345)
    got = st.get_amps()
    k = np.arange(2 * 256, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(12345) + (k + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    vals = (z >> np.uint64(11)).astype(np.float64) / 9007199254740992.0 - 0.5
    compare.assert_exact(got, vals[0::2] + 1j * vals[1::2])


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref/ref_driver not built (no /root/reference here)")
@pytest.mark.parametrize("nodes", [1, 2, 4, 8])
def test_oracle_matches_live_reference_on_a_circuit(nodes):
    """A multi-op circuit (ops chained on one state) through the live reference vs the oracle."""
    rng = np.random.default_rng(100 + nodes)
    k = nodes.bit_length() - 1
    for kind, nq, names in (("sv", 7, cases.SV_OPS), ("dm", 4, [n for n in cases.DM_OPS if n not in ("dm_partialTrace", "dm_expecPauliString")])):
        nbits = nq if kind == "sv" else 2 * nq
        ops = [cases.make_op(rng, names[i % len(names)], nq, k) for i in range(2 * len(names))]
        # keep magnitudes O(1): normalise the random non-unitary gates
        ops = [tuple((a / np.linalg.norm(a, 2) if isinstance(a, np.ndarray) and a.ndim == 2 and a.dtype == np.complex128 else a) for a in op) for op in ops]
        amps = cases.random_state(rng, nbits)
        st = capi.OracleState(kind, nq, nodes)
        st.set_amps(amps)
        for op in ops:
            cases.apply(st, op)
        ref = refrun.run(kind, nq, ops, num_nodes=nodes, init_amps=amps)
        compare.assert_close(st.get_amps(), ref["amps"], what="%s circuit np=%d" % (kind, nodes))
