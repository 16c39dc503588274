"""-m gpu: runs of one-target gates fused into shared passes over HBM (csrc/dfsa_kernels_fused.cu, host/states.hpp gateQueue).
The fused path must give results that are BIT-IDENTICAL to the same gates applied one kernel each (same per-gate arithmetic,
same order) -- compared on the device with == on every double -- and within 1e-12 of the oracle."""
import numpy as np
import pytest

import cases
import compare
import product
from oracle import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dfsa():
    m = product.pkg()
    m.comm_init()
    yield m
    m.set_gate_fusion(True)


def random_gate_run(rng, nq, n, unitary=True):
    ops = []
    for _ in range(n):
        t = int(rng.integers(0, nq))
        g = cases.random_matrix(rng, 2)
        if unitary:
            g = np.linalg.qr(g)[0]
        nc = int(rng.integers(0, 4))
        if nc == 0:
            ops.append(("sv_oneTargGate", t, g))
        else:
            ops.append(("sv_manyCtrlOneTargGate", [int(c) for c in rng.permutation([q for q in range(nq) if q != t])[:nc]], t, g))
    return ops


@pytest.mark.parametrize("nq", [11, 12, 17, 22])
def test_fused_run_is_bit_identical_to_gate_by_gate(dfsa, nq):
    rng = np.random.default_rng(nq)
    for trial in range(3):
        ops = random_gate_run(rng, nq, int(rng.integers(2, 70)))
        states = {}
        for fused in (True, False):
            dfsa.set_gate_fusion(fused)
            st = dfsa.DeviceState("sv", nq)
            st.init_hash(7 + trial)
            for op in ops:
                cases.apply(st, op)
            if fused:
                assert st.pending_gates() == len(ops)          # nothing has been launched yet
            states[fused] = st
        d, ne, _ = states[True].compare(states[False])
        assert ne == 0 and d == 0.0, "fused vs gate-by-gate: %d amplitudes differ (max %.3e)" % (ne, d)
        o = capi.OracleState("sv", nq, 1)
        o.init_hash(7 + trial)
        for op in ops:
            cases.apply(o, op)
        compare.assert_close(states[True].get_amps(), o.get_amps(), tol=1e-12, what="fused run vs oracle")
        for st in states.values():
            st.close()


def test_fused_bench_sweep_26_qubits_bit_identical_and_interleaved_ops_flush(dfsa):
    import bench
    nq = 26
    ops = bench.make_sweep(nq)
    # other kinds of gate in between force the queue out in the right order
    mixed = ops[:20] + [("sv_swapGate", 3, 20), ("sv_phaseGadget", [0, 25], 0.3)] + ops[20:45] + [("sv_manyTargGate", [25, 1, 7], np.linalg.qr(cases.random_matrix(np.random.default_rng(1), 8))[0])] + ops[45:]
    res = {}
    for fused in (True, False):
        dfsa.set_gate_fusion(fused)
        st = dfsa.DeviceState("sv", nq)
        st.init_hash(26)
        for op in mixed:
            cases.apply(st, op)
        res[fused] = st
    d, ne, _ = res[True].compare(res[False])
    assert ne == 0 and d == 0.0
    # and the inverse sweep brings the fused state back
    dfsa.set_gate_fusion(True)
    for op in bench.inverse_ops(mixed):
        cases.apply(res[True], op)
    assert res[True].compare_hash(26)[0] <= 1e-12
    for st in res.values():
        st.close()


def test_comm_synch_launches_pending_gates(dfsa):
    dfsa.set_gate_fusion(True)
    st = dfsa.DeviceState("sv", 14)
    st.init_hash(1)
    st.sv_oneTargGate(3, np.eye(2))
    st.sv_manyCtrlOneTargGate([1], 5, np.eye(2))
    assert st.pending_gates() == 2
    # what the flush will do (host/layout.hpp planFlush through the C API): one run of both gates, on their own index bits
    steps, layout_after = st.plan_pending_flush()
    assert steps == [("gates", [(3, 0), (5, 1 << 1)])] and layout_after == list(range(14))
    dfsa.comm_synch()
    assert st.pending_gates() == 0
    st.close()


@pytest.mark.parametrize("nodes", [2, 8])
def test_fused_runs_at_several_ranks_match_oracle(nodes):
    """prefix controls gate whole ranks, prefix targets flush the queue and go through the exchange path"""
    rng = np.random.default_rng(50 + nodes)
    nq = 15
    ops = random_gate_run(rng, nq, 80)
    amps = cases.random_state(rng, nq)
    o = capi.OracleState("sv", nq, nodes)
    o.set_amps(amps)
    for op in ops:
        cases.apply(o, op)
    want = o.get_amps()
    got = {}
    for fuse in ("1", "0"):
        res = product.run_cases_multirank([dict(kind="sv", nq=nq, ops=ops, amps=amps)], nodes, extra_env={"DFSA_FUSE_GATES": fuse})
        compare.assert_close(res[0]["amps"], want, tol=1e-11, what="fused=%s np=%d" % (fuse, nodes))
        got[fuse] = res[0]["amps"]
    compare.assert_exact(got["1"], got["0"], what="fused vs gate-by-gate at %d ranks" % nodes)
