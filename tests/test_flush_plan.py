"""CPU test of the gate queue's launch plan (host/layout.hpp planFlush, through libdfsa_host.so's dfsa_host_plan_flush): queued
one-target gates run in call order, in runs whose targets all sit on suffix bits; rank-bit qubits come into the shard in as few
relocation steps as the queue allows, evicting the qubits whose next turn is furthest away. The amplitudes are not involved here
(tests/hostsim fuzzes the same code against a dense truth); this checks the plan's shape and its cost."""
import numpy as np
import pytest

import product


def api():
    return product.pkg().api


def sweep(n):
    return [g for q in range(n) for g in ((q, []), (q, [(q + 3) % n, (q + 7) % n]))]


def check_plan_invariants(where, L, gates, steps, where_after):
    """replays the plan on the layout: every gate exactly once, in order, on a suffix bit; relocations pair suffix bits with rank bits"""
    where = list(where)
    nxt = 0
    for kind, body in steps:
        if kind == "relocate":
            assert 1 <= len(body) <= 4
            assert len({a for a, _ in body}) == len(body) and len({b for _, b in body}) == len(body)
            for a, b in body:
                assert a < L <= b < len(where)
                where = [b if w == a else (a if w == b else w) for w in where]
        else:
            assert body, "empty run"
            for t, mask in body:
                lt, lctrls = gates[nxt]
                assert t == where[lt] and t < L
                assert mask == sum(1 << where[c] for c in lctrls)
                nxt += 1
    assert nxt == len(gates)
    assert where == where_after and sorted(where) == list(range(len(where)))
    return where


def test_no_rank_bits_means_one_run():
    gates = sweep(12)
    steps, after = api().plan_flush(list(range(12)), 12, [0] * 12, gates)
    assert [k for k, _ in steps] == ["gates"] and len(steps[0][1]) == len(gates) and after == list(range(12))
    assert api().plan_flush(list(range(12)), 12, [0] * 12, []) == ([], list(range(12)))


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_sweep_pays_about_one_relocation_step_per_layer(k):
    """bench.py's workload: a layer brings the log2(P) rank-bit qubits in TOGETHER -- (1 - 2^-k) * 16A bytes per direction where
    one swap-in per qubit costs k * 8A -- and evicts the qubits it has just finished with. The evicted block walks down the
    register by k qubits per layer; once per trip (every L / k layers) it wraps around and a layer needs two steps, which is also
    what Belady's rule does on a cyclic scan with a cache that is k entries short."""
    n = 20 + k
    L = n - k
    gates = sweep(n)
    where, last_use, clock = list(range(n)), [0] * n, 0
    layers, steps_total, bytes_new = 3 * n, 0, 0.0
    for layer in range(layers):
        for t, _ in gates:                                    # StateVector::touch, at enqueue time
            clock += 1
            last_use[t] = clock
        steps, after = api().plan_flush(where, L, last_use, gates)
        check_plan_invariants(where, L, gates, steps, after)
        relocations = [body for kind, body in steps if kind == "relocate"]
        assert 1 <= len(relocations) <= 2 and all(1 <= len(r) <= k for r in relocations), (layer, steps)
        if layer == 0:
            assert [kind for kind, _ in steps] == ["gates", "relocate", "gates"]
        steps_total += len(relocations)
        bytes_new += sum((1 - 0.5 ** len(r)) * 16 for r in relocations)
        where = after
    assert steps_total <= layers * (1 + 2 * k / L) + 2, steps_total
    # against one 8A swap-in per rank-bit qubit and layer (the minimum of any policy that brings qubits in one at a time)
    assert bytes_new <= layers * k * 8 * {1: 1.08, 2: 0.85, 3: 0.72, 4: 0.62}[k], bytes_new / (layers * k * 8)


def test_random_queues_keep_order_and_never_evict_a_qubit_needed_sooner():
    rng = np.random.default_rng(5)
    for trial in range(300):
        n = int(rng.integers(5, 14))
        k = int(rng.integers(1, min(4, n - 2) + 1))
        L = n - k
        where = [int(x) for x in rng.permutation(n)]
        last_use = [int(x) for x in rng.integers(0, 50, n)]
        gates = []
        for _ in range(int(rng.integers(1, 60))):
            t = int(rng.integers(0, n))
            ctrls = [int(c) for c in rng.permutation([q for q in range(n) if q != t])[: int(rng.integers(0, 4))]]
            gates.append((t, ctrls))
        steps, after = api().plan_flush(where, L, last_use, gates)
        check_plan_invariants(where, L, gates, steps, after)
        # replay again to look at each relocation: the first pair serves the very next gate; any further pair brings in a qubit that is
        # targeted later in the queue, and the qubit it evicts is targeted later still (or never)
        cur, done = list(where), 0
        for kind, body in steps:
            if kind == "gates":
                done += len(body)
                continue
            def next_use(q):
                return next((i for i in range(done, len(gates)) if gates[i][0] == q), None)
            brought = [cur.index(b) for _, b in body]
            evicted = [cur.index(a) for a, _ in body]
            assert brought[0] == gates[done][0]
            for s, v in zip(brought, evicted):
                assert next_use(s) is not None
                assert v != gates[done][0]
                if s != brought[0]:
                    assert next_use(v) is None or next_use(v) > next_use(s)
            for a, b in body:
                cur = [b if w == a else (a if w == b else w) for w in cur]
        # one relocation step at most per gate that found its qubit on a rank bit
        assert sum(1 for kind, _ in steps if kind == "relocate") <= len(gates)
