"""CPU test of the lazy qubit layout's host logic (host/layout.hpp): after any sequence of relocations and rank relabellings,
the restore plan (one relocation step of disjoint suffix<->prefix pairs + index-bit swaps) must put every logical qubit back
on its own index bit -- replayed here on a numpy array whose entries are their own original indices."""
import ctypes as C

import numpy as np
import pytest

import product


def swap_index_bits(arr, a, b):
    """What swapIndexBits / a relocation pair does to the data: new[i] = old[i with bits a and b exchanged]."""
    idx = np.arange(arr.size, dtype=np.int64)
    ba, bb = (idx >> a) & 1, (idx >> b) & 1
    src = idx ^ (((ba ^ bb) << a) | ((ba ^ bb) << b))
    return arr[src]


def relabel(where, a, b):
    return [b if w == a else (a if w == b else w) for w in where]


def restore_plan(where, L):
    h = product.pkg().host_lib()
    h.dfsa_host_plan_restoreLayout.restype = C.c_uint
    n = len(where)
    out = (C.c_uint * (3 * 3 * n))()
    steps = h.dfsa_host_plan_restoreLayout((C.c_uint * n)(*where), n, L, out)
    return [(out[3 * i], out[3 * i + 1], out[3 * i + 2]) for i in range(steps)]


@pytest.mark.parametrize("n,k", [(8, 1), (9, 2), (10, 3), (10, 4)])
def test_restore_plan_returns_every_qubit_to_its_own_index_bit(n, k):
    rng = np.random.default_rng(n * 10 + k)
    L = n - k
    for trial in range(60):
        data = np.arange(1 << n)
        where = list(range(n))                   # where[q] = index bit holding logical qubit q
        for step in range(int(rng.integers(1, 6))):
            kind = rng.integers(0, 3)
            if kind == 0:                        # manyTargGate: disjoint (free suffix bit, prefix bit) pairs trade places
                npairs = int(rng.integers(1, k + 1))
                pre = [int(x) for x in rng.permutation(np.arange(L, n))[:npairs]]
                suf = [int(x) for x in rng.permutation(L)[:npairs]]
                for a, b in zip(suf, pre):
                    data = swap_index_bits(data, a, b)
                    where = relabel(where, a, b)
            elif kind == 1 and k >= 2:           # swapGate of two qubits that both sit on rank bits: relabelling only
                on_prefix = [q for q in range(n) if where[q] >= L]
                q1, q2 = [int(x) for x in rng.permutation(on_prefix)[:2]]
                where[q1], where[q2] = where[q2], where[q1]          # no amplitude moves
            else:                                # swapGate with a suffix qubit: a real index-bit swap, layout unchanged
                a, b = [int(x) for x in rng.permutation(n)[:2]]
                pa, pb = where[a], where[b]
                data = swap_index_bits(data, pa, pb)
        # invariant used by the product: the array indexed by PHYSICAL bits holds the logical state with qubit q on bit where[q].
        # Build the logical state by un-permuting, then check that the plan realises exactly that un-permutation.
        logical = np.empty_like(data)
        idx = np.arange(data.size, dtype=np.int64)
        phys = np.zeros_like(idx)
        for q in range(n):
            phys |= ((idx >> q) & 1) << where[q]
        logical[idx] = data[phys]
        got = data
        w = list(where)
        for kind, a, b in restore_plan(where, L):
            if kind == 0:
                assert a < L <= b, "relocation pairs are (suffix bit, prefix bit)"
            got = swap_index_bits(got, a, b)
            w = relabel(w, a, b)
        assert w == list(range(n))
        assert np.array_equal(got, logical)
        steps = restore_plan(where, L)
        assert sum(1 for s in steps if s[0] == 0) <= 4
        if where == list(range(n)):
            assert steps == []


def relocation_on_layout(where, L, targets):
    h = product.pkg().host_lib()
    n, nt = len(where), len(targets)
    out = (C.c_uint * nt)()
    h.dfsa_host_plan_relocationOnLayout((C.c_uint * n)(*where), n, L, (C.c_uint * nt)(*targets), nt, out)
    return list(out)


@pytest.mark.parametrize("n,k", [(10, 1), (12, 3), (12, 4)])
def test_relocation_lands_where_the_restore_stays_cheap(n, k):
    """dfsa_planRelocationOnLayout (host/distributed_statevector.hpp): every prefix target lands on a distinct free suffix
    bit, suffix targets stay; a displaced qubit whose home is the target's prefix bit is brought home; over a random
    circuit of dense gates the layout never holds a rank relabelling (a qubit whose home is a rank bit sitting on ANOTHER
    rank bit), so the restore plan never needs a full-shard exchange."""
    rng = np.random.default_rng(n + k)
    L = n - k
    for trial in range(200):
        where = list(range(n))
        forced = False                                             # a landing had neither a home slot nor a clean slot to go to
        for gate in range(int(rng.integers(1, 9))):
            logical = [int(x) for x in rng.permutation(n)[: int(rng.integers(1, min(L, 6) + 1))]]
            targets = [where[q] for q in logical]                  # index bits
            placed = relocation_on_layout(where, L, targets)
            assert all(p < L for p in placed) and len(set(placed)) == len(placed)
            before = list(where)
            used = set()
            for t, p in zip(targets, placed):
                if t < L:
                    assert p == t
                    continue
                home_now = before[t]                               # where the qubit whose home is bit t sits
                free = [s_ for s_ in range(L) if s_ not in targets and s_ not in used]
                if home_now in free:
                    assert p == home_now, "a displaced qubit must be brought home"
                elif any(before[s_] == s_ for s_ in free):
                    assert before[p] == p, "a slot holding its own qubit was available"
                else:
                    forced = True
                used.add(p)
                where = relabel(where, p, t)
            if not forced:
                for q in range(L, n):
                    assert where[q] < L or where[q] == q, "rank relabelling in the layout: %r" % (where,)
        if not forced:
            for kind, a, b in restore_plan(where, L):
                assert not (a >= L and b >= L), "restore needs a full-shard exchange: %r" % (where,)
