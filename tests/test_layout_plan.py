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
