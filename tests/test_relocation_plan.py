"""CPU test of the single-shot relocation plan (dfsa_plan_relocate, the host half of dfsa_xk_relocate): gathering
new[j] = shard(owners[sigma(j)])[j with its landing bits := rho] on every rank must equal the reference's sequence of
suffix<->prefix qubit swaps (distributed_statevector.hpp:213-223) applied to the global state, for any number of pairs,
any rank count, any qubit choice; and the plan must be symmetric (if R reads from R', R' reads from R)."""
import ctypes as C

import numpy as np
import pytest

import product


def plan(rank, L, prefix):
    lib = product.pkg().device_lib()
    owners = (C.c_int * 16)()
    rho = C.c_uint()
    rc = lib.dfsa_plan_relocate(rank, L, (C.c_uint32 * len(prefix))(*prefix), len(prefix), owners, C.byref(rho))
    assert rc == 0, lib.dfsa_last_error()
    return list(owners), rho.value


def swap_bits(idx, a, b):
    x = ((idx >> a) ^ (idx >> b)) & 1
    return idx ^ ((x << a) | (x << b))


@pytest.mark.parametrize("log_nodes", [1, 2, 3, 4])
def test_gather_equals_the_sequence_of_swaps(log_nodes):
    rng = np.random.default_rng(log_nodes)
    L, P = 6, 1 << log_nodes
    n = L + log_nodes
    for k in range(1, log_nodes + 1):
        for _ in range(6):
            prefix = [int(x) for x in rng.permutation(np.arange(L, n))[:k]]
            landing = [int(x) for x in rng.permutation(L)[:k]]
            state = rng.standard_normal(1 << n)
            # reference: swap qubit pairs one after the other on the global state (new[i] = old[i with the two bits exchanged])
            want = state.copy()
            for s_q, p_q in zip(landing, prefix):
                want = want[[swap_bits(i, s_q, p_q) for i in range(1 << n)]]
            shards = state.reshape(P, 1 << L)
            s_mask = sum(1 << q for q in landing)
            for rank in range(P):
                owners, rho = plan(rank, L, prefix)
                assert owners[1 << k:] == [-1] * (16 - (1 << k))
                rho_bits = sum(((rho >> i) & 1) << landing[i] for i in range(k))
                got = np.empty(1 << L)
                for j in range(1 << L):
                    sigma = sum(((j >> landing[i]) & 1) << i for i in range(k))
                    got[j] = shards[owners[sigma]][(j & ~s_mask) | rho_bits]
                assert np.array_equal(got, want.reshape(P, 1 << L)[rank]), (rank, prefix, landing)
                # symmetry: the rank I read sigma from reads rho from me; sigma == rho is my own shard
                assert owners[rho] == rank
                for sigma in range(1 << k):
                    theirs, their_rho = plan(owners[sigma], L, prefix)
                    assert their_rho == sigma and theirs[rho] == rank


@pytest.mark.parametrize("log_nodes", [2, 3, 4])
def test_staggered_walk_is_a_bijection_and_a_perfect_matching_at_every_step(log_nodes):
    """The gather kernel (csrc/dfsa_kernels_sv.cu dfsaLaunchRelocate) takes item jj as output index j = jj ^ rhoBits, so that every
    rank starts with its own shard: all ranks sweep jj in the same order at about the same speed, and at the same jj the ranks of
    a group must read from pairwise DIFFERENT owners (each owner's NVLink egress serves one reader at a time) -- and the items
    must still cover every output index exactly once."""
    rng = np.random.default_rng(40 + log_nodes)
    L, P = 5, 1 << log_nodes
    n = L + log_nodes
    for k in range(2, log_nodes + 1):
        for _ in range(4):
            prefix = [int(x) for x in rng.permutation(np.arange(L, n))[:k]]
            landing = [int(x) for x in rng.permutation(L)[:k]]
            group_mask = sum(1 << (p - L) for p in prefix)
            readers = {}                                   # (group id, jj) -> owners read at that item by the ranks of the group
            for rank in range(P):
                owners, rho = plan(rank, L, prefix)
                rho_bits = sum(((rho >> i) & 1) << landing[i] for i in range(k))
                outputs = set()
                for jj in range(1 << L):
                    j = jj ^ rho_bits
                    outputs.add(j)
                    sigma = sum(((j >> landing[i]) & 1) << i for i in range(k))
                    readers.setdefault((rank & ~group_mask, jj), []).append(owners[sigma])
                    if jj == 0:
                        assert owners[sigma] == rank     # the walk starts at home
                assert outputs == set(range(1 << L))
            for (group, jj), who in readers.items():
                assert len(who) == 1 << k and len(set(who)) == 1 << k, (prefix, landing, group, jj, who)
