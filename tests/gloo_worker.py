"""One rank of the CPU-side N>1 check (spawned by tests/test_gloo_host_plans.py with RANK/WORLD_SIZE/MASTER_*).
Uses torch.distributed with the gloo backend -- no GPU -- to verify across real processes that the host-side plans the
drop-in headers act on are consistent between partners: if rank r plans an exchange with rank p, then p plans the same
kind of exchange with r, of the same size (otherwise the pairwise exchange would deadlock or corrupt)."""
import ctypes as C
import hashlib
import os
import pickle
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import bench  # noqa: E402
import cases  # noqa: E402
import product  # noqa: E402
from oracle import capi  # noqa: E402

EXCHANGING = {2, 3, 4, 5}      # ExchangePlan::FullShard, SubCube, HalfContiguous, HalfPacked


def u32(xs):
    xs = [int(x) for x in xs]
    return (C.c_uint * max(1, len(xs)))(*xs), len(xs)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    host = product.pkg().host_lib()
    k = world.bit_length() - 1
    rng = np.random.default_rng(2026)           # same seed everywhere: SPMD, every rank sees the same op list
    nq = 9
    L = nq - k
    plans, ops = [], []
    for i in range(400):
        name = cases.SV_OPS[i % len(cases.SV_OPS)]
        op = cases.make_op(rng, name, nq, k, max_targets=4)
        ops.append(op)
        out = (C.c_ulonglong * 4)()
        if name == "sv_oneTargGate":
            c, n = u32([])
            host.dfsa_host_plan_ctrlOneTarg(rank, L, c, 0, op[1], out)
            plans.append(("x", list(out)))
        elif name == "sv_manyCtrlOneTargGate":
            c, n = u32(op[1])
            host.dfsa_host_plan_ctrlOneTarg(rank, L, c, n, op[2], out)
            plans.append(("x", list(out)))
        elif name == "sv_swapGate":
            host.dfsa_host_plan_swap(rank, L, op[1], op[2], out)
            plans.append(("x", list(out)))
        elif name in ("sv_pauliTensor", "sv_pauliGadget"):
            t, n = u32(op[1])
            p, _ = u32(op[2])
            host.dfsa_host_plan_pauli(rank, L, t, p, n, out)
            plans.append(("p", list(out)))
        elif name == "sv_manyTargGate":
            t, n = u32(op[1])
            placed = (C.c_uint * n)()
            host.dfsa_host_plan_manyTarg(L, t, n, placed)
            # the plan moves exactly the targets the reference's plan moves (the prefix ones) onto distinct free suffix
            # qubits; WHICH free qubits is this build's choice (highest, the reference takes the lowest -- same result)
            ref_plan = capi.plan_manyTarg(L, op[1])
            placed = list(placed)
            assert len(set(placed)) == len(placed) and all(q < L for q in placed), placed
            for tq, mine, theirs in zip(op[1], placed, ref_plan):
                assert (mine == tq) == (theirs == tq) and (mine == tq or (tq >= L and mine not in op[1])), (op[1], placed, ref_plan)
            free = [q for q in range(L - 1, -1, -1) if q not in op[1]]
            assert [m for tq, m in zip(op[1], placed) if m != tq] == free[: sum(1 for tq in op[1] if tq >= L)]
            # every relocation swap is itself a planned exchange
            for a, b in zip(placed, op[1]):
                if a != b:
                    host.dfsa_host_plan_swap(rank, L, a, b, out)
                    plans.append(("x", list(out)))
        else:
            plans.append(("n", []))
    # 1. every rank derived the same op list
    digest = hashlib.sha256(pickle.dumps([(o[0], [np.asarray(a).tolist() if isinstance(a, (np.ndarray, list)) else a for a in o[1:]]) for o in ops])).hexdigest()
    gathered = [None] * world
    dist.all_gather_object(gathered, (digest, plans))
    assert all(g[0] == digest for g in gathered), "ranks disagree on the op list"
    # 2. partner symmetry
    checked = 0
    for idx, (tag, mine) in enumerate(plans):
        if tag == "x":
            kind, pair, num, bit = mine
            if kind in EXCHANGING:
                other = gathered[pair][1][idx][1]
                assert other[0] == kind and other[1] == rank and other[2] == num, (idx, mine, other)
                if kind in (4, 5):
                    assert other[3] != bit          # the partner moves the complementary half
                checked += 1
        elif tag == "p":
            pair = mine[0]
            if pair != rank:
                other = gathered[pair][1][idx][1]
                assert other[0] == rank and other[2] == mine[2] and other[3] == mine[3], (idx, mine, other)
                checked += 1
    assert checked > 20, "too few exchanging ops were exercised"
    # 3. partialTrace planners against the oracle's
    for _ in range(50):
        N = 5
        nt = int(rng.integers(1, N - k + 1))
        targs = sorted(int(x) for x in rng.permutation(N)[:nt])
        t, n = u32(targs)
        re = (C.c_uint * (2 * n))()
        rem = (C.c_uint * (2 * N))()
        host.dfsa_host_plan_partialTrace(N, 2 * N - k, t, n, re, rem)
        o_re, o_rem = capi.plan_partialTrace(N, 2 * N - k, targs)
        assert list(re) == o_re and list(rem)[: 2 * N - 2 * n] == o_rem
    # 4. bench plumbing: identical sweep on every rank, max-over-ranks timing reduction
    sweep = bench.make_sweep(33)
    sig = hashlib.sha256(pickle.dumps([(o[0], o[1] if o[0] == "sv_oneTargGate" else (o[1], o[2])) for o in sweep])).hexdigest()
    sigs = [None] * world
    dist.all_gather_object(sigs, sig)
    assert len(set(sigs)) == 1
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t.item()) == 10.0 + world - 1
    (hb, nb, fl), = bench.op_cost(("sv_oneTargGate", 32, None), "sv", 33, 1)
    assert nb == 16 * 2 ** 32 and hb == 48 * 2 ** 32 and fl == 0
    # a 5-target gate with 2 prefix targets at 4 ranks: the local pass + a relocation before and after, 3/4 of a shard each way
    cost = bench.op_cost(("sv_manyTargGate", [33, 0, 32, 7, 9], None), "sv", 34, 2)
    assert len(cost) == 3 and cost[0] == (32 * 2.0 ** 32, 0.0, 6.0 * 32 * 2.0 ** 32) and cost[1] == cost[2] == (32 * 2.0 ** 32, 0.75 * 16 * 2.0 ** 32, 0.0)
    # the ops of one sweep followed by their inverses are the identity: same number of gates, reversed order
    inv = bench.inverse_ops(sweep)
    assert len(inv) == len(sweep) and inv[0][0] == sweep[-1][0]
    dist.barrier()
    dist.destroy_process_group()
    print("gloo worker %d/%d ok (%d symmetric exchanges checked)" % (rank, world, checked))


if __name__ == "__main__":
    main()
