"""CPU test of the gate-sequence planner behind dfsa_k_gateSequence (csrc/dfsa_kernels_fused.cu): how a run of one-target gates is
split into passes over HBM (batches), which 11 index bits each pass's shared-memory tile spans, and how the gates of a pass
group into octets of three register bits. Invariants: order is preserved (batch and group numbers never decrease), every gate's
target lies in its pass's tile, every tile contains index bits 0..3 (256-byte runs), a group spans exactly three tile bits that
include the targets of all its gates, and the bench sweep on 32 qubits needs 4 passes instead of 64."""
import ctypes as C

import numpy as np
import pytest

import product


class Gate1(C.Structure):
    _fields_ = [("matrix", C.c_double * 8), ("ctrlMask", C.c_uint64), ("target", C.c_uint32), ("reserved", C.c_uint32)]


def plan(gates, L):
    lib = product.pkg().device_lib()
    n = len(gates)
    arr = (Gate1 * n)()
    for i, (t, ctrls) in enumerate(gates):
        arr[i].target = t
        arr[i].ctrlMask = sum(1 << c for c in ctrls)
    batch, group, tiled = (C.c_uint32 * n)(), (C.c_uint32 * n)(), (C.c_uint32 * n)()
    tile_bits, group_bits = (C.c_uint32 * (11 * n))(), (C.c_uint32 * (3 * n))()
    nb = C.c_uint()
    rc = lib.dfsa_plan_gateSequence(arr, n, L, batch, group, tiled, tile_bits, group_bits, C.byref(nb))
    assert rc == 0, lib.dfsa_last_error()
    return list(batch), list(group), list(tiled)[: nb.value], [list(tile_bits[11 * b:11 * b + 11]) for b in range(nb.value)], \
        [list(group_bits[3 * i:3 * i + 3]) for i in range(n)]


def check(gates, L):
    batch, group, tiled, tiles, gbits = plan(gates, L)
    assert batch == sorted(batch)
    for i, (t, ctrls) in enumerate(gates):
        b = batch[i]
        if not tiled[b]:
            assert batch.count(b) == 1
            continue
        tile = tiles[b]
        assert len(set(tile)) == 11 and tile == sorted(tile) and all(q < L for q in tile) and set(range(4)) <= set(tile)
        assert t in tile
        assert len(set(gbits[i])) == 3 and set(gbits[i]) <= set(tile) and t in gbits[i]
        if i and batch[i - 1] == b:
            assert group[i] >= group[i - 1]
            if group[i] == group[i - 1]:
                assert gbits[i] == gbits[i - 1]
    return batch, tiled


def test_bench_sweep_needs_four_passes():
    import bench
    ops = bench.make_sweep(32)
    gates = [(op[1], []) if op[0] == "sv_oneTargGate" else (op[2], op[1]) for op in ops]
    batch, tiled = check(gates, 32)
    assert len(tiled) == 4 and all(tiled)
    assert [batch.count(b) for b in range(4)] == [22, 14, 14, 14]


@pytest.mark.parametrize("L", [11, 14, 27, 33])
def test_random_sequences(L):
    rng = np.random.default_rng(L)
    for trial in range(40):
        n = int(rng.integers(1, 90))
        gates = []
        for _ in range(n):
            t = int(rng.integers(0, L))
            ctrls = [int(c) for c in rng.permutation([q for q in range(L + 3) if q != t])[: int(rng.integers(0, 4))]]
            gates.append((t, ctrls))
        check(gates, L)


def test_small_shards_are_not_tiled_and_bad_gates_are_rejected():
    batch, group, tiled, tiles, gbits = plan([(0, []), (3, [1]), (9, [])], 10)
    assert tiled == [0, 0, 0]
    lib = product.pkg().device_lib()
    arr = (Gate1 * 1)()
    arr[0].target = 5
    arr[0].ctrlMask = 1 << 5                       # controlled on its own target
    out = (C.c_uint32 * 16)()
    nb = C.c_uint()
    assert lib.dfsa_plan_gateSequence(arr, 1, 12, out, out, out, out, out, C.byref(nb)) != 0
