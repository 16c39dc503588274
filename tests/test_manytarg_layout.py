"""CPU test of the host logic behind the tensor-core manyTargGate kernel (local_statevector.hpp:72-99 on the device):
for every kind of target placement the tile plan must cover the targets plus the lowest other index bits, the
shared-memory slab layout must be a bijection of the 512 tile elements onto 16-byte slots, and every access pattern
of the kernel -- the mover's address-order copy, the MMA operand (B-fragment) reads, the MMA result (C-fragment)
writes -- must hit eight different 16-byte bank groups per quarter-warp (no shared-memory bank conflicts)."""
import ctypes as C
import itertools

import numpy as np
import pytest

import product


def plan(targets, log_amps):
    lib = product.pkg().device_lib()
    t = (C.c_uint32 * len(targets))(*targets)
    tile, roles, bit_off = (C.c_uint32 * 9)(), (C.c_uint32 * 9)(), (C.c_uint32 * 9)()
    row_bit, col_bit = (C.c_uint32 * 6)(), (C.c_uint32 * 6)()
    rc = lib.dfsa_plan_manyTargLayout(t, len(targets), log_amps, tile, roles, bit_off, row_bit, col_bit)
    assert rc == 0, lib.dfsa_last_error()
    return list(tile), list(roles), list(bit_off), list(row_bit), list(col_bit)


def bank_groups(offsets):
    return [(o >> 4) & 7 for o in offsets]


def xor_of(bits, contributions):
    out = 0
    for k, c in enumerate(contributions):
        if (bits >> k) & 1:
            out ^= c
    return out


def check(targets, log_amps):
    nt = len(targets)
    tile, roles, bit_off, row_bit, col_bit = plan(targets, log_amps)
    # the tile: ascending index bits = the targets plus the 9 - t lowest other bits
    free = [b for b in range(log_amps) if b not in targets][: 9 - nt]
    assert tile == sorted(list(targets) + free)
    for p, (b, role) in enumerate(zip(tile, roles)):
        assert (targets[role] == b) if role < nt else (free[role - nt] == b), (targets, tile, roles)
        assert (row_bit[role] if role < nt else col_bit[role - nt]) == bit_off[p]
    # bijection onto the 512 16-byte slots of an 8 KiB slab; the two halves of a 32-byte sector stay neighbours
    offs = [xor_of(e, bit_off) for e in range(512)]
    assert sorted(offs) == list(range(0, 8192, 16))
    assert all(offs[e] ^ offs[e ^ 1] == 16 for e in range(512))
    # mover: lanes follow address order (tile bits 0..2 vary within a quarter-warp)
    for hi in range(0, 512, 8):
        assert len(set(bank_groups(offs[hi:hi + 8]))) == 8
    # B-fragment reads: the 8 lanes of a quarter-warp vary gate-row bits 0, 1 (q) and vector bit 0 (g & 1)
    b_read = [xor_of(q, row_bit[:2]) ^ xor_of(g, col_bit[:1]) for g in range(2) for q in range(4)]
    assert len(set(bank_groups(b_read))) == 8, (targets, log_amps, "B reads conflict")
    # C-fragment writes: they vary vector bits 1, 2 (2q) and gate-row bit 0 (g & 1)
    c_write = [xor_of(q, col_bit[1:3]) ^ xor_of(g, row_bit[:1]) for g in range(2) for q in range(4)]
    assert len(set(bank_groups(c_write))) == 8, (targets, log_amps, "C writes conflict")


@pytest.mark.parametrize("nt", [3, 4, 5, 6])
def test_every_placement_on_a_small_shard_is_conflict_free(nt):
    log_amps = 11
    for combo in itertools.combinations(range(log_amps), nt):
        check(list(combo), log_amps)
        check(list(reversed(combo)), log_amps)


@pytest.mark.parametrize("nt", [3, 4, 5, 6])
def test_random_placements_and_orders_on_large_shards(nt):
    rng = np.random.default_rng(nt)
    for log_amps in (9, 20, 33):
        for _ in range(150):
            check([int(x) for x in rng.permutation(log_amps)[:nt]], log_amps)


def test_plan_rejects_what_the_kernel_does_not_serve():
    lib = product.pkg().device_lib()
    out9, out6 = (C.c_uint32 * 9)(), (C.c_uint32 * 6)()
    assert lib.dfsa_plan_manyTargLayout((C.c_uint32 * 2)(0, 1), 2, 12, out9, out9, out9, out6, out6) != 0      # t = 2: stream kernel
    assert lib.dfsa_plan_manyTargLayout((C.c_uint32 * 3)(0, 1, 2), 3, 8, out9, out9, out9, out6, out6) != 0    # shard < one tile
    assert lib.dfsa_plan_manyTargLayout((C.c_uint32 * 3)(0, 1, 1), 3, 12, out9, out9, out9, out6, out6) != 0   # duplicate target
