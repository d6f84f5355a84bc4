"""Host-side model of the shared-memory layout of gemm_dmma_tma_kernel (gemm_hls_b200/csrc/gemm_dmma.cu):
the TMA 128-byte swizzle plus the kernel's slot permutations must (1) hand every m8n8k4 fragment slot the
matrix element the accumulator mapping of the epilogue assumes and (2) be free of bank conflicts.
Pure arithmetic, no GPU: it pins the address formulas the kernel uses."""
import itertools

import pytest

BK, BN, WM, WN, NJ = 32, 128, 2, 4, 4


def perm16(g):
    return (g % 2) + 2 * (g // 4) + 8 * ((g // 2) % 2)


def swizzled(row, byte_in_row):
    """Offset of a tile element inside a [rows][128 B] TMA SWIZZLE_128B tile (1024-byte aligned)."""
    return row * 128 + (((byte_in_row // 16) ^ (row % 8)) * 16) + byte_in_row % 16


def c_row(ta, bm, wr, i, g):
    wrows = bm // WM
    return wr * wrows + 16 * (i // 2) + (4 * (i % 2) + perm16(g) if ta else (i % 2) + 2 * g)


def c_col(wc, j, n):  # column-slot n (0..7) of accumulator tile j
    return wc * 32 + 16 * (j // 2) + 4 * (j % 2) + perm16(n)


def a_address(ta, bm, wr, i, s, g, q):
    wrows = bm // WM
    sw_base = q * 128 + (((perm16(g) // 2) ^ q) * 16) + (perm16(g) % 2) * 8
    if ta:
        base = (wr * wrows // 16) * 4096 + sw_base
        return (base + (i // 2) * 4096 + s * 512) ^ (((2 * (i % 2)) ^ (4 * (s % 2))) * 16)
    base = (wr * wrows + 2 * g) * 128 + (((q // 2) ^ ((2 * g) % 8)) * 16) + (q % 2) * 8
    return (base + (s // 4) * (bm * 128) + (16 * (i // 2) + (i % 2)) * 128) ^ (((2 * (s % 4)) ^ (i % 2)) * 16)


def b_address(wc, j, s, g, q):
    sw_base = q * 128 + (((perm16(g) // 2) ^ q) * 16) + (perm16(g) % 2) * 8
    base = wc * (NJ // 2) * 4096 + sw_base
    return (base + (j // 2) * 4096 + s * 512) ^ (((2 * (j % 2)) ^ (4 * (s % 2))) * 16)


def a_tile_offsets(ta, bm):
    """{smem offset: (row, k)} as the producer's TMA boxes lay the A k-tile out."""
    where = {}
    for r, k in itertools.product(range(bm), range(BK)):
        if ta:  # boxes of [32 k][16 rows], one per 16 rows
            off = (r // 16) * 4096 + swizzled(k, (r % 16) * 8)
        else:   # boxes of [bm rows][16 k], one per 16 k
            off = (k // 16) * bm * 128 + swizzled(r, (k % 16) * 8)
        where[off] = (r, k)
    return where


def b_tile_offsets():
    return {(c // 16) * 4096 + swizzled(k, (c % 16) * 8): (k, c) for k, c in itertools.product(range(BK), range(BN))}


@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("bm", [128, 64])
def test_a_fragment_slots_read_the_rows_the_epilogue_stores(ta, bm):
    where = a_tile_offsets(ta, bm)
    assert len(where) == bm * BK
    mi = bm // (WM * 8)
    for wr, i, s, g, q in itertools.product(range(WM), range(mi), range(BK // 4), range(8), range(4)):
        assert where[a_address(ta, bm, wr, i, s, g, q)] == (c_row(ta, bm, wr, i, g), 4 * s + q)
    rows = sorted(c_row(ta, bm, wr, i, g) for wr, i, g in itertools.product(range(WM), range(mi), range(8)))
    assert rows == list(range(bm))  # every row of the tile is owned exactly once


def test_b_fragment_slots_read_the_columns_the_epilogue_stores():
    where = b_tile_offsets()
    for wc, j, s, g, q in itertools.product(range(WN), range(NJ), range(BK // 4), range(8), range(4)):
        assert where[b_address(wc, j, s, g, q)] == (4 * s + q, c_col(wc, j, g))
    cols = sorted(c_col(wc, j, n) for wc, j, n in itertools.product(range(WN), range(NJ), range(8)))
    assert cols == list(range(BN))
    # a thread's two C values (column-slots 2q, 2q+1) are adjacent and 16-byte aligned: one double2 store
    for wc, j, q in itertools.product(range(WN), range(NJ), range(4)):
        assert c_col(wc, j, 2 * q) % 2 == 0 and c_col(wc, j, 2 * q + 1) == c_col(wc, j, 2 * q) + 1


def _half_warp_conflict_free(addresses):
    banks = [(a % 128) // 8 for a in addresses]  # sixteen 8-byte words of one 128-byte bank row
    return len(set(banks)) == 16


@pytest.mark.parametrize("ta", [False, True])
def test_fragment_loads_are_bank_conflict_free(ta):
    for half in range(2):
        lanes = [(lane // 4, lane % 4) for lane in range(16 * half, 16 * half + 16)]
        for wr, i, s in itertools.product(range(WM), range(8), range(BK // 4)):
            assert _half_warp_conflict_free([a_address(ta, 128, wr, i, s, g, q) for g, q in lanes])
        for wc, j, s in itertools.product(range(WN), range(NJ), range(BK // 4)):
            assert _half_warp_conflict_free([b_address(wc, j, s, g, q) for g, q in lanes])
