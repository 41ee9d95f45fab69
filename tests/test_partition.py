"""Host-side check (no GPU) of how k_matvec cuts a K-quant launch into per-CTA tile ranges and per-warp block ranges
(csrc/matvec.cuh: TileSpace, matvec_launch_shape), through ctb_matvec_partition.  The device code uses the same functions."""
import ctypes as C

import numpy as np
import pytest

Q4_K, Q5_K, Q6_K = 12, 13, 14
COST = {Q4_K: 9, Q5_K: 11, Q6_K: 13}     # relative cost of a tile = bytes per block / 16

SHAPES = [
    # (name, K, [(type, rows), ...])
    ("7b_qkv_mixed", 4096, [(Q4_K, 4096), (Q4_K, 4096), (Q6_K, 4096)]),
    ("7b_qkv_q4", 4096, [(Q4_K, 4096), (Q4_K, 4096), (Q4_K, 4096)]),
    ("7b_attn_out", 4096, [(Q4_K, 4096)]),
    ("7b_gate_up", 4096, [(Q4_K, 11008), (Q4_K, 11008)]),
    ("7b_down_q6", 11008, [(Q6_K, 4096)]),
    ("7b_head", 4096, [(Q6_K, 32000)]),
    ("13b_gate_up", 5120, [(Q4_K, 13824), (Q4_K, 13824)]),
    ("13b_down", 13824, [(Q4_K, 5120)]),
    ("falcon_qkv_up", 4608, [(Q5_K, 4736), (Q5_K, 18432)]),
    ("falcon_down", 18432, [(Q5_K, 4608)]),
    ("tiny", 256, [(Q4_K, 3)]),
    ("ragged", 512, [(Q6_K, 77), (Q4_K, 9)]),
]


def partition(lib, K, segs, n_sm=148):
    types = (C.c_int * len(segs))(*[t for t, _ in segs])
    rows = (C.c_int * len(segs))(*[m for _, m in segs])
    first = (C.c_int * (n_sm + 2))()
    meta = (C.c_int * 8)()
    assert lib.ctb_matvec_partition(types, rows, len(segs), K, n_sm, first, meta) == 0
    grid = meta[0]
    return list(first[:grid + 1]), dict(grid=grid, smem=meta[1], def_max=meta[2], tiles=meta[3], warps=meta[4], rows_per_tile=meta[5], smem_limit=meta[6])


@pytest.mark.parametrize("name,K,segs", SHAPES, ids=[s[0] for s in SHAPES])
def test_cta_ranges_cover_all_tiles_in_order_and_are_cost_balanced(lib, name, K, segs):
    first, m = partition(lib, K, segs)
    rpt = m["rows_per_tile"]
    tiles = [(rows + rpt - 1) // rpt for _, rows in segs]
    assert m["tiles"] == sum(tiles)
    assert 1 <= m["grid"] <= 148 and m["grid"] <= m["tiles"]
    assert first[0] == 0 and first[-1] == m["tiles"]
    assert all(a <= b for a, b in zip(first, first[1:])), "CTA ranges must be contiguous and ordered"
    # cumulative cost at every boundary is within one tile of the ideal split
    cost_of_tile = np.concatenate([np.full(n, COST[t]) for (t, _), n in zip(segs, tiles)])
    cum = np.concatenate([[0], np.cumsum(cost_of_tile)])
    total = cum[-1]
    for c, t in enumerate(first):
        assert abs(cum[t] - total * c / m["grid"]) <= max(COST.values()), (c, t)
    # shared memory: fits the limit, and the parked-terms buffer is at least one block deep
    assert m["smem"] <= m["smem_limit"] and m["def_max"] >= 1


@pytest.mark.parametrize("name,K,segs", SHAPES, ids=[s[0] for s in SHAPES])
def test_warp_ranges_are_equal_and_parking_covers_every_mid_row_segment(lib, name, K, segs):
    """Inside a CTA the blocks of its tiles are cut into `warps` equal contiguous ranges; a range that starts inside a row
    parks at most def_max blocks before it needs its predecessor's state — for the shapes of the bench models the whole
    segment fits (no serialised chain)."""
    first, m = partition(lib, K, segs)
    nb, W = K // 256, m["warps"]
    worst_unparked = 0
    for t0, t1 in zip(first, first[1:]):
        B = (t1 - t0) * nb
        L = -(-B // W) if B else 0
        covered = 0
        for w in range(W):
            s0, e0 = min(B, w * L), min(B, w * L + L)
            covered += e0 - s0
            a0 = s0 % nb
            mid_len = min(nb - a0, e0 - s0) if a0 else 0
            worst_unparked = max(worst_unparked, mid_len - m["def_max"])
        assert covered == B
    if name.startswith(("7b", "13b")):
        assert worst_unparked <= 0, f"a mid-row segment of {worst_unparked + m['def_max']} blocks does not fit the parking buffer"


def test_rejects_what_the_kernel_cannot_run(lib):
    one = (C.c_int * 1)
    first, meta = (C.c_int * 150)(), (C.c_int * 8)()
    assert lib.ctb_matvec_partition(one(Q4_K), one(8), 1, 300, 148, first, meta) != 0     # K not a multiple of 256
    assert lib.ctb_matvec_partition(one(2), one(8), 1, 256, 148, first, meta) != 0        # Q4_0 is not a K-quant
    assert lib.ctb_matvec_partition(one(Q4_K), one(0), 1, 256, 148, first, meta) != 0
