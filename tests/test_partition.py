"""Host-side check (no GPU) of how the step kernel cuts a K-quant mat-vec phase into per-CTA 16-row tile ranges and work items
(csrc/stream.cuh: TileSpace, st_chunk_blocks), through ctb_matvec_partition.  The device code uses the same functions."""
import ctypes as C

import numpy as np
import pytest

Q4_K, Q5_K, Q6_K = 12, 13, 14
COST = {Q4_K: 72, Q5_K: 88, Q6_K: 105}     # relative cost of a tile = bytes per row-block / 2
BLOCK_BYTES = {Q4_K: 144, Q5_K: 176, Q6_K: 210}

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
    chunk = {Q4_K: meta[7] & 255, Q5_K: (meta[7] >> 8) & 255, Q6_K: (meta[7] >> 16) & 255}     # blocks per work item (a build knob)
    return list(first[:grid + 1]), dict(grid=grid, slot=meta[1], alive=meta[2], tiles=meta[3], warps=meta[4], rows_per_tile=meta[5], max_items=meta[6], chunk=chunk)


@pytest.mark.parametrize("name,K,segs", SHAPES, ids=[s[0] for s in SHAPES])
def test_cta_ranges_cover_all_tiles_in_order_and_are_byte_balanced(lib, name, K, segs):
    first, m = partition(lib, K, segs)
    rpt = m["rows_per_tile"]
    assert rpt == 16
    tiles = [(rows + rpt - 1) // rpt for _, rows in segs]
    assert m["tiles"] == sum(tiles)
    assert m["grid"] == 148
    assert first[0] == 0 and first[-1] == m["tiles"]
    assert all(a <= b for a, b in zip(first, first[1:])), "CTA ranges must be contiguous and ordered"
    # cumulative cost at every boundary is within one tile of the ideal split
    cost_of_tile = np.concatenate([np.full(n, COST[t]) for (t, _), n in zip(segs, tiles)])
    cum = np.concatenate([[0], np.cumsum(cost_of_tile)])
    total = cum[-1]
    for c, t in enumerate(first):
        assert abs(cum[t] - total * c / m["grid"]) <= max(COST.values()), (c, t)


@pytest.mark.parametrize("name,K,segs", SHAPES, ids=[s[0] for s in SHAPES])
def test_work_items_fit_a_ring_slot_and_match_the_device_enumeration(lib, name, K, segs):
    """A work item = one 16-row tile x CHUNK[type] blocks; it must fit one ring slot, and the largest CTA's item count the
    library reports must equal the count from walking the tiles here."""
    first, m = partition(lib, K, segs)
    nb = K // 256
    CHUNK = m["chunk"]
    assert all(1 <= c <= 4 for c in CHUNK.values())
    for t in (Q4_K, Q5_K, Q6_K):
        assert 16 * BLOCK_BYTES[t] * CHUNK[t] <= m["slot"]
        assert (16 * BLOCK_BYTES[t]) % 16 == 0      # cp.async.bulk: 16-byte granularity
    type_of_tile = []
    for t, rows in segs:
        type_of_tile += [t] * ((rows + 15) // 16)
    worst = 0
    for t0, t1 in zip(first, first[1:]):
        worst = max(worst, sum(-(-nb // CHUNK[type_of_tile[i]]) for i in range(t0, t1)))
    assert worst == m["max_items"]
    if name.startswith(("7b", "13b", "falcon")):
        # the fold chains of all tiles of a CTA are alive together: one mailbox each
        assert max(b - a for a, b in zip(first, first[1:])) <= 2 * m["alive"]


def test_rejects_what_the_kernel_cannot_run(lib):
    one = (C.c_int * 1)
    first, meta = (C.c_int * 150)(), (C.c_int * 8)()
    assert lib.ctb_matvec_partition(one(Q4_K), one(8), 1, 300, 148, first, meta) != 0     # K not a multiple of 256
    assert lib.ctb_matvec_partition(one(2), one(8), 1, 256, 148, first, meta) != 0        # Q4_0 is not a K-quant
    assert lib.ctb_matvec_partition(one(Q4_K), one(0), 1, 256, 148, first, meta) != 0
