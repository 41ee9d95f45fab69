"""Constraints of the tensor-shard plan: block-aligned K-splits, matching row/K splits, full coverage."""
import pytest

from ctransformers_b200 import tp_plan

L7 = dict(n_embd=4096, n_head=32, n_head_kv=32, n_ff=11008, n_vocab=32000)
L13 = dict(n_embd=5120, n_head=40, n_head_kv=40, n_ff=13824, n_vocab=32000)
L70ISH = dict(n_embd=8192, n_head=64, n_head_kv=8, n_ff=28672, n_vocab=32000)
F7 = dict(n_embd=4608, n_head=72, n_head_kv=1, n_ff=18432, n_vocab=65024)


@pytest.mark.parametrize("shape", [L7, L13, L70ISH, F7], ids=["7b", "13b", "gqa", "falcon-mqa"])
@pytest.mark.parametrize("tp", [1, 2, 4, 8])
def test_plan_is_block_aligned_and_complete(shape, tp):
    sh = tp_plan.plan(tp=tp, **shape)
    assert len(sh) == tp
    hd = shape["n_embd"] // shape["n_head"]
    for a, b in zip(sh, sh[1:]):
        assert a.heads[1] == b.heads[0] and a.attn_k[1] == b.attn_k[0] and a.ff[1] == b.ff[0] and a.vocab[1] == b.vocab[0]
    assert sh[0].heads[0] == 0 and sh[-1].heads[1] == shape["n_head"]
    assert sh[0].ff[0] == 0 and sh[-1].ff[1] == shape["n_ff"]
    assert sh[0].vocab[0] == 0 and sh[-1].vocab[1] == shape["n_vocab"]
    for s in sh:
        # every K-split of wo / w2 is a whole number of Q8_K activation blocks, and equals the producer's row split
        assert s.attn_k[0] % 256 == 0 and s.attn_k[1] % 256 == 0
        assert s.ff[0] % 256 == 0 and s.ff[1] % 256 == 0
        assert s.attn_k == (s.heads[0] * hd, s.heads[1] * hd)
        per_kv = shape["n_head"] // shape["n_head_kv"]
        if s.heads[1] > s.heads[0]:
            assert s.kv_heads == (s.heads[0] // per_kv, (s.heads[1] - 1) // per_kv + 1)
    assert tp_plan.imbalance(sh) <= 1.16       # uneven block counts: 13B at TP 8 is the worst case of the bench shapes


def test_the_survey_s_13b_example():
    """L13's 40 heads are 20 pairs: TP 2/4 -> 20/10 heads per rank, TP 8 -> (6,6,6,6,4,4,4,4)."""
    assert [s.heads[1] - s.heads[0] for s in tp_plan.plan(tp=2, **L13)] == [20, 20]
    assert [s.heads[1] - s.heads[0] for s in tp_plan.plan(tp=4, **L13)] == [10, 10, 10, 10]
    assert [s.heads[1] - s.heads[0] for s in tp_plan.plan(tp=8, **L13)] == [6, 6, 6, 6, 4, 4, 4, 4]
    # w2: K = 13824 = 54 blocks: TP 2 even (27), TP 4 -> 14,14,13,13
    assert [(s.ff[1] - s.ff[0]) // 256 for s in tp_plan.plan(tp=4, **L13)] == [14, 14, 13, 13]


def test_rejects_shapes_that_cannot_be_block_aligned():
    with pytest.raises(ValueError):
        tp_plan.plan(n_embd=4544, n_head=71, n_head_kv=1, n_ff=18176, n_vocab=65024, tp=2)   # true Falcon-7B: 71 heads x 64


@pytest.mark.parametrize("shape", [L7, L13, L70ISH, F7], ids=["7b", "13b", "gqa", "falcon-mqa"])
@pytest.mark.parametrize("tp", [1, 2, 4, 8])
def test_native_shard_arithmetic_equals_the_plan(lib, shape, tp):
    """csrc/engine.cu tp_shard (what the engine slices the weights by) == tp_plan.plan, through ctb_tp_shard (no GPU needed)."""
    import ctypes as C
    sh = tp_plan.plan(tp=tp, **shape)
    for r, s in enumerate(sh):
        out = (C.c_int * 6)()
        assert lib.ctb_tp_shard(shape["n_embd"], shape["n_head"], shape["n_head_kv"], shape["n_ff"], r, tp, out) == 0
        assert tuple(out) == (*s.heads, *s.kv_heads, *s.ff)


def test_native_shard_rejects_shapes_that_do_not_tile(lib):
    import ctypes as C
    out = (C.c_int * 6)()
    assert lib.ctb_tp_shard(4096, 32, 32, 11000, 0, 2, out) != 0     # n_ff not a multiple of 256
    assert lib.ctb_tp_shard(4096, 32, 32, 11008, 2, 2, out) != 0     # rank out of range
    assert lib.ctb_tp_shard(4544, 71, 1, 18176, 0, 2, out) != 0      # true Falcon-7B: head_dim 64 x 71 heads is not whole blocks


def test_oracle_tp_mode_only_reorders_the_row_parallel_sums(tmp_models):
    """oracle/llama_oracle.c tp_world: world 1 is the reference order; world 2 re-associates the fp32 sums of wo / w2 and
    nothing else, so one layer in, the logits are equal to within fp32 rounding noise (and are NOT required to be identical)."""
    import numpy as np
    import modelcases
    import refs
    path, ctx = modelcases.build("llama_gqa_q5km", tmp_models)
    prompt = modelcases.prompt_for("llama_gqa_q5km")[:6]
    a = refs.OracleModel(str(path), ctx)
    b = refs.OracleModel(str(path), ctx)
    b.set_tp(2)
    la, lb = a.eval(prompt).copy(), b.eval(prompt).copy()
    rng = float(la.max() - la.min())
    assert np.abs(la - lb).max() / rng < 5e-2
    c = refs.OracleModel(str(path), ctx)
    c.set_tp(1)
    assert np.array_equal(c.eval(prompt), la)
