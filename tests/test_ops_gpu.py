"""GPU parity tests, op level: every CUDA stage of the hot path, called through the C ABI with host buffers, against
the plain-C oracle (oracle/ggml_oracle.c) on the same seeded inputs.

Bar: bit-exact for integer/byte results (activation quantizers, embedding rows, RoPE, norms); fp32 dot products whose
integer parts are exact are compared to 2e-5 relative (summation order is the only difference)."""
import ctypes as C

import numpy as np
import pytest

import refs
from conftest import ptr
from refs import F16, F32, Q4_0, Q4_K, Q5_K, Q6_K, Q8_0, Q8_K, row_bytes

pytestmark = pytest.mark.gpu

DOT_RTOL = 2e-5   # |gpu - oracle| <= DOT_RTOL * (sum |terms| scale); see _dot_close


def _rand_weights(t, k, m, seed, sigma=0.02):
    from ctransformers_b200 import synth
    return np.ascontiguousarray(synth.random_blocks(t, k, m, sigma, np.random.default_rng(seed))).view(np.uint8)


def _act(rng, k, scale=1.0):
    x = rng.standard_normal(k).astype(np.float32) * scale
    x[rng.integers(0, k, 8)] *= 17.0
    return x


@pytest.mark.parametrize("k", [256, 4096, 11008])
def test_quantize_q8_K_bit_exact(lib, k):
    o = refs.oracle()
    rng = np.random.default_rng(k)
    for trial in range(6):
        x = _act(rng, k, [1e-4, 1.0, 300.0][trial % 3])
        if trial == 3:
            x[:256] = 0.0                      # all-zero block → d = 0
        if trial == 4:
            x[5], x[9] = -3.5, 3.5             # |max| tie: the FIRST one fixes the sign
            x[:256] = np.clip(x[:256], -3.5, 3.5)
        a = np.zeros(row_bytes(Q8_K, k), np.uint8)
        b = np.zeros_like(a)
        o.orc_quantize_row_q8_K(ptr(x), ptr(a), k)
        assert lib.ctb_quantize_row_q8_K(ptr(x), ptr(b), k) == 0
        assert np.array_equal(a, b), f"trial {trial}: {(a != b).sum()} differing bytes"


@pytest.mark.parametrize("k", [32, 4096, 4544])
def test_quantize_q8_0_bit_exact(lib, k):
    o = refs.oracle()
    rng = np.random.default_rng(k + 1)
    for trial in range(4):
        x = _act(rng, k, [1e-3, 1.0, 50.0, 1.0][trial])
        if trial == 3:
            x[:32] = 0.0
        a = np.zeros(row_bytes(Q8_0, k), np.uint8)
        b = np.zeros_like(a)
        o.orc_quantize_row_q8_0(ptr(x), ptr(a), k)
        assert lib.ctb_quantize_row_q8_0(ptr(x), ptr(b), k) == 0
        assert np.array_equal(a, b)


def _dot_close(got, want, scale):
    err = np.abs(got - want)
    assert np.all(err <= DOT_RTOL * scale + 1e-30), f"max err {err.max():.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("t", [Q4_K, Q5_K, Q6_K, Q4_0, Q8_0])
@pytest.mark.parametrize("k,m", [(256, 3), (4096, 64), (11008, 33)])
def test_mul_mat_vs_oracle(lib, t, k, m):
    o = refs.oracle()
    rng = np.random.default_rng(t * 1000 + k)
    w = _rand_weights(t, k, m, seed=t + k)
    n = 2
    x = np.stack([_act(rng, k), _act(rng, k, 0.05)])
    want = np.zeros((n, m), np.float32)
    got = np.zeros((n, m), np.float32)
    assert o.orc_mul_mat(t, ptr(w), ptr(x), ptr(want), k, m, n) == 0
    assert lib.ctb_mul_mat(t, ptr(w), ptr(x), ptr(got), k, m, n) == 0
    # scale of the accumulated magnitude: per-row sum of |terms| is bounded by ~ sqrt(k)*|w||x|; use max |want| + rms
    scale = np.abs(want).max() + np.sqrt((want ** 2).mean())
    _dot_close(got, want, scale)


@pytest.mark.parametrize("t", [F16, F32])
def test_mul_mat_float_weights(lib, t):
    k, m = 512, 40
    rng = np.random.default_rng(t)
    wf = (rng.standard_normal((m, k)) * 0.05).astype(np.float32)
    w = wf.astype(np.float16) if t == F16 else wf
    x = _act(rng, k)[None]
    got = np.zeros((1, m), np.float32)
    assert lib.ctb_mul_mat(t, ptr(w), ptr(x), ptr(got), k, m, 1) == 0
    xa = x.astype(np.float16).astype(np.float32) if t == F16 else x
    want = (w.astype(np.float32) @ xa[0].astype(np.float32))
    assert np.allclose(got[0], want, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("mode", [1, 2])
def test_norm_bit_exact(lib, mode):
    o = refs.oracle()
    n = 4096
    rng = np.random.default_rng(mode)
    x, w, b = _act(rng, n, 3.0), (1 + 0.1 * rng.standard_normal(n)).astype(np.float32), (0.1 * rng.standard_normal(n)).astype(np.float32)
    want, got = np.zeros(n, np.float32), np.zeros(n, np.float32)
    if mode == 1:
        o.orc_rms_norm_mul(ptr(x), ptr(w), ptr(want), n, 1e-5)
        assert lib.ctb_norm(1, ptr(x), ptr(w), None, ptr(got), n, 1e-5) == 0
    else:
        o.orc_layer_norm_mul_add(ptr(x), ptr(w), ptr(b), ptr(want), n, 1e-5)
        assert lib.ctb_norm(2, ptr(x), ptr(w), ptr(b), ptr(got), n, 1e-5) == 0
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), f"{(want != got).sum()} of {n} differ, max {np.abs(want - got).max()}"


@pytest.mark.parametrize("mode,hd", [(0, 128), (2, 64), (0, 64)])
def test_rope_bit_exact(lib, mode, hd):
    o = refs.oracle()
    rng = np.random.default_rng(hd + mode)
    for pos in (0, 1, 37, 511):
        x = rng.standard_normal((8, hd)).astype(np.float32)
        want, got = x.copy(), x.copy()
        o.orc_rope(ptr(want), 8, hd, pos, mode, 10000.0, 1.0)
        assert lib.ctb_rope(ptr(got), 8, hd, pos, mode, 10000.0, 1.0) == 0
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), f"pos {pos}: max diff {np.abs(want - got).max()}"


@pytest.mark.parametrize("n_head,n_kv,hd,T", [(4, 4, 128, 1), (4, 4, 128, 300), (8, 1, 64, 77), (8, 2, 128, 512)])
def test_attention_vs_oracle(lib, n_head, n_kv, hd, T):
    o = refs.oracle()
    rng = np.random.default_rng(T + hd)
    q = rng.standard_normal((n_head, hd)).astype(np.float32)
    kc = (rng.standard_normal((T, n_kv, hd)) * 0.7).astype(np.float16)
    vc = rng.standard_normal((n_kv, T, hd)).astype(np.float16)
    scale = 1.0 / np.sqrt(hd)
    got = np.zeros((n_head, hd), np.float32)
    assert lib.ctb_attention(ptr(q), ptr(kc), ptr(vc), ptr(got), n_head, n_kv, hd, T, scale) == 0
    want = np.zeros((n_head, hd), np.float32)
    for h in range(n_head):
        kvh = h // (n_head // n_kv)
        kslice = np.ascontiguousarray(kc[:, kvh, :])                 # [T][hd]
        vslice = np.ascontiguousarray(vc[kvh].T)                     # oracle wants [hd][T]
        o.orc_attn_head(ptr(q[h]), ptr(kslice), hd, ptr(vslice), T, hd, T, scale, ptr(want[h]))
    # fp16 rounding points are reproduced; only fp32 summation order differs.  A differently-ordered score sum can flip an
    # fp16 rounding of (s - max) or of p once in a while, which moves one term by 2^-11 relative.
    assert np.allclose(got, want, rtol=2e-3, atol=2e-4), np.abs(got - want).max()
    assert np.abs(got - want).max() <= 1e-3 * np.abs(want).max() + 1e-6


@pytest.mark.parametrize("t", [Q4_K, Q5_K, Q4_0])
def test_ffn_gate_vs_oracle(lib, t):
    o = refs.oracle()
    k, m = 4096, 96
    rng = np.random.default_rng(t)
    w1, w3 = _rand_weights(t, k, m, 1), _rand_weights(t, k, m, 2)
    x = _act(rng, k)
    g, u = np.zeros(m, np.float32), np.zeros(m, np.float32)
    o.orc_mul_mat(t, ptr(w1), ptr(x), ptr(g), k, m, 1)
    o.orc_mul_mat(t, ptr(w3), ptr(x), ptr(u), k, m, 1)
    s = np.zeros(m, np.float32)
    o.orc_silu(ptr(g), ptr(s), m)
    want = s * u
    got = np.zeros(m, np.float32)
    assert lib.ctb_ffn_gate(t, ptr(w1), ptr(w3), ptr(x), ptr(got), k, m) == 0
    assert np.allclose(got, want, rtol=2e-3, atol=1e-5 * np.abs(want).max())


@pytest.mark.parametrize("t", [Q4_K, Q5_K, Q6_K, Q4_0, Q8_0, F16, F32])
def test_get_row_bit_exact(lib, t):
    o = refs.oracle()
    k, rows = 512, 9
    if t == F32:
        tab = np.random.default_rng(1).standard_normal((rows, k)).astype(np.float32)
        want = tab
    elif t == F16:
        tab = np.random.default_rng(1).standard_normal((rows, k)).astype(np.float16)
        want = tab.astype(np.float32)
    else:
        tab = _rand_weights(t, k, rows, 5, sigma=1.0)
        want = np.zeros((rows, k), np.float32)
        getattr(o, "orc_dequantize_row_" + refs.TYPE_NAME[t])(ptr(tab), ptr(want), rows * k)
    for r in (0, 4, rows - 1):
        got = np.zeros(k, np.float32)
        assert lib.ctb_get_row(t, ptr(tab), k, rows, r, ptr(got)) == 0
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(want[r]).view(np.uint32))


@pytest.mark.skipif(not refs.have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("t", [Q4_K, Q5_K, Q6_K, Q4_0, Q8_0])
@pytest.mark.parametrize("k", [512, 1024, 2816])
def test_mul_mat_real_quantized_weights(lib, t, k):
    """Weights produced by the reference's quantizer (all scale/min bit patterns occur, unlike the random-block generator)."""
    o = refs.oracle()
    rng = np.random.default_rng(k + t)
    m = 48
    w = refs.ref_quantize(t, (rng.standard_normal((m, k)) * 0.05 + 0.01).astype(np.float32))
    x = _act(rng, k)[None]
    want, got = np.zeros((1, m), np.float32), np.zeros((1, m), np.float32)
    assert o.orc_mul_mat(t, ptr(w), ptr(x), ptr(want), k, m, 1) == 0
    assert lib.ctb_mul_mat(t, ptr(w), ptr(x), ptr(got), k, m, 1) == 0
    _dot_close(got, want, np.abs(want).max() + np.sqrt((want ** 2).mean()))
