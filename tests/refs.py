"""ctypes loaders for the two CHECKERS (test infrastructure, never the product):

  * ``oracle()``  — oracle/_build/liboracle.so, our plain-C restatement (oracle/ggml_oracle.c)
  * ``ref()``     — oracle/_ref/libctransformers_ref.so, the unmodified reference compiled from
                    /root/reference by oracle/Makefile (present when it was built in the container;
                    it travels to the GPU box with the snapshot)
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "_build" / "liboracle.so"
REF_SO = ROOT / "oracle" / "_ref" / "libctransformers_ref.so"

# ggml type ids (ggml.h enum ggml_type)
F32, F16, Q4_0, Q8_0, Q4_K, Q5_K, Q6_K, Q8_K = 0, 1, 2, 8, 12, 13, 14, 15
BLOCK = {Q4_0: (32, 18), Q8_0: (32, 34), Q4_K: (256, 144), Q5_K: (256, 176), Q6_K: (256, 210), Q8_K: (256, 292),
         F32: (1, 4), F16: (1, 2)}
TYPE_NAME = {Q4_0: "q4_0", Q8_0: "q8_0", Q4_K: "q4_K", Q5_K: "q5_K", Q6_K: "q6_K"}


def row_bytes(t, k):
    bs, sz = BLOCK[t]
    assert k % bs == 0
    return k // bs * sz


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        if not ORACLE_SO.exists():
            subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "oracle"])
        o = C.CDLL(str(ORACLE_SO))
        fp, vp, i = C.POINTER(C.c_float), C.c_void_p, C.c_int
        o.orc_fp16_to_fp32.restype = C.c_float
        o.orc_fp16_to_fp32.argtypes = [C.c_uint16]
        o.orc_fp32_to_fp16.restype = C.c_uint16
        o.orc_fp32_to_fp16.argtypes = [C.c_float]
        for n in ("orc_vec_dot_q4_0_q8_0", "orc_vec_dot_q8_0_q8_0", "orc_vec_dot_q4_K_q8_K", "orc_vec_dot_q5_K_q8_K",
                  "orc_vec_dot_q6_K_q8_K"):
            getattr(o, n).restype = C.c_float
            getattr(o, n).argtypes = [i, vp, vp]
        o.orc_mul_mat.restype = i
        o.orc_mul_mat.argtypes = [i, vp, vp, vp, i, i, i]
        o.orc_rms_norm_mul.argtypes = [vp, vp, vp, i, C.c_float]
        o.orc_layer_norm_mul_add.argtypes = [vp, vp, vp, vp, i, C.c_float]
        o.orc_rope.argtypes = [vp, i, i, i, i, C.c_float, C.c_float]
        o.orc_rope_table.argtypes = [vp, i, i, C.c_float, C.c_float]
        o.orc_attn_head.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, i, i, C.c_float, vp]
        _oracle = o
    return _oracle


class _Traits(C.Structure):
    _fields_ = [("type_name", C.c_char_p), ("blck_size", C.c_int), ("type_size", C.c_size_t), ("is_quantized", C.c_bool),
                ("to_float", C.c_void_p), ("from_float", C.c_void_p), ("from_float_reference", C.c_void_p),
                ("vec_dot", C.c_void_p), ("vec_dot_type", C.c_int)]


def have_ref():
    return REF_SO.exists()


def ref():
    """The compiled reference (AVX2 build).  Also initialises ggml's fp16 tables once."""
    global _ref
    if _ref is None:
        r = C.CDLL(str(REF_SO))
        r.ggml_internal_get_type_traits.restype = _Traits
        r.ggml_internal_get_type_traits.argtypes = [C.c_int]

        class _IP(C.Structure):
            _fields_ = [("mem_size", C.c_size_t), ("mem_buffer", C.c_void_p), ("no_alloc", C.c_bool)]
        r.ggml_init.restype = C.c_void_p
        r.ggml_init.argtypes = [_IP]
        r.ggml_free.argtypes = [C.c_void_p]
        r.ggml_free(r.ggml_init(_IP(1 << 20, None, False)))  # builds table_silu_f16 & co (ggml.c:4319-4333)
        r.ggml_quantize_chunk.restype = C.c_size_t
        r.ggml_quantize_chunk.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _ref = r
    return _ref


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def ref_traits(t):
    tr = ref().ggml_internal_get_type_traits(t)
    to_float = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int)(tr.to_float) if tr.to_float else None
    from_float = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int)(tr.from_float) if tr.from_float else None
    vec_dot = C.CFUNCTYPE(None, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)(tr.vec_dot) if tr.vec_dot else None
    return dict(to_float=to_float, from_float=from_float, vec_dot=vec_dot, vec_dot_type=tr.vec_dot_type)


def ref_quantize(t, x):
    """Quantize f32 rows [M,K] with the reference's own quantizer (ggml.c:19319 ggml_quantize_chunk)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    m, k = x.shape
    out = np.zeros(m * row_bytes(t, k), dtype=np.uint8)
    hist = np.zeros(16, dtype=np.int64)
    n = ref().ggml_quantize_chunk(t, ptr(x), ptr(out), 0, m * k, ptr(hist))
    assert n == out.size, (n, out.size)
    return out


def ref_quantize_act(t, x):
    """Quantize an activation row with the reference's from_float for type t (Q8_K / Q8_0)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(row_bytes(t, x.size), dtype=np.uint8)
    ref_traits(t)["from_float"](ptr(x), ptr(out), x.size)
    return out


def ref_vec_dot(t, k, wrow, act):
    s = np.zeros(1, dtype=np.float32)
    ref_traits(t)["vec_dot"](k, ptr(s), ptr(wrow), ptr(act))
    return float(s[0])
