"""Locate and bind libctransformers.so (the B200 build).

Mirror of the reference's FFI layer (ctransformers/lib.py:9-73 `find_library`, ctransformers/llm.py:117-208
`load_library`): same 17 prototypes, plus the additive ``ctb_*`` entry points of include/ctransformers_b200.h.
There is exactly one library flavour here — the in-tree sm_100a build — and no fallback: a missing library is
an error, never a silent switch to another code path.
"""
import ctypes as C
from pathlib import Path
from typing import Optional

LIB_DIR = Path(__file__).resolve().parent / "lib"
LIB_NAME = "libctransformers.so"


class ConfigStruct(C.Structure):
    """By-value argument of ctransformers_llm_create (reference: models/llm.h:6-11)."""
    _fields_ = [("context_length", C.c_int), ("gpu_layers", C.c_int), ("mmap", C.c_bool), ("mlock", C.c_bool)]


def find_library(path: Optional[str] = None) -> str:
    """An explicit path wins (like the reference, which returns unknown strings verbatim); otherwise the in-tree build."""
    if path:
        return str(path)
    lib = LIB_DIR / LIB_NAME
    if not lib.is_file():
        raise OSError(
            f"{lib} has not been built. Build it with `python -m ctransformers_b200.build` "
            "(needs nvcc; compiles for sm_100a). There is no CPU fallback library."
        )
    return str(lib)


_P = C.c_void_p
_IP = C.POINTER(C.c_int)
_FP = C.POINTER(C.c_float)

# name -> (restype, argtypes); part 1 = the reference FFI (models/llm.cc:32-138)
PROTOTYPES = {
    "ctransformers_llm_create": (_P, [C.c_char_p, C.c_char_p, ConfigStruct]),
    "ctransformers_llm_delete": (None, [_P]),
    "ctransformers_llm_tokenize": (C.c_int, [_P, C.c_char_p, C.c_bool, _IP]),
    "ctransformers_llm_detokenize": (C.c_char_p, [_P, C.c_int]),
    "ctransformers_llm_is_eos_token": (C.c_bool, [_P, C.c_int]),
    "ctransformers_llm_eos_token_id": (C.c_int, [_P]),
    "ctransformers_llm_bos_token_id": (C.c_int, [_P]),
    "ctransformers_llm_vocab_size": (C.c_int, [_P]),
    "ctransformers_llm_context_length": (C.c_int, [_P]),
    "ctransformers_llm_architecture": (C.c_char_p, [_P]),
    "ctransformers_llm_batch_eval": (C.c_bool, [_P, _IP, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ctransformers_llm_logits_data": (_FP, [_P]),
    "ctransformers_llm_logits_size": (C.c_int, [_P]),
    "ctransformers_llm_embeddings_data": (_FP, [_P]),
    "ctransformers_llm_embeddings_size": (C.c_int, [_P]),
    "ctransformers_llm_sample": (C.c_int, [_P, _IP, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]),
    "ctransformers_llm_reset": (None, [_P]),
}
# part 2 = additive entry points
EXTRA_PROTOTYPES = {
    "ctb_abi_version": (C.c_int, []),
    "ctb_llm_last_eval_ms": (C.c_double, [_P]),
    "ctb_llm_launches_per_token": (C.c_long, [_P]),
    "ctb_llm_speculative_hits": (C.c_long, [_P]),
    "ctb_llm_weight_bytes_per_token": (C.c_ulonglong, [_P]),
    "ctb_llm_trace_step": (C.c_long, [_P, C.c_int, C.c_int, C.POINTER(C.c_ulonglong), C.c_long]),
    "ctb_llm_load_ms": (C.c_double, [_P]),
    "ctb_llm_device_samples": (C.c_long, [_P]),
    "ctb_llm_set_stream": (None, [_P, C.c_void_p]),
    "ctb_tp_unique_id": (C.c_int, [_P, C.c_int]),
    "ctb_llm_create_tp": (_P, [C.c_char_p, C.c_char_p, ConfigStruct, C.c_int, C.c_int, _P]),
    "ctb_tp_shard": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _IP]),
    "ctb_llm_decode_greedy": (C.c_double, [_P, C.c_int, C.c_int, C.c_int, _IP]),
    "ctb_llm_profile_step": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double), _IP]),
    "ctb_llm_time_matvec_only": (C.c_double, [_P, C.c_int, C.POINTER(C.c_long)]),
    "ctb_llm_time_matvec_kinds": (C.c_double, [_P, C.c_int, C.POINTER(C.c_long), C.c_uint]),
    "ctb_mul_mat": (C.c_int, [C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int]),
    "ctb_quantize_row_q8_K": (C.c_int, [_P, _P, C.c_int]),
    "ctb_quantize_row_q8_0": (C.c_int, [_P, _P, C.c_int]),
    "ctb_norm": (C.c_int, [C.c_int, _P, _P, _P, _P, C.c_int, C.c_float]),
    "ctb_rope": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]),
    "ctb_attention": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]),
    "ctb_ffn_gate": (C.c_int, [C.c_int, _P, _P, _P, _P, C.c_int, C.c_int]),
    "ctb_matvec_partition": (C.c_int, [_IP, _IP, C.c_int, C.c_int, C.c_int, _IP, _IP]),
    "ctb_get_row": (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_int, _P]),
    "ctb_vocab_load": (_P, [C.c_char_p]),
    "ctb_vocab_free": (None, [_P]),
    "ctb_vocab_size": (C.c_int, [_P]),
    "ctb_vocab_tokenize": (C.c_int, [_P, C.c_char_p, C.c_bool, _IP, C.c_int]),
    "ctb_vocab_piece": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int]),
    "ctb_sample": (C.c_int, [_FP, C.c_int, _IP, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]),
}


def load_library(path: Optional[str] = None):
    lib = C.CDLL(find_library(path))
    for table, required in ((PROTOTYPES, True), (EXTRA_PROTOTYPES, False)):
        for name, (res, args) in table.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                if required:
                    raise OSError(f"{path or LIB_NAME} does not export '{name}'")
                continue  # a reference-built library has no ctb_* symbols; that is fine when passed via lib=
            fn.restype, fn.argtypes = res, args
    return lib
