"""CPU checks of the drop-in boundary: the library loads and exports every symbol include/*.h declares, and the
product path refuses to run (loudly) without a GPU instead of falling back to anything."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "ctransformers_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:ctransformers_llm|ctb)_[a-z0-9_A-Z]+)\s*\(", text)))


def test_header_declares_the_17_reference_functions():
    ref = [s for s in declared_symbols() if s.startswith("ctransformers_llm_")]
    assert len(ref) == 17, ref


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.ctb_abi_version() == 1


def test_python_prototypes_cover_the_header():
    from ctransformers_b200.lib import EXTRA_PROTOTYPES, PROTOTYPES
    assert sorted(list(PROTOTYPES) + list(EXTRA_PROTOTYPES)) == declared_symbols()


def test_no_cpu_fallback_without_gpu(lib, tmp_models):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ctransformers_b200 import AutoModelForCausalLM, synth
    path = tmp_models / "nogpu.gguf"
    synth.write_llama(path, synth.LlamaShape(n_vocab=512, n_embd=256, n_head=4, n_head_kv=4, n_ff=512, n_layer=1), "Q4_K_M")
    with pytest.raises(RuntimeError):
        AutoModelForCausalLM.from_pretrained(str(path))
    import numpy as np
    x = np.zeros(256, np.float32)
    y = np.zeros(292, np.uint8)
    assert lib.ctb_quantize_row_q8_K(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), 256) != 0


def test_unsupported_model_type_returns_null(lib, tmp_models):
    from ctransformers_b200.lib import ConfigStruct
    p = tmp_models / "not_gguf.bin"
    p.write_bytes(b"ggml" + b"\0" * 64)
    assert lib.ctransformers_llm_create(str(p).encode(), b"gpt2", ConfigStruct(-1, 0, True, False)) is None


def test_missing_library_is_an_error(tmp_path, monkeypatch):
    import ctransformers_b200.lib as L
    monkeypatch.setattr(L, "LIB_DIR", tmp_path)
    with pytest.raises(OSError):
        L.find_library()
