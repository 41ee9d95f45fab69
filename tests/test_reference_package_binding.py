"""The drop-in claim of INTEGRATION.md §1, tested from the reference's side: the UNMODIFIED reference Python package
(/root/reference/ctransformers, imported in place — present in the build container only, so this test is skipped on the GPU box)
binds this repo's shared library through its own `load_library(lib=...)` (reference ctransformers/llm.py:117-210: every
ctransformers_llm_* symbol gets argtypes/restype there, a missing export raises AttributeError) and drives `LLM(...)` into
`ctransformers_llm_create`.  Without a GPU the create call must fail loudly (no CPU fallback) with the reference's own error
(llm.py:254-257 "Failed to create LLM"), not crash; with a GPU (a container that has both the reference tree and a GPU) it generates four greedy tokens."""
import subprocess
import sys
from pathlib import Path

import pytest

from ctransformers_b200.lib import find_library

REF = Path("/root/reference")

SCRIPT = r"""
import sys
sys.path.insert(0, "/root/reference")
import ctransformers                                     # the reference package, unmodified
from ctransformers.llm import load_library, LLM
assert ctransformers.__file__.startswith("/root/reference/"), ctransformers.__file__
so, model = sys.argv[1], sys.argv[2]
lib = load_library(so)                                   # binds all 17 entry points or raises AttributeError
for name in ("create", "delete", "tokenize", "detokenize", "is_eos_token", "eos_token_id", "bos_token_id", "vocab_size", "context_length",
             "architecture", "batch_eval", "logits_data", "logits_size", "embeddings_data", "embeddings_size", "sample", "reset"):
    assert getattr(lib, "ctransformers_llm_" + name).restype is not object
try:
    llm = LLM(model_path=model, model_type="gguf", lib=so)
except RuntimeError as e:
    print("CREATE_FAILED:", e)
else:
    toks = llm.tokenize("hello world")
    out = []
    for t in llm.generate(toks, top_k=1, batch_size=8):
        out.append(t)
        if len(out) == 4:
            break
    print("CREATED", llm.vocab_size, llm.context_length, out)
"""


@pytest.mark.skipif(not (REF / "ctransformers" / "llm.py").exists(), reason="/root/reference is only present in the build container")
def test_unmodified_reference_package_binds_this_library(tmp_models):
    import modelcases
    path, _ = modelcases.build("llama_tiny_q4km", tmp_models)
    so = find_library(None)
    r = subprocess.run([sys.executable, "-c", SCRIPT, so, str(path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import torch
    if torch.cuda.is_available():
        assert "CREATED 1024" in r.stdout, r.stdout + r.stderr
    else:
        # no GPU here: the library refuses loudly and the reference package reports it the way it reports any failed load
        assert "CREATE_FAILED: Failed to create LLM" in r.stdout, r.stdout + r.stderr
        assert "no CUDA device available" in r.stderr
