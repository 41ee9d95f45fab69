#!/usr/bin/env python
"""BASELINE.json configs[0]: GPT-2 117M-shaped Q4_0 GGML file, ctx 128, 96-token prompt + 32 new tokens, on the reference's CPU
build (oracle/_ref) driven through this repository's Python surface.  Plumbing check with a number attached; no GPU involved.

    python tools/config1_gpt2.py [threads] > profiles/<round>_config1_gpt2_reference.json
"""
import json
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import refs  # noqa: E402
from ctransformers_b200 import AutoModelForCausalLM, synth  # noqa: E402

threads = int(sys.argv[1]) if len(sys.argv) > 1 else min(os.cpu_count() or 1, 8)
with tempfile.TemporaryDirectory() as tmp:
    shape = synth.GPT2Shape(n_ctx=128)          # 50257 / 768 / 12 heads / 12 layers, context 128
    path = synth.write_gpt2_ggml(Path(tmp) / "gpt2-117m-shaped.q4_0.bin", shape, "Q4_0", seed=0)
    llm = AutoModelForCausalLM.from_pretrained(str(path), model_type="gpt2", lib=str(refs.REF_SO), threads=threads)
    ids = np.random.default_rng(1).integers(0, shape.n_vocab, 96).tolist()
    t0 = time.perf_counter()
    llm.eval(ids, batch_size=8, threads=threads)
    t_prompt = time.perf_counter() - t0
    toks = []
    t0 = time.perf_counter()
    for _ in range(32):
        t = llm.sample(top_k=1, repetition_penalty=1.0, seed=0)
        toks.append(int(t))
        if len(toks) < 32:
            llm.eval([t], threads=threads)
    t_gen = time.perf_counter() - t0
print(json.dumps({"config": "GPT-2 117M-shaped Q4_0 GGML, ctx=128, 96-token prompt, 32 new tokens, greedy", "impl": "reference (oracle/_ref, CPU)",
                  "threads": threads, "host_cpus": os.cpu_count(), "prompt_tokens_per_s": 96 / t_prompt, "decode_tokens_per_s": 32 / t_gen,
                  "tokens": toks[:8]}))
