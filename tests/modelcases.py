"""Shared definitions of the small synthetic models the end-to-end tests (and the golden generator) use."""
from pathlib import Path

import numpy as np

from ctransformers_b200 import synth

CASES = {
    # name: (arch, shape, ftype, ctx)
    "llama_tiny_q4km": ("llama", synth.LlamaShape(n_vocab=1024, n_embd=256, n_head=4, n_head_kv=4, n_ff=768, n_layer=3, n_ctx_train=256), "Q4_K_M", 96),
    "llama_gqa_q5km": ("llama", synth.LlamaShape(n_vocab=2048, n_embd=1024, n_head=8, n_head_kv=2, n_ff=2816, n_layer=2, n_ctx_train=256), "Q5_K_M", 96),
    "llama_tiny_q4_0": ("llama", synth.LlamaShape(n_vocab=1024, n_embd=256, n_head=4, n_head_kv=4, n_ff=768, n_layer=2, n_ctx_train=256), "Q4_0", 64),
    "llama_tiny_q8_0": ("llama", synth.LlamaShape(n_vocab=1024, n_embd=256, n_head=2, n_head_kv=2, n_ff=512, n_layer=2, n_ctx_train=256), "Q8_0", 64),
    "falcon_tiny_q5km": ("falcon", synth.FalconShape(n_vocab=1024, n_embd=512, n_head=8, n_head_kv=1, n_ff=2048, n_layer=2, n_ctx_train=256), "Q5_K_M", 96),
    # wide enough that every mat-vec launch has many row tiles per CTA and QKV mixes Q4_K with Q6_K (layer 0 "use_more_bits")
    "llama_wide_q4km": ("llama", synth.LlamaShape(n_vocab=1536, n_embd=2048, n_head=16, n_head_kv=16, n_ff=5632, n_layer=2, n_ctx_train=256), "Q4_K_M", 96),
    "llama_tiny_q5_0": ("llama", synth.LlamaShape(n_vocab=1024, n_embd=256, n_head=4, n_head_kv=4, n_ff=768, n_layer=2, n_ctx_train=256), "Q5_0", 64),
    "falcon_tiny_q4_0": ("falcon", synth.FalconShape(n_vocab=1024, n_embd=256, n_head=4, n_head_kv=1, n_ff=1024, n_layer=2, n_ctx_train=256), "Q4_0", 64),
}
PROMPT_LEN = 37   # with 24 new tokens the context reaches 61: the f16 dot of V·P then uses its SIMD lanes AND its scalar tail
N_NEW = 24


def build(name, directory, quantizer=None):
    arch, shape, ftype, ctx = CASES[name]
    path = Path(directory) / f"{name}.gguf"
    if not path.exists():
        (synth.write_llama if arch == "llama" else synth.write_falcon)(path, shape, ftype, seed=11, quantizer=quantizer)
    return path, ctx


def prompt_for(name):
    arch, shape, _, _ = CASES[name]
    rng = np.random.default_rng(5)
    lo = 259 if arch == "llama" else 0
    ids = rng.integers(lo, shape.n_vocab, PROMPT_LEN).tolist()
    if arch == "llama":
        ids[0] = 1
    return ids


def run_greedy(llm, prompt, n_new, batch_size=8):
    """prompt eval (chunked like the reference default) then n_new greedy steps; returns logits after the prompt,
    embeddings after the prompt, the greedy tokens and the logits after the last step."""
    llm.eval(prompt, batch_size=batch_size)
    first_logits = np.array(llm.logits, dtype=np.float32)
    first_embd = np.array(llm.embeddings, dtype=np.float32)
    toks, gaps = [], []
    for _ in range(n_new):
        lg = np.array(llm.logits, dtype=np.float32)
        top2 = np.sort(lg)[-2:]
        gaps.append(float(top2[1] - top2[0]))
        t = llm.sample(top_k=1, repetition_penalty=1.0, seed=0)
        toks.append(int(t))
        llm.eval([t])
    return first_logits, first_embd, toks, np.array(llm.logits, dtype=np.float32), gaps
