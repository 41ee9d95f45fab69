#!/usr/bin/env python
"""Workload for ncu: bench.py's model, a short prompt, then a few single-token decode steps.

    ncu -k regex:'k_matvec|k_attn|k_embed' --launch-skip $((NP*161+1)) -c 162 ... python tools/prof_decode.py NP NSTEPS

(7B shape, per token: 1 k_embed + 32 x (4 k_matvec + 1 k_attn) = 161 launches, + 1 k_matvec for the logits of the last
prompt token and of every decode step.)  Never a source of bench numbers.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from ctransformers_b200 import AutoModelForCausalLM  # noqa: E402

n_prompt = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
path = bench.ensure_model(0, 1, lambda: None)
llm = AutoModelForCausalLM.from_pretrained(str(path), context_length=bench.CTX)
ids = bench.prompt_ids()[:n_prompt]
llm.eval(ids, batch_size=256)
tok = llm.sample(top_k=1, repetition_penalty=1.0, seed=0)
for _ in range(n_steps):
    llm.eval([tok])
    tok = llm.sample(top_k=1, repetition_penalty=1.0, seed=0)
print("ok", tok)
