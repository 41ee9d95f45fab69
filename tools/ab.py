#!/usr/bin/env python
"""A/B timing of library builds on bench.py's workload: python tools/ab.py name=path.so ...  (device-timed greedy decode)."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from ctransformers_b200 import AutoModelForCausalLM  # noqa: E402

path = bench.ensure_model(0, 1, lambda: None)
ids = bench.prompt_ids()
for spec in sys.argv[1:]:
    name, lib = spec.split("=", 1)
    llm = AutoModelForCausalLM.from_pretrained(str(path), lib=lib, context_length=bench.CTX)
    llm.eval(ids, batch_size=256)
    first = llm.sample(top_k=1, repetition_penalty=1.0, seed=0)
    out = (C.c_int * 256)()
    llm.ctb_llm_decode_greedy(first, bench.PROMPT, 8, out)
    best = 1e9
    for rep in range(3):
        ms = llm.ctb_llm_decode_greedy(int(out[7]), bench.PROMPT + 8, 64, out)
        best = min(best, ms / 64)
    n = C.c_long(0)
    mv = llm.ctb_llm_time_matvec_only(32, C.byref(n))
    kinds = []
    if hasattr(llm, "ctb_llm_time_matvec_kinds"):
        for k, nm in enumerate(["qkv", "wo", "up", "down", "out"]):
            ms_k = llm.ctb_llm_time_matvec_kinds(16, C.byref(n), 1 << k)
            kinds.append(f"{nm} {1e3 * ms_k / max(1, n.value):.2f}us x{n.value}")
    print("   per launch, same-kind launches back to back:", ", ".join(kinds))
    print(f"{name}: {1e3 / best:.1f} tok/s  step {best:.4f} ms  matvec-only {mv:.4f} ms/step  first tokens {list(out[:4])}", flush=True)
    del llm
