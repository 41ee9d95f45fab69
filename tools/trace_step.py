#!/usr/bin/env python
"""Phase anatomy of one fused decode step (device %globaltimer stamps, ctb_llm_trace_step): per phase kind the median over
CTAs of {wait at the barrier, input staging, wait for the first weight item, item loop}, and the spread between CTAs.

    python tools/trace_step.py [out.json]
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from ctransformers_b200 import AutoModelForCausalLM  # noqa: E402

path = bench.ensure_model(0, 1, lambda: None)
llm = AutoModelForCausalLM.from_pretrained(str(path), context_length=bench.CTX)
ids = bench.prompt_ids()
llm.eval(ids, batch_size=256)
tok = llm.sample(top_k=1, repetition_penalty=1.0, seed=0)
cap = 8 * 200 * 160 + 1024
buf = (C.c_ulonglong * cap)()
n = llm.ctb_llm_trace_step(tok, bench.PROMPT, buf, cap)
assert n > 0, n
a = np.frombuffer(buf, dtype=np.uint64)
meta = a[:2 * n].reshape(n, 2).astype(int)
G = 148
st = a[2 * n:2 * n + n * G * 8].reshape(n, G, 8).astype(np.int64)
names = {0: "matvec", 1: "attn", 2: "embed", 3: "pick"}
mvk = ["qkv", "wo", "up", "down", "head"]
rows = []
t_prev_end = st[0, :, 0].min()
for i in range(n):
    kind = names[meta[i, 0]] + (":" + mvk[meta[i, 1]] if meta[i, 0] == 0 else "")
    t0, t1, t2, t3 = st[i, :, 0], st[i, :, 1], st[i, :, 2], st[i, :, 3]
    rows.append(dict(i=i, kind=kind, start_first=int(t0.min() - t_prev_end), start_spread=int(t0.max() - t0.min()),
                     stage=float(np.median(t1 - t0)) if meta[i, 0] == 0 else 0.0,
                     first_item=float(np.median(t2 - t1)) if meta[i, 0] == 0 else 0.0,
                     body=float(np.median(t3 - np.where(t2 > 0, t2, t0))), end_spread=int(t3.max() - t3.min()),
                     total=int(t3.max() - t0.min()),
                     # grid barrier in front of the phase: last CTA enters -> its arrive is out -> first / last CTA sees everybody -> fence done
                     bar_enter_spread=int(st[i, :, 4].max() - st[i, :, 4].min()) if i else 0,
                     bar_arrive=float(np.median(st[i, :, 5] - st[i, :, 4])) if i else 0.0,
                     bar_seen_after_last=int(st[i, :, 6].min() - st[i, :, 5].max()) if i else 0,
                     bar_seen_spread=int(st[i, :, 6].max() - st[i, :, 6].min()) if i else 0,
                     bar_fence=float(np.median(st[i, :, 7] - st[i, :, 6])) if i else 0.0,
                     bar_to_start=float(np.median(st[i, :, 0] - st[i, :, 7])) if i else 0.0))
    t_prev_end = t3.max()
agg = {}
for r in rows:
    agg.setdefault(r["kind"], []).append(r)
print(f"{'phase':14s} {'n':>3s} {'total':>8s} {'gap':>7s} {'spread0':>8s} {'stage':>7s} {'item0':>7s} {'body':>7s} {'spread1':>8s}   (ns, medians over the phases of a kind)")
summ = {}
for k, rs in agg.items():
    med = lambda f: float(np.median([r[f] for r in rs]))
    summ[k] = {f: med(f) for f in ("total", "start_first", "start_spread", "stage", "first_item", "body", "end_spread")}
    print(f"{k:14s} {len(rs):3d} {med('total'):8.0f} {med('start_first'):7.0f} {med('start_spread'):8.0f} {med('stage'):7.0f} {med('first_item'):7.0f} {med('body'):7.0f} {med('end_spread'):8.0f}")
print(f"{'barrier':14s} {'enter_spr':>9s} {'arrive':>7s} {'seen-last':>9s} {'seen_spr':>8s} {'fence':>6s} {'->start':>7s}")
for k, rs in agg.items():
    med = lambda f: float(np.median([r[f] for r in rs]))
    print(f"{k:14s} {med('bar_enter_spread'):9.0f} {med('bar_arrive'):7.0f} {med('bar_seen_after_last'):9.0f} {med('bar_seen_spread'):8.0f} {med('bar_fence'):6.0f} {med('bar_to_start'):7.0f}")
step = int(st[:, :, 3].max() - st[0, :, 0].min())
print("step span", step / 1e3, "us")
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps({"step_ns": step, "by_kind": summ, "phases": rows}, indent=1))
