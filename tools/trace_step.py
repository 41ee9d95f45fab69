#!/usr/bin/env python
"""In-graph anatomy of the k_matvec launches of one decode step (device %globaltimer stamps, see ctb_llm_trace_step).

    python tools/trace_step.py [out.json]

Prints, per projection kind, the medians over the step's launches of: gap from the previous launch's last warp end to this
launch's dependency release, staging time (dependency released -> input quantized), streaming time (staged -> last warp end),
spread between the first and the last warp to finish.
"""
import ctypes as C
import json
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from ctransformers_b200 import AutoModelForCausalLM  # noqa: E402

path = bench.ensure_model(0, 1, lambda: None)
import os
llm = AutoModelForCausalLM.from_pretrained(str(path), context_length=bench.CTX, **({'lib': os.environ['CTB_LIB']} if os.environ.get('CTB_LIB') else {}))
ids = bench.prompt_ids()
llm.eval(ids, batch_size=256)
tok = llm.sample(top_k=1, repetition_penalty=1.0, seed=0)
W = 16
per = 148 * (4 + W) + 2
cap = 200 * per
buf = (C.c_ulonglong * cap)()
n = llm.ctb_llm_trace_step(tok, bench.PROMPT, buf, cap)
assert n > 0, n
names = ["qkv", "wo", "up", "down", "out"]
rows = []
prev_end = None
for i in range(n):
    o = i * per
    kind, ncta = int(buf[o]), int(buf[o + 1])
    ent, rel, stg, ends, sta = [], [], [], [], []
    for c in range(ncta):
        b = o + 2 + c * (4 + W)
        if buf[b] == 0:
            continue
        ent.append(buf[b]); rel.append(buf[b + 1]); stg.append(buf[b + 2]); sta.append(buf[b + 3])
        ends += [buf[b + 4 + w] for w in range(W) if buf[b + 4 + w]]
    r = {"kind": names[kind], "ctas": len(ent), "first_entry": min(ent), "release": statistics.median(rel), "release_max": max(rel),
         "staged": statistics.median(stg), "staged_max": max(stg), "first_end": min(ends), "median_end": statistics.median(ends), "last_end": max(ends)}
    r["gap_prev_end_to_release_us"] = None if prev_end is None else (r["release"] - prev_end) / 1e3
    r["staging_us"] = (r["staged"] - r["release"]) / 1e3
    r["stats_us"] = (statistics.median(sta) - r["release"]) / 1e3   # dependency released -> norm statistics known
    r["stream_us"] = (r["last_end"] - r["staged"]) / 1e3
    r["end_spread_us"] = (r["last_end"] - r["first_end"]) / 1e3
    r["median_end_to_last_us"] = (r["last_end"] - r["median_end"]) / 1e3
    r["early_entry_us"] = (r["release"] - r["first_entry"]) / 1e3
    prev_end = r["last_end"]
    if 8 <= i < 13:          # raw stamps of one layer's launches (+ the next QKV), relative to the launch's median release
        base = r["release"]
        r["raw"] = [[int(buf[o + 2 + c * (4 + W) + j]) - int(base) if buf[o + 2 + c * (4 + W) + j] else None for j in range(4 + W)] for c in range(ncta)]
    rows.append(r)
print(f"{n} k_matvec launches traced; step span {(rows[-1]['last_end'] - rows[0]['first_entry']) / 1e3:.1f} us")
print(f"{'kind':6} {'n':>3} {'gap prev end->release':>22} {'staging':>9} {'(stats)':>8} {'stream':>9} {'end spread':>11} {'median->last end':>17} {'entered early by':>17}")
for k in names:
    rs = [r for r in rows if r["kind"] == k and r["gap_prev_end_to_release_us"] is not None]
    if not rs:
        continue
    med = lambda f: statistics.median(r[f] for r in rs)
    print(f"{k:6} {len(rs):3d} {med('gap_prev_end_to_release_us'):22.2f} {med('staging_us'):9.2f} {med('stats_us'):8.2f} {med('stream_us'):9.2f} {med('end_spread_us'):11.2f} {med('median_end_to_last_us'):17.2f} {med('early_entry_us'):17.2f}")
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(rows))
