#!/usr/bin/env python
"""Turn one round's gpurun_out/<tag>/ captures into the tracked summaries under profiles/.

    python tools/make_profiles.py r01

Inputs (written on the GPU box by the commands quoted in profiles/<tag>_README.md):
  bench_n1.json, bench_reference.json   bench.py lines (not under a profiler)
  launches_step.csv                     ncu --metrics gpu__time_duration.sum,dram__bytes_* ... of one whole decode step
  prof_full.ncu-rep                     ncu --set full --import-source on of the first kernels of a decode step
"""
import csv
import json
import shutil
import subprocess
import sys
from collections import OrderedDict, defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = ROOT / "gpurun_out" / tag
out = ROOT / "profiles"
out.mkdir(exist_ok=True)

for name in ("bench_n1.json", "bench_reference.json", "smi.csv", "host.txt"):
    if (src / name).exists():
        shutil.copy(src / name, out / f"{tag}_{name}")

# ---- launch list of one decode step
rows = list(csv.reader(open(src / "launches_step.csv")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[h]
ki, gi, bi = H.index("Kernel Name"), H.index("Grid Size"), H.index("Block Size")
launches = OrderedDict()
for r in rows[h + 1:]:
    if len(r) < len(H) or not r[0].isdigit():
        continue
    d = launches.setdefault(int(r[0]), {"kernel": r[ki], "grid": r[gi], "block": r[bi]})
    d[r[-3]] = float(r[-1].replace(",", ""))
short = lambda k: k.split("(")[0].replace("void ", "").replace("ctb::", "")
with open(out / f"{tag}_launches_step.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "time_us", "dram_read_MB", "dram_write_MB", "warp_inst", "issue_active_pct"])
    for i, d in launches.items():
        w.writerow([i, short(d["kernel"]), d["grid"], round(d.get("gpu__time_duration.sum", 0) / 1e3, 3), round(d.get("dram__bytes_read.sum", 0) / 1e6, 3),
                    round(d.get("dram__bytes_write.sum", 0) / 1e6, 3), int(d.get("smsp__inst_executed.sum", 0)), d.get("smsp__issue_active.avg.pct_of_peak_sustained_active")])
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in launches.values():
    k = "k_matvec" if "k_matvec" in d["kernel"] else short(d["kernel"])
    a = agg[k]
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0) / 1e3; a[2] += d.get("dram__bytes_read.sum", 0); a[3] += d.get("dram__bytes_write.sum", 0)
total_us = sum(a[1] for a in agg.values())
mv = agg["k_matvec"]
traffic = {"source": f"profiles/{tag}_launches_step.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum, one decode step, per launch)",
           "launches": mv[0], "dram_bytes_per_step": mv[2] + mv[3], "dram_bytes_per_launch_avg": (mv[2] + mv[3]) / max(1, mv[0])}
(out / "k_matvec_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")

bench = json.loads((src / "bench_n1.json").read_text()) if (src / "bench_n1.json").exists() else {}
ref = json.loads((src / "bench_reference.json").read_text()) if (src / "bench_reference.json").exists() else {}
md = [f"# {tag}: one decode step of the bench workload under ncu (cold caches, serialised launches, no PDL overlap)", "",
      "| kernel | launches | sum of launch times (us) | share | dram read (MB) | dram write (MB) |", "|---|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    md.append(f"| {k} | {a[0]} | {a[1]:.1f} | {100 * a[1] / total_us:.1f} % | {a[2] / 1e6:.1f} | {a[3] / 1e6:.2f} |")
md += ["", f"Sum over the step: {total_us:.0f} us under ncu.  These per-launch times are cold-cache and serialised; only the SHARE is comparable with bench.py."]
if bench:
    r = bench["roofline"]
    e = r["eager_ms_per_step_by_kind"]
    tot = e["matvec"] + e["attention"] + e["other"]
    md += ["", "## bench.py (same build, not under a profiler)", "",
           f"* value {bench['value']:.1f} tokens/s ({bench['ms_per_step']:.4f} ms/step, device-timed graph replays), e2e {bench['e2e']['value']:.1f} tokens/s, clocks {bench['clocks']}",
           f"* k_matvec roofline: {r['achieved']:.0f} GB/s of {r['peak']:.0f} GB/s = {100 * r['frac']:.1f} %  ({r['launches_per_step']} launches, {r['algorithmic_bytes_per_launch'] / 1e6:.2f} MB and {r['avg_launch_us']:.2f} us per launch on average; ncu traffic {traffic['dram_bytes_per_launch_avg'] / 1e6:.2f} MB per launch)",
           f"* whole step: {r['step']['achieved']:.0f} GB/s = {100 * r['step']['frac']:.1f} % of peak",
           f"* kernel share of the step, eager pass with an event after every kernel: matvec {100 * e['matvec'] / tot:.1f} %, attention {100 * e['attention'] / tot:.1f} %, other {100 * e['other'] / tot:.1f} %  (ncu share above: matvec {100 * mv[1] / total_us:.1f} %)"]
    if "cpu_baseline" in bench:
        md.append(f"* cpu_baseline: {bench['cpu_baseline']['value']:.2f} tokens/s — {bench['cpu_baseline']['sample']}")
if ref:
    md.append(f"* --impl reference: {ref['value']:.2f} tokens/s — {ref['cpu_baseline']['sample']}; thread sweep (s/token): {ref['cpu_baseline']['thread_sweep_s_per_token']}")
(out / f"{tag}_step_summary.md").write_text("\n".join(md) + "\n")

# ---- full capture: key metrics per captured kernel + top stall lines of the source page
rep = src / "prof_full.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    Hh = rr[0]
    keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
    keep += [c for c in Hh if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")]
    idx = [Hh.index(c) for c in keep if c in Hh]
    with open(out / f"{tag}_ncu_full_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([Hh[i] for i in idx]); w.writerow([rr[1][i] for i in idx])
        for r in rr[2:]:
            w.writerow([r[i] for i in idx])
    srcp = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    sr = list(csv.reader(srcp.splitlines()))
    starts = [i for i, r in enumerate(sr) if r and r[0] == "Kernel Name"] + [len(sr)]
    lines = [f"# {tag}: warp-stall sampling per kernel (ncu --set full --import-source on), top instructions by samples", ""]
    seen = set()
    for a, b in zip(starts[:-1], starts[1:]):
        name = sr[a][1]
        Hs = sr[a + 1]; ci = {c: i for i, c in enumerate(Hs)}
        data = sr[a + 2:b]
        tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
        inst = sum(int(r[ci["Instructions Executed"]] or 0) for r in data)
        key = (name, tot, inst)
        if key in seen:
            continue
        seen.add(key)
        kinds = [c for c in Hs if c.startswith("stall_") and "Not Issued" not in c]
        ag = {k: sum(int(r[ci[k]] or 0) for r in data) for k in kinds}
        lines += [f"## {name}", f"samples {tot}, warp instructions {inst}, SASS lines {len(data)}", "",
                  "stall mix: " + ", ".join(f"{k[6:]} {100 * v / max(1, tot):.1f}%" for k, v in sorted(ag.items(), key=lambda kv: -kv[1]) if v > 0.02 * tot), "",
                  "| sass line | samples | long_sb | short_sb | wait | no_inst | instruction |", "|---|---|---|---|---|---|---|"]
        top = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]] or 0))[:12]
        for i in sorted(top):
            r = data[i]
            lines.append(f"| {i} | {r[ci['# Samples']]} | {r[ci['stall_long_sb']]} | {r[ci['stall_short_sb']]} | {r[ci['stall_wait']]} | {r[ci['stall_no_inst']]} | `{r[ci['Source']].strip()[:80]}` |")
        lines.append("")
    (out / f"{tag}_ncu_stalls.md").write_text("\n".join(lines) + "\n")
print("profiles written:", sorted(p.name for p in out.iterdir()))
