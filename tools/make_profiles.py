#!/usr/bin/env python
"""Turn a round's gpurun_out/ captures into the tracked summaries under profiles/.

    python tools/make_profiles.py r02 gpurun_out/r02_full.ncu-rep

Inputs (written on the GPU box by the commands quoted in profiles/<tag>_README.md):
  <report>.ncu-rep      ncu --set full --import-source on of k_pstep / k_step launches of tools/prof_decode.py
Outputs: profiles/<tag>_ncu_metrics.csv (one row per captured launch), profiles/<tag>_ncu_stalls.md (stall mix, opcode mix and the
hottest SASS lines per kernel), profiles/<tag>_sass_evidence.txt (the TMA / mbarrier / tensor-core instructions of the step kernel),
profiles/k_step_traffic.json (DRAM bytes of the decode launch: bench.py's roofline.traffic).
"""
import collections
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
rep = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "gpurun_out" / f"{tag}_full.ncu-rep"
out = ROOT / "profiles"
out.mkdir(exist_ok=True)


def ncu(*args):
    return subprocess.run(["ncu", "-i", str(rep), *args], capture_output=True, text=True, check=True).stdout


# ---- raw metrics, one row per launch
rows = list(csv.reader(ncu("--page", "raw", "--csv").splitlines()))
H, U = rows[0], rows[1]
KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max"] + [h for h in H if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
idx = [H.index(k) for k in KEEP if k in H]
with open(out / f"{tag}_ncu_metrics.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([H[i] for i in idx])
    w.writerow([U[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])

# the decode launch = the longest k_step launch captured
ki, ti = H.index("Kernel Name"), H.index("gpu__time_duration.sum")
step_rows = [r for r in rows[2:] if "k_step" in r[ki]]
if step_rows:
    dec = max(step_rows, key=lambda r: float(r[ti]))
    rd, wr = float(dec[H.index("dram__bytes_read.sum")]), float(dec[H.index("dram__bytes_write.sum")])
    unit_r, unit_w = U[H.index("dram__bytes_read.sum")], U[H.index("dram__bytes_write.sum")]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = rd * scale[unit_r] + wr * scale[unit_w]
    (out / "k_step_traffic.json").write_text(json.dumps({
        "source": f"profiles/{tag}_ncu_metrics.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one whole decode launch of k_step: "
                  "129 mat-vec phases + 32 attention phases + embedding + pick at a context of ~45)",
        "dram_bytes_per_step": total, "matvec_phases": 129, "dram_bytes_per_matvec_phase_avg": total / 129,
        "algorithmic_weight_bytes_per_step": 4005470208, "ratio": total / 4005470208}, indent=1) + "\n")

# ---- per-kernel stall / opcode summaries from the source page
txt = ncu("--page", "source", "--csv", "--print-source", "sass")
blocks = re.split(r'(?m)^"Kernel Name",', txt)[1:]
md = [f"# {tag}: warp-stall sampling and instruction mix per captured launch (ncu --set full --import-source on; tools/prof_decode.py 40 3)", ""]
evidence = []
best_i = 0.0
seen = collections.Counter()
done = set()
for blk in blocks:
    lines = blk.splitlines()
    name = lines[0].strip('",')
    rws = list(csv.reader(lines[1:]))
    hdr, data = rws[0], [r for r in rws[1:] if len(r) == len(rws[0])]
    ix = {h: i for i, h in enumerate(hdr)}
    num = lambda x: float(x) if x not in ("", None) else 0.0
    tot_i = sum(num(r[ix["Instructions Executed"]]) for r in data)
    tot_s = sum(num(r[ix["# Samples"]]) for r in data)
    if (name, tot_i, tot_s) in done:   # the source page lists every launch once per view
        continue
    done.add((name, tot_i, tot_s))
    seen[name] += 1
    md += [f"## {name}  (capture {seen[name]}: {tot_i:.0f} warp instructions, {tot_s:.0f} samples, {len(data)} SASS lines)", ""]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    mix = sorted(((sum(num(r[ix[s]]) for r in data), s) for s in stalls), reverse=True)[:8]
    md += ["stall mix: " + ", ".join(f"{s[6:]} {100 * v / max(tot_s, 1):.1f}%" for v, s in mix), ""]
    byop = collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        byop[(m.group(2).split(".")[0] if m else "?")] += num(r[ix["Instructions Executed"]])
    md += ["opcode mix (executed warp instructions): " + ", ".join(f"{op} {100 * c / max(tot_i, 1):.1f}%" for op, c in byop.most_common(14)), ""]
    md += ["| samples | executed | top stalls | instruction |", "|---|---|---|---|"]
    for r in sorted(data, key=lambda r: -num(r[ix["# Samples"]]))[:14]:
        st = sorted(((num(r[ix[s]]), s[6:]) for s in stalls if num(r[ix[s]]) > 0), reverse=True)[:2]
        md.append(f"| {num(r[ix['# Samples']]):.0f} | {num(r[ix['Instructions Executed']]):.0f} | {', '.join(f'{n} {v:.0f}' for v, n in st)} | `{r[ix['Source']].strip()[:70]}` |")
    md.append("")
    if "k_step" in name and tot_i > best_i:   # the decode launch = the k_step capture with the most instructions
        best_i = tot_i
        evidence = []
        for r in data:
            s_ = r[ix["Source"]]
            if re.search(r"UBLKCP|SYNCS|IMMA|UTMA|RED\.|LDGSTS|BAR\.SYNC", s_):
                evidence.append(f"{r[ix['Address']][-6:]}  executed {num(r[ix['Instructions Executed']]):>10.0f}  {s_.strip()}")
(out / f"{tag}_ncu_stalls.md").write_text("\n".join(md) + "\n")
(out / f"{tag}_sass_evidence.txt").write_text(
    "SASS of ctb::k_step (sm_100a) as profiled: the TMA bulk copies (UBLKCP = cp.async.bulk), mbarrier operations (SYNCS), the legacy-path int8\n"
    "tensor-core instructions (IMMA.16832.U8.S8 = mma.sync.m16n8k32), the cp.async descriptor prefetch (LDGSTS), named barriers and the grid-barrier RED.\n\n"
    + "\n".join(evidence) + "\n")
print("wrote", [p.name for p in out.glob(f"{tag}_*")], "k_step_traffic.json")
