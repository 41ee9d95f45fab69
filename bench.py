#!/usr/bin/env python
"""bench.py — decode tokens/s, Llama-2-7B-shaped Q4_K_M GGUF, batch 1 (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one single-token decode pass of the whole hot path (129 mat-vec launches over 225 weight tensors + 32 attention launches at a context of
256..511 tokens).  Workload: synthetic 7B-shaped model (random valid quant blocks in the reference's Q4_K_M tensor mix,
3.8 GB of weights ≫ the 126 MB L2, so every step streams its inputs from HBM — no L2 flush needed), 256-token prompt
prefilled untimed, then W warm-up + K timed decode steps.

  value     device-timed: K steps replayed as CUDA graphs with the token fed back on the device (k_argmax), CUDA events on the
            launching stream, max over ranks; tokens/s summed over ranks (replicas: one sequence per GPU, weak scaling)
  e2e       the same K steps through the public API (llm.eval([tok]) + llm.sample(top_k=1)): token H2D, logits D2H and the
            host sampler inside the timed region
  roofline  dominant kernel k_matvec (HBM-bound): the GGUF bytes of the weights its 129 launches of a step read (4005.4 MB) ÷ their
            duration, measured live by replaying exactly those launches as a CUDA graph between CUDA events on the engine's
            stream; peak = MEASURED_PEAKS.json hbm_gbs; traffic = ncu dram bytes per launch (profiles/k_matvec_traffic.json).
            roofline.step = the same for the whole step (weights + KV + logits bytes ÷ device-timed step)
  cpu_baseline / --impl reference: the UNMODIFIED reference (oracle/_ref/libctransformers_ref.so) on the host cores, same
            model file, same prompt, bounded sample, thread count swept (ggml's spin-wait pool collapses when oversubscribed).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CTX = 512
PROMPT = 256
MODEL_DIR = Path(os.environ.get("CTB_MODEL_DIR", "/tmp/ctb_models"))
REF_SO = ROOT / "oracle" / "_ref" / "libctransformers_ref.so"


# --workload: the default is the BASELINE.json headline (configs[1]); "falcon7b" is configs[3] (not a driver bench line)
WORKLOADS = {
    "llama2-7b": dict(file="llama2-7b-shaped.Q4_K_M.synthetic.gguf", arch="llama", shape="LLAMA2_7B", ftype="Q4_K_M", lo=259,
                      metric="decode tokens/s Llama-2-7B Q4_K_M b=1", dtype="int8 (q4_K/q6_K weights x q8_K activations, dp4a), fp32 combine",
                      name="Llama-2-7B-shaped Q4_K_M GGUF"),
    "falcon7b": dict(file="falcon-7b-shaped.Q5_K_M.synthetic.gguf", arch="falcon", shape="FALCON_7B_SHAPED", ftype="Q5_K_M", lo=0,
                     metric="decode tokens/s Falcon-7B Q5_K_M b=1", dtype="int8 (q5_K/q6_K/q8_0 weights x q8_K/q8_0 activations, dp4a), fp32 combine",
                     name="Falcon-7B-shaped (n_embd 4608, multi-query) Q5_K_M GGUF"),
}
WL = WORKLOADS["llama2-7b"]


def model_path():
    return MODEL_DIR / WL["file"]


def ensure_model(rank, world, barrier):
    from ctransformers_b200 import synth
    p = model_path()
    if rank == 0 and not p.exists():
        MODEL_DIR.mkdir(parents=True, exist_ok=True)
        tmp = p.with_suffix(".tmp")
        (synth.write_llama if WL["arch"] == "llama" else synth.write_falcon)(tmp, getattr(synth, WL["shape"]), WL["ftype"], seed=0)
        tmp.rename(p)
    barrier()
    return p


def prompt_ids():
    import numpy as np
    from ctransformers_b200 import synth
    ids = np.random.default_rng(1).integers(WL["lo"], getattr(synth, WL["shape"]).n_vocab, PROMPT).tolist()
    if WL["arch"] == "llama":
        ids[0] = 1
    return ids


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], threading.Event(), None

    def run(self):
        # one long-running nvidia-smi in loop mode (a sample every 50 ms) instead of one process per sample
        cmd = ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "50"]
        try:
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if line.strip():
                    self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass
        finally:
            try:
                self.proc.kill()
            except Exception:
                pass

    def summary(self):
        self.stop_flag.set()
        try:
            self.proc.kill()
        except Exception:
            pass
        self.join(timeout=6)
        if not self.rows:   # loop mode unavailable: one direct query as a last resort
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def hbm_peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kv_bytes_per_step(n_layer, n_embd_gqa, t_avg):
    return 2 * n_layer * n_embd_gqa * t_avg * 2 + 2 * n_layer * n_embd_gqa * 2


def pick_threads(llm, tok, cores):
    """ggml's spinning thread pool degrades badly when oversubscribed; sweep like BASELINE.md's plan and keep the fastest."""
    forced = os.environ.get("CTB_REF_THREADS")
    if forced:
        return max(1, min(cores, int(forced))), {}
    cands = sorted({c for c in (4, 8, 16, 24, 32, 48, 64, cores // 2, cores) if 1 <= c <= cores})
    timing = {}
    for c in cands:
        t0 = time.perf_counter()
        for _ in range(2):
            llm.eval([tok], threads=c)
        timing[c] = (time.perf_counter() - t0) / 2
        if timing[c] > 4 * min(timing.values()):
            break                      # far past the optimum, more threads only get slower
    return min(timing, key=timing.get), {str(k): round(v, 4) for k, v in timing.items()}


def run_reference(args, rank, world, barrier):
    """The reference's own CPU implementation on the host cores (rank 0 only)."""
    if rank != 0:
        return
    if not REF_SO.exists():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libctransformers_ref.so was not built (needs /root/reference at build time)"}))
        return
    from ctransformers_b200 import AutoModelForCausalLM
    cores = os.cpu_count() or 1
    p = ensure_model(0, 1, lambda: None)
    llm = AutoModelForCausalLM.from_pretrained(str(p), lib=str(REF_SO), context_length=CTX, threads=min(cores, 16))
    ids = prompt_ids()
    llm.eval(ids, batch_size=256, threads=min(cores, 32))           # untimed prefill, one chunk
    tok = llm.sample(top_k=1, repetition_penalty=1.0, seed=0)
    threads, sweep = pick_threads(llm, tok, cores)
    steps = max(1, min(args.steps, CTX - PROMPT - args.warmup - 2 * len(sweep) - 1))
    for _ in range(args.warmup):
        llm.eval([tok], threads=threads); tok = llm.sample(top_k=1, repetition_penalty=1.0, seed=0)
    t0 = time.perf_counter()
    for _ in range(steps):
        llm.eval([tok], threads=threads); tok = llm.sample(top_k=1, repetition_penalty=1.0, seed=0)
    dt = time.perf_counter() - t0
    v = steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic", "config": workload_config(1),
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": threads, "kind": "reference", "host_cores": cores, "thread_sweep_s_per_token": sweep,
                         "sample": f"{steps} decode steps at context {PROMPT}+ after a {PROMPT}-token prompt, llm.eval+llm.sample, unmodified reference CPU build (AVX2), {threads} threads (best of the sweep)"},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


METRIC = WL["metric"]
DTYPE = WL["dtype"]


def workload_config(n):
    return {"workload": WL["name"] + " (synthetic random quant blocks), batch=1 decode, ctx=512, 256-token prompt then decode",
            "global_batch": n, "ctx": CTX, "prompt": PROMPT, "parallelism": f"replicas x{n} (one sequence per GPU, no collective)",
            "l2": "inputs (GBs of weights per step) exceed the 126 MB L2; no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=192)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="llama2-7b", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    global WL, METRIC, DTYPE
    WL = WORKLOADS[args.workload]
    METRIC, DTYPE = WL["metric"], WL["dtype"]

    from ctransformers_b200 import replicas
    who = replicas.Rank.from_env()
    rank, world, local = who.rank, who.world, who.local
    if args.impl == "reference":
        if not replicas.reference_rank_runs(who):
            return                       # under torchrun only rank 0 times the CPU reference
        return run_reference(args, rank, world, lambda: None)

    import numpy as np
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    group = replicas.Group(who, backend="nccl", device=f"cuda:{local}")

    def barrier():
        group.barrier(torch.cuda.synchronize)

    max_over_ranks = group.max

    from ctransformers_b200 import AutoModelForCausalLM, synth
    path = ensure_model(rank, world, barrier)
    llm = AutoModelForCausalLM.from_pretrained(str(path), context_length=CTX)
    shape = getattr(synth, WL["shape"])
    ids = prompt_ids()
    steps = max(1, min(args.steps, CTX - PROMPT - args.warmup - 1))
    W = max(args.warmup, 3)

    def prefill():
        llm._context = []
        llm.eval(ids, batch_size=256)
        return llm.sample(top_k=1, repetition_penalty=1.0, seed=0)

    # ------------------------------------------------------------------ device-timed ("value")
    first = prefill()
    out = (C.c_int * (W + steps))()
    assert llm.ctb_llm_decode_greedy(first, PROMPT, W, out) >= 0                      # warm-up steps at n_past = 256..
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ms = llm.ctb_llm_decode_greedy(int(out[W - 1]), PROMPT + W, steps, out)           # K timed steps, CUDA events inside
    barrier()
    clocks = sampler.summary()
    assert ms > 0
    ms = max_over_ranks(ms)
    tokens_dev = list(out[:steps])
    replicas_agree = all(t == tokens_dev for t in group.gather_ints(tokens_dev))   # every replica decodes the same sequence
    value = replicas.aggregate_tokens_per_s(world, steps, ms)

    # ------------------------------------------------------------------ end to end through the public API ("e2e")
    tok = prefill()
    for _ in range(W):
        llm.eval([tok]); tok = llm.sample(top_k=1, repetition_penalty=1.0, seed=0)
    barrier()
    t0 = time.perf_counter()
    e2e_tokens = []
    for _ in range(steps):
        llm.eval([tok])                                   # H2D {token, n_past}; D2H logits + hidden state; stream sync
        tok = llm.sample(top_k=1, repetition_penalty=1.0, seed=0)
        e2e_tokens.append(tok)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = replicas.aggregate_tokens_per_s(world, steps, e2e_s * 1e3)

    # ------------------------------------------------------------------ roofline
    peak, peak_src = hbm_peak()
    wbytes = int(llm.ctb_llm_weight_bytes_per_token())
    t_avg = PROMPT + W + steps / 2
    gqa = shape.n_embd // shape.n_head * shape.n_head_kv
    step_bytes = wbytes + kv_bytes_per_step(shape.n_layer, gqa, t_avg) + shape.n_vocab * 4
    achieved = step_bytes / (ms / 1e3 / steps) / 1e9
    ms_kind = (C.c_double * 4)()
    cnt_kind = (C.c_int * 4)()
    prof_steps = 4
    for i in range(prof_steps):
        llm.ctb_llm_profile_step(tokens_dev[i], PROMPT + W + steps // 2 + i - prof_steps, ms_kind, cnt_kind)
    # dominant kernel = k_matvec: its launches of one step alone, replayed as a CUDA graph between two CUDA events on the
    # engine's stream (after everything else, so the KV cache it leaves behind does not matter)
    n_mv = C.c_long(0)
    mv_ms = llm.ctb_llm_time_matvec_only(32, C.byref(n_mv))
    mv_ms = max_over_ranks(mv_ms)
    n_mv = max(1, n_mv.value)
    mv_achieved = wbytes / (mv_ms / 1e3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "k_matvec_traffic.json"
    if tf.exists() and args.workload == "llama2-7b":
        traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch_avg")
    roofline = {
        "bound": "hbm", "kernel": "k_matvec", "achieved": mv_achieved, "peak": peak, "unit": "GB/s", "frac": mv_achieved / peak, "traffic": traffic,
        "peak_source": peak_src, "launches_per_step": n_mv, "algorithmic_bytes_per_launch": wbytes / n_mv, "avg_launch_us": 1e3 * mv_ms / n_mv,
        "how": "weight bytes of the step's k_matvec launches / their duration: the same launches (no attention, embedding, argmax) replayed as a CUDA graph, CUDA events on the launching stream, 32 replays",
        "step": {"achieved": achieved, "frac": achieved / peak, "bytes_per_step": step_bytes, "weight_bytes_per_step": wbytes,
                 "frac_vs_3.9GB_weights_only": (3.9e9 / (ms / 1e3 / steps) / 1e9) / peak,
                 "how": "weights + KV + logits bytes of a whole decode step / device-timed step (all kernels)"},
        "eager_ms_per_step_by_kind": {"matvec": ms_kind[0] / prof_steps, "attention": ms_kind[1] / prof_steps, "other": ms_kind[3] / prof_steps,
                                      "how": f"eager pass, CUDA event after every kernel, {prof_steps} steps (kernel share of the step)"},
    }

    result = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": steps, "warmup": W,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic", "config": workload_config(world),
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": 16, "d2h_bytes_per_step": shape.n_vocab * 4 + shape.n_embd * 4,
                "how": "llm.eval([tok]) + llm.sample(top_k=1) per step, wall clock between device syncs",
                "lookahead_hits": int(llm.ctb_llm_speculative_hits())},
        "gpu_launches": int(llm.ctb_llm_launches_per_token()) * steps,
        "roofline": roofline,
        "greedy_tokens_match_e2e": tokens_dev[:steps] == e2e_tokens[:steps], "replicas_agree": replicas_agree,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline and REF_SO.exists():
        del llm
        cores = os.cpu_count() or 1
        ref = AutoModelForCausalLM.from_pretrained(str(path), lib=str(REF_SO), context_length=CTX, threads=min(cores, 16))
        n_prompt, n_dec = 32, 12
        ref.eval(ids[:n_prompt], batch_size=32, threads=min(cores, 32))
        t = ref.sample(top_k=1, repetition_penalty=1.0, seed=0)
        threads, sweep = pick_threads(ref, t, cores)
        t0 = time.perf_counter()
        for _ in range(n_dec):
            t = ref.sample(top_k=1, repetition_penalty=1.0, seed=0)
            ref.eval([t], threads=threads)
        dt = time.perf_counter() - t0
        result["cpu_baseline"] = {"value": n_dec / dt, "unit": "tokens/s", "cores": threads, "kind": "reference", "host_cores": cores, "thread_sweep_s_per_token": sweep,
                                  "sample": f"{n_dec} decode steps after a {n_prompt}-token prompt (context ≈{n_prompt + 2 * len(sweep) + n_dec}), same model file, unmodified reference CPU build (AVX2), {threads} threads (best of the sweep)"}
    if rank == 0:
        print(json.dumps(result))
    group.close()


if __name__ == "__main__":
    main()
