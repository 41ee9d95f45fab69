"""world_size-2 gloo test (CPU) of the tensor-sharded mode's host plumbing (ctransformers_b200/tp.py + the C ABI around it):
rank 0 draws the 128-byte communicator id, torch.distributed carries it, both ranks compute their shard ranges, and — there
being no GPU here — ctb_llm_create_tp refuses loudly on both (no CPU fallback), which the Python mirror turns into the
reference's RuntimeError.  The GPU side of the mode is tests/test_tp_gpu.py / tools/tp_check.py."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

from test_replicas import _free_port

ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.environ["CTB_ROOT"])
sys.path.insert(0, os.path.join(os.environ["CTB_ROOT"], "tests"))
import torch.distributed as dist
from ctransformers_b200 import LLM, Config, tp_plan
from ctransformers_b200.lib import load_library
from ctransformers_b200.tp import tensor_parallel_ticket
import modelcases
dist.init_process_group("gloo")
rank, world, uid = tensor_parallel_ticket()
lib = load_library()
out6 = (C.c_int * 6)()
shape = modelcases.CASES["llama_wide_q4km"][1]
assert lib.ctb_tp_shard(shape.n_embd, shape.n_head, shape.n_head_kv, shape.n_ff, rank, world, out6) == 0
plan = tp_plan.plan(shape.n_embd, shape.n_head, shape.n_head_kv, shape.n_ff, shape.n_vocab, world)[rank]
path, ctx = modelcases.build("llama_wide_q4km", os.environ["CTB_OUT"]) if rank == 0 else (None, None)
dist.barrier()
path, ctx = modelcases.build("llama_wide_q4km", os.environ["CTB_OUT"])
try:
    LLM(str(path), config=Config(context_length=ctx), tp=(rank, world, uid))
    created = True
except RuntimeError as e:
    created = str(e)
res = {"rank": rank, "world": world, "uid": uid.hex(), "shard": list(out6), "plan": [*plan.heads, *plan.kv_heads, *plan.ff], "created": created}
open(os.path.join(os.environ["CTB_OUT"], f"tp_rank{rank}.json"), "w").write(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
"""


def test_two_ranks_share_one_id_and_split_the_model(tmp_path):
    import torch
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, CTB_ROOT=str(ROOT), CTB_OUT=str(tmp_path))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = [json.loads((tmp_path / f"tp_rank{k}.json").read_text()) for k in range(2)]
    assert [d["rank"] for d in res] == [0, 1] and all(d["world"] == 2 for d in res)
    assert res[0]["uid"] == res[1]["uid"] and len(res[0]["uid"]) == 256 and set(res[0]["uid"]) != {"0"}   # one non-trivial 128-byte id
    for d in res:
        assert d["shard"] == d["plan"]                     # native shard arithmetic == tp_plan on every rank
    assert res[0]["shard"][1] == res[1]["shard"][0] and res[0]["shard"][5] == res[1]["shard"][4]           # heads and n_ff ranges are adjacent
    if not torch.cuda.is_available():
        for d in res:
            assert isinstance(d["created"], str) and "Failed to create LLM" in d["created"]                # no GPU: refused, not emulated
        assert "no CUDA device available" in r.stderr
