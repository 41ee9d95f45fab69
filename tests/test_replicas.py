"""world_size-2 gloo test (CPU) of the replica plumbing bench.py uses under torchrun."""
import os
import socket
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import json, os, sys
sys.path.insert(0, os.environ["CTB_ROOT"])
from ctransformers_b200 import replicas
who = replicas.Rank.from_env()
g = replicas.Group(who, backend="gloo", device="cpu")
g.barrier()
ms_local = 10.0 + 5.0 * who.rank                      # rank 1 is the slow one
ms = g.max(ms_local)
toks = g.gather_ints([7, 8, 9 + 0 * who.rank])
g.barrier()
out = {"rank": who.rank, "world": who.world, "ms": ms, "value": replicas.aggregate_tokens_per_s(who.world, 30, ms),
       "same_tokens": all(t == toks[0] for t in toks), "ref_runs": replicas.reference_rank_runs(who)}
open(os.path.join(os.environ["CTB_OUT"], f"rank{who.rank}.json"), "w").write(json.dumps(out))
g.close()
"""


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_barrier_max_and_aggregate(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, CTB_ROOT=str(ROOT), CTB_OUT=str(tmp_path))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    res = [json.loads((tmp_path / f"rank{k}.json").read_text()) for k in range(2)]   # one file per rank: stdout of the ranks may interleave
    assert [d["rank"] for d in res] == [0, 1] and all(d["world"] == 2 for d in res)
    for d in res:
        assert d["ms"] == 15.0                         # max over ranks, on every rank
        assert abs(d["value"] - 2 * 30 / 0.015) < 1e-6  # whole-job tokens/s = replicas x steps / slowest rank
        assert d["same_tokens"]
    assert [d["ref_runs"] for d in res] == [True, False]


def test_single_rank_group_is_a_noop():
    from ctransformers_b200 import replicas
    g = replicas.Group(replicas.Rank(0, 1, 0))
    g.barrier()
    assert g.max(3.5) == 3.5 and g.gather_ints([1, 2]) == [[1, 2]]
    assert replicas.aggregate_tokens_per_s(1, 64, 128.0) == 500.0
