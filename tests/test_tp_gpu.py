"""Tensor-sharded mode on real GPUs (needs >= 2; skipped otherwise): tools/tp_check.py under torchrun, 2 ranks.
Bit-exact against the oracle's tensor-parallel summation mode (see the tool's docstring)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("case", ["llama_wide_q4km", "llama_gqa_q5km"])
def test_two_rank_shards_are_bit_exact_against_the_tp_oracle(case, tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           str(ROOT / "tools" / "tp_check.py"), "--case", case, "--parity-steps", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = next((l for l in r.stdout.splitlines() if l.startswith("TPCHECK ")), None)
    assert line, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(line[len("TPCHECK "):])
    assert res["ok"] and res["parity"]["vs_tp_oracle_bit_exact"] and res["parity"]["ranks_agree_bitwise"], res
    assert r.returncode == 0
