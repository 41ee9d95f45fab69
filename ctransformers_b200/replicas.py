"""Multi-GPU plumbing for the replica mode (one process per GPU, one full model and one sequence each, no data-path collective).

Batch-1 decode is a single dependent chain, so N GPUs run N independent sequences; the only communication
is the rendezvous, a barrier around the timed region and a MAX reduction of the per-rank device time.  `torch.distributed`
does that: backend "nccl" on GPUs, "gloo" in the CPU tests.
"""
import contextlib
import os
import sys
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence


@dataclass
class Rank:
    rank: int
    world: int
    local: int

    @classmethod
    def from_env(cls) -> "Rank":
        return cls(int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0)))


@contextlib.contextmanager
def _stdout_to_stderr():
    """NCCL prints its version banner on file descriptor 1 when the first communicator is created; bench.py's stdout must
    carry exactly one JSON line, so the banner is sent to stderr."""
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


class Group:
    """Barrier / max-over-ranks / gather on top of torch.distributed; a no-op group when world == 1."""

    def __init__(self, who: Rank, backend: Optional[str] = None, device: Optional[str] = None):
        self.who, self.dist, self.device = who, None, device or "cpu"
        if who.world > 1:
            import torch
            import torch.distributed as dist
            backend = backend or ("nccl" if self.device.startswith("cuda") else "gloo")
            kw = {"device_id": torch.device(self.device)} if backend == "nccl" else {}
            with _stdout_to_stderr():
                if not dist.is_initialized():
                    dist.init_process_group(backend, rank=who.rank, world_size=who.world, **kw)
                self.dist = dist
                self.barrier()                 # creates the communicator now, while stdout is diverted

    def barrier(self, sync: Optional[Callable[[], None]] = None) -> None:
        if self.dist:
            self.dist.barrier()
        if sync:
            sync()

    def max(self, x: float) -> float:
        if not self.dist:
            return float(x)
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_ints(self, xs: Sequence[int]) -> List[List[int]]:
        """every rank's list on every rank (used to check that replicas produced the same greedy tokens)"""
        if not self.dist:
            return [list(xs)]
        import torch
        mine = torch.tensor(list(xs), dtype=torch.int64, device=self.device)
        outs = [torch.empty_like(mine) for _ in range(self.who.world)]
        self.dist.all_gather(outs, mine)
        return [o.tolist() for o in outs]

    def close(self) -> None:
        if self.dist and self.dist.is_initialized():
            self.dist.destroy_process_group()
            self.dist = None


def aggregate_tokens_per_s(world: int, steps: int, ms_max: float) -> float:
    """whole-job throughput of `world` replicas that each decoded `steps` tokens; the slowest rank's device time counts"""
    return world * steps / (ms_max / 1e3)


def reference_rank_runs(who: Rank) -> bool:
    """--impl reference: rank 0 alone times the CPU reference; the other ranks exit without work"""
    return who.rank == 0
