"""Host plumbing of the tensor-sharded mode (BASELINE.json configs[4]): one process per GPU, launched with
torchrun; `torch.distributed` only carries the 128-byte NCCL id from rank 0 to the others — the per-layer exchanges are NCCL
all-reduces issued by the native engine on its own stream (csrc/engine.cu: Engine::tp_all_reduce).

    import torch.distributed as dist
    dist.init_process_group("gloo")            # or "nccl"
    llm = LLM(path, tp=tensor_parallel_ticket())

Every rank must then make the same calls in the same order (same prompt, same seed): all ranks hold the same logits.
"""
import ctypes as C

from .lib import load_library


def tensor_parallel_ticket(lib=None):
    """(rank, world, unique_id_bytes) for LLM(..., tp=...), agreed through the default torch.distributed process group."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0 and world > 1:
        raw = C.create_string_buffer(128)
        n = load_library(lib).ctb_tp_unique_id(raw, 128)
        if n != 128:
            raise RuntimeError("ctb_tp_unique_id failed (is libnccl.so.2 loadable?)")
        buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
    if world > 1:
        on_gpu = dist.get_backend() == "nccl"
        t = buf.cuda() if on_gpu else buf
        dist.broadcast(t, 0)
        buf = t.cpu()
    return rank, world, bytes(buf.numpy().tobytes())
