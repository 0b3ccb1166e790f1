"""Frame sharding and the one collective of the hot path.

The reference expresses multi-GPU Stage 3 as DDP + DistributedSampler over frame pairs
(lab4d/dataloader/data_utils.py:56-61, lab4d/engine/train_utils.py:15-28, lab4d/train.py:28-36).  Frames are
independent units: every rank holds the canonical surfels, renders its own frames (warp -> rasterize -> loss ->
backward) and the canonical-surfel gradients are summed across ranks once per optimizer step.

    shard_frames()   which global frame indices this rank renders (rank r takes r, r+G, ...)
    FlatGrads        one flat fp32 buffer that *is* the .grad of every surfel parameter, so the exchange is a
                     single NCCL all-reduce over NVLink (gloo on CPU in the tests), no per-tensor buckets.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend: str | None = None, device: torch.device | None = None) -> tuple[int, int]:
    """Initialise torch.distributed from the torchrun environment (env://).  No-op for world size 1."""
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        if backend == "nccl":
            # the collective's CTAs must be resident on every rank before it can progress: a high-priority communication
            # stream lets them in ahead of the queued compute CTAs when an all-reduce overlaps the next step's kernels
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
            kw["pg_options"] = opts
        dist.init_process_group(backend=backend, init_method="env://", rank=rank, world_size=world, **kw)
    return rank, world


def shard_frames(num_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin frame assignment: rank r renders frames r, r+world, ...."""
    return list(range(rank, num_frames, world))


class FlatGrads:
    """Gradients of a list of parameters laid out back to back in ONE fp32 buffer; each p.grad is a view."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = list(params)
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros((n,), dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero_(self):
        self.flat.zero_()

    def allreduce_(self, average_over: int | None = None):
        """Sum over ranks (one collective); optionally divide by the global number of frames."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        if average_over:
            self.flat.mul_(1.0 / float(average_over))
        return self.flat

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4
