"""Data-parallel plumbing for the snapshot-sharded encoder (reference: ``train.py:31-48, 309-310, 353-354``).

The path shards naturally: snapshots are independent given the shared graph and the 1,056 parameters, so each rank
processes its own slice of the batch with NO data-path collective; the only exchange is the mean of the parameter
gradients after backward (what DDP does for the reference).  The module also works unchanged inside
``torch.nn.parallel.DistributedDataParallel``; :class:`FlatGradAllReduce` is the lean equivalent used by the benchmark:
all gradients live in ONE flat buffer, so the exchange is a single 4.2 KB ``all_reduce`` (NCCL over NVLink on GPUs, gloo
in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """``torchrun`` bootstrap (train.py:31-46): returns ``(rank, world_size, local_rank)``; a no-op single-process
    set-up when the launcher variables are absent."""
    if "RANK" not in os.environ or "WORLD_SIZE" not in os.environ:
        return 0, 1, 0
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        kwargs = {"device_id": torch.device("cuda", local_rank)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local_rank


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice ``[start, stop)`` of ``total`` samples for ``rank`` (the first ``total % world``
    ranks take one extra)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class FlatGradAllReduce:
    """Keeps the gradients of ``params`` as views into one flat fp32 buffer and averages it across ranks with a
    single collective."""

    def __init__(self, params: Iterable[torch.nn.Parameter], module: torch.nn.Module | None = None):
        """``module``: when given, every ``GATv2Conv`` inside it is switched to ``fused_grad_accumulation`` -- its backward then
        adds the parameter gradients straight onto the flat buffer's views (no autograd accumulation kernels)."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        if module is not None:
            from .gatv2 import GATv2Conv

            for m in module.modules():
                if isinstance(m, GATv2Conv):
                    m.fused_grad_accumulation = True

    def zero_(self):
        self.flat.zero_()
        off = 0
        for p in self.params:  # re-attach views in case an optimizer dropped them (set_to_none)
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def all_reduce_mean(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            if dist.get_backend() == "nccl":  # the mean is taken inside the collective: one launch
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
                self.flat.div_(dist.get_world_size())
        return self.flat
