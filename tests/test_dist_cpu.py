"""CPU tests of the multi-rank plumbing (gloo, world_size 2): the flat-buffer gradient all-reduce and the batch sharding
reproduce the unsharded gradients (oracle arithmetic stands in for the kernels: no GPU here)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gatv2_oracle as G
from helpers import random_graph


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from tec_mollm_b200 import dist as tdist

    r, w, _ = tdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    B, L, N, F, H, C = 4, 3, 12, 5, 2, 3
    ei = random_graph(N, 40, seed=1)
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(B, L, N, F, generator=gen, dtype=torch.float64)
    gy = torch.randn(B, L, N, H * C, generator=gen, dtype=torch.float64)
    p = G.init_params(F, C, H, seed=3, dtype=torch.float64)
    params = [torch.nn.Parameter(v.clone().float()) for v in p.values()]
    flat = tdist.FlatGradAllReduce(params)
    assert flat.flat.numel() == sum(v.numel() for v in p.values())
    lo, hi = tdist.shard_range(B, rank, world)
    # snapshot index s = l*B + b (tec_mollm.py:84): shard the batch dim, keep all L time steps
    xs = x[lo:hi].permute(1, 0, 2, 3).reshape(-1, N, F)
    gs = gy[lo:hi].permute(1, 0, 2, 3).reshape(-1, N, H * C)
    _, grads = G.fwd_bwd(xs, ei, p, H, C, gs)
    flat.zero_()
    for prm, k in zip(params, p.keys()):
        prm.grad.add_(grads[k].float() * world)   # mean over ranks of (world * local sum) == global sum
    flat.all_reduce_mean()
    if rank == 0:
        xf = x.permute(1, 0, 2, 3).reshape(-1, N, F)
        gf = gy.permute(1, 0, 2, 3).reshape(-1, N, H * C)
        _, full = G.fwd_bwd(xf, ei, p, H, C, gf)
        err = max(((prm.grad.double() - full[k]).abs().max() / full[k].abs().max().clamp_min(1e-30)).item()
                  for prm, k in zip(params, p.keys()))
        ret["err"] = err
        ret["numel"] = flat.flat.numel()
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_all_reduce_world2():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["err"] < 1e-6, ret["err"]
    assert ret["numel"] == 2 * (6 * 5 + 6) + 6 + 6


def test_shard_range_covers_everything():
    from tec_mollm_b200.dist import shard_range

    for total in (1, 2, 7, 8, 128):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
