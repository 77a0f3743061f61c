#!/usr/bin/env python3
"""Names the collective kernel behind the 4.2 KB gradient all-reduce and its device time (no nsys in the image: torch.profiler /
CUPTI on rank 0).   torchrun --nproc-per-node N tools/allreduce_probe.py [--batch-per-gpu B]"""
import argparse, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tec_mollm_b200 import SpatialEncoder, graph
from tec_mollm_b200 import dist as tdist
ap = argparse.ArgumentParser(); ap.add_argument("--batch-per-gpu", type=int, default=16); args = ap.parse_args()
rank, world, lr = tdist.init_from_env("nccl")
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
ei, _ = graph.build_graph(np.linspace(15, 55, 41), np.linspace(70, 140, 71), 150.0, device=dev)
S, N = args.batch_per_gpu * 48, 2911
enc = SpatialEncoder(22, 11, heads=2, dropout=0.1, snapshot_mode="shared").to(dev).train()
flat = tdist.FlatGradAllReduce(enc.parameters(), module=enc)
x = torch.randn(S, N, 22, device=dev, requires_grad=True); gy = torch.randn(S, N, 22, device=dev)
def step():
    flat.zero_(); x.grad = None; enc(x, ei).backward(gy); flat.all_reduce_mean()
for _ in range(10): step()
torch.cuda.synchronize(); dist.barrier(device_ids=[lr])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(30): step()
    torch.cuda.synchronize()
if rank == 0:
    rows = [(e.key, e.count, e.device_time_total / max(1, e.count)) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1] * r[2])
    print(f"# device kernels of 30 training steps on rank 0 of {world} (B = {args.batch_per_gpu} per GPU), torch.profiler\n")
    print("| kernel | launches | avg us |\n|---|---|---|")
    for k, c, us in rows[:14]:
        print(f"| `{k[:110]}` | {c} | {us:.1f} |")
dist.barrier(device_ids=[lr]); dist.destroy_process_group()
