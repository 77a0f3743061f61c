#!/usr/bin/env python3
"""BASELINE config 2 stand-in + the step hygiene of SURVEY.md 8f N4: a TEC-MoLLM-shaped training loop at B = 2 per GPU, bf16
autocast, gradient accumulation 6 (train.py:182,186), single GPU or DDP under torchrun.

Only the spatial block (SpatioTemporalEmbedding + SpatialEncoder) is this repository's code; everything around it is a plain-torch
stand-in with the reference's tensor shapes (src/model/tec_mollm.py:60-125), NOT a reimplementation: two multi-scale Conv1d blocks
(k = 3/5/7, GroupNorm(1), GELU, 1x1 stride 2), 4-step patches -> d_llm = 768, a random-init 3-layer GPT-2 (pretrained weights and
peft are not available offline) with hand-written LoRA r = 32 on c_attn, a 2-layer MLP head, HuberLoss, AdamW.

Two loop bodies over the same model, so the difference is the harness alone:
  --hygiene reference : train.py:57-112 as written -- gradients all-reduced on EVERY micro-step (no no_sync()), `empty_cache()`
                        and `loss.item()` after every micro-step, GradScaler around a bf16 autocast;
  --hygiene clean     : `no_sync()` on the non-boundary micro-steps, the loss accumulated on the device and read once per
                        optimizer step, no allocator flush, no scaler (bf16 needs none).
Reports optimizer-step time, samples/s (global) and the spatial block's share.

    python tools/config2_step.py                                   # one GPU, both modes, fused and unfused glue
    torchrun --nproc-per-node 8 tools/config2_step.py --ddp        # DDP over NCCL
"""
import argparse
import contextlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tec_mollm_b200 import SpatialEncoder, SpatioTemporalEmbedding, graph  # noqa: E402
from tec_mollm_b200 import dist as tdist  # noqa: E402

B, L, N, L_OUT, D_LLM = 2, 48, 2911, 12, 768


class MSBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.convs = nn.ModuleList([nn.Sequential(nn.Conv1d(cin, cout, k, padding=(k - 1) // 2), nn.GroupNorm(1, cout), nn.GELU())
                                    for k in (3, 5, 7)])
        self.final = nn.Conv1d(3 * cout, cout, 1, stride=2)

    def forward(self, x):
        return self.final(torch.cat([c(x) for c in self.convs], dim=1))


class LoRA(nn.Module):  # y = base(x) + (x A) B * (alpha / r) around GPT-2's Conv1D c_attn
    def __init__(self, base, r=32, alpha=32):
        super().__init__()
        self.base = base
        nin, nout = base.weight.shape
        self.A = nn.Parameter(torch.randn(nin, r) * 0.01)
        self.Bm = nn.Parameter(torch.zeros(r, nout))
        self.scale = alpha / r

    def forward(self, x):
        return self.base(x) + (x @ self.A) @ self.Bm * self.scale


class StandIn(nn.Module):
    def __init__(self, fused: bool, temporal: str = "torch"):
        super().__init__()
        from transformers import GPT2Config, GPT2Model
        from tec_mollm_b200 import MultiScaleConvEmbedder
        self.fused = fused
        self.spatio_temporal_embedding = SpatioTemporalEmbedding(16, num_nodes=N)     # tec_mollm.py:25-29 (drop-in)
        self.spatial_encoder = SpatialEncoder(22, 11, heads=2, snapshot_mode="shared")  # tec_mollm.py:33-37 (drop-in)
        # TemporalEncoder conv embedder (modules.py:62-91): plain torch ops, or this repository's drop-in
        self.temporal = MultiScaleConvEmbedder(22, [64, 128], [2, 2]) if temporal == "fused" else nn.Sequential(MSBlock(22, 64), MSBlock(64, 128))
        self.patch = nn.Linear(128 * 4, D_LLM)
        self.llm = GPT2Model(GPT2Config(n_layer=3, n_positions=64))
        for p in self.llm.parameters():
            p.requires_grad_(False)
        for blk in self.llm.h:
            blk.attn.c_attn = LoRA(blk.attn.c_attn)
        self.head = nn.Sequential(nn.Linear(3 * D_LLM, 3 * D_LLM // 4), nn.GELU(), nn.Dropout(0.1), nn.Linear(3 * D_LLM // 4, L_OUT))

    def forward(self, x, time_features, edge_index):
        x = self.spatio_temporal_embedding(x, time_features)                           # (B, L, N, 22): one fused kernel
        if self.fused:
            xt = self.spatial_encoder.forward_block(x, edge_index)                     # (B*N, L, 22)
        else:                                                                          # tec_mollm.py:84-106 as written
            xg = x.permute(1, 0, 2, 3).reshape(-1, N, 22)
            xs = xg + self.spatial_encoder(xg, edge_index, None)
            xt = xs.view(L, B, N, 22).permute(1, 2, 0, 3).reshape(-1, L, 22)
        h = self.temporal(xt.transpose(1, 2))                                          # (B*N, 128, 12)
        h = self.patch(h.transpose(1, 2).reshape(B * N, 3, 4 * 128))                   # (B*N, 3, 768)
        h = self.llm(inputs_embeds=h).last_hidden_state
        h = nn.functional.dropout(h, 0.1, self.training)
        return self.head(h.reshape(B * N, -1)).view(B, N, L_OUT).permute(0, 2, 1).unsqueeze(-1)


def run(fused, hygiene, accum, opt_steps, ddp, dev, rank, world, temporal="torch"):
    torch.manual_seed(0)
    model = StandIn(fused, temporal).to(dev).train()
    net = nn.parallel.DistributedDataParallel(model, device_ids=[dev.index]) if ddp else model   # train.py:354
    lat, lon = np.linspace(15, 55, 41), np.linspace(70, 140, 71)
    ei, _ = graph.build_graph(lat, lon, 150.0, device=dev)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)
    scaler = torch.amp.GradScaler("cuda", enabled=hygiene == "reference")              # train.py:350
    gen = torch.Generator().manual_seed(1 + rank)
    # host batches, as the DataLoader hands them over (train.py:58-60): x (B, L, N, 6), y, (B, L, 4) time features
    xs = [torch.randn(B, L, N, 6, generator=gen).pin_memory() for _ in range(accum)]
    tfs = [torch.stack([torch.randint(0, 12, (B, L), generator=gen), torch.randint(0, 366, (B, L), generator=gen),
                        torch.randint(0, 13, (B, L), generator=gen), torch.randint(0, 4, (B, L), generator=gen)], -1).float().pin_memory()
           for _ in range(accum)]
    ys = [torch.randn(B, L_OUT, N, 1, generator=gen).pin_memory() for _ in range(accum)]
    loss_fn = nn.HuberLoss()

    def optimizer_step():
        total = torch.zeros((), device=dev)
        for i in range(accum):
            x, tf, y = xs[i].to(dev, non_blocking=True), tfs[i].to(dev, non_blocking=True), ys[i].to(dev, non_blocking=True)
            boundary = i == accum - 1
            sync = contextlib.nullcontext() if (hygiene == "reference" or boundary or not ddp) else net.no_sync()
            with sync, torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(x, tf.unsqueeze(-2).expand(B, L, N, 4), ei)                   # train.py:64-65, :75
                loss = loss_fn(out.float(), y) / accum
            with sync:
                scaler.scale(loss).backward()
            if hygiene == "reference":
                del x, y, tf, out
                torch.cuda.empty_cache()                                               # train.py:85
                total += loss.item() * accum                                           # train.py:112 (D2H sync per micro-step)
            else:
                total += loss.detach()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], max_norm=1.0)
        scaler.step(opt)
        scaler.update()
        opt.zero_grad(set_to_none=True)
        return float(total)  # one read per optimizer step

    for _ in range(2):
        optimizer_step()
    torch.cuda.synchronize(dev)
    if ddp:
        dist.barrier(device_ids=[dev.index])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(opt_steps):
        optimizer_step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / opt_steps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if ddp:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()

    # spatial block alone (embedding + encoder + glue), same tensors, same autocast
    xin = xs[0].to(dev)
    tf = tfs[0].to(dev)
    gz = torch.randn(B * N, L, 22, device=dev)

    def block():
        model.spatio_temporal_embedding.zero_grad(set_to_none=True)
        model.spatial_encoder.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            x = model.spatio_temporal_embedding(xin, tf)
            if fused:
                z = model.spatial_encoder.forward_block(x, ei)
            else:
                xg = x.permute(1, 0, 2, 3).reshape(-1, N, 22)
                z = (xg + model.spatial_encoder(xg, ei, None)).view(L, B, N, 22).permute(1, 2, 0, 3).reshape(-1, L, 22)
        z.backward(gz)

    for _ in range(3):
        block()
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(20):
        block()
    e1.record()
    torch.cuda.synchronize(dev)
    ms_block = e0.elapsed_time(e1) / 20
    return {"what": "config2_train_loop_stand_in", "hygiene": hygiene, "ddp_world": world if ddp else 1, "batch_per_gpu": B,
            "accumulation_steps": accum, "temporal_conv_embedder": "tec_mollm_b200.MultiScaleConvEmbedder" if temporal == "fused" else "torch ops",
            "spatial_block": "fused embedding + forward_block" if fused else "fused embedding + reference glue (torch ops)",
            "ms_per_optimizer_step": ms, "ms_per_micro_step": ms / accum, "samples_per_s": world * B * accum / (ms * 1e-3),
            "ms_spatial_block_fwd_bwd": ms_block, "spatial_share_of_micro_step": ms_block / (ms / accum),
            "note": "bf16 autocast, eager; everything outside the spatial block is a plain-torch stand-in with the reference's shapes"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ddp", action="store_true")
    ap.add_argument("--accum", type=int, default=6)
    ap.add_argument("--opt-steps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local_rank = tdist.init_from_env("nccl") if args.ddp else (0, 1, 0)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    for hygiene, fused, temporal in (("reference", False, "torch"), ("clean", False, "torch"), ("clean", True, "torch"), ("clean", True, "fused")):
        r = run(fused, hygiene, args.accum, args.opt_steps, args.ddp, dev, rank, world, temporal)
        if rank == 0:
            print(json.dumps(r), flush=True)
    if args.ddp:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
