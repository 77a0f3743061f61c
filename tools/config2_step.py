#!/usr/bin/env python3
"""BASELINE config 2 stand-in: one full TEC-MoLLM-shaped training step at B = 2, bf16 autocast, to report the
step time and the spatial block's share of it.  Only the spatial block is this repository's code; everything around it is a
plain-torch stand-in with the reference's tensor shapes (src/model/tec_mollm.py:60-125), NOT a reimplementation: embedding concat
(6 + 16 channels), two multi-scale Conv1d blocks (k = 3/5/7, GroupNorm(1), GELU, 1x1 stride 2), 4-step patches -> d_llm = 768,
a random-init 3-layer GPT-2 (pretrained weights and peft are not available offline) with hand-written LoRA r = 32 on c_attn,
a 2-layer MLP head, HuberLoss, AdamW on the trainable parameters.
   gpurun -- 'python tools/config2_step.py > gpurun_out/config2.jsonl'
"""
import json
import os
import sys

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tec_mollm_b200 import SpatialEncoder, graph  # noqa: E402

dev = torch.device("cuda", 0)
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
    def __init__(self, fused: bool):
        super().__init__()
        from transformers import GPT2Config, GPT2Model
        self.fused = fused
        self.node_emb = nn.Embedding(N, 16)
        self.tod_emb = nn.Embedding(12, 16)
        self.spatial_encoder = SpatialEncoder(22, 11, heads=2)
        self.temporal = nn.Sequential(MSBlock(22, 64), MSBlock(64, 128))
        self.patch = nn.Linear(128 * 4, D_LLM)
        self.llm = GPT2Model(GPT2Config(n_layer=3, n_positions=64))
        for p in self.llm.parameters():
            p.requires_grad_(False)
        for blk in self.llm.h:
            blk.attn.c_attn = LoRA(blk.attn.c_attn)
        self.head = nn.Sequential(nn.Linear(3 * D_LLM, 3 * D_LLM // 4), nn.GELU(), nn.Dropout(0.1), nn.Linear(3 * D_LLM // 4, L_OUT))

    def forward(self, x, tod, edge_index):
        emb = self.node_emb.weight.view(1, 1, N, 16) + self.tod_emb(tod)               # (B, L, N, 16)
        x = torch.cat([x, emb.expand(B, L, N, 16)], dim=-1)                            # (B, L, N, 22)
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


def run(fused):
    torch.manual_seed(0)
    model = StandIn(fused).to(dev).train()
    lat, lon = np.linspace(15, 55, 41), np.linspace(70, 140, 71)
    ei, _ = graph.build_graph(lat, lon, 150.0, device=dev)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)
    x = torch.randn(B, L, N, 6, device=dev)
    tod = torch.randint(0, 12, (B, L, N), device=dev)
    target = torch.randn(B, L_OUT, N, 1, device=dev)
    loss_fn = nn.HuberLoss()
    ev = {}

    def step(mark=False):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if mark:
                ev["a"] = torch.cuda.Event(enable_timing=True); ev["a"].record()
            out = model(x, tod, ei)
            loss = loss_fn(out.float(), target)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # spatial block alone, same tensors, same autocast
    xin = torch.randn(B, L, N, 22, device=dev, requires_grad=True)
    gz = torch.randn(B * N, L, 22, device=dev)

    def block():
        xin.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if fused:
                z = model.spatial_encoder.forward_block(xin, ei)
            else:
                xg = xin.permute(1, 0, 2, 3).reshape(-1, N, 22)
                z = (xg + model.spatial_encoder(xg, ei, None)).view(L, B, N, 22).permute(1, 2, 0, 3).reshape(-1, L, 22)
        z.backward(gz)

    for _ in range(3):
        block()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        block()
    e1.record()
    torch.cuda.synchronize()
    ms_block = e0.elapsed_time(e1) / steps
    return {"what": "config2_train_step_stand_in", "B": B, "spatial_block": "forward_block (fused glue)" if fused else "reference glue (torch ops) around the encoder",
            "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "ms_spatial_block_fwd_bwd": ms_block, "spatial_share": ms_block / ms,
            "note": "bf16 autocast, eager (no CUDA graph); everything outside the spatial block is a plain-torch stand-in with the reference's shapes"}


if __name__ == "__main__":
    for fused in (False, True):
        print(json.dumps(run(fused)), flush=True)
