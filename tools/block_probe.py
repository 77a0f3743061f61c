#!/usr/bin/env python3
"""the reference's spatial block (tec_mollm.py:84-106: permute copy, encoder, residual add, permute
copy) written with torch ops around the encoder, against SpatialEncoder.forward_block (no input copy, residual + output
permute in one pass, one pass in backward).  fwd+bwd, CUDA events, 3 warm-ups, working set larger than L2.
   gpurun -- 'python tools/block_probe.py > gpurun_out/block_probe.jsonl'
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tec_mollm_b200 import SpatialEncoder, graph  # noqa: E402

dev = torch.device("cuda", 0)
L, N, F, H, C = 48, 2911, 22, 2, 11


def timed(fn, steps=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    lat, lon = np.linspace(15, 55, 41), np.linspace(70, 140, 71)
    ei, _ = graph.build_graph(lat, lon, 150.0, device=dev)
    for B in (8, 32, 128):
        enc = SpatialEncoder(F, C, heads=H, dropout=0.1).to(dev).train()
        x = torch.randn(B, L, N, F, device=dev).requires_grad_(True)
        gz = torch.randn(B * N, L, F, device=dev)

        def clear():
            x.grad = None
            for p in enc.parameters():
                p.grad = None

        def glue():
            clear()
            xg = x.permute(1, 0, 2, 3).reshape(-1, N, F)
            xs = xg + enc(xg, ei, None)
            xs.view(L, B, N, F).permute(1, 2, 0, 3).reshape(-1, L, F).backward(gz)

        def fused():
            clear()
            enc.forward_block(x, ei).backward(gz)

        def bare():
            clear()
            enc(x, ei).backward(gz_bl)

        gz_bl = torch.randn(B, L, N, F, device=dev)
        t_bare, t_glue, t_fused = timed(bare), timed(glue), timed(fused)
        rows = B * L * N
        print(json.dumps({"what": "spatial_block_fwd_bwd", "B": B, "ms_encoder_only": t_bare, "ms_reference_glue": t_glue,
                          "ms_forward_block": t_fused, "glue_overhead_ms": t_glue - t_bare, "fused_overhead_ms": t_fused - t_bare,
                          "fused_glue_GBps": rows * F * 4 * 6 / ((t_fused - t_bare) * 1e-3) / 1e9,
                          "note": "fused glue moves 6 x C x 4 B/row: fwd 3 (x, y, z), bwd transposition 2, +1 read for dx accumulating in place (bulk reduction store)"}))
        del enc, x, gz, gz_bl
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
