#!/usr/bin/env python3
"""Phase breakdown of BASELINE config 4 (300 km graph, H=4, C=11) on one GPU."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tec_mollm_b200 import SpatialEncoder, gatv2
dev = torch.device("cuda", 0)
ei = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "graph_cn300.npz"))["edge_index"]).to(dev)
for (H, B) in ((4, 16), (2, 16)):
    S, N, F, C = B * 48, 2911, 22, 11
    enc = SpatialEncoder(F, C, heads=H, dropout=0.1).to(dev).train()
    x = torch.randn(S, N, F, device=dev).requires_grad_(True)
    gy = torch.randn(S, N, H * C, device=dev)
    def step():
        x.grad = None; enc.zero_grad(set_to_none=True)
        enc(x, ei).backward(gy)
    for _ in range(3): step()
    torch.cuda.synchronize()
    gatv2.PHASE_EVENTS = []
    for _ in range(3): step()
    torch.cuda.synchronize()
    ev, gatv2.PHASE_EVENTS = gatv2.PHASE_EVENTS, None
    ph = {}; prev = None
    for name, e in ev:
        if prev is not None and name in ("proj_fwd", "edge_fwd", "edge_bwd", "proj_bwd"):
            ph[name] = ph.get(name, 0.0) + prev.elapsed_time(e) / 3
        prev = e
    plan = enc.gat_conv.plan_for(ei, N)
    print(f"300 km H={H} B={B}: window fwd {plan.max_window} bwd {plan.max_window_bwd} tiles {plan.num_tiles}/{plan.num_tiles_bwd}", {k: round(v, 3) for k, v in ph.items()})
