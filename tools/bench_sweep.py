#!/usr/bin/env python3
"""Secondary measurements for BASELINE.json's other configurations (the bench.py line stays config 3, B=128):

  * batch sweep B = 2 / 8 / 32 / 128 x 48 snapshots, default shape (configs 1 and 3), fp32 and bf16-autocast;
  * config 4: 300 km graph (E = 79,443), H = 4, C = 11;
  * config 5: 64,800-node global 1-degree grid: haversine build (pairs/s) + GATv2 fwd+bwd at B = 4 x 48 snapshots;
  * graph builder on the 2911-node grid (150 / 300 km).

One GPU, CUDA events on the launching stream, 3 warm-ups, inputs larger than L2 except where noted.  Prints one JSON
object per measurement (stdout); run with  gpurun -- 'python tools/bench_sweep.py > gpurun_out/sweep.jsonl'.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tec_mollm_b200 import SpatialEncoder, graph  # noqa: E402

L_IN = 48
dev = torch.device("cuda", 0)


def time_encoder(ei, N, F, H, C, B, autocast=False, steps=5, warmup=3, dropout=0.1):
    S = B * L_IN
    enc = SpatialEncoder(F, C, heads=H, dropout=dropout).to(dev).train()
    x = torch.randn(S, N, F, device=dev).requires_grad_(True)
    gy = torch.randn(S, N, H * C, device=dev)

    def step():
        x.grad = None
        for p in enc.parameters():
            p.grad = None
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = enc(x, ei)
        else:
            y = enc(x, ei)
        y.backward(gy)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    E = enc.gat_conv.plan_for(ei, N).num_edges
    del x, gy, enc
    torch.cuda.empty_cache()
    return ms, S * E / (ms * 1e-3), E


def grid(lat0, lat1, nlat, lon0, lon1, nlon):
    return np.linspace(lat0, lat1, nlat), np.linspace(lon0, lon1, nlon)


def time_graph(lat, lon, thr, reps=3):
    graph.build_graph(lat, lon, thr, device=dev)  # warm-up (module load, allocator)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        ei, ew = graph.build_graph(lat, lon, thr, device=dev)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    n = len(lat) * len(lon)
    return best, n, ei


def main():
    out = []
    lat, lon = grid(15, 55, 41, 70, 140, 71)  # the synthetic 41 x 71 grid
    for thr in (150.0, 300.0):
        t, n, ei = time_graph(lat, lon, thr)
        out.append({"what": "graph_build", "nodes": n, "threshold_km": thr, "edges": int(ei.size(1)), "seconds": t,
                    "pairs_per_s": n * n / t, "note": "host call incl. count + scan + fill + guard-band re-check, wall clock"})
        if thr == 150.0:
            ei150 = ei
        else:
            ei300 = ei
    for B in (2, 8, 32, 128):
        for ac in (False, True):
            ms, v, E = time_encoder(ei150, 2911, 22, 2, 11, B, autocast=ac)
            out.append({"what": "gatv2_fwd_bwd", "config": "150 km, F=22 H=2 C=11", "B": B, "autocast_bf16": ac, "ms_per_step": ms,
                        "edge_msgs_per_s": v, "samples_per_s": B / (ms * 1e-3), "edges_per_snapshot": E,
                        "note": "B=2 working set (0.4 GB) is larger than L2" })
    for B in (8, 32):
        ms, v, E = time_encoder(ei300, 2911, 22, 4, 11, B)
        out.append({"what": "gatv2_fwd_bwd", "config": "300 km, F=22 H=4 C=11 (BASELINE config 4)", "B": B, "autocast_bf16": False,
                    "ms_per_step": ms, "edge_msgs_per_s": v, "samples_per_s": B / (ms * 1e-3), "edges_per_snapshot": E})
    glat, glon = np.arange(-89.5, 90.0, 1.0), np.arange(-179.5, 180.0, 1.0)  # 180 x 360 cell-centred global grid
    t, n, eig = time_graph(glat, glon, 150.0, reps=2)
    out.append({"what": "graph_build", "nodes": n, "threshold_km": 150.0, "edges": int(eig.size(1)), "seconds": t,
                "pairs_per_s": n * n / t, "note": "BASELINE config 5 grid (64,800 nodes): the reference's dense (N,N) fp64 matrix would be 33.6 GB"})
    ms, v, E = time_encoder(eig, n, 22, 2, 11, 4, steps=3)
    out.append({"what": "gatv2_fwd_bwd", "config": "global 1-degree grid 64,800 nodes (BASELINE config 5, one GPU's share)", "B": 4,
                "autocast_bf16": False, "ms_per_step": ms, "edge_msgs_per_s": v, "samples_per_s": 4 / (ms * 1e-3), "edges_per_snapshot": E})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
