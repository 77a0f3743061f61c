#!/usr/bin/env python3
"""Where does the time go at B = 2 (the reference's training batch)?  Host wall clock vs device time per fwd+bwd, eager and as
a replayed CUDA graph."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tec_mollm_b200 import SpatialEncoder
dev = torch.device("cuda", 0)
ei = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "graph_cn150.npz"))["edge_index"]).to(dev)
for B in (2, 8):
    S, N, F, H, C = B * 48, 2911, 22, 2, 11
    enc = SpatialEncoder(F, C, heads=H, dropout=0.1).to(dev).train()
    x = torch.randn(S, N, F, device=dev).requires_grad_(True)
    gy = torch.randn(S, N, H * C, device=dev)
    def step():
        x.grad = None
        enc.zero_grad(set_to_none=True)
        enc(x, ei).backward(gy)
    for _ in range(20): step()
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n): step()
    host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n
    # CUDA graph (eval-mode dropout seed is drawn on the host, so capture with a fixed seed: dropout off)
    enc.eval()
    side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3): step()
    torch.cuda.current_stream(dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    x.grad = None; enc.zero_grad(set_to_none=True)
    with torch.cuda.graph(g, stream=side):
        enc(x, ei).backward(gy)
    for _ in range(10): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"B={B}: eager host-issue {host*1e6:.0f} us/step, eager wall {wall*1e6:.0f} us/step, CUDA-graph replay {e0.elapsed_time(e1)/n*1e3:.0f} us/step")

# where the eager host time goes (B = 2): cProfile of 300 steps, top entries by cumulative time
if os.environ.get("PROBE_PROFILE", "1") == "1":
    import cProfile, pstats, io
    B = 2
    S, N, F, H, C = B * 48, 2911, 22, 2, 11
    enc = SpatialEncoder(F, C, heads=H, dropout=0.1).to(dev).train()
    x = torch.randn(S, N, F, device=dev).requires_grad_(True)
    gy = torch.randn(S, N, H * C, device=dev)
    def step2():
        x.grad = None
        enc.zero_grad(set_to_none=True)
        enc(x, ei).backward(gy)
    for _ in range(20): step2()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300): step2()
    pr.disable()
    torch.cuda.synchronize()
    sio = io.StringIO()
    pstats.Stats(pr, stream=sio).sort_stats("cumulative").print_stats(28)
    print(sio.getvalue())
