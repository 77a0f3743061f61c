#!/usr/bin/env python3
"""Where does the global grid's time go?  fwd+bwd phase times of the encoder (B = 4 x 48 snapshots, fp32, dropout 0.1) on the
180 x 360 global grid and on the same grid without its polar rows (|lat| <= 87.5, 85.5): the rows next to the poles carry 45 %
of the edges (degree 486) in 2 % of the nodes."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tec_mollm_b200 import SpatialEncoder, _lib, graph
dev = torch.device("cuda", 0)
lib = _lib.lib()
for name, lat in (("global 180x360", np.linspace(-89.5, 89.5, 180)), ("|lat|<=87.5 (176 rows)", np.linspace(-87.5, 87.5, 176)),
                  ("|lat|<=85.5 (172 rows)", np.linspace(-85.5, 85.5, 172)), ("|lat|<=60.5 (122 rows)", np.linspace(-60.5, 60.5, 122))):
    lon = np.linspace(-179.5, 179.5, 360)
    ei, _ = graph.build_graph(lat, lon, 150.0, device=dev)
    N, S = lat.size * lon.size, 4 * 48
    enc = SpatialEncoder(22, 11, heads=2, dropout=0.1, snapshot_mode="shared").to(dev).train()
    x = torch.randn(S, N, 22, device=dev, requires_grad=True)
    gy = torch.randn(S, N, 22, device=dev)
    plan = enc.gat_conv.plan_for(ei, N)
    for _ in range(3):
        x.grad = None; enc.zero_grad(set_to_none=True); enc(x, ei).backward(gy)
    torch.cuda.synchronize()
    lib.tecgat_phase_timing(1); ms4 = (ctypes.c_double * 4)(); lib.tecgat_phase_times(ms4)
    n = 5
    for _ in range(n):
        x.grad = None; enc.zero_grad(set_to_none=True); enc(x, ei).backward(gy)
    torch.cuda.synchronize(); lib.tecgat_phase_times(ms4); lib.tecgat_phase_timing(0)
    E = plan.num_edges
    print(json.dumps({"grid": name, "nodes": N, "edges_incl_self": E, "max_in_degree": plan.max_in_degree, "sliding_window": plan.sliding_window,
                      "ms": {k: round(ms4[i] / n, 3) for i, k in enumerate(("proj_fwd", "edge_fwd", "edge_bwd", "proj_bwd"))},
                      "edge_msgs_per_s": S * E / (sum(ms4) / n * 1e-3)}), flush=True)
    del x, gy, enc
    torch.cuda.empty_cache()
