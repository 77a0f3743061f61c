#!/usr/bin/env python
"""GPU diagnostics (not a test): localises parity outliers.  Run on the B200 box."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_golden, rel_err  # noqa: E402
from oracle import gatv2_oracle as G  # noqa: E402
from oracle import graph_oracle as go  # noqa: E402

dev = torch.device("cuda", 0)


def proj(x, wl, bl, wr, br, impl, dtype=0):
    from tec_mollm_b200 import _lib

    R, F = x.shape
    HC = wl.shape[0]
    st = torch.float32 if dtype == 0 else torch.bfloat16
    xl = torch.full((R, HC), float("nan"), device=dev, dtype=st)
    xr = torch.full((R, HC), float("nan"), device=dev, dtype=st)
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.call("tecgat_project_fwd", p(x), p(wl), p(bl), p(wr), p(br), p(xl), p(xr), R, F, HC, dtype, impl,
              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return xl, xr


def diag_projection():
    print("=== projection TC vs FFMA vs fp64, multi-tile per CTA")
    for R, F, HC in [(128 * 1000 + 17, 22, 22), (128 * 700, 22, 44), (128 * 2000 + 5, 10, 10)]:
        gen = torch.Generator().manual_seed(R)
        x = torch.randn(R, F, generator=gen).to(dev)
        wl, wr = (torch.randn(HC, F, generator=gen) * 0.3).to(dev), (torch.randn(HC, F, generator=gen) * 0.3).to(dev)
        bl, br = torch.randn(HC, generator=gen).to(dev), torch.randn(HC, generator=gen).to(dev)
        ref_l = (x.double() @ wl.double().t() + bl.double())
        ref_r = (x.double() @ wr.double().t() + br.double())
        for name, impl in (("ffma", 1), ("tc", 0)):
            xl, xr = proj(x, wl, bl, wr, br, impl)
            for nm, out, ref in (("xl", xl, ref_l), ("xr", xr, ref_r)):
                d = (out.double() - ref).abs()
                bad = torch.isnan(d) | (d > 1e-5 * ref.abs().max())
                rows = torch.nonzero(bad.any(1)).flatten()
                print(f"R={R} F={F} HC={HC} {name} {nm}: max rel {rel_err(out, ref):.3e}; bad rows {rows.numel()}"
                      + (f" first {rows[:8].tolist()} tiles {sorted(set((rows // 128).tolist()))[:12]}" if rows.numel() else ""))


def run_enc(F, H, Cc, p, x, gy, ei, impl_env):
    os.environ["TECGAT_PROJ"] = impl_env
    from tec_mollm_b200 import SpatialEncoder

    enc = SpatialEncoder(F, Cc, heads=H, dropout=0.0).to(dev).eval()
    enc.load_state_dict({f"gat_conv.{k}": v.float() for k, v in p.items()}, strict=True)
    xg = x.float().to(dev).requires_grad_(True)
    y = enc(xg, ei.to(dev))
    y.backward(gy.float().to(dev))
    grads = {"x": xg.grad.detach().clone()}
    grads.update({k[len("gat_conv."):]: q.grad.detach().clone() for k, q in enc.named_parameters()})
    return y.detach(), grads


def report(name, a, ref, shape_hint=None):
    d = (a.double().cpu() - ref).abs()
    thr = 1e-5 * ref.abs().max()
    bad = torch.nonzero(d > thr)
    l2 = (d.norm() / ref.norm()).item()
    print(f"  {name}: max rel {d.max().item() / ref.abs().max().item():.3e}  rel L2 {l2:.3e}  entries over tol {bad.size(0)} / {d.numel()}")
    if bad.size(0):
        print("   first offenders (index, |diff|, ref):", [(b.tolist(), float(d[tuple(b)]), float(ref[tuple(b)])) for b in bad[:6]])
    return bad


def diag_dense_h4():
    print("=== cn300 / H=4 (backward fallback path), TC vs FFMA forward")
    g = load_golden("graph_cn300.npz")
    ei = torch.from_numpy(g["edge_index"])
    S, N, F, H, Cc = 2, 2911, 22, 4, 11
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(S, N, F, generator=gen, dtype=torch.float64)
    gy = torch.randn(S, N, H * Cc, generator=gen, dtype=torch.float64)
    p = G.init_params(F, Cc, H, seed=7, dtype=torch.float64)
    p["bias"] = torch.randn(H * Cc, generator=gen, dtype=torch.float64) * 0.1
    y_ref, g_ref = G.fwd_bwd(x, ei, p, H, Cc, gy)
    for impl in ("ffma", "tc"):
        y, grads = run_enc(F, H, Cc, p, x, gy, ei, impl)
        print(f" impl={impl}")
        report("y", y, y_ref)
        bad = report("dx", grads["x"], g_ref["x"])
        for k in G.PARAM_NAMES:
            report(k, grads[k], g_ref[k])
        if bad.size(0):
            # kink hypothesis: is some pre-activation s_ij of the offending destination within fp32 noise of zero?
            s_idx, node = int(bad[0][0]), int(bad[0][1])
            xl = (x[s_idx] @ p["lin_l.weight"].t() + p["lin_l.bias"])
            xr = (x[s_idx] @ p["lin_r.weight"].t() + p["lin_r.bias"])
            src_in = torch.cat([ei[0][ei[1] == node], torch.tensor([node])])
            dst_out = torch.cat([ei[1][ei[0] == node], torch.tensor([node])])
            s_in = (xl[src_in] + xr[node]).abs().min().item()
            s_out = (xl[node] + xr[dst_out]).abs().min().item()
            print(f"   offender snapshot {s_idx} node {node}: min |s| over in-edges {s_in:.3e}, over out-edges {s_out:.3e}")


def diag_global():
    print("=== global grid forward")
    from tec_mollm_b200 import graph

    lat, lon = go.synthetic_grid("global")
    ei, _ = graph.build_graph(lat, lon, 150.0, device=dev)
    S, N, F, H, Cc = 2, 64800, 22, 2, 11
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(S, N, F, generator=gen)
    gy = torch.randn(S, N, H * Cc, generator=gen)
    p = G.init_params(F, Cc, H, seed=4)
    y_ref, g_ref = G.fwd_bwd(x[:1].double(), ei.cpu(), {k: v.double() for k, v in p.items()}, H, Cc, gy[:1].double())
    for impl in ("ffma", "tc"):
        y, grads = run_enc(F, H, Cc, p, x, gy, ei.cpu(), impl)
        print(f" impl={impl}")
        bad = report("y[0]", y[:1], y_ref)
        if bad.size(0):
            nodes = sorted(set(bad[:, 1].tolist()))
            print("   offending nodes:", len(nodes), "lat rows:", sorted(set(n // 360 for n in nodes))[:20])
        report("dx[0]", grads["x"][:1], g_ref["x"])


if __name__ == "__main__":
    which = sys.argv[1:] or ["proj", "h4", "global"]
    if "proj" in which:
        diag_projection()
    if "h4" in which:
        diag_dense_h4()
    if "global" in which:
        diag_global()
