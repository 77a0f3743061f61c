"""Shared helpers for the parity tests (checker side only)."""
import ctypes as C
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return np.load(os.path.join(GOLD, name))


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    """max |a - ref| / max |ref|  (the 'relative' of BASELINE.json's parity bar)."""
    a = a.detach().double().cpu()
    ref = ref.detach().double().cpu()
    denom = ref.abs().max().item()
    return (a - ref).abs().max().item() / (denom if denom > 0 else 1.0)


def random_graph(num_nodes, num_edges, seed, self_loops=True, duplicates=True, isolated=()):
    """Arbitrary directed multigraph: not symmetric, not sorted, optional self loops / duplicate edges /
    isolated nodes (the edge cases SURVEY.md Appendix A lists)."""
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, num_nodes, (2, num_edges), generator=g)
    if duplicates and num_edges >= 4:
        ei[:, -2:] = ei[:, :2]
    if self_loops and num_edges >= 6:
        ei[1, 2:4] = ei[0, 2:4]
    if not self_loops:
        ei = ei[:, ei[0] != ei[1]]
    for n in isolated:
        ei = ei[:, (ei[0] != n) & (ei[1] != n)]
    return ei


def kernel_dropout_mask(plan, snapshots, heads, p, seed, mode="shared"):
    """The keep-mask the CUDA kernels use, re-expressed in the oracle's edge order.

    Kernel counter = snapshot * E + csr_slot; oracle ("shared" mode, block-diagonal edge list) orders edges as
    [kept edges of snapshot 0, ..., kept edges of snapshot S-1, self loops of all S*N rows].
    Returns a float tensor (S*E, heads) in the oracle's order.
    """
    from tec_mollm_b200 import _lib

    E = plan.num_edges
    N = plan.num_nodes
    kept = plan.kept_edges
    _, _, eid = plan.export()
    keep = np.empty((snapshots * E, heads), dtype=np.uint8)
    _lib.call("tecgat_dropout_mask_host", C.c_uint64(seed), 0, snapshots * E, heads, C.c_float(p), E,
              C.c_void_p(keep.ctypes.data))
    keep = keep.reshape(snapshots, E, heads)
    eid = eid.astype(np.int64)
    is_self = eid >= kept
    if mode == "shared":
        out = np.empty((snapshots * E, heads), dtype=np.float32)
        for s in range(snapshots):
            # kept edge with PyG id e of snapshot s sits at s*kept + e; self loop of node i at S*kept + s*N + i
            pos = np.where(is_self, snapshots * kept + s * N + (eid - kept), s * kept + eid)
            out[pos] = keep[s]
    else:  # literal: the oracle's edge list is [kept edges (snapshot 0 only), self loops of all S*N rows]
        out = np.empty((kept + snapshots * N, heads), dtype=np.float32)
        out[np.where(is_self, kept + (eid - kept), eid)] = keep[0]
        for s in range(1, snapshots):
            out[kept + s * N + (eid[is_self] - kept)] = keep[s][is_self]
    return torch.from_numpy(out)


def kernel_projection(x2d, params, heads, out_channels, device, dtype_code=0):
    """xl, xr exactly as the module's forward computes them (same C-ABI call, same implementation choice): the
    kernels are bit-reproducible, so this is what the fused edge kernels saw."""
    from tec_mollm_b200 import _lib
    from tec_mollm_b200.gatv2 import _proj_impl

    R, F = x2d.shape
    HC = heads * out_channels
    d = lambda t: t.detach().float().contiguous().to(device)
    xd = d(x2d)
    wl, bl, wr, br = d(params["lin_l.weight"]), d(params["lin_l.bias"]), d(params["lin_r.weight"]), d(params["lin_r.bias"])
    st = torch.float32 if dtype_code == 0 else torch.bfloat16
    xl = torch.empty((R, HC), device=device, dtype=st)
    xr = torch.empty((R, HC), device=device, dtype=st)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(device):
        _lib.call("tecgat_project_fwd", ptr(xd), ptr(wl), ptr(bl), ptr(wr), ptr(br), ptr(xl), ptr(xr), R, F, HC,
                  dtype_code, _proj_impl(), C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        torch.cuda.synchronize(device)
    return xl.cpu(), xr.cpu()


def oracle_with_kernel_branches(x, ei, params, heads, out_channels, gy, device, snapshot_mode="shared", edge_mask=None,
                                p=0.0, amb_tol=1e-5):
    """fp64 oracle forward + closed-form backward in which every LeakyReLU DERIVATIVE takes the branch the kernels took.

    GATv2's gradient is discontinuous where a pre-activation s = xl_j + xr_i crosses zero.  When |s| is below the fp32
    resolution of xl / xr, every fp32 implementation -- PyG's included -- follows its own rounding, so 'the' reference
    gradient is only defined up to that choice.  This helper (1) recomputes the kernels' decisions from their own fp32
    xl / xr (one IEEE fp32 add, exactly the kernels' FADD), (2) asserts they differ from the fp64 decisions only where
    |s| < amb_tol, and (3) returns the exact fp64 gradient for those decisions: the comparison that follows is then
    strict (1e-5) on every entry.  The forward value is continuous in s and needs none of this."""
    from oracle import gatv2_oracle as G

    S, N, F = x.shape
    H, Cc = heads, out_channels
    x2d = x.reshape(-1, F).double()
    p64 = {k: v.double() for k, v in params.items()}
    ei_full = G.expand_shared(ei, N, S) if snapshot_mode == "shared" else ei
    xl32, xr32 = kernel_projection(x.reshape(-1, F), params, H, Cc, device)
    e2 = G.remove_then_add_self_loops(ei_full, S * N)
    s32 = xl32.view(-1, H, Cc)[e2[0]] + xr32.view(-1, H, Cc)[e2[1]]
    pos = s32 > 0
    out = G.gatv2_backward_manual(x2d, ei_full, p64, H, Cc, gy.reshape(-1, H * Cc).double(), edge_mask=edge_mask, p=p,
                                  pos_mask=pos, return_preact=True)
    s64 = out["_extra"]["s"]
    flipped = pos != (s64 > 0)
    n_flip = int(flipped.sum())
    if n_flip:
        assert s64[flipped].abs().max().item() < amb_tol, "a branch decision differs where |s| is NOT tiny"
        assert n_flip <= max(4, 1e-5 * s64.numel()), f"{n_flip} branch flips"
    y = out["_extra"]["y"].view(S, N, H * Cc)
    grads = {k: v for k, v in out.items() if k != "_extra"}
    grads["x"] = grads["x"].view(S, N, F)
    return y, grads, n_flip
