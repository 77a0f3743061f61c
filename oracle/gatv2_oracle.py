"""CPU oracle: op-for-op restatement of PyG ``GATv2Conv`` as the reference uses it.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED BY THE
REFERENCE: ``torch_geometric`` is an un-vendored, un-pinned dependency
(``/root/reference/README.md:57``) that cannot be installed here, and the
reference has no test for this path.  What is restated, and from where:

* call sites            ``/root/reference/src/model/modules.py:329-336`` (ctor:
                        ``GATv2Conv(in, out, heads, dropout, concat=True,
                        add_self_loops=True)``) and ``:352-358`` (forward on the
                        flattened ``(S*N, F)`` input with the one-graph
                        ``edge_index``);
* algorithm             upstream ``torch_geometric/nn/conv/gatv2_conv.py``
                        (``forward`` / ``edge_update`` / ``message``),
                        ``torch_geometric/utils/_softmax.py`` (detached
                        scatter-max, exp, scatter-sum ``+ 1e-16``),
                        ``torch_geometric/utils/loop.py`` (remove then append
                        self loops), ``torch_geometric/nn/inits.py`` (glorot,
                        uniform) -- restated in SURVEY.md Appendix A.

Everything is plain torch on CPU; dtype follows the inputs (fp64 = "truth",
fp32 = same-precision reference, ``torch.autocast('cpu', bfloat16)`` around
``gatv2_forward`` = the reference's bf16-autocast dtype flow, train.py:68).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

PARAM_NAMES = ("lin_l.weight", "lin_l.bias", "lin_r.weight", "lin_r.bias", "att", "bias")


# --------------------------------------------------------------------------
# parameter init (torch_geometric.nn.inits.glorot / uniform, Linear.reset_parameters)
# --------------------------------------------------------------------------
def _glorot_(t: torch.Tensor, gen: Optional[torch.Generator]) -> None:
    stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-stdv, stdv, generator=gen)


def _uniform_(size: int, t: torch.Tensor, gen: Optional[torch.Generator]) -> None:
    bound = 1.0 / math.sqrt(size)
    with torch.no_grad():
        t.uniform_(-bound, bound, generator=gen)


def init_params(in_channels: int, out_channels: int, heads: int, seed: int = 0,
                dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Parameters with PyG's names, shapes and init distributions.

    Draw order follows ``GATv2Conv.reset_parameters``: lin_l (weight glorot,
    bias U(+-1/sqrt(in))), lin_r (same), att glorot over its last two dims
    ``(H, C)``, output bias zeros.
    """
    gen = torch.Generator().manual_seed(seed)
    hc = heads * out_channels
    p = {
        "lin_l.weight": torch.empty(hc, in_channels),
        "lin_l.bias": torch.empty(hc),
        "lin_r.weight": torch.empty(hc, in_channels),
        "lin_r.bias": torch.empty(hc),
        "att": torch.empty(1, heads, out_channels),
        "bias": torch.zeros(hc),
    }
    _glorot_(p["lin_l.weight"], gen)
    _uniform_(in_channels, p["lin_l.bias"], gen)
    _glorot_(p["lin_r.weight"], gen)
    _uniform_(in_channels, p["lin_r.bias"], gen)
    _glorot_(p["att"], gen)
    return {k: v.to(dtype) for k, v in p.items()}


# --------------------------------------------------------------------------
# edge list surgery (torch_geometric.utils.loop)
# --------------------------------------------------------------------------
def remove_then_add_self_loops(edge_index: torch.Tensor, num_rows: int) -> torch.Tensor:
    """``remove_self_loops`` followed by ``add_self_loops``: the surviving edges keep
    their order and one ``(i, i)`` edge per row is appended LAST."""
    keep = edge_index[0] != edge_index[1]
    ei = edge_index[:, keep]
    loop = torch.arange(num_rows, dtype=edge_index.dtype)
    return torch.cat([ei, torch.stack([loop, loop])], dim=1)


def expand_shared(edge_index: torch.Tensor, num_nodes: int, snapshots: int) -> torch.Tensor:
    """Block-diagonal replication of a one-graph edge list over ``snapshots`` copies
    (the *intended* semantics of tec_mollm.py:84-89)."""
    offs = torch.arange(snapshots, dtype=edge_index.dtype).view(-1, 1, 1) * num_nodes
    return (edge_index.unsqueeze(0) + offs).permute(1, 0, 2).reshape(2, -1)


# --------------------------------------------------------------------------
# forward (op for op)
# --------------------------------------------------------------------------
def gatv2_forward(x: torch.Tensor, edge_index: torch.Tensor, params: Dict[str, torch.Tensor],
                  heads: int, out_channels: int, negative_slope: float = 0.2,
                  edge_mask: Optional[torch.Tensor] = None, p: float = 0.0,
                  return_attention: bool = False):
    """``GATv2Conv.forward(x2d, edge_index)`` with concat=True, add_self_loops=True.

    ``x``: (Nn, F); ``edge_index``: (2, E0) int64 with ``edge_index[0]`` = source j and
    ``edge_index[1]`` = target i.  ``edge_mask``: optional (E0' + Nn, H) keep-mask over
    the *post-surgery* edge order (kept edges, then self loops); with it
    ``alpha <- alpha * mask / (1 - p)`` (training-mode dropout with an explicit mask).
    """
    nn_rows = x.size(0)
    H, C = heads, out_channels
    xl = F.linear(x, params["lin_l.weight"], params["lin_l.bias"]).view(nn_rows, H, C)
    xr = F.linear(x, params["lin_r.weight"], params["lin_r.bias"]).view(nn_rows, H, C)
    ei = remove_then_add_self_loops(edge_index, nn_rows)
    src, dst = ei[0], ei[1]
    xj = xl.index_select(0, src)
    xi = xr.index_select(0, dst)
    z = F.leaky_relu(xi + xj, negative_slope)
    e = (z * params["att"]).sum(dim=-1)                                  # (E, H)
    idx = dst.view(-1, 1).expand(-1, H)
    m = torch.full((nn_rows, H), float("-inf"), dtype=e.dtype)
    m = m.scatter_reduce(0, idx, e.detach(), reduce="amax", include_self=True)
    ex = (e - m.index_select(0, dst)).exp()
    den = torch.zeros((nn_rows, H), dtype=e.dtype).scatter_add(0, idx, ex) + 1e-16
    alpha = ex / den.index_select(0, dst)
    alpha_pre = alpha
    if edge_mask is not None:
        alpha = alpha * edge_mask.to(alpha.dtype) / (1.0 - p)
    msg = xj * alpha.unsqueeze(-1)                                       # (E, H, C)
    out = torch.zeros((nn_rows, H, C), dtype=msg.dtype).index_add(0, dst, msg)
    y = out.view(nn_rows, H * C) + params["bias"]
    if return_attention:
        return y, (ei, alpha_pre, alpha)
    return y


def spatial_encoder_forward(x: torch.Tensor, edge_index: torch.Tensor, params: Dict[str, torch.Tensor],
                            heads: int, out_channels: int, snapshot_mode: str = "shared",
                            edge_mask: Optional[torch.Tensor] = None, p: float = 0.0) -> torch.Tensor:
    """``SpatialEncoder.forward`` (modules.py:340-359): ``x`` (S, N, F) -> (S, N, H*C).

    ``snapshot_mode="literal"`` is the call exactly as written (edges only reach
    snapshot 0); ``"shared"`` applies the one-graph edge list to every snapshot.
    """
    S, N, Fin = x.shape
    x2d = x.reshape(-1, Fin)
    if snapshot_mode == "shared":
        ei = expand_shared(edge_index, N, S)
    elif snapshot_mode == "literal":
        ei = edge_index
    else:
        raise ValueError(snapshot_mode)
    y = gatv2_forward(x2d, ei, params, heads, out_channels, edge_mask=edge_mask, p=p)
    return y.view(S, N, heads * out_channels)


# --------------------------------------------------------------------------
# hand-derived backward -- used to self-check the oracle
# and to document the formulas the CUDA backward implements
# --------------------------------------------------------------------------
def gatv2_backward_manual(x, edge_index, params, heads, out_channels, grad_y,
                          negative_slope: float = 0.2, edge_mask=None, p: float = 0.0, pos_mask=None,
                          return_preact: bool = False):
    """Closed-form gradients.  ``pos_mask`` (E, H, C) bool, optional, overrides the LeakyReLU branch decision
    ``s > 0`` used for the DERIVATIVE (the forward value is continuous and keeps its own sign): the gradient of GATv2 is
    discontinuous where a pre-activation ``s = xl_j + xr_i`` crosses zero, and when ``|s|`` is below the fp32 resolution
    of ``xl``/``xr`` any fp32 implementation (PyG's included) picks the branch its own rounding dictates.  The GPU tests
    pass the branch decisions implied by the kernels' own fp32 ``xl``/``xr`` and separately assert that they differ from
    the fp64 decisions only where ``|s|`` is tiny."""
    nn_rows = x.size(0)
    H, C = heads, out_channels
    Wl, bl, Wr, br = params["lin_l.weight"], params["lin_l.bias"], params["lin_r.weight"], params["lin_r.bias"]
    att = params["att"]
    xl = (x @ Wl.t() + bl).view(nn_rows, H, C)
    xr = (x @ Wr.t() + br).view(nn_rows, H, C)
    ei = remove_then_add_self_loops(edge_index, nn_rows)
    src, dst = ei[0], ei[1]
    s = xl[src] + xr[dst]
    z = torch.where(s > 0, s, s * negative_slope)
    e = (z * att).sum(-1)
    idx = dst.view(-1, 1).expand(-1, H)
    m = torch.full((nn_rows, H), float("-inf"), dtype=e.dtype).scatter_reduce(0, idx, e, reduce="amax")
    ex = (e - m[dst]).exp()
    den = torch.zeros((nn_rows, H), dtype=e.dtype).scatter_add(0, idx, ex) + 1e-16
    alpha = ex / den[dst]
    q = torch.ones_like(alpha) if edge_mask is None else edge_mask.to(alpha.dtype) / (1.0 - p)
    out = torch.zeros((nn_rows, H, C), dtype=x.dtype).index_add(0, dst, xl[src] * (alpha * q).unsqueeze(-1))
    g = grad_y.view(nn_rows, H, C)
    delta = (g * out).sum(-1)                                            # (Nn, H): g_i . out_i
    gx = (g[dst] * xl[src]).sum(-1)                                      # (E, H)
    de = alpha * (q * gx - delta[dst])
    pos = (s > 0) if pos_mask is None else pos_mask
    ds = de.unsqueeze(-1) * att * torch.where(pos, torch.ones_like(s), torch.full_like(s, negative_slope))
    d_att = (de.unsqueeze(-1) * z).sum(0, keepdim=True)
    d_xl = torch.zeros_like(xl).index_add(0, src, (alpha * q).unsqueeze(-1) * g[dst] + ds)
    d_xr = torch.zeros_like(xr).index_add(0, dst, ds)
    d_xl2, d_xr2 = d_xl.view(nn_rows, H * C), d_xr.view(nn_rows, H * C)
    y = out.view(nn_rows, H * C) + params["bias"]
    extra = {"y": y, "d_xl": d_xl2, "d_xr": d_xr2}
    if return_preact:
        extra.update({"s": s, "edges": ei})
    return {
        "_extra": extra,
        "x": d_xl2 @ Wl + d_xr2 @ Wr,
        "lin_l.weight": d_xl2.t() @ x, "lin_l.bias": d_xl2.sum(0),
        "lin_r.weight": d_xr2.t() @ x, "lin_r.bias": d_xr2.sum(0),
        "att": d_att, "bias": grad_y.sum(0),
    }


# --------------------------------------------------------------------------
# independent dense formulation (adjacency counts) -- second self-check
# --------------------------------------------------------------------------
def gatv2_forward_dense(x, edge_index, params, heads, out_channels, negative_slope: float = 0.2):
    """O(Nn^2) formulation with an edge-multiplicity matrix; no scatter ops.  Small inputs only."""
    nn_rows = x.size(0)
    H, C = heads, out_channels
    xl = (x @ params["lin_l.weight"].t() + params["lin_l.bias"]).view(nn_rows, H, C)
    xr = (x @ params["lin_r.weight"].t() + params["lin_r.bias"]).view(nn_rows, H, C)
    cnt = torch.zeros(nn_rows, nn_rows, dtype=x.dtype)                   # cnt[i, j] = #edges j -> i
    for j, i in edge_index.t().tolist():
        if i != j:
            cnt[i, j] += 1
    cnt += torch.eye(nn_rows, dtype=x.dtype)
    s = xr.unsqueeze(1) + xl.unsqueeze(0)                                # (i, j, H, C)
    z = torch.where(s > 0, s, negative_slope * s)
    e = (z * params["att"].view(1, 1, H, C)).sum(-1)                     # (i, j, H)
    e = e.masked_fill(cnt.unsqueeze(-1) == 0, float("-inf"))
    m = e.max(dim=1, keepdim=True).values
    ex = (e - m).exp() * cnt.unsqueeze(-1)
    alpha = ex / (ex.sum(1, keepdim=True) + 1e-16)
    out = torch.einsum("ijh,jhc->ihc", alpha, xl)
    return out.reshape(nn_rows, H * C) + params["bias"]


# --------------------------------------------------------------------------
# fwd+bwd helper used by the parity tests and the CPU baseline
# --------------------------------------------------------------------------
def fwd_bwd(x: torch.Tensor, edge_index: torch.Tensor, params: Dict[str, torch.Tensor], heads: int,
            out_channels: int, grad_out: torch.Tensor, snapshot_mode: str = "shared",
            edge_mask: Optional[torch.Tensor] = None, p: float = 0.0,
            autocast_bf16: bool = False) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """Forward + autograd backward through the restatement.  Returns ``(y, grads)`` with
    ``grads`` keyed by ``"x"`` and the six PyG parameter names."""
    xg = x.detach().clone().requires_grad_(True)
    pg = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    if autocast_bf16:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            y = spatial_encoder_forward(xg, edge_index, pg, heads, out_channels, snapshot_mode, edge_mask, p)
    else:
        y = spatial_encoder_forward(xg, edge_index, pg, heads, out_channels, snapshot_mode, edge_mask, p)
    y.backward(grad_out.to(y.dtype))
    grads = {"x": xg.grad}
    grads.update({k: v.grad for k, v in pg.items()})
    return y.detach(), grads
