"""Drop-in ``GATv2Conv`` for the TEC-MoLLM SpatialEncoder, backed by ``libtecgat.so`` (sm_100a).

Mirrors the operator the reference instantiates at ``src/model/modules.py:329-336``
(``torch_geometric.nn.GATv2Conv(in, out, heads=, dropout=, concat=True, add_self_loops=True)``)
and calls at ``:356`` (``forward(x2d, edge_index)``): same constructor, same parameter names and
shapes (``lin_l.weight``, ``lin_l.bias``, ``lin_r.weight``, ``lin_r.bias``, ``att``, ``bias`` -- so
reference checkpoints load with ``strict=True``), same init distributions, same math.

All compute runs in hand-written CUDA through the C ABI (``include/tecgat.h``); this file only
owns tensors, the cached graph plan and the autograd wiring.  No Triton, no dispatch, no CPU path:
a non-CUDA input raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib

__all__ = ["GATv2Conv", "GraphPlan", "tile_nodes_for"]

_COMPILED_CHANNELS = (1, 2, 3, 4, 5, 6, 7, 8, 11, 12, 16, 24, 32)  # TG_FOR_EACH_C in csrc/edge_common.cuh


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)  # the handle without building a Stream object (~10x cheaper)


def _stream(device: torch.device):
    if _raw_stream is not None and device.index is not None:
        return C.c_void_p(_raw_stream(device.index))
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def tile_nodes_for(heads: int, backward: bool = False, out_channels: Optional[int] = None) -> int:
    """Destination nodes per tile of the persistent edge kernels: one lane per (node, head), heads padded to a power
    of two.  Forward: 15 consumer warps + one producer warp (512 threads, 128 registers).  Backward: 7 consumer warps + one
    producer warp (256 threads, 253 registers), or -- for the shapes compiled with a fixed head count (heads = 2,
    out_channels 5 or 11) -- 8 consumer warps + a producer warpgroup that hands its registers over (384 threads,
    ``setmaxnreg`` 232 / 40)."""
    if heads < 1 or heads > 32:
        raise ValueError(f"heads={heads} unsupported (1..32)")
    hp = 1
    while hp < heads:
        hp *= 2
    npw = 32 // hp  # nodes per warp
    wide = backward and heads == 2 and out_channels in (5, 11)
    warps = (8 if wide else 7) if backward else 15
    knob = os.environ.get("TECGAT_TILE_BWD" if backward else "TECGAT_TILE_FWD")  # tuning knob (benchmarks only)
    t = int(knob) if knob else warps * npw
    return max(npw, min(warps * npw, (t // npw) * npw))


def _proj_impl() -> int:
    """Tensor-core projections are the product path; ``TECGAT_PROJ=ffma`` selects the CUDA-core
    cross-check kernels (tests only)."""
    return _lib.PROJ_FFMA if os.environ.get("TECGAT_PROJ", "tc").lower() == "ffma" else _lib.PROJ_TC


class GraphPlan:
    """Immutable device-side plan of one ``edge_index`` (see csrc/plan.cu).  Replaces the per-forward
    ``remove_self_loops``/``add_self_loops`` of PyG with a one-time build."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, tile_nodes: int, tile_nodes_bwd: Optional[int] = None):
        if not edge_index.is_cuda:
            raise RuntimeError("tec_mollm_b200: edge_index must be a CUDA tensor (no CPU path)")
        if edge_index.dim() != 2 or edge_index.size(0) != 2 or edge_index.dtype != torch.int64:
            raise ValueError(f"edge_index must be int64 of shape (2, E); got {edge_index.dtype} {tuple(edge_index.shape)}")
        ei = edge_index.contiguous()
        self.device = ei.device
        self.num_nodes = int(num_nodes)
        self.tile_nodes = int(tile_nodes)
        self.tile_nodes_bwd = int(tile_nodes_bwd if tile_nodes_bwd is not None else tile_nodes)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.call("tecgat_plan_create", _ptr(ei), ei.size(1), self.num_nodes, self.tile_nodes, self.tile_nodes_bwd,
                      _stream(self.device), C.byref(handle))
        self._h = handle
        info = (C.c_int64 * 12)()
        _lib.call("tecgat_plan_info", self._h, info)
        (self.num_edges, self.max_in_degree, self.max_out_degree, self.num_tiles, _, self.max_window, _,
         self.kept_edges, self.num_tiles_bwd, _, self.max_window_bwd, sw) = [int(v) for v in info]
        self.sliding_window = bool(sw)  # banded graph: the backward can run the sliding-window kernel

    @property
    def handle(self):
        return self._h

    def export(self):
        """Host copies (numpy int32): rowptr (N+1), col (E), eid (E) -- see tecgat_plan_export."""
        import numpy as np

        rowptr = np.empty(self.num_nodes + 1, dtype=np.int32)
        col = np.empty(self.num_edges, dtype=np.int32)
        eid = np.empty(self.num_edges, dtype=np.int32)
        _lib.call("tecgat_plan_export", self._h, C.c_void_p(rowptr.ctypes.data), C.c_void_p(col.ctypes.data),
                  C.c_void_p(eid.ctypes.data))
        return rowptr, col, eid

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().tecgat_plan_destroy(h)
            except Exception:
                pass


class _on_device:
    """``torch.cuda.device(dev)`` only when the calling thread is on another device (the context manager costs ~5 us per
    use; autograd's worker threads normally already sit on the gradient's device)."""

    __slots__ = ("ctx",)

    def __init__(self, dev: torch.device):
        self.ctx = None if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


class _GATv2Function(torch.autograd.Function):
    """forward = ``tecgat_forward`` (projection + fused edge kernel); backward = ``tecgat_backward`` (fused edge backward +
    projection backward + one fixed-order finish of all six parameter gradients): five launches per training step.
    Saved for backward: x, xl, xr, y, stat (PyG saves 4-6 (S*E, H, C) tensors)."""

    @staticmethod
    def forward(ctx, x2d, wl, bl, wr, br, att, bias, plan: GraphPlan, S, H, Cc, slope, p, seed, seed_t, mode, dtype, impl,
                block=None, owner=None):
        # block = (B, L): x2d is the (B, L, N, F) tensor flattened; return the whole spatial block of tec_mollm.py:84-106,
        # z[b, n, l, :] = x[b, l, n, :] + y[b, l, n, :], instead of y
        dev = x2d.device
        R, F = x2d.shape
        HC = H * Cc
        st_dtype = torch.float32 if dtype == _lib.F32 else torch.bfloat16
        with _on_device(dev):
            stream = _stream(dev)
            xl = torch.empty((R, HC), device=dev, dtype=st_dtype)
            xr = torch.empty((R, HC), device=dev, dtype=st_dtype)
            y = torch.empty((R, HC), device=dev, dtype=torch.float32)
            stat = torch.empty((R, H), device=dev, dtype=torch.float32)
            _lib.call("tecgat_forward", plan.handle, _ptr(x2d), _ptr(wl), _ptr(bl), _ptr(wr), _ptr(br), _ptr(att), _ptr(bias),
                      _ptr(xl), _ptr(xr), _ptr(y), _ptr(stat), S, F, H, Cc, slope, p, seed, _ptr(seed_t), mode, dtype, impl, stream)
            out = y
            if block is not None:
                B, L = block
                out = torch.empty((B, R // S, L, HC), device=dev, dtype=torch.float32)
                _lib.call("tecgat_residual_permute_fwd", _ptr(x2d), _ptr(y), _ptr(out), B, L, R // S, HC, stream)
        ctx.save_for_backward(x2d, wl, wr, att, bias, xl, xr, y, stat)
        ctx.plan = plan
        ctx.cfg = (S, H, Cc, slope, p, seed, mode, dtype, impl)
        ctx.seed_t = seed_t
        ctx.block = block
        ctx.owner = owner
        return out

    @staticmethod
    def backward(ctx, gy):
        x2d, wl, wr, att, bias, xl, xr, y, stat = ctx.saved_tensors
        S, H, Cc, slope, p, seed, mode, dtype, impl = ctx.cfg
        plan: GraphPlan = ctx.plan
        dev = x2d.device
        R, F = x2d.shape
        HC = H * Cc
        gy = gy.contiguous()
        if gy.dtype != torch.float32:
            gy = gy.float()
        block = ctx.block
        if block is None and gy.data_ptr() % 16:  # bulk-TMA sources are 16-byte aligned
            gy = gy.clone()
        L = _lib.lib()
        with _on_device(dev):  # autograd worker threads do not always inherit the current device
            stream = _stream(dev)
            if block is not None:  # gy arrives as (B, N, L, HC): one transposition gives the gradient of y AND of the residual
                B, Lb = block
                g = torch.empty((R, HC), device=dev, dtype=torch.float32)
                _lib.call("tecgat_residual_permute_bwd", _ptr(gy), _ptr(g), B, Lb, R // S, HC, stream)
                gy = g
            dxl = torch.empty_like(xl)
            dxr = torch.empty_like(xr)
            need_dx = ctx.needs_input_grad[0]
            aligned = not (x2d.data_ptr() | xl.data_ptr() | xr.data_ptr() | dxl.data_ptr() | dxr.data_ptr() | gy.data_ptr()) & 15
            fused_ok = aligned and L.tecgat_backward_fused_supported(F, HC, impl) == 1
            # block: dx = g + dxl Wl + dxr Wr accumulates in place into g (nobody reads gy after the edge kernel)
            dx_acc = block is not None and need_dx and fused_ok
            dx = gy if dx_acc else (torch.empty_like(x2d) if need_dx else None)
            # parameter gradients: straight onto the owner's .grad storage when it asked for that (fused_grad_accumulation) and
            # every .grad exists -- replaces autograd's six accumulation kernels; otherwise fresh tensors returned to autograd
            owner = ctx.owner
            grads = None
            if fused_ok and owner is not None and owner.fused_grad_accumulation:
                params = (owner.lin_l.weight, owner.lin_l.bias, owner.lin_r.weight, owner.lin_r.bias, owner.att, owner.bias)
                gs = [q.grad for q in params]
                if all(t is not None and t.dtype == torch.float32 and t.is_contiguous() and t.device == dev for t in gs):
                    grads = gs
            acc = grads is not None
            if not acc:
                grads = [torch.empty_like(wl), torch.empty((HC,), device=dev, dtype=torch.float32), torch.empty_like(wr),
                         torch.empty((HC,), device=dev, dtype=torch.float32),
                         torch.empty((1, H, Cc), device=dev, dtype=torch.float32),
                         torch.empty((HC,), device=dev, dtype=torch.float32)]
            dwl, dbl, dwr, dbr, datt, dbias = grads
            ws = torch.empty((max(1, L.tecgat_backward_workspace(plan.handle, S, F, H, Cc, impl)),), device=dev, dtype=torch.uint8)
            _lib.call("tecgat_backward", plan.handle, _ptr(x2d), _ptr(wl), _ptr(wr), _ptr(att), _ptr(bias), _ptr(xl), _ptr(xr),
                      _ptr(y), _ptr(stat), _ptr(gy), _ptr(dxl), _ptr(dxr), _ptr(dx), int(dx_acc), _ptr(dwl), _ptr(dbl), _ptr(dwr),
                      _ptr(dbr), _ptr(datt), _ptr(dbias), int(acc), _ptr(ws), S, F, H, Cc, slope, p, seed, _ptr(ctx.seed_t), mode,
                      dtype, impl, stream)
            if block is not None and need_dx and not dx_acc:
                dx += gy
        if acc:
            return (dx,) + (None,) * 19
        return (dx, dwl, dbl, dwr, dbr, datt, dbias) + (None,) * 13


class _GATv2PairsFunction(torch.autograd.Function):
    """``H`` heads as ``H / 2`` independent two-head layers on parameter slices (``GATv2Conv._split_head_pairs``) in ONE autograd
    node: every pair's edge kernel also writes its columns of the ``(rows, H*C)`` output (``tecgat_forward_into``: no concat
    pass), the backward accumulates ``dx`` across the pairs inside the projection's epilogue and writes each pair's parameter
    gradients straight into its slice of the full-size gradient tensors."""

    @staticmethod
    def forward(ctx, x2d, wl, bl, wr, br, att, bias, plan: GraphPlan, S, H, Cc, slope, p, seed_ts, mode, dtype, impl):
        dev = x2d.device
        R, F = x2d.shape
        HC, C2 = H * Cc, 2 * Cc
        st_dtype = torch.float32 if dtype == _lib.F32 else torch.bfloat16
        out = torch.empty((R, HC), device=dev, dtype=torch.float32)
        saved = []
        with _on_device(dev):
            stream = _stream(dev)
            for g in range(H // 2):
                xl = torch.empty((R, C2), device=dev, dtype=st_dtype)
                xr = torch.empty((R, C2), device=dev, dtype=st_dtype)
                y = torch.empty((R, C2), device=dev, dtype=torch.float32)   # dense copy: the backward's bulk-TMA source
                stat = torch.empty((R, 2), device=dev, dtype=torch.float32)
                o = g * C2
                _lib.call("tecgat_forward_into", plan.handle, _ptr(x2d), C.c_void_p(wl.data_ptr() + 4 * o * F), C.c_void_p(bl.data_ptr() + 4 * o),
                          C.c_void_p(wr.data_ptr() + 4 * o * F), C.c_void_p(br.data_ptr() + 4 * o), C.c_void_p(att.data_ptr() + 4 * o),
                          C.c_void_p(bias.data_ptr() + 4 * o), _ptr(xl), _ptr(xr), _ptr(y), _ptr(stat), S, F, 2, Cc, slope, p, 0,
                          _ptr(seed_ts[g]) if seed_ts else None, mode, dtype, impl, C.c_void_p(out.data_ptr() + 4 * o), HC, stream)
                saved += [xl, xr, y, stat]
        ctx.save_for_backward(x2d, wl, wr, att, bias, *saved)
        ctx.plan = plan
        ctx.cfg = (S, H, Cc, slope, p, mode, dtype, impl)
        ctx.seed_ts = seed_ts
        return out

    @staticmethod
    def backward(ctx, gy):
        x2d, wl, wr, att, bias, *saved = ctx.saved_tensors
        S, H, Cc, slope, p, mode, dtype, impl = ctx.cfg
        plan: GraphPlan = ctx.plan
        dev = x2d.device
        R, F = x2d.shape
        HC, C2 = H * Cc, 2 * Cc
        if gy.dtype != torch.float32:
            gy = gy.float()
        L = _lib.lib()
        need_dx = ctx.needs_input_grad[0]
        fused_ok = L.tecgat_backward_fused_supported(F, C2, impl) == 1 and x2d.data_ptr() % 16 == 0
        dwl = torch.empty((HC, F), device=dev, dtype=torch.float32)
        dwr = torch.empty((HC, F), device=dev, dtype=torch.float32)
        dbl = torch.empty((HC,), device=dev, dtype=torch.float32)
        dbr = torch.empty((HC,), device=dev, dtype=torch.float32)
        datt = torch.empty((1, H, Cc), device=dev, dtype=torch.float32)
        dbias = torch.empty((HC,), device=dev, dtype=torch.float32)
        dx = torch.empty_like(x2d) if need_dx else None
        with _on_device(dev):  # autograd worker threads do not always inherit the current device
            stream = _stream(dev)
            ws = torch.empty((max(1, L.tecgat_backward_workspace(plan.handle, S, F, 2, Cc, impl)),), device=dev, dtype=torch.uint8)
            for g in range(H // 2):
                xl, xr, y, stat = saved[4 * g:4 * g + 4]
                o = g * C2
                g_g = gy[:, o:o + C2].contiguous()   # the edge kernel stages windows of dense rows
                dxl = torch.empty_like(xl)
                dxr = torch.empty_like(xr)
                # dx: pair 0 writes, later pairs accumulate in the projection's epilogue (or through a temporary + add)
                acc = int(need_dx and g > 0 and fused_ok)
                dx_g = dx if (g == 0 or acc or not need_dx) else torch.empty_like(x2d)
                if fused_ok:  # the gradient finish writes plain floats: straight into the slices
                    tg = [C.c_void_p(dwl.data_ptr() + 4 * o * F), C.c_void_p(dbl.data_ptr() + 4 * o), C.c_void_p(dwr.data_ptr() + 4 * o * F),
                          C.c_void_p(dbr.data_ptr() + 4 * o), C.c_void_p(datt.data_ptr() + 4 * o), C.c_void_p(dbias.data_ptr() + 4 * o)]
                    tmp = None
                else:
                    tmp = [torch.empty((C2, F), device=dev), torch.empty((C2,), device=dev), torch.empty((C2, F), device=dev),
                           torch.empty((C2,), device=dev), torch.empty((C2,), device=dev), torch.empty((C2,), device=dev)]
                    tg = [_ptr(t) for t in tmp]
                _lib.call("tecgat_backward", plan.handle, _ptr(x2d), C.c_void_p(wl.data_ptr() + 4 * o * F), C.c_void_p(wr.data_ptr() + 4 * o * F),
                          C.c_void_p(att.data_ptr() + 4 * o), C.c_void_p(bias.data_ptr() + 4 * o), _ptr(xl), _ptr(xr), _ptr(y), _ptr(stat),
                          _ptr(g_g), _ptr(dxl), _ptr(dxr), _ptr(dx_g), acc, tg[0], tg[1], tg[2], tg[3], tg[4], tg[5], 0, _ptr(ws),
                          S, F, 2, Cc, slope, p, 0, _ptr(ctx.seed_ts[g]) if ctx.seed_ts else None, mode, dtype, impl, stream)
                if tmp is not None:
                    dwl[o:o + C2].copy_(tmp[0]); dbl[o:o + C2].copy_(tmp[1]); dwr[o:o + C2].copy_(tmp[2]); dbr[o:o + C2].copy_(tmp[3])
                    datt.view(-1)[o:o + C2].copy_(tmp[4]); dbias[o:o + C2].copy_(tmp[5])
                if need_dx and dx_g is not dx:
                    dx += dx_g
        return (dx, dwl, dbl, dwr, dbr, datt, dbias) + (None,) * 10


class _Linear(nn.Module):
    """Parameter holder with ``torch_geometric.nn.dense.Linear``'s names (``weight``, ``bias``) and
    init (glorot weight, U(+-1/sqrt(in)) bias)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.weight.size(-2) + self.weight.size(-1)))
        with torch.no_grad():
            self.weight.uniform_(-a, a)
            b = 1.0 / math.sqrt(self.in_channels)
            self.bias.uniform_(-b, b)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, bias=True"


class GATv2Conv(nn.Module):
    """``GATv2Conv(in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0,
    add_self_loops=True, edge_dim=None, fill_value='mean', bias=True, share_weights=False)``.

    Only the configuration the reference uses is implemented in CUDA (``concat=True``,
    ``add_self_loops=True``, ``edge_dim=None``, ``bias=True``, ``share_weights=False``); any other
    value raises ``NotImplementedError`` instead of silently doing something else.
    """

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True,
                 edge_dim: Optional[int] = None, fill_value="mean", bias: bool = True,
                 share_weights: bool = False, **kwargs):
        super().__init__()
        if kwargs:
            raise NotImplementedError(f"GATv2Conv: unsupported arguments {sorted(kwargs)}")
        if not isinstance(in_channels, int):
            raise NotImplementedError("GATv2Conv: bipartite (tuple) in_channels are not supported")
        if not concat or not add_self_loops or edge_dim is not None or not bias or share_weights:
            raise NotImplementedError(
                "GATv2Conv (tec_mollm_b200): only concat=True, add_self_loops=True, edge_dim=None, bias=True, "
                "share_weights=False are implemented (the reference's configuration, modules.py:329-336)")
        if not 0.0 <= dropout < 1.0:
            raise ValueError(f"dropout must be in [0, 1); got {dropout}")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, float(negative_slope), float(dropout)
        self.add_self_loops, self.edge_dim, self.fill_value, self.share_weights = add_self_loops, edge_dim, fill_value, False
        self.lin_l = _Linear(in_channels, heads * out_channels)
        self.lin_r = _Linear(in_channels, heads * out_channels)
        self.att = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.empty(heads * out_channels))
        if heads < 1 or heads > 32:
            raise NotImplementedError(f"GATv2Conv (tec_mollm_b200): heads={heads} outside the compiled range 1..32")
        if out_channels not in _COMPILED_CHANNELS:
            raise NotImplementedError(f"GATv2Conv (tec_mollm_b200): out_channels={out_channels} is not among the compiled "
                                      f"channel counts {_COMPILED_CHANNELS}")
        self._plans = {}
        # opt-in: backward adds the parameter gradients straight onto existing .grad storage and returns None for them
        # (no autograd accumulation kernels).  Leave it off under DistributedDataParallel, whose hooks need autograd's own
        # accumulation; tec_mollm_b200.dist.FlatGradAllReduce switches it on.
        self.fused_grad_accumulation = False
        self._last_seed = None
        self._rng_state = None  # device int64[2] = {seed, counter}: the attention-dropout stream (see _dropout_state)
        self._tile_nodes = tile_nodes_for(heads)
        self._tile_nodes_bwd = tile_nodes_for(heads, backward=True, out_channels=out_channels)
        hp = 1
        while hp < heads:
            hp *= 2
        if 2 * heads * out_channels > self._tile_nodes_bwd * hp:  # the backward's gradient flush maps 2*H*C columns onto its threads
            raise NotImplementedError(f"GATv2Conv (tec_mollm_b200): heads*out_channels = {heads * out_channels} exceeds what the "
                                      f"backward kernel's {self._tile_nodes_bwd * hp} consumer threads reduce")
        self.reset_parameters()

    def __getstate__(self):  # graph plans hold device handles: never pickled / deep-copied
        state = self.__dict__.copy()
        state["_plans"] = {}
        state["_rng_state"] = None
        state["_last_seed"] = None
        return state

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()
        a = math.sqrt(6.0 / (self.att.size(-2) + self.att.size(-1)))
        with torch.no_grad():
            self.att.uniform_(-a, a)
            self.bias.zero_()

    # ---- plan cache: the reference passes the same edge_index tensor every step (train.py:292-294, 388) ----
    def plan_for(self, edge_index: torch.Tensor, num_nodes: int, tiles: Optional[Tuple[int, int]] = None) -> GraphPlan:
        tiles = tiles or (self._tile_nodes, self._tile_nodes_bwd)
        key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), int(num_nodes), edge_index.device, tiles)
        hit = self._plans.get(key)
        if hit is not None:
            return hit[0]
        plan = GraphPlan(edge_index, num_nodes, tiles[0], tiles[1])
        if len(self._plans) >= 8:
            self._plans.pop(next(iter(self._plans)))
        self._plans[key] = (plan, edge_index)  # keep the tensor alive so its data_ptr cannot be recycled
        return plan

    def _dropout_state(self, device) -> torch.Tensor:
        """Device-resident state {seed, counter} of the attention-dropout stream.  The seed is drawn ONCE from torch's CPU
        generator (reproducible under ``torch.manual_seed``, no device sync); every training forward then advances the
        counter on the device (``tecgat_seed_advance``), so nothing on the host is consumed per step and a captured CUDA graph
        draws a fresh mask on every replay."""
        st = self._rng_state
        if st is None or st.device != device:
            seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
            st = torch.tensor([seed, 0], dtype=torch.int64, device=device)
            self._rng_state = st
        return st

    def forward_snapshots(self, x: torch.Tensor, edge_index: torch.Tensor, snapshots: int, num_nodes: int,
                          snapshot_mode: str = "shared", seed: Optional[int] = None, block=None) -> torch.Tensor:
        """``x``: (snapshots*num_nodes, in_channels); the one-graph ``edge_index`` (indices < num_nodes) is
        applied per ``snapshot_mode``.  Returns (snapshots*num_nodes, heads*out_channels) fp32; with ``block=(B, L)``
        (snapshots = B*L laid out (B, L)) the whole spatial block ``(B, num_nodes, L, heads*out_channels)``."""
        if not x.is_cuda:
            raise RuntimeError("tec_mollm_b200.GATv2Conv: input must be a CUDA tensor (there is no CPU fallback)")
        if x.dim() != 2 or x.size(1) != self.in_channels or x.size(0) != snapshots * num_nodes:
            raise ValueError(f"x must be ({snapshots * num_nodes}, {self.in_channels}); got {tuple(x.shape)}")
        if snapshot_mode not in ("shared", "literal"):
            raise ValueError(f"snapshot_mode must be 'shared' or 'literal'; got {snapshot_mode!r}")
        dtype = _lib.F32
        if torch.is_autocast_enabled("cuda"):
            ac = torch.get_autocast_dtype("cuda")
            if ac != torch.bfloat16:
                raise NotImplementedError(f"autocast dtype {ac} unsupported (bfloat16 only, train.py:68)")
            dtype = _lib.BF16
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        p = self.dropout if self.training else 0.0
        mode = _lib.MODE_SHARED if snapshot_mode == "shared" else _lib.MODE_LITERAL
        if block is None and seed is None and self._split_head_pairs():
            return self._forward_head_pairs(x, edge_index, int(snapshots), num_nodes, float(p), mode, dtype)
        plan = self.plan_for(edge_index, num_nodes)
        seed_t = None
        if p > 0.0 and seed is None:  # an explicit ``seed`` (tests: reproduce the mask on the host) is passed by value
            seed_t = torch.empty(1, dtype=torch.int64, device=x.device)
            with _on_device(x.device):
                _lib.call("tecgat_seed_advance", _ptr(self._dropout_state(x.device)), _ptr(seed_t), _stream(x.device))
            self._last_seed = seed_t  # tests rebuild the kernels' mask on the host from it
        f32 = lambda t: t if t.dtype == torch.float32 else t.float()
        return _GATv2Function.apply(
            x, f32(self.lin_l.weight), f32(self.lin_l.bias), f32(self.lin_r.weight), f32(self.lin_r.bias),
            f32(self.att), f32(self.bias), plan, int(snapshots), self.heads, self.out_channels,
            self.negative_slope, float(p), int(seed or 0), seed_t, mode, dtype, _proj_impl(), block, self)

    # ---- more than two heads: the heads of a GATv2 layer only meet in the concatenation, so H heads = H / 2 independent
    # two-head layers on slices of the parameters.  Rows of 2 * C channels fit the fully staged edge kernels where H * C
    # channels do not (BASELINE config 4: H = 4, C = 11 -- the backward's three 176-byte row windows exceed shared memory and
    # it gathers rows from L2 instead), and the two-head shapes are the ones compiled with fixed tile sizes.
    def _split_head_pairs(self) -> bool:
        knob = os.environ.get("TECGAT_HEAD_SPLIT")  # tests: "0" keeps the one-launch kernels for any head count
        ok = self.heads > 2 and self.heads % 2 == 0 and self.out_channels in (5, 11)  # the compiled two-head shapes
        return ok and knob != "0"

    def _forward_head_pairs(self, x, edge_index, snapshots, num_nodes, p, mode, dtype):
        plan = self.plan_for(edge_index, num_nodes, (tile_nodes_for(2), tile_nodes_for(2, backward=True, out_channels=self.out_channels)))
        f32c = lambda t: (t if t.dtype == torch.float32 else t.float()).contiguous()
        seed_ts = None
        if p > 0.0:  # every pair draws its own seed: the kernels number the (snapshot, head) streams inside one call
            seed_ts = []
            with _on_device(x.device):
                for _ in range(self.heads // 2):
                    seed_t = torch.empty(1, dtype=torch.int64, device=x.device)
                    _lib.call("tecgat_seed_advance", _ptr(self._dropout_state(x.device)), _ptr(seed_t), _stream(x.device))
                    seed_ts.append(seed_t)
        return _GATv2PairsFunction.apply(x, f32c(self.lin_l.weight), f32c(self.lin_l.bias), f32c(self.lin_r.weight), f32c(self.lin_r.bias),
                                         f32c(self.att), f32c(self.bias), plan, snapshots, self.heads, self.out_channels,
                                         self.negative_slope, p, seed_ts, mode, dtype, _proj_impl())

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, edge_attr=None, return_attention_weights=None):
        """PyG semantics for one 2-D input: ``num_nodes = x.size(0)`` rows, edges as given (so calling it the way
        modules.py:356 does reproduces the reference literally)."""
        if edge_attr is not None or return_attention_weights is not None:
            raise NotImplementedError("GATv2Conv (tec_mollm_b200): edge_attr / return_attention_weights are not supported")
        return self.forward_snapshots(x, edge_index, 1, x.size(0), "literal")

    def __repr__(self):
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})"
