"""Drop-in ``SpatioTemporalEmbedding`` (reference: ``src/model/modules.py:211-266``) -- the producer side of the spatial block
(SURVEY.md 8f N1).

Same constructor ``(d_emb, num_nodes=2911, num_years=13)``, same five ``nn.Embedding`` members (``node_embedding``,
``tod_embedding``, ``doy_embedding``, ``year_embedding``, ``season_embedding`` -> reference checkpoints load ``strict=True``),
same ``forward(x, time_features) -> (B, L, N, C_in + d_emb)``, bit-identical values.  The five gathers, four adds and the concat
run as ONE kernel (``tecgat_embed_fwd``), and the backward reduces the gradient into the five tables without atomics
(``tecgat_embed_bwd``).  What this buys end to end: only the raw ``(B, L, N, C_in)`` features and the ``(B, L, 4)`` time
indices cross the host link (train.py:58-65), 6/22 of the bytes of the embedded tensor.

``time_features`` may be the reference's ``(B, L, N, 4)`` tensor (it is an ``expand`` over the nodes, train.py:65) or the
``(B, L, 4)`` tensor it is expanded from; values are truncated like the reference's ``.long()`` (modules.py:250-253).
"""
from __future__ import annotations

import ctypes as C
import logging

import torch
from torch import nn

from . import _lib
from .gatv2 import _on_device, _ptr, _stream


def snapshot_time_indices(time_features: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """``(B, L, N, 4)`` (uniform over N) or ``(B, L, 4)`` -> contiguous int32 ``(B*L, 4)`` on the same device."""
    tf = time_features
    if tf.dim() == 4:
        if tf.size(2) not in (1, num_nodes) or tf.size(3) != 4:
            raise ValueError(f"time_features must be (B, L, N, 4) or (B, L, 4); got {tuple(tf.shape)}")
        if tf.size(2) > 1 and tf.stride(2) != 0:
            # a materialised per-node tensor: the fused kernel takes one index row per snapshot, so make sure it IS uniform
            if not bool((tf == tf[:, :, :1, :]).all()):
                raise NotImplementedError("tec_mollm_b200.SpatioTemporalEmbedding: time_features differ across nodes; the fused "
                                          "path implements the reference's data flow (per-(batch, step) features, train.py:64-65)")
        tf = tf[:, :, 0, :]
    elif tf.dim() != 3 or tf.size(2) != 4:
        raise ValueError(f"time_features must be (B, L, N, 4) or (B, L, 4); got {tuple(tf.shape)}")
    return tf.reshape(-1, 4).to(torch.int32).contiguous()  # float -> int truncates toward zero, like .long()


class _EmbedFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, tf, node, tod, doy, year, season):
        dev = x.device
        B, L, N, Cr = x.shape
        De = node.size(1)
        out = torch.empty((B, L, N, Cr + De), device=dev, dtype=torch.float32)
        with _on_device(dev):
            _lib.call("tecgat_embed_fwd", _ptr(x), _ptr(tf), _ptr(node), _ptr(tod), _ptr(doy), _ptr(year), _ptr(season), _ptr(out),
                      B * L, N, Cr, De, tod.size(0), doy.size(0), year.size(0), season.size(0), _stream(dev))
        ctx.save_for_backward(tf)
        ctx.dims = (B * L, N, Cr, De, node.size(0), tod.size(0), doy.size(0), year.size(0), season.size(0))
        return out

    @staticmethod
    def backward(ctx, ge):
        (tf,) = ctx.saved_tensors
        S, N, Cr, De, n_node, n_tod, n_doy, n_year, n_season = ctx.dims
        dev = ge.device
        ge = ge.contiguous()
        if ge.dtype != torch.float32:
            ge = ge.float()
        grads = [None] * 5
        if any(ctx.needs_input_grad[2:]):
            dnode = torch.empty((N, De), device=dev, dtype=torch.float32)  # `node` arrives sliced to the N rows in use
            tabs = [torch.empty((r, De), device=dev, dtype=torch.float32) for r in (n_tod, n_doy, n_year, n_season)]
            ws = torch.empty((max(1, _lib.lib().tecgat_embed_bwd_workspace(S, N, De)),), device=dev, dtype=torch.uint8)
            with _on_device(dev):
                _lib.call("tecgat_embed_bwd", _ptr(ge), _ptr(tf), _ptr(dnode), _ptr(tabs[0]), _ptr(tabs[1]), _ptr(tabs[2]),
                          _ptr(tabs[3]), _ptr(ws), S, N, Cr, De, n_tod, n_doy, n_year, n_season, 0, _stream(dev))
            grads = [dnode] + tabs
        dx = ge[..., :Cr].contiguous() if ctx.needs_input_grad[0] else None
        return (dx, None) + tuple(grads)


class SpatioTemporalEmbedding(nn.Module):
    """Creates learnable embeddings for nodes and multiple time features (one fused sm_100a kernel each way)."""

    def __init__(self, d_emb: int, num_nodes: int = 2911, num_years: int = 13):
        super().__init__()
        self.d_emb = d_emb
        self.node_embedding = nn.Embedding(num_embeddings=num_nodes, embedding_dim=d_emb)
        self.tod_embedding = nn.Embedding(num_embeddings=12, embedding_dim=d_emb)
        self.doy_embedding = nn.Embedding(num_embeddings=366, embedding_dim=d_emb)
        self.year_embedding = nn.Embedding(num_embeddings=num_years, embedding_dim=d_emb)
        self.season_embedding = nn.Embedding(num_embeddings=4, embedding_dim=d_emb)
        logging.info("SpatioTemporalEmbedding module initialized with year and season embeddings.")

    def forward(self, x: torch.Tensor, time_features: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4:
            raise ValueError(f"x must be (B, L, N, C_in); got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("tec_mollm_b200.SpatioTemporalEmbedding needs CUDA tensors (there is no CPU path)")
        B, L, N, _ = x.shape
        if N != self.node_embedding.num_embeddings:
            # the reference embeds arange(num_nodes of x) (modules.py:245-246): fewer nodes than the table is legal there
            if N > self.node_embedding.num_embeddings:
                raise ValueError(f"x has {N} nodes but the node table only {self.node_embedding.num_embeddings}")
        if self.d_emb != 16 and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("tec_mollm_b200.SpatioTemporalEmbedding: the backward kernel is built for d_emb = 16")
        tf = snapshot_time_indices(time_features.to(x.device), N)
        xc = x.contiguous()
        if xc.dtype != torch.float32:
            xc = xc.float()
        f32 = lambda t: t if t.dtype == torch.float32 else t.float()
        node_w = f32(self.node_embedding.weight)
        if N != node_w.size(0):
            node_w = node_w[:N]
        return _EmbedFunction.apply(xc, tf, node_w, f32(self.tod_embedding.weight), f32(self.doy_embedding.weight),
                                    f32(self.year_embedding.weight), f32(self.season_embedding.weight))
