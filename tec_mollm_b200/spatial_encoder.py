"""Drop-in ``SpatialEncoder`` (reference: ``src/model/modules.py:315-359``).

Same constructor ``(in_channels, out_channels, heads=2, dropout=0.1)``, same attributes (``.gat_conv`` owning
the PyG-named parameters, ``.output_channels``), same ``forward(x, edge_index, edge_weight=None)`` taking
``x`` of shape ``(B*L, N, C_in)`` and returning ``(B*L, N, heads*out_channels)`` (extra leading dimensions, e.g. the
``(B, L, N, C_in)`` tensor the reference permutes from, are accepted and kept).

One extra keyword, ``snapshot_mode``: ``"shared"`` (default) applies the one-graph ``edge_index`` to every one of the
``B*L`` snapshots -- the semantics the reference intends (tec_mollm.py:86-88) and BASELINE.json measures;
``"literal"`` reproduces the call exactly as written (modules.py:353-356: the flattened ``(B*L*N, C)`` input meets an
``edge_index`` whose indices are all ``< N``, so only snapshot 0 sees edges).
"""
from __future__ import annotations

import logging

import torch
from torch import nn

from .gatv2 import GATv2Conv


class SpatialEncoder(nn.Module):
    """Captures spatial dependencies using a GATv2 layer (hand-written sm_100a kernels)."""

    _warned_default_mode = False

    def __init__(self, in_channels: int, out_channels: int, heads: int = 2, dropout: float = 0.1,
                 snapshot_mode: str = None):
        super().__init__()
        if snapshot_mode is None:
            snapshot_mode = "shared"
            if not SpatialEncoder._warned_default_mode:  # once per process: the default is NOT what the reference computes
                SpatialEncoder._warned_default_mode = True
                logging.warning(
                    "tec_mollm_b200.SpatialEncoder: snapshot_mode defaults to 'shared' (the graph is applied to every one of the "
                    "B*L snapshots, the semantics tec_mollm.py:86-88 intends). The reference's own call (modules.py:353-356) gives "
                    "edges to snapshot 0 only; pass snapshot_mode='literal' to reproduce a reference-trained checkpoint's outputs.")
        if snapshot_mode not in ("shared", "literal"):
            raise ValueError(f"snapshot_mode must be 'shared' or 'literal'; got {snapshot_mode!r}")
        self.gat_conv = GATv2Conv(in_channels, out_channels, heads=heads, dropout=dropout, concat=True,
                                  add_self_loops=True)
        self.output_channels = out_channels * heads
        self.snapshot_mode = snapshot_mode
        logging.info(f"SpatialEncoder initialized with out_channels={self.output_channels}")

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, edge_weight: torch.Tensor = None) -> torch.Tensor:
        # edge_weight is accepted and ignored, exactly like the reference (modules.py:347,355)
        if x.dim() < 3:
            raise ValueError(f"x must be (B*L, N, C_in) or (..., N, C_in); got {tuple(x.shape)}")
        # Snapshots are independent given the shared graph, so ANY leading layout works: the reference's (L*B, N, C)
        # (tec_mollm.py:84) or the (B, L, N, C) tensor it is permuted from -- passing that one directly saves the caller's
        # permute copy; the output keeps the leading dimensions.
        lead, (num_nodes, in_channels) = x.shape[:-2], x.shape[-2:]
        snapshots = 1
        for d in lead:
            snapshots *= int(d)
        if len(lead) > 1 and self.snapshot_mode == "literal":
            raise ValueError("snapshot_mode='literal' reproduces the reference's flattened call: pass (B*L, N, C_in)")
        x2d = x.reshape(-1, in_channels)
        y = self.gat_conv.forward_snapshots(x2d, edge_index, snapshots, num_nodes, self.snapshot_mode)
        return y.view(*lead, num_nodes, self.output_channels)

    def forward_block(self, x: torch.Tensor, edge_index: torch.Tensor, edge_weight: torch.Tensor = None) -> torch.Tensor:
        """The reference's whole spatial block (tec_mollm.py:84-106) in one call: ``x`` is the embedded ``(B, L, N, C)``
        tensor; returns ``(B*N, L, C)``, the TemporalEncoder's input, equal to

            x_for_gnn = x.permute(1, 0, 2, 3).reshape(-1, N, C)
            x_spatial = x_for_gnn + spatial_encoder(x_for_gnn, edge_index, edge_weight)
            x_spatial.view(L, B, N, C).permute(1, 2, 0, 3).reshape(-1, L, C)

        without the permute copy of :84 (the encoder reads the snapshots where they lie) and with the residual add and
        the permute of :94-100 fused into one pass; in backward one transposition yields the gradient of both branches and
        the projection's backward accumulates ``dx`` into it in place.  Dropout draws differ from the three-line form only in
        which snapshot gets which counter."""
        if x.dim() != 4:
            raise ValueError(f"x must be (B, L, N, C); got {tuple(x.shape)}")
        if x.size(-1) != self.output_channels:
            raise ValueError(f"the residual needs in_channels == heads*out_channels; got {x.size(-1)} vs {self.output_channels}")
        if not x.is_cuda:
            raise RuntimeError("SpatialEncoder.forward_block needs CUDA tensors (there is no CPU path)")
        if self.snapshot_mode != "shared":
            raise ValueError("forward_block implements the shared-graph semantics only")
        B, L, N, Cc = x.shape
        z = self.gat_conv.forward_snapshots(x.reshape(-1, Cc), edge_index, B * L, N, "shared", block=(B, L))
        return z.view(B * N, L, Cc)

    def graphed(self, sample_x: torch.Tensor, edge_index: torch.Tensor, block: bool = False):
        """Opt-in CUDA-graph mode for small batches (the reference trains at B = 2, train.py:182, where the eager step is
        host-launch bound): returns a callable ``f(x) -> y`` whose forward and backward each replay one captured CUDA graph
        (``torch.cuda.make_graphed_callables``).  ``x`` must keep ``sample_x``'s shape, dtype and ``requires_grad``;
        ``edge_index`` is bound at capture.  Parameter gradients reach ``.grad`` as usual; attention dropout stays correct
        under replay because its seed lives on the device (``tecgat_seed_advance``).  ``block=True`` graphs
        :meth:`forward_block` instead of :meth:`forward`."""
        if not sample_x.is_cuda:
            raise RuntimeError("SpatialEncoder.graphed needs CUDA tensors (there is no CPU path)")
        if self.gat_conv.fused_grad_accumulation:
            raise RuntimeError("graphed(): fused_grad_accumulation writes onto .grad storage that a captured graph would pin; "
                               "switch it off (the graphed callable returns gradients through autograd)")
        enc = self

        class _Bound(nn.Module):
            def __init__(self):
                super().__init__()
                self.enc = enc

            def forward(self, x):
                return enc.forward_block(x, edge_index) if block else enc(x, edge_index)

        with torch.no_grad():  # the plan is built outside capture (plan creation synchronises once)
            num_nodes = sample_x.shape[-2]
            self.gat_conv.plan_for(edge_index, num_nodes)
            if self.training and self.gat_conv.dropout > 0:
                self.gat_conv._dropout_state(sample_x.device)
        return torch.cuda.make_graphed_callables(_Bound(), (sample_x,))
