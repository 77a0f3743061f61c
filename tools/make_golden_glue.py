#!/usr/bin/env python
"""Golden fixtures pinned to the REFERENCE'S OWN model code (run in the build container only; /root/reference is read-only
and absent on the GPU box).

``src/model/modules.py`` imports ``peft`` and ``torch_geometric`` at module top (both absent here), so empty stub modules go
into ``sys.modules`` first -- exactly like the ``h5py`` stub of tools/make_golden.py.  ``torch_geometric.nn.GATv2Conv`` is
stubbed by an ``nn.Module`` that runs the oracle restatement of PyG's operator (oracle/gatv2_oracle.py), so everything AROUND
the operator is the reference's unmodified code:

  glue_embedding.npz      SpatioTemporalEmbedding.forward (modules.py:230-264) + autograd table gradients
  glue_spatial_block.npz  TEC_MoLLM.forward lines 75-106 (tec_mollm.py): embedding -> permute -> SpatialEncoder.forward
                          (modules.py:340-359, the flattened "literal" call) -> residual -> permute, forward and backward, with
                          the temporal encoder / LLM / head replaced by shape-preserving stand-ins that record x_temporal
  temporal_block.npz      Multi_Scale_Conv_Block x2 (MultiScaleConvEmbedder, modules.py:13-91) forward + all gradients

    python tools/make_golden_glue.py
"""
import os
import sys
import types

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import gatv2_oracle as gat  # noqa: E402


class OracleGATv2Conv(nn.Module):
    """PyG's GATv2Conv surface over the oracle restatement (parameter names and init order as upstream)."""

    def __init__(self, in_channels, out_channels, heads=1, dropout=0.0, concat=True, add_self_loops=True, **kw):
        super().__init__()
        assert concat and add_self_loops and not kw
        self.heads, self.out_channels, self.dropout = heads, out_channels, dropout
        p = gat.init_params(in_channels, out_channels, heads, seed=123, dtype=torch.float64)
        self.lin_l = nn.Linear(in_channels, heads * out_channels).double()
        self.lin_r = nn.Linear(in_channels, heads * out_channels).double()
        with torch.no_grad():
            self.lin_l.weight.copy_(p["lin_l.weight"]); self.lin_l.bias.copy_(p["lin_l.bias"])
            self.lin_r.weight.copy_(p["lin_r.weight"]); self.lin_r.bias.copy_(p["lin_r.bias"])
        self.att = nn.Parameter(p["att"].clone())
        self.bias = nn.Parameter(torch.randn(heads * out_channels, dtype=torch.float64, generator=torch.Generator().manual_seed(5)) * 0.1)

    def forward(self, x, edge_index):
        assert not self.training or self.dropout == 0.0
        params = {"lin_l.weight": self.lin_l.weight, "lin_l.bias": self.lin_l.bias, "lin_r.weight": self.lin_r.weight,
                  "lin_r.bias": self.lin_r.bias, "att": self.att, "bias": self.bias}
        return gat.gatv2_forward(x, edge_index, params, self.heads, self.out_channels)


def import_reference():
    for name in ("peft", "torch_geometric", "torch_geometric.nn", "h5py"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["peft"].get_peft_model = lambda *a, **k: None
    sys.modules["peft"].LoraConfig = object
    sys.modules["torch_geometric.nn"].GATv2Conv = OracleGATv2Conv
    sys.modules["torch_geometric"].nn = sys.modules["torch_geometric.nn"]
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import src.model.modules as M
    import src.model.tec_mollm as TM
    return M, TM


def main():
    M, TM = import_reference()
    os.makedirs(GOLD, exist_ok=True)
    gen = torch.Generator().manual_seed(2024)

    # ---- 1. SpatioTemporalEmbedding --------------------------------------------------------------------------------
    B, L, N, Cr, De = 2, 3, 50, 6, 16
    torch.manual_seed(11)
    emb = M.SpatioTemporalEmbedding(De, num_nodes=N, num_years=13)           # fp32, like the reference
    x = torch.randn(B, L, N, Cr, generator=gen)
    tf = torch.stack([torch.randint(0, 12, (B, L), generator=gen), torch.randint(0, 366, (B, L), generator=gen),
                      torch.randint(0, 13, (B, L), generator=gen), torch.randint(0, 4, (B, L), generator=gen)], dim=-1).float()
    tf[0, 0] = tf[1, 2]  # two snapshots sharing every index: table rows that accumulate more than one snapshot
    out = emb(x, tf.unsqueeze(-2).expand(B, L, N, 4))                         # train.py:65
    g = torch.randn(out.shape, generator=gen)
    out.backward(g)
    np.savez_compressed(
        os.path.join(GOLD, "glue_embedding.npz"), x=x.numpy(), tf=tf.numpy(), out=out.detach().numpy(), g=g.numpy(),
        **{f"w_{k}": v.detach().numpy() for k, v in emb.state_dict().items()},
        **{f"g_{k}": v.grad.numpy() for k, v in emb.named_parameters()})

    # ---- 2. TEC_MoLLM.forward lines 75-106 with the reference's SpatialEncoder around the oracle operator ----------------
    B, L, N = 2, 4, 63
    gs = np.load(os.path.join(GOLD, "graph_small150.npz"))
    ei = torch.from_numpy(gs["edge_index"])
    model = object.__new__(TM.TEC_MoLLM)
    nn.Module.__init__(model)
    torch.manual_seed(12)
    model.spatio_temporal_embedding = M.SpatioTemporalEmbedding(De, num_nodes=N, num_years=13).double()
    model.spatial_encoder = M.SpatialEncoder(in_channels=Cr + De, out_channels=11, heads=2).double().eval()
    seen = {}

    class Rec(nn.Module):  # stands in for TemporalEncoder: records x_temporal, keeps it differentiable
        def forward(self, xt):
            seen["x_temporal"] = xt
            return xt

    class Llm(nn.Module):
        def forward(self, inputs_embeds, attention_mask):
            return inputs_embeds

    class Head(nn.Module):
        def forward(self, h):
            return h.sum(-1)

    model.temporal_encoder, model.llm_backbone, model.prediction_head = Rec(), Llm(), Head()
    model.eval()
    x = torch.randn(B, L, N, Cr, generator=gen, dtype=torch.float64).requires_grad_(True)
    tf = torch.stack([torch.randint(0, 12, (B, L), generator=gen), torch.randint(0, 366, (B, L), generator=gen),
                      torch.randint(0, 13, (B, L), generator=gen), torch.randint(0, 4, (B, L), generator=gen)], dim=-1).double()
    final = model(x, tf.unsqueeze(-2).expand(B, L, N, 4), ei, None)             # the reference's forward, unmodified
    xt = seen["x_temporal"]
    gz = torch.randn(xt.shape, generator=gen, dtype=torch.float64)
    xt.backward(gz)
    named = dict(model.named_parameters())
    np.savez_compressed(
        os.path.join(GOLD, "glue_spatial_block.npz"), x=x.detach().numpy(), tf=tf.numpy(), edge_index=ei.numpy(),
        x_temporal=xt.detach().numpy(), gz=gz.numpy(), g_x=x.grad.numpy(), final_shape=np.array(final.shape),
        **{f"w_{k}": v.detach().numpy() for k, v in named.items()},
        **{f"g_{k}": v.grad.numpy() for k, v in named.items()})

    # ---- 3. the TemporalEncoder's convolutional embedder (two Multi_Scale_Conv_Blocks) --------------------------------
    torch.manual_seed(13)
    conv = M.MultiScaleConvEmbedder(22, [64, 128], [2, 2]).double()
    Bn, Lt = 6, 48
    # inputs and weights are fp32-representable (stored as fp32, exact); the arithmetic runs in fp64; results are stored as fp32
    # (6e-8 relative: far inside the 1e-5 gate) to keep the fixture small
    xt = torch.randn(Bn, Lt, 22, generator=gen).double().requires_grad_(True)
    h = conv(xt.permute(0, 2, 1))                                               # modules.py:143-146
    gh = torch.randn(h.shape, generator=gen).double()
    h.backward(gh)
    f32 = lambda t: t.detach().numpy().astype(np.float32)
    np.savez_compressed(
        os.path.join(GOLD, "temporal_block.npz"), x=f32(xt), y=f32(h), gy=f32(gh), g_x=f32(xt.grad),
        **{f"w_{k}": f32(v) for k, v in conv.named_parameters()},
        **{f"g_{k}": f32(v.grad) for k, v in conv.named_parameters()})
    for f in ("glue_embedding.npz", "glue_spatial_block.npz", "temporal_block.npz"):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
