"""GPU parity of the spatial block's producer / consumer glue against fixtures computed by the REFERENCE'S OWN classes
(tools/make_golden_glue.py): SpatioTemporalEmbedding (modules.py:211-266) and TEC_MoLLM.forward lines 75-106 around the
reference's SpatialEncoder."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _load_emb(g, prefix, device, num_nodes):
    from tec_mollm_b200 import SpatioTemporalEmbedding

    emb = SpatioTemporalEmbedding(16, num_nodes=num_nodes, num_years=13).to(device)
    sd = {k[len(prefix):]: torch.from_numpy(g[k]).float() for k in g.files if k.startswith(prefix)}
    emb.load_state_dict(sd, strict=True)  # the reference's parameter names
    return emb


def test_embedding_matches_the_reference_class(cuda_device):
    g = load_golden("glue_embedding.npz")
    x, tf = torch.from_numpy(g["x"]).to(cuda_device), torch.from_numpy(g["tf"]).to(cuda_device)
    B, L, N, _ = x.shape
    emb = _load_emb(g, "w_", cuda_device, N)
    ref = torch.from_numpy(g["out"])
    for tfi in (tf.unsqueeze(-2).expand(B, L, N, 4), tf, tf.unsqueeze(-2).expand(B, L, N, 4).contiguous()):  # train.py:65 and its source
        out = emb(x, tfi)
        assert torch.equal(out.cpu(), ref)                       # the reference's association order: bit-identical
    out.backward(torch.from_numpy(g["g"]).to(cuda_device))
    for k, q in emb.named_parameters():
        e = rel_err(q.grad, torch.from_numpy(g["g_" + k]))
        assert e <= 1e-6, f"{k}: {e:.3e}"
    bad = tf.unsqueeze(-2).expand(B, L, N, 4).clone()
    bad[0, 0, 3, 1] += 1
    with pytest.raises(NotImplementedError):
        emb(x, bad)
    with pytest.raises(RuntimeError):
        emb(x.cpu(), tf.cpu())


@pytest.mark.parametrize("raw_channels", [6, 4, 5])
def test_embedding_forward_other_channel_counts(cuda_device, raw_channels):
    """The reference's shape (6 raw channels) takes the float2 kernel, anything else the scalar one: both must be the
    reference's op sequence (modules.py:259-264) bit for bit, including an unaligned (odd-offset) input view."""
    from tec_mollm_b200 import SpatioTemporalEmbedding

    B, L, N = 2, 5, 301
    torch.manual_seed(11)
    emb = SpatioTemporalEmbedding(16, num_nodes=N).to(cuda_device)
    buf = torch.randn(B * L * N * raw_channels + 1, device=cuda_device)
    tf = torch.stack([torch.randint(0, 12, (B, L)), torch.randint(0, 366, (B, L)), torch.randint(0, 13, (B, L)),
                      torch.randint(0, 4, (B, L))], dim=-1).float().to(cuda_device)
    idx = tf.long()
    for off in (0, 1):  # off = 1: a 4-byte-aligned view (the float2 kernel must not be chosen for it)
        x = buf[off:off + B * L * N * raw_channels].view(B, L, N, raw_channels)
        with torch.no_grad():
            t = (emb.tod_embedding(idx[..., 0]) + emb.doy_embedding(idx[..., 1]) + emb.year_embedding(idx[..., 2])
                 + emb.season_embedding(idx[..., 3]))
            node = emb.node_embedding(torch.arange(N, device=cuda_device))
            ref = torch.cat([x, node.view(1, 1, N, 16) + t.unsqueeze(2)], dim=-1)
            out = emb(x, tf)
        assert torch.equal(out, ref)


def test_embedding_backward_is_deterministic_and_accumulates_over_snapshots(cuda_device):
    """Full-size rows (2911 nodes, 96 snapshots): gradients against torch's own embedding autograd, and bit-reproducible."""
    from tec_mollm_b200 import SpatioTemporalEmbedding

    B, L, N = 2, 48, 2911
    torch.manual_seed(3)
    emb = SpatioTemporalEmbedding(16, num_nodes=N).to(cuda_device)
    x = torch.randn(B, L, N, 6, device=cuda_device)
    tf = torch.stack([torch.randint(0, 12, (B, L)), torch.randint(0, 366, (B, L)), torch.randint(0, 13, (B, L)),
                      torch.randint(0, 4, (B, L))], dim=-1).float().to(cuda_device)
    gout = torch.randn(B, L, N, 22, device=cuda_device)
    runs = []
    for _ in range(2):
        emb.zero_grad(set_to_none=True)
        emb(x, tf).backward(gout)
        runs.append({k: q.grad.clone() for k, q in emb.named_parameters()})
    for k in runs[0]:
        assert torch.equal(runs[0][k], runs[1][k]), k
    idx = tf.long()
    ge = gout[..., 6:].double()
    ref = {"node_embedding.weight": ge.sum((0, 1))}
    gT = ge.sum(2)
    for j, (name, rows) in enumerate((("tod", 12), ("doy", 366), ("year", 13), ("season", 4))):
        t = torch.zeros(rows, 16, dtype=torch.float64, device=cuda_device)
        t.index_add_(0, idx[..., j].reshape(-1), gT.reshape(-1, 16))
        ref[f"{name}_embedding.weight"] = t
    for k, v in ref.items():
        assert rel_err(runs[0][k], v) <= 1e-6, k


def test_spatial_block_matches_the_reference_forward(cuda_device):
    """tec_mollm.py:75-106 as the reference runs it (its SpatialEncoder flattens (L*B, N, C) and passes the one-graph
    edge_index: the 'literal' mode), forward and every gradient, against the fixture made by the reference's own code."""
    from tec_mollm_b200 import SpatialEncoder

    g = load_golden("glue_spatial_block.npz")
    x = torch.from_numpy(g["x"]).float().to(cuda_device).requires_grad_(True)
    tf = torch.from_numpy(g["tf"]).float().to(cuda_device)
    ei = torch.from_numpy(g["edge_index"]).to(cuda_device)
    B, L, N, _ = x.shape
    emb = _load_emb(g, "w_spatio_temporal_embedding.", cuda_device, N)
    enc = SpatialEncoder(22, 11, heads=2, snapshot_mode="literal").to(cuda_device).eval()
    enc.load_state_dict({k[len("w_spatial_encoder."):]: torch.from_numpy(g[k]).float() for k in g.files
                         if k.startswith("w_spatial_encoder.")}, strict=True)
    xe = emb(x, tf.unsqueeze(-2).expand(B, L, N, 4))
    xg = xe.permute(1, 0, 2, 3).reshape(-1, N, 22)                               # tec_mollm.py:84
    xs = xg + enc(xg, ei, None)                                                  # :89, :94
    xt = xs.view(L, B, N, 22).permute(1, 2, 0, 3).reshape(-1, L, 22)             # :100, :106
    assert rel_err(xt, torch.from_numpy(g["x_temporal"])) <= 1e-5
    xt.backward(torch.from_numpy(g["gz"]).float().to(cuda_device))
    assert rel_err(x.grad, torch.from_numpy(g["g_x"])) <= 1e-5
    for k, q in list(emb.named_parameters()):
        e = rel_err(q.grad, torch.from_numpy(g["g_spatio_temporal_embedding." + k]))
        assert e <= 1e-5, f"embedding {k}: {e:.3e}"
    for k, q in enc.named_parameters():
        e = rel_err(q.grad, torch.from_numpy(g["g_spatial_encoder." + k]))
        assert e <= 1e-5, f"encoder {k}: {e:.3e}"


def test_forward_block_from_raw_features_shared_mode(cuda_device):
    """The fused spatial block fed by the fused embedding (only raw features + (B, L, 4) indices on the input side) equals
    the three reference lines around the encoder in shared mode, output and gradients."""
    from tec_mollm_b200 import SpatialEncoder, SpatioTemporalEmbedding

    B, L, N = 2, 5, 2911
    ei = torch.from_numpy(load_golden("graph_cn150.npz")["edge_index"]).to(cuda_device)
    torch.manual_seed(9)
    emb = SpatioTemporalEmbedding(16, num_nodes=N).to(cuda_device)
    enc = SpatialEncoder(22, 11, heads=2, dropout=0.0, snapshot_mode="shared").to(cuda_device).eval()
    x = torch.randn(B, L, N, 6, device=cuda_device)
    tf = torch.stack([torch.randint(0, 12, (B, L)), torch.randint(0, 366, (B, L)), torch.randint(0, 13, (B, L)),
                      torch.randint(0, 4, (B, L))], dim=-1).float().to(cuda_device)
    gz = torch.randn(B * N, L, 22, device=cuda_device)
    res = []
    for fused in (False, True):
        emb.zero_grad(set_to_none=True)
        enc.zero_grad(set_to_none=True)
        xe = emb(x, tf)
        if fused:
            z = enc.forward_block(xe, ei)
        else:
            xg = xe.permute(1, 0, 2, 3).reshape(-1, N, 22)
            z = (xg + enc(xg, ei)).view(L, B, N, 22).permute(1, 2, 0, 3).reshape(-1, L, 22)
        z.backward(gz)
        res.append((z.detach().clone(), {k: q.grad.clone() for m in (emb, enc) for k, q in m.named_parameters()}))
    assert rel_err(res[1][0], res[0][0]) <= 1e-6
    for k in res[0][1]:
        assert rel_err(res[1][1][k], res[0][1][k]) <= 2e-6, k
