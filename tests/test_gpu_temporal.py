"""GPU parity of the TemporalEncoder's convolutional embedder against the fixture computed by the
REFERENCE'S OWN MultiScaleConvEmbedder (tools/make_golden_glue.py, fp64): the stacked 7-tap library convolution + the fused
GroupNorm/GELU/concat/stride pass (csrc/temporal.cu) + the 1x1 contraction."""
import pytest
import torch

from helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _load(device):
    from tec_mollm_b200 import MultiScaleConvEmbedder

    g = load_golden("temporal_block.npz")
    m = MultiScaleConvEmbedder(22, [64, 128], [2, 2]).to(device)
    m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w_")}, strict=True)  # reference names
    return g, m


def test_conv_embedder_fp32_matches_the_reference_class(cuda_device):
    g, m = _load(cuda_device)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False   # the fp32 contract (1e-5)
    try:
        x = torch.from_numpy(g["x"]).to(cuda_device).requires_grad_(True)
        y = m(x.permute(0, 2, 1))                                                     # modules.py:143-146
        assert y.shape == (x.size(0), 128, 12)
        e = rel_err(y, torch.from_numpy(g["y"]))
        assert e <= 1e-5, f"y: {e:.3e}"
        y.backward(torch.from_numpy(g["gy"]).to(cuda_device))
        e = rel_err(x.grad, torch.from_numpy(g["g_x"]))
        assert e <= 1e-5, f"dx: {e:.3e}"
        for k, q in m.named_parameters():
            e = rel_err(q.grad, torch.from_numpy(g["g_" + k]))
            assert e <= 1e-5, f"{k}: {e:.3e}"
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_conv_embedder_bf16_autocast_and_determinism(cuda_device):
    """bf16 autocast (train.py:68): within 1e-2 of the fp64 fixture on the output, 3e-2 on the gradients (two bf16 convolutions
    deep), and two runs bit-identical (no atomics in the fused pass)."""
    g, m = _load(cuda_device)
    runs = []
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        x = torch.from_numpy(g["x"]).to(cuda_device).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x.permute(0, 2, 1))
        y.float().backward(torch.from_numpy(g["gy"]).to(cuda_device))
        runs.append((y.detach().float().clone(), x.grad.clone(), {k: q.grad.clone() for k, q in m.named_parameters()}))
    assert rel_err(runs[0][0], torch.from_numpy(g["y"])) <= 1e-2
    assert rel_err(runs[0][1], torch.from_numpy(g["g_x"])) <= 3e-2
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    gam = "embedder.0.convs.1.1.weight"
    assert torch.equal(runs[0][2][gam], runs[1][2][gam])


def test_fused_gn_gelu_pass_against_torch_ops(cuda_device):
    """The fused pass alone, odd sizes and stride 1 / 2 / 3, against torch's group_norm + gelu + slicing in fp64."""
    import torch.nn.functional as F
    from tec_mollm_b200.temporal import _GnGeluStride

    torch.manual_seed(5)
    for (n, br, c, length, stride) in ((7, 3, 64, 48, 2), (5, 3, 128, 24, 2), (3, 2, 10, 37, 3), (4, 1, 33, 50, 1)):
        y = torch.randn(n, br * c, length, device=cuda_device, requires_grad=True)
        gamma = torch.randn(br, c, device=cuda_device, requires_grad=True)
        beta = torch.randn(br, c, device=cuda_device, requires_grad=True)
        z = _GnGeluStride.apply(y, gamma, beta, br, stride, 1e-5, torch.float32)
        gz = torch.randn_like(z)
        z.backward(gz)
        y64, g64, b64 = (t.detach().double().requires_grad_(True) for t in (y, gamma, beta))
        outs = [F.gelu(F.group_norm(y64[:, j * c:(j + 1) * c], 1, g64[j], b64[j], eps=1e-5)) for j in range(br)]
        ref = torch.cat(outs, dim=1)[:, :, ::stride]
        ref.backward(gz.double())
        assert rel_err(z, ref) <= 2e-6
        assert rel_err(y.grad, y64.grad) <= 1e-5
        assert rel_err(gamma.grad, g64.grad) <= 1e-5 and rel_err(beta.grad, b64.grad) <= 1e-5
