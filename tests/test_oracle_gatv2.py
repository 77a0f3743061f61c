"""CPU tests of the GATv2 oracle (checker infrastructure): self-consistency + frozen golden vectors.

The reference pins nothing for this path, so the oracle is cross-checked three ways in fp64:
hand-derived backward vs autograd, an independent dense formulation vs the scatter formulation, and the
reference's literal-call property (rows >= N reduce to W_l x + b_l + bias)."""
import numpy as np
import pytest
import torch

from oracle import gatv2_oracle as G
from helpers import load_golden, random_graph


def _params(F, H, C, seed=3, dtype=torch.float64):
    p = G.init_params(F, C, H, seed=seed, dtype=dtype)
    p["bias"] = torch.randn(H * C, generator=torch.Generator().manual_seed(seed + 1), dtype=dtype) * 0.1
    return p


@pytest.mark.parametrize("F,H,C", [(7, 2, 5), (22, 2, 11), (6, 4, 3), (5, 1, 8)])
def test_manual_backward_matches_autograd(F, H, C):
    N = 40
    ei = random_graph(N, 200, seed=0, isolated=(5,))
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(N, F, generator=gen, dtype=torch.float64)
    gy = torch.randn(N, H * C, generator=gen, dtype=torch.float64)
    p = _params(F, H, C)
    E = int((ei[0] != ei[1]).sum()) + N
    mask = (torch.rand(E, H, generator=gen) > 0.3).double()
    xg = x.clone().requires_grad_(True)
    pg = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    G.gatv2_forward(xg, ei, pg, H, C, edge_mask=mask, p=0.3).backward(gy)
    man = G.gatv2_backward_manual(x, ei, p, H, C, gy, edge_mask=mask, p=0.3)
    assert torch.allclose(man["x"], xg.grad, rtol=0, atol=1e-12)
    for k in G.PARAM_NAMES:
        assert torch.allclose(man[k], pg[k].grad, rtol=0, atol=1e-11), k


def test_dense_formulation_agrees():
    N, F, H, C = 30, 6, 2, 4
    ei = random_graph(N, 150, seed=2)
    x = torch.randn(N, F, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    p = _params(F, H, C)
    assert torch.allclose(G.gatv2_forward_dense(x, ei, p, H, C), G.gatv2_forward(x, ei, p, H, C), rtol=0, atol=1e-13)


def test_self_loop_surgery_and_edge_cases():
    ei = torch.tensor([[0, 1, 1, 2, 2, 0], [1, 1, 0, 2, 0, 1]])  # two self loops, duplicate (0->1)
    out = G.remove_then_add_self_loops(ei, 4)
    assert out.tolist() == [[0, 1, 2, 0, 0, 1, 2, 3], [1, 0, 0, 1, 0, 1, 2, 3]]
    # an isolated node attends only to itself: y = xl + bias
    x = torch.randn(4, 3, dtype=torch.float64)
    p = _params(3, 2, 2)
    y = G.gatv2_forward(x, ei, p, 2, 2)
    xl = x @ p["lin_l.weight"].t() + p["lin_l.bias"]
    assert torch.allclose(y[3], xl[3] + p["bias"], atol=1e-14)


def test_literal_mode_is_bug_compatible():
    """modules.py:353-356 as written: only snapshot 0 sees edges; every other row is W_l x + b_l + bias."""
    S, N, F, H, C = 3, 12, 5, 2, 3
    ei = random_graph(N, 40, seed=4)
    x = torch.randn(S, N, F, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    p = _params(F, H, C)
    y_lit = G.spatial_encoder_forward(x, ei, p, H, C, "literal")
    y_sh = G.spatial_encoder_forward(x, ei, p, H, C, "shared")
    assert torch.allclose(y_lit[0], y_sh[0], atol=1e-14)
    xl = x @ p["lin_l.weight"].t() + p["lin_l.bias"]
    assert torch.allclose(y_lit[1:], xl[1:] + p["bias"], atol=1e-14)
    assert not torch.allclose(y_lit[1:], y_sh[1:], atol=1e-6)


def test_shared_mode_equals_per_snapshot_calls():
    S, N, F, H, C = 4, 15, 6, 2, 5
    ei = random_graph(N, 60, seed=6)
    x = torch.randn(S, N, F, generator=torch.Generator().manual_seed(7), dtype=torch.float64)
    p = _params(F, H, C)
    y = G.spatial_encoder_forward(x, ei, p, H, C, "shared")
    for s in range(S):
        assert torch.allclose(y[s], G.gatv2_forward(x[s], ei, p, H, C), atol=1e-14)


def test_batch_sharding_sums_to_unsharded_gradients():
    """snapshots are independent, so per-shard parameter gradients add up exactly."""
    S, N, F, H, C = 6, 10, 4, 2, 3
    ei = random_graph(N, 30, seed=8)
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(S, N, F, generator=gen, dtype=torch.float64)
    gy = torch.randn(S, N, H * C, generator=gen, dtype=torch.float64)
    p = _params(F, H, C)
    _, full = G.fwd_bwd(x, ei, p, H, C, gy)
    acc = {k: torch.zeros_like(v) for k, v in p.items()}
    for lo, hi in ((0, 2), (2, 6)):
        _, part = G.fwd_bwd(x[lo:hi], ei, p, H, C, gy[lo:hi])
        assert torch.allclose(part["x"], full["x"][lo:hi], atol=1e-13)
        for k in acc:
            acc[k] += part[k]
    for k in acc:
        assert torch.allclose(acc[k], full[k], atol=1e-12), k


@pytest.mark.parametrize("name", ["f22h2c11", "f10h2c5", "f22h4c11"])
def test_frozen_golden_vectors(name):
    g = load_golden(f"gatv2_{name}.npz")
    F, H, C = int(g["F"]), int(g["H"]), int(g["C"])
    params = {k: torch.from_numpy(g[f"p_{k}"]) for k in G.PARAM_NAMES}
    x, gy, ei = torch.from_numpy(g["x"]), torch.from_numpy(g["gy"]), torch.from_numpy(g["edge_index"])
    for mode in ("shared", "literal"):
        y, grads = G.fwd_bwd(x, ei, params, H, C, gy, snapshot_mode=mode)
        assert np.allclose(y.numpy(), g[f"y_{mode}"], rtol=0, atol=1e-12)
        for k, v in grads.items():
            assert np.allclose(v.numpy(), g[f"g_{mode}_{k}"], rtol=0, atol=1e-10), (mode, k)


def test_autocast_dtype_flow():
    """Under autocast(bf16) the Linear outputs / add / leaky_relu are bf16, everything after the att product is fp32."""
    N, F, H, C = 20, 22, 2, 11
    ei = random_graph(N, 80, seed=10)
    x = torch.randn(N, F, generator=torch.Generator().manual_seed(11))
    p = _params(F, H, C, dtype=torch.float32)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        y = G.gatv2_forward(x, ei, p, H, C)
    assert y.dtype == torch.float32
    ref = G.gatv2_forward(x.double(), ei, {k: v.double() for k, v in p.items()}, H, C)
    err = (y.double() - ref).abs().max() / ref.abs().max()
    assert 1e-5 < err < 2e-2


def test_init_distributions():
    p = G.init_params(22, 11, 2, seed=0)
    a = (6.0 / (22 + 22)) ** 0.5
    assert p["lin_l.weight"].shape == (22, 22) and p["lin_l.weight"].abs().max() <= a
    assert p["lin_l.bias"].abs().max() <= 22 ** -0.5
    assert p["att"].shape == (1, 2, 11) and p["att"].abs().max() <= (6.0 / 13) ** 0.5
    assert torch.count_nonzero(p["bias"]) == 0
    assert sum(v.numel() for v in p.values()) == 1056
