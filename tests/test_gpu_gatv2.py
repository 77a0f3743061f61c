"""GPU parity tests (run with `-m gpu` on the B200): CUDA path through the C ABI vs the CPU oracle.

Tolerances are BASELINE.json's: fp32 outputs and gradients <= 1e-5 relative (max |diff| / max |ref|, against the
fp64 oracle), bf16-autocast <= 1e-2 (against the oracle run under CPU autocast, i.e. PyG's dtype flow).

fp32 gradients are compared with `oracle_with_kernel_branches` (helpers.py): the LeakyReLU derivative is discontinuous at
s = 0, so where |s| is below fp32 resolution the reference gradient is only defined up to the branch an fp32
implementation's rounding picks; the helper pins the oracle to the kernels' (verified-ambiguous) choices and the
comparison is then strict on every entry."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from helpers import kernel_dropout_mask, load_golden, oracle_with_kernel_branches, random_graph, rel_err
from oracle import gatv2_oracle as G
from oracle import graph_oracle as go

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5
TOL_BF16 = 1e-2


def _encoder(F, H, C, params, device, mode="shared", dropout=0.0):
    from tec_mollm_b200 import SpatialEncoder

    enc = SpatialEncoder(F, C, heads=H, dropout=dropout, snapshot_mode=mode).to(device)
    sd = {f"gat_conv.{k}": v.float() for k, v in params.items()}
    enc.load_state_dict(sd, strict=True)
    return enc


def _run_cuda(enc, x, ei, gy, autocast=False, seed=None):
    dev = next(enc.parameters()).device
    xg = x.float().to(dev).requires_grad_(True)
    enc.zero_grad(set_to_none=True)
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = enc(xg, ei.to(dev))
    else:
        y = enc(xg, ei.to(dev))
    y.backward(gy.float().to(dev))
    grads = {"x": xg.grad}
    grads.update({k[len("gat_conv."):]: p.grad for k, p in enc.named_parameters()})
    return y.detach(), grads


def _check(y, grads, y_ref, g_ref, tol, what=""):
    assert y.dtype == torch.float32
    e = rel_err(y, y_ref)
    assert e <= tol, f"{what} y: rel err {e:.3e} > {tol}"
    for k, ref in g_ref.items():
        e = rel_err(grads[k], ref)
        assert e <= tol, f"{what} grad {k}: rel err {e:.3e} > {tol}"


def _rand_case(S, N, F, H, C, seed, dtype=torch.float64):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(S, N, F, generator=gen, dtype=dtype)
    gy = torch.randn(S, N, H * C, generator=gen, dtype=dtype)
    p = G.init_params(F, C, H, seed=seed + 1, dtype=dtype)
    p["bias"] = torch.randn(H * C, generator=gen, dtype=dtype) * 0.1
    return x, gy, p


# ---------------------------------------------------------------------------------------------------------
# plan
# ---------------------------------------------------------------------------------------------------------
def test_plan_matches_pyg_edge_surgery(cuda_device):
    from tec_mollm_b200 import GraphPlan

    N = 50
    ei = random_graph(N, 300, seed=0, isolated=(7, 49))
    plan = GraphPlan(ei.to(cuda_device), N, 32)
    rowptr, col, eid = plan.export()
    ref = G.remove_then_add_self_loops(ei, N)
    assert plan.num_edges == ref.size(1) and plan.kept_edges == ref.size(1) - N
    src, dst = ref[0].numpy(), ref[1].numpy()
    assert sorted(eid.tolist()) == list(range(ref.size(1)))          # a permutation of PyG's edge ids
    for i in range(N):
        ids = eid[rowptr[i]:rowptr[i + 1]]
        assert np.all(dst[ids] == i)                                  # destination-sorted
        assert np.array_equal(col[rowptr[i]:rowptr[i + 1]], src[ids])
        assert ids[0] == plan.kept_edges + i                          # every row starts with its self loop ...
        assert np.all(np.diff(ids[1:]) > 0)                           # ... then the kept edges in PyG's (stable) order
    assert plan.max_in_degree == int(np.diff(rowptr).max())


def test_plan_rejects_bad_input(cuda_device):
    from tec_mollm_b200 import GraphPlan

    with pytest.raises(RuntimeError, match="outside"):
        GraphPlan(torch.tensor([[0, 9], [1, 2]], device=cuda_device), 5, 32)
    with pytest.raises(ValueError):
        GraphPlan(torch.zeros(3, 4, dtype=torch.int64, device=cuda_device), 5, 32)
    with pytest.raises(RuntimeError, match="CUDA"):
        GraphPlan(torch.tensor([[0], [1]]), 5, 32)


# ---------------------------------------------------------------------------------------------------------
# projections: tensor-core path vs CUDA-core path vs fp64 matmul
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", ["ffma", "tc"])
@pytest.mark.parametrize("R,F,HC", [(1000, 22, 22), (128 * 5, 10, 10), (77, 22, 44), (2911 * 3, 22, 22), (300, 7, 6)])
def test_projection_forward(cuda_device, impl, R, F, HC):
    from tec_mollm_b200 import _lib

    gen = torch.Generator().manual_seed(R + F)
    x = torch.randn(R, F, generator=gen)
    wl, wr = torch.randn(HC, F, generator=gen) * 0.3, torch.randn(HC, F, generator=gen) * 0.3
    bl, br = torch.randn(HC, generator=gen), torch.randn(HC, generator=gen)
    d = lambda t: t.to(cuda_device)
    xl = torch.full((R, HC), float("nan"), device=cuda_device)
    xr = torch.full((R, HC), float("nan"), device=cuda_device)
    code = _lib.PROJ_TC if impl == "tc" else _lib.PROJ_FFMA
    xd, wld, bld, wrd, brd = d(x), d(wl), d(bl), d(wr), d(br)
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.call("tecgat_project_fwd", p(xd), p(wld), p(bld), p(wrd), p(brd), p(xl), p(xr), R, F, HC, _lib.F32, code,
              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref_l = x.double() @ wl.double().t() + bl.double()
    ref_r = x.double() @ wr.double().t() + br.double()
    assert rel_err(xl, ref_l) <= 2e-6, f"xl {rel_err(xl, ref_l):.3e}"
    assert rel_err(xr, ref_r) <= 2e-6, f"xr {rel_err(xr, ref_r):.3e}"
    # bf16 contract
    xl16 = torch.empty((R, HC), device=cuda_device, dtype=torch.bfloat16)
    xr16 = torch.empty((R, HC), device=cuda_device, dtype=torch.bfloat16)
    _lib.call("tecgat_project_fwd", p(xd), p(wld), p(bld), p(wrd), p(brd), p(xl16), p(xr16), R, F, HC, _lib.BF16, code,
              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rel_err(xl16.float(), ref_l) <= 1e-2 and rel_err(xr16.float(), ref_r) <= 1e-2


@pytest.mark.parametrize("impl", ["ffma", "tc"])
@pytest.mark.parametrize("R,F,HC,need_dx", [(1000, 22, 22, True), (128 * 700 + 13, 22, 22, True), (77, 10, 10, True), (128 * 40 + 5, 22, 44, True),
                                            (128 * 40, 22, 22, False), (300, 7, 6, True), (128 * 3 + 1, 22, 44, True)])
def test_projection_backward(cuda_device, impl, R, F, HC, need_dx):
    """dx = dxl Wl + dxr Wr, dW = d^T x, db = sum d: tensor-core (3xTF32, TMEM-accumulated dW) and CUDA-core kernels vs
    fp64 matmul; many tiles per CTA, ragged last tile, dx skipped when the input needs no gradient."""
    from tec_mollm_b200 import _lib

    gen = torch.Generator().manual_seed(R * 7 + F)
    x = torch.randn(R, F, generator=gen)
    dxl, dxr = torch.randn(R, HC, generator=gen), torch.randn(R, HC, generator=gen)
    wl, wr = torch.randn(HC, F, generator=gen) * 0.3, torch.randn(HC, F, generator=gen) * 0.3
    code = _lib.PROJ_TC if impl == "tc" else _lib.PROJ_FFMA
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    ref_dx = dxl.double() @ wl.double() + dxr.double() @ wr.double()
    refs = {"dwl": dxl.double().t() @ x.double(), "dwr": dxr.double().t() @ x.double(),
            "dbl": dxl.double().sum(0), "dbr": dxr.double().sum(0)}
    for dtype, st, tol in ((_lib.F32, torch.float32, 3e-6), (_lib.BF16, torch.bfloat16, 1e-2)):
        d = lambda t: t.to(cuda_device)
        xd, wld, wrd = d(x), d(wl), d(wr)
        dl, dr = d(dxl).to(st), d(dxr).to(st)
        dx = torch.full((R, F), float("nan"), device=cuda_device) if need_dx else None
        outs = {k: torch.full(v.shape, float("nan"), device=cuda_device) for k, v in refs.items()}
        ws = torch.empty(max(1, _lib.lib().tecgat_project_bwd_workspace(R, F, HC, code)), dtype=torch.uint8, device=cuda_device)
        _lib.call("tecgat_project_bwd", p(dl), p(dr), p(xd), p(wld), p(wrd), p(dx), p(outs["dwl"]), p(outs["dbl"]),
                  p(outs["dwr"]), p(outs["dbr"]), p(ws), R, F, HC, dtype, code,
                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        if dtype == _lib.BF16:  # the bf16 contract rounds the operands first: compare against the rounded operands
            dl64, dr64 = dl.double().cpu(), dr.double().cpu()
            r_dx = dl64 @ wl.double() + dr64 @ wr.double()
            r = {"dwl": dl64.t() @ x.double(), "dwr": dr64.t() @ x.double(), "dbl": dl64.sum(0), "dbr": dr64.sum(0)}
        else:
            r_dx, r = ref_dx, refs
        if need_dx:
            assert rel_err(dx, r_dx) <= tol, f"dx {rel_err(dx, r_dx):.3e}"
        for k in r:
            assert rel_err(outs[k], r[k]) <= tol, f"{k} {rel_err(outs[k], r[k]):.3e} (dtype {dtype})"


# ---------------------------------------------------------------------------------------------------------
# fused path vs oracle
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["f22h2c11", "f10h2c5", "f22h4c11"])
@pytest.mark.parametrize("mode", ["shared", "literal"])
def test_golden_fixtures_fp32(cuda_device, name, mode):
    g = load_golden(f"gatv2_{name}.npz")
    F, H, Cc = int(g["F"]), int(g["H"]), int(g["C"])
    params = {k: torch.from_numpy(g[f"p_{k}"]) for k in G.PARAM_NAMES}
    x, gy, ei = torch.from_numpy(g["x"]), torch.from_numpy(g["gy"]), torch.from_numpy(g["edge_index"])
    enc = _encoder(F, H, Cc, params, cuda_device, mode).eval()
    y, grads = _run_cuda(enc, x, ei, gy)
    y_ref = torch.from_numpy(g[f"y_{mode}"])
    assert rel_err(y, y_ref) <= TOL_F32                                  # frozen golden output
    y_o, g_ref, flips = oracle_with_kernel_branches(x, ei, params, H, Cc, gy, cuda_device, mode)
    # with zero ambiguous pre-activations the branch-pinned oracle IS the golden oracle; a flipped branch (|s| < 1e-5, see
    # helpers.oracle_with_kernel_branches) moves the forward value by at most ~|s| * (1 - slope)
    assert rel_err(y_o, y_ref) <= (1e-8 if flips == 0 else 1e-6)   # fp64 CPU arithmetic differs between host CPUs at ~1e-9
    _check(y, grads, y_ref, g_ref, TOL_F32, f"{name}/{mode}")
    if flips == 0:                                                       # no ambiguous pre-activation: frozen golden grads
        g_gold = {k: torch.from_numpy(g[f"g_{mode}_{k}"]) for k in ("x",) + G.PARAM_NAMES}
        _check(y, grads, y_ref, g_gold, TOL_F32, f"{name}/{mode}/golden")


@pytest.mark.parametrize("S,N,F,H,C,E", [
    (3, 40, 7, 2, 5, 200),      # duplicates, self loops, isolated node, asymmetric
    (2, 300, 22, 2, 11, 3000),  # several tiles, window larger than a tile
    (1, 17, 6, 1, 8, 60),       # single head, even C
    (4, 90, 5, 4, 3, 500),      # four heads
    (2, 64, 9, 2, 16, 400),
    (2, 33, 4, 3, 4, 150),      # heads not a power of two
    (3, 50, 6, 1, 5, 300),      # H*C odd: rows are not pair aligned (scalar-load instantiation), 20-byte rows
    (2, 45, 8, 3, 3, 250),      # H*C = 9, padded head lanes
    (5, 70, 22, 2, 11, 0),      # no edges but several snapshots: odd row offsets of every window
])
def test_random_graphs_fp32(cuda_device, S, N, F, H, C, E):
    ei = random_graph(N, E, seed=N + E, isolated=(3,))
    x, gy, p = _rand_case(S, N, F, H, C, seed=S * 100 + N)
    enc = _encoder(F, H, C, p, cuda_device).eval()
    y, grads = _run_cuda(enc, x, ei, gy)
    y_ref, g_ref, _ = oracle_with_kernel_branches(x, ei, p, H, C, gy, cuda_device)
    assert rel_err(y_ref, G.spatial_encoder_forward(x, ei, p, H, C)) <= 1e-12   # closed form == op-for-op oracle
    _check(y, grads, y_ref, g_ref, TOL_F32, f"S{S}N{N}F{F}H{H}C{C}")


@pytest.mark.parametrize("S,N,F,H,C,E", [(2, 300, 22, 2, 11, 3000), (2, 90, 5, 4, 3, 500), (1, 3000, 22, 2, 11, 24000)])
def test_gather_path_when_windows_are_not_staged(cuda_device, monkeypatch, S, N, F, H, C, E):
    """Tiles whose row window does not fit a shared-memory stage gather neighbour rows from global memory.  The last
    case gets there naturally (random graph: every tile's window spans all 3000 rows); the others force it."""
    if N < 3000:
        monkeypatch.setenv("TECGAT_EDGE_NOSTAGE", "1")
    ei = random_graph(N, E, seed=N + E + 7, isolated=(5,))
    x, gy, p = _rand_case(S, N, F, H, C, seed=S * 10 + N)
    enc = _encoder(F, H, C, p, cuda_device).eval()
    y, grads = _run_cuda(enc, x, ei, gy)
    y_ref, g_ref, _ = oracle_with_kernel_branches(x, ei, p, H, C, gy, cuda_device)
    _check(y, grads, y_ref, g_ref, TOL_F32, f"gather S{S}N{N}")


@pytest.mark.parametrize("S,N,F,H,C,E", [(3, 300, 22, 2, 11, 3000), (2, 90, 5, 4, 3, 500), (2, 257, 10, 1, 5, 1200)])
def test_backward_semi_staging(cuda_device, monkeypatch, S, N, F, H, C, E):
    """Rows too wide for shared memory (H*C = 44 on the 300 km graph, see test_dense_graph_four_heads) stage only the ELL
    slab and the (delta, stat) planes and gather xl / xr / g / y rows from L2; forced here on small graphs, with dropout."""
    monkeypatch.setenv("TECGAT_EDGE_SEMI", "1")
    ei = random_graph(N, E, seed=N + E + 11, isolated=(4,))
    x, gy, p = _rand_case(S, N, F, H, C, seed=S * 7 + N)
    enc = _encoder(F, H, C, p, cuda_device).eval()
    y, grads = _run_cuda(enc, x, ei, gy)
    y_ref, g_ref, _ = oracle_with_kernel_branches(x, ei, p, H, C, gy, cuda_device)
    _check(y, grads, y_ref, g_ref, TOL_F32, f"semi S{S}N{N}")


def test_softmax_shift_retry_on_huge_score_spread(cuda_device):
    """The kernels shift the softmax by the self-loop score; rows where another score exceeds it by more than ~2^100
    are redone with the exact maximum.  att scaled by 400 makes score differences of several hundred."""
    S, N, F, H, C, E = 2, 120, 22, 2, 11, 900
    ei = random_graph(N, E, seed=99, isolated=(2,))
    x, gy, p = _rand_case(S, N, F, H, C, seed=77)
    p["att"] = p["att"] * 400.0
    enc = _encoder(F, H, C, p, cuda_device).eval()
    y, grads = _run_cuda(enc, x, ei, gy)
    y_ref, g_ref, _ = oracle_with_kernel_branches(x, ei, p, H, C, gy, cuda_device)
    assert torch.isfinite(y).all()
    assert rel_err(y, y_ref) <= 2e-5      # alpha is a near one-hot: exp2 argument errors scale with |score| ~ 1e3
    for k in ("x", "att", "lin_l.weight", "lin_r.weight"):
        assert torch.isfinite(grads[k]).all() and rel_err(grads[k], g_ref[k]) <= 1e-3, k


def test_empty_edge_list_and_single_node(cuda_device):
    """No edges at all: every node attends only to its self loop (y = W_l x + b_l + bias)."""
    S, N, F, H, C = 2, 9, 5, 2, 3
    x, gy, p = _rand_case(S, N, F, H, C, seed=1)
    ei = torch.zeros(2, 0, dtype=torch.int64)
    enc = _encoder(F, H, C, p, cuda_device).eval()
    y, grads = _run_cuda(enc, x, ei, gy)
    y_ref, g_ref, _ = oracle_with_kernel_branches(x, ei, p, H, C, gy, cuda_device)
    _check(y, grads, y_ref, g_ref, TOL_F32, "empty")
    x1, gy1, _ = _rand_case(1, 1, F, H, C, seed=2)
    y, grads = _run_cuda(enc, x1, ei, gy1)
    y_ref, g_ref, _ = oracle_with_kernel_branches(x1, ei, p, H, C, gy1, cuda_device)
    _check(y, grads, y_ref, g_ref, TOL_F32, "single node")


def test_reference_graph_fp32_and_determinism(cuda_device):
    """The 2911-node, 150 km graph (reference golden edge list), default shape F=22, H=2, C=11."""
    g = load_golden("graph_cn150.npz")
    ei = torch.from_numpy(g["edge_index"])
    S, N, F, H, C = 4, 2911, 22, 2, 11
    x, gy, p = _rand_case(S, N, F, H, C, seed=5)
    enc = _encoder(F, H, C, p, cuda_device).eval()
    y, grads = _run_cuda(enc, x, ei, gy)
    y_ref, g_ref, _ = oracle_with_kernel_branches(x, ei, p, H, C, gy, cuda_device)
    _check(y, grads, y_ref, g_ref, TOL_F32, "cn150")
    y2, grads2 = _run_cuda(enc, x, ei, gy)
    assert torch.equal(y, y2)                                   # atomic-free => bit-reproducible
    for k in grads:
        assert torch.equal(grads[k], grads2[k]), k


@pytest.mark.parametrize("split", ["0", "1"])
def test_dense_graph_four_heads(cuda_device, monkeypatch, split):
    """BASELINE config 4: 300 km graph (76,532 edges, max degree 38), H=4, C=11 -- through the one-launch four-head kernels
    (split = 0) and as two independent head pairs on parameter slices (the default for this shape)."""
    monkeypatch.setenv("TECGAT_HEAD_SPLIT", split)
    g = load_golden("graph_cn300.npz")
    ei = torch.from_numpy(g["edge_index"])
    S, N, F, H, C = 2, 2911, 22, 4, 11
    x, gy, p = _rand_case(S, N, F, H, C, seed=6)
    enc = _encoder(F, H, C, p, cuda_device).eval()
    y, grads = _run_cuda(enc, x, ei, gy)
    y_ref, g_ref, _ = oracle_with_kernel_branches(x, ei, p, H, C, gy, cuda_device)
    _check(y, grads, y_ref, g_ref, TOL_F32, "cn300/h4")


def test_head_pairs_training_and_autocast(cuda_device, monkeypatch):
    """Four heads as two head pairs: dropout draws a different stream per pair (p -> 0 reproduces the deterministic result, the
    two pairs' masks differ), bf16 autocast within 1e-2 of the fp32 path, literal mode equals the one-launch kernels."""
    S, N, F, H, C = 3, 200, 22, 4, 11
    ei = random_graph(N, 900, seed=41, isolated=(3,))
    x, gy, p = _rand_case(S, N, F, H, C, seed=42, dtype=torch.float32)
    outs = {}
    for split in ("0", "1"):
        monkeypatch.setenv("TECGAT_HEAD_SPLIT", split)
        for mode in ("shared", "literal"):
            enc = _encoder(F, H, C, p, cuda_device, mode=mode).eval()
            outs[(split, mode)] = _run_cuda(enc, x, ei, gy)
    for mode in ("shared", "literal"):
        y0, g0 = outs[("0", mode)]
        y1, g1 = outs[("1", mode)]
        assert rel_err(y1, y0) <= 2e-6, mode
        for k in g0:
            assert rel_err(g1[k], g0[k]) <= 5e-6, (mode, k)
    monkeypatch.setenv("TECGAT_HEAD_SPLIT", "1")
    enc = _encoder(F, H, C, p, cuda_device, dropout=0.5).train()
    xg = x.to(cuda_device)
    with torch.no_grad():
        ya, yb = enc(xg, ei.to(cuda_device)), enc(xg, ei.to(cuda_device))
    assert not torch.equal(ya, yb)                                   # fresh masks per call
    y_eval = outs[("1", "shared")][0]
    d = (ya - y_eval).abs().reshape(S * N, 2, 2 * C).amax(dim=(0, 2))   # both pairs are perturbed by their own masks
    assert (d > 1e-3).all()
    enc = _encoder(F, H, C, p, cuda_device).eval()
    y16, g16 = _run_cuda(enc, x, ei, gy, autocast=True)
    assert rel_err(y16, y_eval) <= 1e-2
    # bf16 gradients sit 5e-2 .. 7e-2 from fp64 in PyG's own autocast dtype flow too (test_gpu_round2's training-shape test gates
    # them against that flow); here only that the pair path is not worse than that
    assert rel_err(g16["x"], outs[("1", "shared")][1]["x"]) <= 1e-1


@pytest.mark.parametrize("mode", ["shared", "literal"])
def test_dropout_with_the_kernels_own_mask(cuda_device, mode):
    """Training-mode attention dropout: the oracle is fed exactly the keep-mask the kernels derive from
    (seed, snapshot, CSR slot, head), so forward AND backward must agree to fp32 tolerance."""
    S, N, F, H, C, p_drop = 3, 60, 22, 2, 11, 0.25
    ei = random_graph(N, 400, seed=12)
    x, gy, p = _rand_case(S, N, F, H, C, seed=13)
    enc = _encoder(F, H, C, p, cuda_device, mode, dropout=p_drop).train()
    y, grads = _run_cuda(enc, x, ei, gy)
    seed = int(enc.gat_conv._last_seed.item()) & (2 ** 64 - 1)  # drawn on the device (tecgat_seed_advance)
    plan = next(iter(enc.gat_conv._plans.values()))[0]
    mask = kernel_dropout_mask(plan, S, H, p_drop, seed, mode).double()
    assert 0.15 < 1.0 - mask.mean().item() < 0.35
    y_ref, g_ref, _ = oracle_with_kernel_branches(x, ei, p, H, C, gy, cuda_device, mode, edge_mask=mask, p=p_drop)
    y_ad, _ = G.fwd_bwd(x, ei, p, H, C, gy, snapshot_mode=mode, edge_mask=mask, p=p_drop)
    assert rel_err(y_ref, y_ad) <= 1e-12
    _check(y, grads, y_ref, g_ref, TOL_F32, f"dropout/{mode}")
    enc.eval()
    y_eval, _ = _run_cuda(enc, x, ei, gy)
    y_ref0, _ = G.fwd_bwd(x, ei, p, H, C, gy, snapshot_mode=mode)
    assert rel_err(y_eval, y_ref0) <= TOL_F32                   # eval() disables dropout


@pytest.mark.parametrize("S,N,F,H,C,E", [(3, 63, 22, 2, 11, 0), (2, 200, 10, 2, 5, 1500)])
def test_bf16_autocast(cuda_device, S, N, F, H, C, E):
    """bf16-autocast contract (train.py:68).  The forward must sit within 1e-2 of PyG's dtype flow (the oracle run under
    CPU autocast).  PyG's autocast BACKWARD accumulates the gathered gradients in bf16 and is itself 4-15 % away from the
    fp64 truth (measured: lin_r.weight 0.15), so for gradients the gate is the truth: ours must be within 1e-2 of fp64,
    or as close to it as PyG-autocast is (25 % noise margin: both store d xl / d xr in bf16), and never further from
    PyG-autocast than twice PyG-autocast's own distance from the truth."""
    if E == 0:
        ei = torch.from_numpy(load_golden("graph_small150.npz")["edge_index"])
    else:
        ei = random_graph(N, E, seed=21)
    x, gy, p = _rand_case(S, N, F, H, C, seed=22, dtype=torch.float32)
    enc = _encoder(F, H, C, p, cuda_device).eval()
    y, grads = _run_cuda(enc, x, ei, gy, autocast=True)
    y_ac, g_ac = G.fwd_bwd(x, ei, p, H, C, gy, autocast_bf16=True)       # PyG's dtype flow on CPU
    y64, g64 = G.fwd_bwd(x.double(), ei, {k: v.double() for k, v in p.items()}, H, C, gy.double())
    assert y.dtype == torch.float32
    e = rel_err(y, y_ac)
    assert e <= TOL_BF16, f"y vs autocast oracle: {e:.3e}"
    assert rel_err(y, y64) <= TOL_BF16
    for k in g64:
        ours, theirs = rel_err(grads[k], g64[k]), rel_err(g_ac[k], g64[k])
        print(f"bf16 grad {k}: ours vs fp64 {ours:.3e}; PyG-autocast vs fp64 {theirs:.3e}")
        assert ours <= max(TOL_BF16, 1.25 * theirs), f"grad {k}: {ours:.3e} (PyG-autocast itself: {theirs:.3e})"
        assert rel_err(grads[k], g_ac[k]) <= max(TOL_BF16, 2.0 * theirs), k


def test_full_size_properties(cuda_device):
    """BASELINE config 1 size (B=2 x 48 snapshots x 2911 nodes): size-independent properties instead of a full oracle run.
    (1) snapshot permutation equivariance, bit-exact; (2) a replicated snapshot gives replicated outputs and input grads;
    (3) parameter gradients add over snapshot shards; (4) oracle parity on a random subset of snapshots."""
    g = load_golden("graph_cn150.npz")
    ei = torch.from_numpy(g["edge_index"])
    S, N, F, H, C = 96, 2911, 22, 2, 11
    x, gy, p = _rand_case(S, N, F, H, C, seed=31, dtype=torch.float32)
    enc = _encoder(F, H, C, p, cuda_device).eval()
    y, grads = _run_cuda(enc, x, ei, gy)
    perm = torch.randperm(S, generator=torch.Generator().manual_seed(1))
    y_p, grads_p = _run_cuda(enc, x[perm], ei, gy[perm])
    assert torch.equal(y_p.cpu(), y.cpu()[perm])
    assert torch.equal(grads_p["x"].cpu(), grads["x"].cpu()[perm])
    xr = x.clone(); xr[1] = xr[0]
    gr = gy.clone(); gr[1] = gr[0]
    y_r, grads_r = _run_cuda(enc, xr, ei, gr)
    assert torch.equal(y_r[0], y_r[1]) and torch.equal(grads_r["x"][0], grads_r["x"][1])
    half = S // 2
    _, ga = _run_cuda(enc, x[:half], ei, gy[:half])
    ga = {k: v.clone() for k, v in ga.items()}
    _, gb = _run_cuda(enc, x[half:], ei, gy[half:])
    for k in G.PARAM_NAMES:
        assert rel_err(ga[k] + gb[k], grads[k]) <= 2e-6, k
    sub = [0, 17, 95]
    y_ref, g_ref, _ = oracle_with_kernel_branches(x[sub], ei, p, H, C, gy[sub], cuda_device)
    assert rel_err(y.cpu()[sub], y_ref) <= TOL_F32
    assert rel_err(grads["x"].cpu()[sub], g_ref["x"]) <= TOL_F32


def test_gatv2conv_direct_call_is_pyg_semantics(cuda_device):
    """GATv2Conv.forward(x2d, edge_index) on the flattened input == the reference's literal call (modules.py:356)."""
    from tec_mollm_b200 import GATv2Conv

    S, N, F, H, C = 3, 30, 8, 2, 4
    ei = random_graph(N, 120, seed=40)
    x, gy, p = _rand_case(S, N, F, H, C, seed=41)
    conv = GATv2Conv(F, C, heads=H, dropout=0.0, concat=True, add_self_loops=True).to(cuda_device).eval()
    conv.load_state_dict({k: v.float() for k, v in p.items()}, strict=True)
    y = conv(x.reshape(-1, F).float().to(cuda_device), ei.to(cuda_device))
    y_ref = G.spatial_encoder_forward(x, ei, p, H, C, "literal").reshape(-1, H * C)
    assert rel_err(y, y_ref) <= TOL_F32


def test_works_inside_grad_scaler_and_optimizer_step(cuda_device):
    """train.py:68-110 shape of use: autocast + GradScaler + AdamW step; loss must go down on a toy target."""
    from tec_mollm_b200 import SpatialEncoder

    torch.manual_seed(0)
    S, N, F, H, C = 4, 63, 22, 2, 11
    ei = torch.from_numpy(load_golden("graph_small150.npz")["edge_index"]).to(cuda_device)
    enc = SpatialEncoder(F, C, heads=H, dropout=0.1).to(cuda_device).train()
    opt = torch.optim.AdamW(enc.parameters(), lr=1e-2)
    scaler = torch.amp.GradScaler("cuda")
    x = torch.randn(S, N, F, device=cuda_device)
    target = torch.randn(S, N, H * C, device=cuda_device)
    losses = []
    for _ in range(30):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = x + enc(x, ei)                                  # the residual of tec_mollm.py:94
            loss = torch.nn.functional.huber_loss(out, target)
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(enc.parameters(), 1.0)
        scaler.step(opt)
        scaler.update()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_cuda_graph_capture_and_replay(cuda_device):
    """The whole forward + backward is capturable in a CUDA graph once the plan exists (no hidden syncs, no host
    callbacks, workspaces from the capturing allocator), and replays reproduce the eager results bit for bit."""
    ei = torch.from_numpy(load_golden("graph_small150.npz")["edge_index"]).to(cuda_device)
    N = int(ei.max().item()) + 1
    S, F, H, C = 6, 22, 2, 11
    x, gy, p = _rand_case(S, N, F, H, C, seed=5, dtype=torch.float32)
    enc = _encoder(F, H, C, p, cuda_device).eval()
    xs = x.to(cuda_device).requires_grad_(True)
    gs = gy.to(cuda_device)
    side = torch.cuda.Stream(cuda_device)
    side.wait_stream(torch.cuda.current_stream(cuda_device))
    with torch.cuda.stream(side):          # warm-up on a side stream, as torch.cuda.graph requires: builds the plan,
        for _ in range(3):                 # sets the kernel attributes, creates the AccumulateGrad nodes on this stream
            xs.grad = None
            enc.zero_grad(set_to_none=True)
            enc(xs, ei).backward(gs)
        ref = {"y": enc(xs, ei).detach().clone(), "x": xs.grad.clone(), **{k: q.grad.clone() for k, q in enc.named_parameters()}}
    torch.cuda.current_stream(cuda_device).wait_stream(side)
    xs.grad = None
    enc.zero_grad(set_to_none=True)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        yg = enc(xs, ei)
        yg.backward(gs)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize(cuda_device)
    assert torch.equal(yg, ref["y"]) and torch.equal(xs.grad, ref["x"])
    for k, q in enc.named_parameters():
        assert torch.equal(q.grad, ref[k]), k


def test_four_dimensional_input_keeps_leading_dims(cuda_device):
    """(B, L, N, C) straight from the embedding (no permute copy): same numbers as the reference's (L*B, N, C) call,
    snapshot for snapshot."""
    ei = torch.from_numpy(load_golden("graph_small150.npz")["edge_index"]).to(cuda_device)
    N = int(ei.max().item()) + 1
    B, L, F, H, C = 2, 3, 22, 2, 11
    x, gy, p = _rand_case(B * L, N, F, H, C, seed=9, dtype=torch.float32)
    enc = _encoder(F, H, C, p, cuda_device).eval()
    x4 = x.view(B, L, N, F).to(cuda_device)
    y4 = enc(x4, ei)
    assert y4.shape == (B, L, N, H * C)
    y3 = enc(x4.permute(1, 0, 2, 3).reshape(L * B, N, F), ei)          # tec_mollm.py:84
    assert torch.equal(y3.view(L, B, N, H * C).permute(1, 0, 2, 3), y4)


@pytest.mark.gpu
@pytest.mark.parametrize("B,L,H,C", [(2, 3, 2, 11), (3, 48, 2, 11), (2, 5, 1, 3), (1, 4, 4, 11)])
def test_forward_block_equals_the_reference_glue(cuda_device, monkeypatch, B, L, H, C):
    """forward_block (fused residual + permute, tec_mollm.py:84-106) against the reference's own three lines run with
    torch ops around the same encoder: bit-identical output and input gradient, equal parameter gradients."""
    monkeypatch.setenv("TECGAT_HEAD_SPLIT", "0")  # bit-identity holds between the SAME kernels (forward_block never splits head pairs)
    ei = torch.from_numpy(load_golden("graph_small150.npz")["edge_index"]).to(cuda_device)
    N = int(ei.max().item()) + 1
    F = H * C
    x, _, p = _rand_case(B * L, N, F, H, C, seed=21, dtype=torch.float32)
    enc = _encoder(F, H, C, p, cuda_device).eval()
    gz = torch.randn(B * N, L, F, generator=torch.Generator().manual_seed(5)).to(cuda_device)
    xa = x.view(B, L, N, F).to(cuda_device).requires_grad_(True)
    za = enc.forward_block(xa, ei)
    assert za.shape == (B * N, L, F)
    za.backward(gz)
    ga = {k: q.grad.clone() for k, q in enc.named_parameters()}
    for q in enc.parameters():
        q.grad = None
    xb = x.view(B, L, N, F).to(cuda_device).requires_grad_(True)
    x_for_gnn = xb.permute(1, 0, 2, 3).reshape(-1, N, F)                       # tec_mollm.py:84
    x_spatial = x_for_gnn + enc(x_for_gnn, ei, None)                           # :89-94
    zb = x_spatial.view(L, B, N, F).permute(1, 2, 0, 3).reshape(-1, L, F)      # :100-106
    zb.backward(gz)
    assert torch.equal(za, zb)
    assert torch.equal(xa.grad, xb.grad)
    for k, q in enc.named_parameters():  # the snapshots are summed in a different order (b*L + l instead of l*B + b)
        assert rel_err(ga[k], q.grad) <= 2e-6, k


@pytest.mark.gpu
def test_forward_block_rejects_bad_input(cuda_device):
    from tec_mollm_b200 import SpatialEncoder
    enc = SpatialEncoder(22, 11, heads=2).to(cuda_device)
    ei = torch.tensor([[0, 1], [1, 0]], device=cuda_device)
    with pytest.raises(ValueError):
        enc.forward_block(torch.zeros(2, 3, 22, device=cuda_device), ei)
    with pytest.raises(ValueError):
        enc.forward_block(torch.zeros(1, 2, 3, 10, device=cuda_device), ei)
    with pytest.raises(RuntimeError):
        enc.forward_block(torch.zeros(1, 2, 3, 22), ei.cpu())
