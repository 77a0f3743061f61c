#!/usr/bin/env python3
"""N3 probe: the stacked 7-tap branch convolution of Multi_Scale_Conv_Block (modules.py:13-60) as a library convolution (NCL,
cuDNN) against the same contraction written as im2col + one dense GEMM on a channels-last tensor (the layout the TemporalEncoder's
caller already holds, tec_mollm.py:106).  fwd+bwd, 5822 sequences (B = 2), fp32 and bf16 autocast."""
import json
import sys

import torch
import torch.nn.functional as F

dev = torch.device("cuda", 0)
n = 5822


def timeit(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def im2col_cl(x, k):
    """x: (n, L, C) channels-last -> (n*L, k*C) rows of k consecutive (zero-padded) positions."""
    nn_, L, C = x.shape
    xp = F.pad(x, (0, 0, k // 2, k // 2))
    return xp.unfold(1, k, 1).permute(0, 1, 3, 2).reshape(nn_ * L, k * C)  # unfold -> (n, L, C, k)


for cin, cout, L in ((22, 192, 48), (64, 384, 24)):
    for autocast in (False, True):
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        x_ncl = torch.randn(n, cin, L, device=dev, requires_grad=True)
        w = torch.randn(cout, cin, 7, device=dev, requires_grad=True) * 0.1
        w = w.detach().requires_grad_(True)
        b = torch.randn(cout, device=dev, requires_grad=True)
        x_cl = x_ncl.detach().transpose(1, 2).contiguous().requires_grad_(True)
        w_g = w.detach().permute(2, 1, 0).reshape(7 * cin, cout).contiguous().requires_grad_(True)  # (k*C, cout)
        gy = torch.randn(n, cout, L, device=dev)
        gy_cl = gy.transpose(1, 2).reshape(n * L, cout).contiguous()

        def conv():
            for t in (x_ncl, w, b):
                t.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                y = F.conv1d(x_ncl, w, b, padding=3)
            y.backward(gy.to(y.dtype))

        def gemm():
            for t in (x_cl, w_g, b):
                t.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                y = F.linear(im2col_cl(x_cl, 7), w_g.t(), b)
            y.backward(gy_cl.to(y.dtype))

        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            ya = F.conv1d(x_ncl, w, b, padding=3).float()
            yb = F.linear(im2col_cl(x_cl, 7), w_g.t(), b).float().view(n, L, cout).transpose(1, 2)
        err = float((ya - yb).abs().max() / ya.abs().max())
        print(json.dumps({"cin": cin, "cout": cout, "L": L, "bf16_autocast": autocast, "conv1d_ms": timeit(conv), "im2col_gemm_ms": timeit(gemm),
                          "max_rel_diff": err}), flush=True)
