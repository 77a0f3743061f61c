"""Drop-in ``Multi_Scale_Conv_Block`` / ``MultiScaleConvEmbedder`` of the TemporalEncoder (reference:
``src/model/modules.py:13-91``).

Same constructors, same parameter names (``convs.{0,1,2}.0`` = Conv1d, ``convs.{0,1,2}.1`` = GroupNorm, ``final_conv``) so a
reference checkpoint loads ``strict=True``, same math.  What changes is the execution:

* the three branch convolutions (k = 3 / 5 / 7, "same" padding) run as ONE library convolution with the kernels zero-padded to
  seven taps and stacked along the output channels -- its output already is the reference's ``torch.cat`` layout;
* GroupNorm(1, C) + GELU + concat + the stride of the final 1x1 convolution run as one hand-written pass each way
  (``tecgat_gn_gelu_fwd`` / ``tecgat_gn_gelu_bwd``, csrc/temporal.cu): only the positions the strided 1x1 convolution reads are
  written, the backward is atomic-free;
* the final 1x1 convolution then is a plain dense contraction on the compacted tensor.

The dense contractions stay library calls (cuDNN / cuBLAS through torch); under ``torch.autocast(bf16)`` they run in bf16 and the
fused pass writes bf16 for the final convolution, exactly the dtype flow autocast gives the reference (GroupNorm and GELU in
fp32 on the convolution's bf16 output).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .gatv2 import _on_device, _ptr, _stream


def _code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return _lib.F32
    if dt == torch.bfloat16:
        return _lib.BF16
    raise NotImplementedError(f"tec_mollm_b200.temporal: dtype {dt} unsupported (float32 / bfloat16)")


class _GnGeluStride(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, gamma, beta, branches, stride, eps, out_dtype):
        n, ch, length = y.shape
        c = ch // branches
        y = y.contiguous()
        dev = y.device
        lo = (length + stride - 1) // stride
        z = torch.empty((n, ch, lo), device=dev, dtype=out_dtype)
        mean = torch.empty((n, branches), device=dev, dtype=torch.float32)
        rstd = torch.empty((n, branches), device=dev, dtype=torch.float32)
        with _on_device(dev):
            _lib.call("tecgat_gn_gelu_fwd", _ptr(y), _ptr(gamma), _ptr(beta), _ptr(z), _ptr(mean), _ptr(rstd), n, branches, c, length,
                      stride, eps, _code(y.dtype), _code(out_dtype), _stream(dev))
        ctx.save_for_backward(y, gamma, beta, mean, rstd)
        ctx.cfg = (branches, stride, c, length)
        return z

    @staticmethod
    def backward(ctx, dz):
        y, gamma, beta, mean, rstd = ctx.saved_tensors
        branches, stride, c, length = ctx.cfg
        n = y.size(0)
        dev = y.device
        dz = dz.contiguous()
        dy = torch.empty_like(y)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(beta)
        ws = torch.empty((max(1, _lib.lib().tecgat_gn_gelu_bwd_workspace(n, branches, c)),), device=dev, dtype=torch.uint8)
        with _on_device(dev):
            _lib.call("tecgat_gn_gelu_bwd", _ptr(y), _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(rstd), _ptr(dz), _ptr(dy), _ptr(dgamma),
                      _ptr(dbeta), _ptr(ws), n, branches, c, length, stride, _code(y.dtype), _code(dz.dtype), _stream(dev))
        return dy, dgamma, dbeta, None, None, None, None


class Multi_Scale_Conv_Block(nn.Module):  # noqa: N801 - the reference's class name (modules.py:13)
    def __init__(self, in_channels: int, out_channels: int, stride: int, kernel_sizes: list = [3, 5, 7]):  # noqa: B006
        super().__init__()
        if any(k % 2 == 0 for k in kernel_sizes):
            raise NotImplementedError("odd kernel sizes only ('same' padding, modules.py:25)")
        self.kernel_sizes, self.stride, self.out_channels = list(kernel_sizes), stride, out_channels
        self.convs = nn.ModuleList([
            nn.Sequential(nn.Conv1d(in_channels, out_channels, kernel_size=k, padding=(k - 1) // 2), nn.GroupNorm(1, out_channels), nn.GELU())
            for k in kernel_sizes])
        self.final_conv = nn.Conv1d(out_channels * len(kernel_sizes), out_channels, kernel_size=1, stride=stride)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``x``: (B, C_in, L) -> (B, C_out, ceil(L / stride))."""
        if not x.is_cuda:
            raise RuntimeError("tec_mollm_b200.Multi_Scale_Conv_Block needs CUDA tensors (there is no CPU path)")
        kmax = max(self.kernel_sizes)
        w = torch.cat([F.pad(seq[0].weight, ((kmax - k) // 2, (kmax - k) // 2)) for seq, k in zip(self.convs, self.kernel_sizes)], dim=0)
        b = torch.cat([seq[0].bias for seq in self.convs], dim=0)
        y = F.conv1d(x, w, b, padding=(kmax - 1) // 2)                        # the three branches, already concatenated
        gamma = torch.stack([seq[1].weight for seq in self.convs]).float()
        beta = torch.stack([seq[1].bias for seq in self.convs]).float()
        autocast = torch.is_autocast_enabled("cuda")
        out_dtype = torch.get_autocast_dtype("cuda") if autocast else torch.float32
        if y.dtype not in (torch.float32, torch.bfloat16):
            y = y.float()
        z = _GnGeluStride.apply(y, gamma, beta, len(self.convs), self.stride, float(self.convs[0][1].eps), out_dtype)
        return F.conv1d(z, self.final_conv.weight, self.final_conv.bias)       # 1x1, stride already applied


class MultiScaleConvEmbedder(nn.Module):
    """modules.py:62-91: the stack of Multi_Scale_Conv_Blocks of the TemporalEncoder."""

    def __init__(self, in_channels: int, channel_list: list, strides: list):
        super().__init__()
        assert len(channel_list) == len(strides), "Channel list and strides list must have the same length."
        layers, current = [], in_channels
        for out_channels, stride in zip(channel_list, strides):
            layers.append(Multi_Scale_Conv_Block(current, out_channels, stride))
            current = out_channels
        self.embedder = nn.Sequential(*layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.embedder(x)
