#!/usr/bin/env python3
"""Bench leg of SURVEY.md 8f N3: the TemporalEncoder's conv embedder (modules.py:62-91) at BASELINE config 2's shape
(B = 2 -> B*N = 5,822 sequences of 48 steps x 22 channels), forward + backward.

  * `torch_ops`  : the reference's op sequence (3 Conv1d + GroupNorm + GELU per block, cat, strided 1x1) as torch runs it;
  * `drop_in`    : tec_mollm_b200.MultiScaleConvEmbedder (one stacked 7-tap library convolution + the fused pass + 1x1);
  * `fused_pass` : the hand-written pass alone with its HBM roofline (algorithmic bytes: forward reads y and writes z at the
                   strided positions; backward reads y and d z and writes d y), CUDA-event timed;
  * `cpu_baseline`: the reference's op sequence on the host cores, a bounded sample of the sequences.
One JSON line per contract (fp32 with TF32 off, bf16 autocast).   gpurun -- 'python tools/temporal_bench.py > gpurun_out/temporal.jsonl'
"""
import json
import os
import sys
import time

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tec_mollm_b200 import MultiScaleConvEmbedder  # noqa: E402
from tec_mollm_b200.temporal import _GnGeluStride  # noqa: E402


class MSBlock(nn.Module):  # the reference's structure (modules.py:13-60) with torch ops
    def __init__(self, cin, cout):
        super().__init__()
        self.convs = nn.ModuleList([nn.Sequential(nn.Conv1d(cin, cout, k, padding=(k - 1) // 2), nn.GroupNorm(1, cout), nn.GELU())
                                    for k in (3, 5, 7)])
        self.final_conv = nn.Conv1d(3 * cout, cout, 1, stride=2)

    def forward(self, x):
        return self.final_conv(torch.cat([c(x) for c in self.convs], dim=1))


def timed(fn, dev, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda", 0)
    Bn, L, C = 2 * 2911, 48, 22
    peak = 6538.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    torch.manual_seed(0)
    ref = nn.Sequential(MSBlock(22, 64), MSBlock(64, 128)).to(dev)
    ours = MultiScaleConvEmbedder(22, [64, 128], [2, 2]).to(dev)
    ours.embedder.load_state_dict(ref.state_dict())
    x = torch.randn(Bn, C, L, device=dev, requires_grad=True)
    gy = torch.randn(Bn, 128, 12, device=dev)
    cpu = nn.Sequential(MSBlock(22, 64), MSBlock(64, 128))
    ns = 256
    xc = torch.randn(ns, C, L, requires_grad=True)
    gc = torch.randn(ns, 128, 12)
    t0 = time.perf_counter()
    for _ in range(3):
        cpu(xc).backward(gc)
    cpu_s = (time.perf_counter() - t0) / 3
    for autocast in (False, True):
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False

        def step(m):
            def fn():
                x.grad = None
                m.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    y = m(x)
                y.float().backward(gy)
            return fn

        ms_ref, ms_ours = timed(step(ref), dev), timed(step(ours), dev)
        # the fused pass alone, both blocks' shapes
        passes = {}
        for name, (ch, ln) in (("block1", (64, 48)), ("block2", (128, 24))):
            ydt = torch.bfloat16 if autocast else torch.float32
            y = torch.randn(Bn, 3 * ch, ln, device=dev).to(ydt).requires_grad_(True)
            gamma = torch.ones(3, ch, device=dev, requires_grad=True)
            beta = torch.zeros(3, ch, device=dev, requires_grad=True)
            holder = {}

            def fwd():
                holder["z"] = _GnGeluStride.apply(y, gamma, beta, 3, 2, 1e-5, ydt)

            fwd()
            gz = torch.randn_like(holder["z"])

            def both():
                y.grad = None
                fwd()
                holder["z"].backward(gz)

            ms_f = timed(fwd, dev, 50)
            ms_fb = timed(both, dev, 50)
            es = y.element_size()
            bytes_f = Bn * 3 * ch * (ln * es + (ln // 2) * es)
            bytes_b = Bn * 3 * ch * (ln * es + (ln // 2) * es + ln * es)
            passes[name] = {"fwd_ms": ms_f, "bwd_ms": ms_fb - ms_f,
                            "fwd": {"bound": "hbm", "achieved": bytes_f / (ms_f * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": bytes_f / (ms_f * 1e-3) / 1e9 / peak, "algorithmic_bytes": bytes_f},
                            "bwd": {"bound": "hbm", "achieved": bytes_b / ((ms_fb - ms_f) * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": bytes_b / ((ms_fb - ms_f) * 1e-3) / 1e9 / peak, "algorithmic_bytes": bytes_b}}
        print(json.dumps({
            "what": "temporal_conv_embedder_fwd_bwd", "sequences": Bn, "contract": "bf16 autocast" if autocast else "fp32 (TF32 off)",
            "torch_ops_ms": ms_ref, "drop_in_ms": ms_ours, "speedup": ms_ref / ms_ours, "sequences_per_s": Bn / (ms_ours * 1e-3),
            "fused_pass": passes,
            "cpu_baseline": {"value": ns / cpu_s, "unit": "sequences/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{ns} of {Bn} sequences, fp32, the reference's op sequence (torch CPU), fwd+bwd {cpu_s * 1e3:.0f} ms"}}), flush=True)


if __name__ == "__main__":
    main()
