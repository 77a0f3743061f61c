#!/usr/bin/env python3
"""Run the forward projection repeatedly on the golden f22h2c11 input and report bitwise differences between runs."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import kernel_projection, load_golden  # noqa: E402
from oracle import gatv2_oracle as G  # noqa: E402

g = load_golden("gatv2_f22h2c11.npz")
F, H, C = int(g["F"]), int(g["H"]), int(g["C"])
params = {k: torch.from_numpy(g[f"p_{k}"]) for k in G.PARAM_NAMES}
x = torch.from_numpy(g["x"])
dev = torch.device("cuda", 0)
ref = None
for i in range(30):
    xl, xr = kernel_projection(x.reshape(-1, F), params, H, C, dev)
    cur = torch.cat([xl, xr], 1)
    if ref is None:
        ref = cur
    else:
        nd = int((cur != ref).sum())
        if nd:
            idx = (cur != ref).nonzero()[:5].tolist()
            print(f"run {i}: {nd} entries differ, e.g. {idx}, max abs diff {float((cur - ref).abs().max()):.3e}")
print("rows", x.reshape(-1, F).shape[0], "done")
