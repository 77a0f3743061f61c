"""Shared helpers for the parity tests (checker side only)."""
import ctypes as C
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return np.load(os.path.join(GOLD, name))


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    """max |a - ref| / max |ref|  (the 'relative' of BASELINE.json's parity bar)."""
    a = a.detach().double().cpu()
    ref = ref.detach().double().cpu()
    denom = ref.abs().max().item()
    return (a - ref).abs().max().item() / (denom if denom > 0 else 1.0)


def random_graph(num_nodes, num_edges, seed, self_loops=True, duplicates=True, isolated=()):
    """Arbitrary directed multigraph: not symmetric, not sorted, optional self loops / duplicate edges /
    isolated nodes (the edge cases SURVEY.md Appendix A lists)."""
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, num_nodes, (2, num_edges), generator=g)
    if duplicates and num_edges >= 4:
        ei[:, -2:] = ei[:, :2]
    if self_loops and num_edges >= 6:
        ei[1, 2:4] = ei[0, 2:4]
    if not self_loops:
        ei = ei[:, ei[0] != ei[1]]
    for n in isolated:
        ei = ei[:, (ei[0] != n) & (ei[1] != n)]
    return ei


def kernel_dropout_mask(plan, snapshots, heads, p, seed, mode="shared"):
    """The keep-mask the CUDA kernels use, re-expressed in the oracle's edge order.

    Kernel counter = snapshot * E + csr_slot; oracle ("shared" mode, block-diagonal edge list) orders edges as
    [kept edges of snapshot 0, ..., kept edges of snapshot S-1, self loops of all S*N rows].
    Returns a float tensor (S*E, heads) in the oracle's order.
    """
    from tec_mollm_b200 import _lib

    E = plan.num_edges
    N = plan.num_nodes
    kept = plan.kept_edges
    _, _, eid = plan.export()
    keep = np.empty((snapshots * E, heads), dtype=np.uint8)
    _lib.call("tecgat_dropout_mask_host", C.c_uint64(seed), 0, snapshots * E, heads, C.c_float(p),
              C.c_void_p(keep.ctypes.data))
    keep = keep.reshape(snapshots, E, heads)
    eid = eid.astype(np.int64)
    is_self = eid >= kept
    if mode == "shared":
        out = np.empty((snapshots * E, heads), dtype=np.float32)
        for s in range(snapshots):
            # kept edge with PyG id e of snapshot s sits at s*kept + e; self loop of node i at S*kept + s*N + i
            pos = np.where(is_self, snapshots * kept + s * N + (eid - kept), s * kept + eid)
            out[pos] = keep[s]
    else:  # literal: the oracle's edge list is [kept edges (snapshot 0 only), self loops of all S*N rows]
        out = np.empty((kept + snapshots * N, heads), dtype=np.float32)
        out[np.where(is_self, kept + (eid - kept), eid)] = keep[0]
        for s in range(1, snapshots):
            out[kept + s * N + (eid[is_self] - kept)] = keep[s][is_self]
    return torch.from_numpy(out)
