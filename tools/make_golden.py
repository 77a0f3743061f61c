#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (run in the build container only).

Graph fixtures come from the REFERENCE'S OWN functions, imported unmodified from
/root/reference/src/graph/graph_constructor.py (h5py is absent, so an empty stub module is
put in sys.modules first -- the functions used here never touch it).  GATv2 fixtures come
from the oracle restatement (the reference cannot be imported for that path: torch_geometric
and peft are absent), so they freeze the oracle rather than pin it; see oracle/__init__.py.

    python tools/make_golden.py
"""
import hashlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def reference_graph(lat, lon, thr):
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import src.graph.graph_constructor as gc  # the reference, unmodified

    D = gc.calculate_haversine_distance_matrix(lat, lon)
    A = gc.construct_binary_adjacency(D, thr)
    deg = gc.compute_degree_matrix(A)
    norm = gc.symmetrically_normalize_adjacency(A)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "graph_A.pt")
        gc.convert_to_pyg_and_save(norm, path)
        saved = torch.load(path)
    return D, A, deg, saved["edge_index"].numpy(), saved["edge_weight"].numpy()


def main():
    os.makedirs(GOLD, exist_ok=True)
    from oracle import graph_oracle as go
    from oracle import gatv2_oracle as gat

    # ---- graph goldens -------------------------------------------------------------
    lat_cn, lon_cn = go.synthetic_grid("cn")
    cases = {
        "cn150": (lat_cn, lon_cn, 150.0),
        "cn300": (lat_cn, lon_cn, 300.0),
        "small150": (np.arange(40.0, 47.0, 1.0), np.arange(100.0, 109.0, 1.0), 150.0),
        "small300": (np.arange(40.0, 47.0, 1.0), np.arange(100.0, 109.0, 1.0), 300.0),
        # irregular axes + a threshold that isolates some nodes (degree 0 -> weight path inf->0)
        "ragged120": (np.array([10.0, 10.5, 12.0, 30.0, 31.2, 80.0]), np.array([0.0, 1.0, 2.5, 90.0, 179.0]), 120.0),
    }
    for name, (lat, lon, thr) in cases.items():
        D, A, deg, ei, ew = reference_graph(lat, lon, thr)
        assert ei.dtype == np.int64 and ew.dtype == np.float32
        keep_D = D.shape[0] <= 64
        np.savez_compressed(
            os.path.join(GOLD, f"graph_{name}.npz"),
            lat=lat, lon=lon, thr=np.float64(thr), edge_index=ei, edge_weight=ew,
            degree=np.diag(deg).astype(np.int64),
            D_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(D).tobytes()).digest(), dtype=np.uint8),
            D=(D if keep_D else np.zeros((0, 0))),
            D_row0=D[0].copy(),
        )
        print(f"graph_{name}: N={D.shape[0]} E={ei.shape[1]} maxdeg={np.diag(deg).max()}")

    # ---- GATv2 goldens (oracle, fp64 + fp32) ----------------------------------------
    g = np.load(os.path.join(GOLD, "graph_small150.npz"))
    ei = torch.from_numpy(g["edge_index"])
    N = int(g["lat"].size * g["lon"].size)
    for name, (F_in, H, C, S) in {"f22h2c11": (22, 2, 11, 3), "f10h2c5": (10, 2, 5, 2), "f22h4c11": (22, 4, 11, 2)}.items():
        gen = torch.Generator().manual_seed(1234)
        x = torch.randn(S, N, F_in, generator=gen, dtype=torch.float64)
        gy = torch.randn(S, N, H * C, generator=gen, dtype=torch.float64)
        params = gat.init_params(F_in, C, H, seed=7, dtype=torch.float64)
        params["bias"] = torch.randn(H * C, generator=gen, dtype=torch.float64) * 0.1
        out = {}
        for mode in ("shared", "literal"):
            y, grads = gat.fwd_bwd(x, ei, params, H, C, gy, snapshot_mode=mode)
            out[f"y_{mode}"] = y.numpy()
            for k, v in grads.items():
                out[f"g_{mode}_{k}"] = v.numpy()
        np.savez_compressed(
            os.path.join(GOLD, f"gatv2_{name}.npz"), x=x.numpy(), gy=gy.numpy(), edge_index=ei.numpy(),
            F=F_in, H=H, C=C, S=S, N=N, **{f"p_{k}": v.numpy() for k, v in params.items()}, **out)
        print(f"gatv2_{name}: S={S} N={N}")


if __name__ == "__main__":
    main()
