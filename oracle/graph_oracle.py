"""CPU oracle for the haversine adjacency builder (``/root/reference/src/graph/graph_constructor.py``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PINNED: checked against golden
vectors produced by the reference's own functions (``tools/make_golden.py`` ->
``tests/golden/graph_*.npz``) in ``tests/test_oracle_graph.py``.

The arithmetic of the reference lives in scikit-learn (``haversine_distances``,
``graph_constructor.py:5,56``) and scipy.sparse (``:4,112-128``); both are in this image
(sklearn 1.9.0, scipy 1.18.1 -- the reference pins no version, README.md:57), so the
restatement calls the same third-party routines in the same order and adds

* a row-blocked variant (``haversine_distances(X_blk, X)``; bit-equal to the
  ``pdist``-mirror path) so the 64,800-node grid of BASELINE config 5
  never needs the 33.6 GB dense matrix, and
* a sparse normalisation that never forms the dense ``(N, N)`` int64 matrix.
"""
from __future__ import annotations

from math import radians
from typing import Tuple

import numpy as np
from scipy.sparse import coo_matrix, diags
from sklearn.metrics.pairwise import haversine_distances

EARTH_RADIUS_KM = 6371.0  # graph_constructor.py:53


def node_coords_rad(lat: np.ndarray, lon: np.ndarray) -> np.ndarray:
    """graph_constructor.py:46-50 -- lat-major node order, ``math.radians`` per element."""
    lon_grid, lat_grid = np.meshgrid(lon, lat)
    coords = np.vstack([lat_grid.ravel(), lon_grid.ravel()]).T
    return np.array([[radians(c[0]), radians(c[1])] for c in coords])


def haversine_matrix(lat: np.ndarray, lon: np.ndarray) -> np.ndarray:
    """graph_constructor.py:34-59 -- dense (N, N) fp64 distance matrix in km."""
    return haversine_distances(node_coords_rad(lat, lon)) * EARTH_RADIUS_KM


def binary_adjacency(distance_matrix: np.ndarray, thr_km: float = 150.0) -> np.ndarray:
    """graph_constructor.py:61-81 -- inclusive threshold, int64, zero diagonal."""
    adj = (distance_matrix <= thr_km).astype(int)
    np.fill_diagonal(adj, 0)
    return adj


def sym_normalize(adj) -> coo_matrix:
    """graph_constructor.py:99-128 -- ``D^-1/2 A D^-1/2`` through scipy.sparse, COO out
    (row-major order: rows ascending, columns ascending inside a row)."""
    adj_sparse = coo_matrix(adj)
    deg = np.array(adj_sparse.sum(axis=1)).flatten()
    with np.errstate(divide="ignore"):
        inv_sqrt = 1.0 / np.sqrt(deg)
    inv_sqrt[np.isinf(inv_sqrt)] = 0
    d = diags(inv_sqrt)
    return d.dot(adj_sparse).dot(d).tocoo()


def to_edge_arrays(norm: coo_matrix) -> Tuple[np.ndarray, np.ndarray]:
    """graph_constructor.py:141-144 -- ``edge_index`` int64 (2, E) = (row, col);
    ``edge_weight`` fp64 -> fp32."""
    edge_index = np.vstack((norm.row, norm.col)).astype(np.int64)
    edge_weight = norm.data.astype(np.float32)
    return edge_index, edge_weight


def graph_edges_dense(lat: np.ndarray, lon: np.ndarray, thr_km: float = 150.0):
    """The reference pipeline ``__main__`` (graph_constructor.py:165-214) end to end."""
    D = haversine_matrix(lat, lon)
    return to_edge_arrays(sym_normalize(binary_adjacency(D, thr_km)))


def graph_edges_blocked(coords_rad: np.ndarray, thr_km: float = 150.0, block: int = 1024,
                        row_range=None):
    """Same result as :func:`graph_edges_dense` without the dense matrix: distances are
    evaluated one row block at a time with the same sklearn routine (bit-equal values).
    ``row_range=(r0, r1)`` restricts the *rows* for which edges are emitted (used by the
    bounded CPU-baseline sample); degrees then cover only those rows, so edge weights are
    returned only when ``row_range`` is None."""
    n = coords_rad.shape[0]
    r0, r1 = (0, n) if row_range is None else row_range
    rows, cols = [], []
    for b0 in range(r0, r1, block):
        b1 = min(b0 + block, r1)
        d = haversine_distances(coords_rad[b0:b1], coords_rad) * EARTH_RADIUS_KM
        mask = d <= thr_km
        rr, cc = np.nonzero(mask)
        rr = rr + b0
        keep = rr != cc
        rows.append(rr[keep])
        cols.append(cc[keep])
    row = np.concatenate(rows).astype(np.int64)
    col = np.concatenate(cols).astype(np.int64)
    edge_index = np.vstack((row, col))
    if row_range is not None:
        return edge_index, None
    deg = np.bincount(row, minlength=n).astype(np.float64)
    with np.errstate(divide="ignore"):
        inv_sqrt = 1.0 / np.sqrt(deg)
    inv_sqrt[np.isinf(inv_sqrt)] = 0
    # scipy evaluates (d_r * a) * d_c with a = 1 (fp64): same two multiplies here.
    w = ((inv_sqrt[row] * 1.0) * inv_sqrt[col]).astype(np.float32)
    return edge_index, w


def synthetic_grid(kind: str = "cn"):
    """The synthetic lat/lon axes used everywhere: ``"cn"`` = the
    41 x 71, 1-degree China-region grid standing in for the absent HDF5 coordinates;
    ``"global"`` = 180 x 360 cell-centred 1-degree grid (64,800 nodes)."""
    if kind == "cn":
        return np.arange(15.0, 56.0, 1.0), np.arange(70.0, 141.0, 1.0)
    if kind == "global":
        return np.arange(-89.5, 90.0, 1.0), np.arange(-179.5, 180.0, 1.0)
    raise ValueError(kind)
