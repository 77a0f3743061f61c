"""Graph builder: mirror of ``/root/reference/src/graph/graph_constructor.py`` on the GPU.

The five reference functions keep their names, signatures and return types
(``calculate_haversine_distance_matrix`` :34, ``construct_binary_adjacency`` :61,
``compute_degree_matrix`` :83, ``symmetrically_normalize_adjacency`` :99,
``convert_to_pyg_and_save`` :130, plus ``get_coordinates_from_data`` :15) so the reference's
``__main__`` pipeline runs unchanged against this module; the dense ``(N, N)`` objects they trade in
are produced by the CUDA kernels and only materialised because that API demands them.

The product path is :func:`build_graph` / :func:`build_and_save_graph`: coordinates -> ``edge_index`` /
``edge_weight`` directly (``tecgraph_edges_count`` + ``tecgraph_edges_fill``), never forming an
``(N, N)`` matrix, with the reference's exact edge set, order (scipy COO row-major) and fp32 weights.
"""
from __future__ import annotations

import ctypes as C
import logging
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib

EARTH_RADIUS_KM = 6371.0  # graph_constructor.py:53


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("tec_mollm_b200.graph: a CUDA device is required (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def node_coordinates_rad(lat: np.ndarray, lon: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Per-node (lat, lon) in radians, lat-major node order ``i_lat * len(lon) + i_lon``
    (graph_constructor.py:46-50; ``np.radians`` is bit-identical to the reference's per-element ``math.radians``)."""
    lon_grid, lat_grid = np.meshgrid(np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64))
    return np.radians(lat_grid.ravel()), np.radians(lon_grid.ravel())


def build_graph_from_nodes(lat_rad, lon_rad, distance_threshold_km: float = 150.0, device=None,
                           radius_km: float = EARTH_RADIUS_KM, return_stats: bool = False):
    """Edges between arbitrary nodes given per-node coordinates in radians (fp64).  Returns
    ``(edge_index int64 (2, E), edge_weight float32 (E))`` on the device."""
    dev = _device(device)
    lat_t = torch.as_tensor(np.ascontiguousarray(lat_rad, dtype=np.float64)).to(dev)
    lon_t = torch.as_tensor(np.ascontiguousarray(lon_rad, dtype=np.float64)).to(dev)
    n = lat_t.numel()
    if lon_t.numel() != n or n == 0:
        raise ValueError("lat_rad and lon_rad must be non-empty and of equal length")
    ctx, total, amb = C.c_void_p(), C.c_int64(), C.c_int64()
    with torch.cuda.device(dev):
        _lib.call("tecgraph_edges_count", _ptr(lat_t), _ptr(lon_t), n, float(distance_threshold_km), float(radius_km),
                  _stream(dev), C.byref(ctx), C.byref(total), C.byref(amb))
        try:
            edge_index = torch.empty((2, total.value), dtype=torch.int64, device=dev)
            edge_weight = torch.empty((total.value,), dtype=torch.float32, device=dev)
            _lib.call("tecgraph_edges_fill", ctx, _ptr(edge_index), _ptr(edge_weight), _stream(dev))
            torch.cuda.current_stream(dev).synchronize()  # the context owns device buffers the kernel reads
            st4 = (C.c_double * 4)()
            if return_stats:
                _lib.call("tecgraph_ctx_stats", ctx, st4)
        finally:
            _lib.lib().tecgraph_ctx_destroy(ctx)
    if return_stats:
        return edge_index, edge_weight, {"guard_band_pairs": int(amb.value), "num_nodes": n, "count_kernel_ms": st4[0],
                                         "fill_kernel_ms": st4[1], "kernel_ms": st4[0] + st4[1], "evaluated_pairs": int(st4[2])}
    return edge_index, edge_weight


def build_graph(lat: np.ndarray, lon: np.ndarray, distance_threshold_km: float = 150.0, device=None,
                return_stats: bool = False):
    """Reference pipeline (graph_constructor.py:165-214) in one call: 1-D ``lat`` / ``lon`` axes in degrees ->
    ``edge_index``, ``edge_weight`` on the device, bit-identical to the reference's ``graph_A.pt`` contents."""
    la, lo = node_coordinates_rad(lat, lon)
    return build_graph_from_nodes(la, lo, distance_threshold_km, device, return_stats=return_stats)


def build_and_save_graph(lat, lon, output_path: str, distance_threshold_km: float = 150.0, device=None):
    """Writes the same ``{'edge_index', 'edge_weight'}`` dict the reference saves (graph_constructor.py:147)."""
    ei, ew = build_graph(lat, lon, distance_threshold_km, device)
    torch.save({"edge_index": ei.cpu(), "edge_weight": ew.cpu()}, output_path)
    logging.info(f"Graph data saved successfully. Edges: {ei.shape[1]}")
    return ei, ew


# ------------------------------------------------------------------------------------------------
# reference-compatible API (dense objects)
# ------------------------------------------------------------------------------------------------
def get_coordinates_from_data(file_paths: list):
    """graph_constructor.py:15-32.  The HDF5 loader belongs to the reference's data layer (out of scope here);
    it is used when importable, otherwise this raises."""
    try:
        from src.data.data_loader import load_and_split_data  # the reference's own loader
    except Exception as exc:  # pragma: no cover - depends on the user's environment
        raise RuntimeError("get_coordinates_from_data needs the reference's src.data.data_loader (h5py + HDF5 files); "
                           "pass latitude / longitude arrays to build_graph() instead") from exc
    data = load_and_split_data(file_paths)
    if not data:
        logging.error("Failed to load data to get coordinates.")
        return None, None
    if "latitude" in data["train"] and "longitude" in data["train"]:
        return data["train"]["latitude"], data["train"]["longitude"]
    logging.error("Latitude or Longitude not found in the loaded data.")
    return None, None


def calculate_haversine_distance_matrix(lat: np.ndarray, lon: np.ndarray, device=None) -> np.ndarray:
    """graph_constructor.py:34-59: dense (N, N) fp64 distance matrix in km, computed on the GPU
    (``tecgraph_distance_rows``).  Values agree with scikit-learn's to ~1 ulp (CUDA vs glibc libm); use
    :func:`build_graph` when the bit-exact edge set is what matters."""
    dev = _device(device)
    la, lo = node_coordinates_rad(lat, lon)
    n = la.size
    logging.info(f"Calculating pairwise Haversine distances for {n} nodes...")
    lat_t, lon_t = torch.from_numpy(la).to(dev), torch.from_numpy(lo).to(dev)
    out = torch.empty((n, n), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        for r0 in range(0, n, 32768):
            r1 = min(n, r0 + 32768)
            _lib.call("tecgraph_distance_rows", _ptr(lat_t), _ptr(lon_t), n, r0, r1, EARTH_RADIUS_KM,
                      C.c_void_p(out.data_ptr() + r0 * n * 8), _stream(dev))
    distance_matrix = out.cpu().numpy()
    logging.info(f"Calculated distance matrix with shape: {distance_matrix.shape}")
    return distance_matrix


def construct_binary_adjacency(distance_matrix: np.ndarray, distance_threshold_km: float = 150.0) -> np.ndarray:
    """graph_constructor.py:61-81: inclusive threshold, int64, zero diagonal."""
    logging.info(f"Constructing binary adjacency matrix with threshold {distance_threshold_km} km...")
    adj_matrix = (distance_matrix <= distance_threshold_km).astype(int)
    np.fill_diagonal(adj_matrix, 0)
    logging.info(f"Constructed binary adjacency matrix with {np.sum(adj_matrix)} edges.")
    return adj_matrix


def compute_degree_matrix(adj_matrix: np.ndarray) -> np.ndarray:
    """graph_constructor.py:83-97."""
    return np.diag(np.sum(adj_matrix, axis=1))


def symmetrically_normalize_adjacency(adj_matrix: np.ndarray):
    """graph_constructor.py:99-128: ``D^-1/2 A D^-1/2`` as a ``scipy.sparse.coo_matrix`` in row-major order, values
    ``(d_r * a) * d_c`` in fp64 (the association scipy's two sparse products produce)."""
    from scipy.sparse import coo_matrix

    adj = np.asarray(adj_matrix)
    row, col = np.nonzero(adj)
    vals = adj[row, col].astype(np.float64)
    deg = adj.sum(axis=1).astype(np.float64)
    with np.errstate(divide="ignore"):
        inv_sqrt = 1.0 / np.sqrt(deg)
    inv_sqrt[np.isinf(inv_sqrt)] = 0
    data = (inv_sqrt[row] * vals) * inv_sqrt[col]
    return coo_matrix((data, (row, col)), shape=adj.shape)


def convert_to_pyg_and_save(normalized_adj, output_path: str):
    """graph_constructor.py:130-149: same tensors, same dict, same ``torch.save``."""
    edge_index = torch.tensor(np.vstack((normalized_adj.row, normalized_adj.col)), dtype=torch.long)
    edge_weight = torch.tensor(normalized_adj.data, dtype=torch.float)
    torch.save({"edge_index": edge_index, "edge_weight": edge_weight}, output_path)
    logging.info(f"Graph data saved successfully. Edges: {edge_index.shape[1]}")


def load_graph(path: str, device=None):
    """train.py:292-294: ``torch.load(graph_path)`` then ``.to(device)``."""
    g = torch.load(path)
    dev = _device(device)
    return g["edge_index"].to(dev), g["edge_weight"].to(dev)
