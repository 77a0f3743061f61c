"""CPU tests pinning the graph oracle to the REFERENCE: golden vectors under tests/golden/graph_*.npz were produced
by the reference's own functions (tools/make_golden.py imports /root/reference/src/graph/graph_constructor.py)."""
import ctypes
import hashlib
import os
import subprocess

import numpy as np
import pytest

from oracle import graph_oracle as go
from helpers import ROOT, load_golden

CASES = ["cn150", "cn300", "small150", "small300", "ragged120"]


@pytest.fixture(scope="module")
def havref():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libhavref.so"))
    lib.havref_dist_km.restype = ctypes.c_double
    lib.havref_dist_km.argtypes = [ctypes.c_double] * 4
    P = ctypes.POINTER(ctypes.c_double)
    lib.havref_rows.argtypes = [P, P, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, P]
    return lib


@pytest.mark.parametrize("name", CASES)
def test_dense_pipeline_matches_reference_golden(name):
    g = load_golden(f"graph_{name}.npz")
    ei, ew = go.graph_edges_dense(g["lat"], g["lon"], float(g["thr"]))
    assert ei.dtype == np.int64 and ew.dtype == np.float32
    assert np.array_equal(ei, g["edge_index"])            # bit-exact edge set AND order
    assert np.array_equal(ew, g["edge_weight"])           # bit-exact fp32 weights
    D = go.haversine_matrix(g["lat"], g["lon"])
    assert hashlib.sha256(np.ascontiguousarray(D).tobytes()).digest() == g["D_sha256"].tobytes()
    assert np.array_equal(D[0], g["D_row0"])


@pytest.mark.parametrize("name", CASES)
def test_blocked_variant_matches_reference_golden(name):
    g = load_golden(f"graph_{name}.npz")
    coords = go.node_coords_rad(g["lat"], g["lon"])
    ei, ew = go.graph_edges_blocked(coords, float(g["thr"]), block=257)
    assert np.array_equal(ei, g["edge_index"]) and np.array_equal(ew, g["edge_weight"])


def test_structural_asserts_of_the_reference_self_test():
    """The asserts of graph_constructor.py:169-223, on the synthetic 41 x 71 grid."""
    lat, lon = go.synthetic_grid("cn")
    D = go.haversine_matrix(lat, lon)
    assert D.shape == (2911, 2911) and np.allclose(D, D.T) and np.all(np.diag(D) == 0)
    A = go.binary_adjacency(D)
    assert np.all((A == 0) | (A == 1)) and np.all(np.diag(A) == 0) and A.sum() == 20924
    norm = go.sym_normalize(A)
    dense = norm.toarray()
    assert np.allclose(dense, dense.T) and norm.min() >= 0 and norm.max() <= 1
    ei, ew = go.to_edge_arrays(norm)
    assert ei.shape == (2, norm.nnz) and ew.shape == (norm.nnz,)
    assert ei[:, :5].tolist() == [[0, 0, 1, 1, 1], [1, 71, 0, 2, 72]]
    deg = np.bincount(ei[1], minlength=2911)
    hist = {int(k): int(v) for k, v in zip(*np.unique(deg, return_counts=True))}
    assert hist == {2: 2, 3: 87, 4: 625, 5: 44, 6: 85, 7: 67, 8: 1518, 9: 14, 10: 469}


def test_sklearn_known_answer():
    """scikit-learn's docstring vector: Ezeiza <-> Charles de Gaulle = 11099.54 km at R = 6371."""
    from math import radians
    from sklearn.metrics.pairwise import haversine_distances

    bsas = [radians(-34.83333), radians(-58.5166646)]
    paris = [radians(49.0083899664), radians(2.53844117956)]
    d = haversine_distances([bsas, paris]) * 6371000 / 1000
    assert abs(d[0, 1] - 11099.54035582) < 1e-6


def test_c_restatement_is_bit_equal_to_sklearn(havref):
    lat, lon = go.synthetic_grid("cn")
    c = go.node_coords_rad(lat, lon)
    la, lo = np.ascontiguousarray(c[:, 0]), np.ascontiguousarray(c[:, 1])
    n = la.size
    out = np.empty((64, n))
    P = ctypes.POINTER(ctypes.c_double)
    havref.havref_rows(la.ctypes.data_as(P), lo.ctypes.data_as(P), n, 1000, 1064, out.ctypes.data_as(P))
    from sklearn.metrics.pairwise import haversine_distances

    ref = haversine_distances(c[1000:1064], c) * 6371.0
    assert np.array_equal(out, ref)
    assert havref.havref_dist_km(0.5, 1.0, 0.5, 1.0) == 0.0


def test_global_grid_counts_on_a_row_sample():
    """BASELINE config 5 grid (180 x 360 cell-centred): degree structure on a few latitude rows."""
    lat, lon = go.synthetic_grid("global")
    coords = go.node_coords_rad(lat, lon)
    assert coords.shape[0] == 64800
    ei, _ = go.graph_edges_blocked(coords, 150.0, block=360, row_range=(89 * 360, 91 * 360))  # the two equator rows
    deg = np.bincount(ei[0] - 89 * 360, minlength=720)
    assert deg.min() == 4 and deg.max() == 4
    ei, _ = go.graph_edges_blocked(coords, 150.0, block=360, row_range=(179 * 360, 180 * 360))  # row next to the pole
    assert np.bincount(ei[0] - 179 * 360).max() == 486
