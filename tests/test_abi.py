"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol the header declares;
host-only entry points behave; the Python mirror keeps the reference's surface (no compute calls: no GPU here)."""
import ctypes as C
import inspect
import os
import re

import numpy as np
import pytest
import torch

from helpers import ROOT


@pytest.fixture(scope="module")
def lib():
    from tec_mollm_b200 import build, _lib

    build.build()
    return _lib.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tecgat.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tecgat_\w+|tecgraph_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from tec_mollm_b200 import _lib

    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tecgat.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, "ctypes signature table and header disagree"
    assert lib.tecgat_abi_version() == _lib.ABI_VERSION == 4


def test_library_is_sm100a_native(lib):
    """The shipped cubin is sm_100a and carries tcgen05 / bulk-TMA instructions (checked when cuobjdump exists)."""
    import shutil
    import subprocess
    from tec_mollm_b200 import _lib

    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass and "LDTM" in sass, "tcgen05.mma / tcgen05.ld missing from the projection kernel"
    assert "UBLKCP" in sass, "bulk-TMA copies missing"


def test_dropout_mask_host_statistics(lib):
    from tec_mollm_b200 import _lib

    n, H, p = 200000, 2, 0.1
    keep = np.empty((n, H), dtype=np.uint8)
    E = 23835
    _lib.call("tecgat_dropout_mask_host", C.c_uint64(1234), 0, n, H, C.c_float(p), E, C.c_void_p(keep.ctypes.data))
    rate = 1.0 - keep.mean()
    assert abs(rate - p) < 4e-3
    assert abs(keep[:, 0].mean() - keep[:, 1].mean()) < 6e-3
    keep2 = np.empty_like(keep)
    _lib.call("tecgat_dropout_mask_host", C.c_uint64(1234), 0, n, H, C.c_float(p), E, C.c_void_p(keep2.ctypes.data))
    assert np.array_equal(keep, keep2)                      # counter-based: reproducible
    _lib.call("tecgat_dropout_mask_host", C.c_uint64(1235), 0, n, H, C.c_float(p), E, C.c_void_p(keep2.ctypes.data))
    assert (keep != keep2).mean() > 0.1                     # seed matters
    # windows of the counter stream are consistent
    part = np.empty((1000, H), dtype=np.uint8)
    _lib.call("tecgat_dropout_mask_host", C.c_uint64(1234), 5000, 1000, H, C.c_float(p), E, C.c_void_p(part.ctypes.data))
    assert np.array_equal(part, keep[5000:6000])
    # every snapshot gets its own stream, heads of a pair are decorrelated
    snaps = keep[: 8 * E].reshape(8, E, H)
    assert (snaps[0] != snaps[1]).mean() > 0.1
    both = (1 - keep[:, 0].astype(float)) * (1 - keep[:, 1].astype(float))
    assert abs(both.mean() - p * p) < 2e-3


def test_dropout_streams_are_not_shifted_copies(lib):
    """with one 32-bit key per (snapshot, head) stream, two streams whose keys differ by a multiple of the counter
    stride were the SAME mask sequence shifted by a few slots.  Every draw of a launch now has its own counter
    ((snapshot * heads + head) * edges + slot): no stream equals a shifted copy of another, and the agreement of any two streams at
    small shifts is what independent draws give; the high half
    of the 64-bit seed matters."""
    from tec_mollm_b200 import _lib

    E, S, H, p = 4096, 96, 2, 0.25
    keep = np.empty((S * E, H), dtype=np.uint8)
    _lib.call("tecgat_dropout_mask_host", C.c_uint64(0xDEADBEEF12345678), 0, S * E, H, C.c_float(p), E, C.c_void_p(keep.ctypes.data))
    streams = keep.reshape(S, E, H).transpose(0, 2, 1).reshape(S * H, E).astype(np.int8)
    assert len({s.tobytes() for s in streams}) == S * H                 # all distinct
    expect = p * p + (1 - p) * (1 - p)                                   # P(two independent keep bits agree)
    a = streams[:48]
    worst = 0.0
    for shift in range(0, 9):
        b = streams[48:96, shift:]
        agree = (a[:, None, : E - shift] == b[None, :, :]).mean(axis=2)   # (48, 48) stream pairs
        worst = max(worst, float(np.abs(agree - expect).max()))
    assert worst < 0.05, worst                                           # a shifted copy would agree ~100 %
    hi = np.empty((E, H), dtype=np.uint8)
    _lib.call("tecgat_dropout_mask_host", C.c_uint64(0x0000000012345678), 0, E, H, C.c_float(p), E, C.c_void_p(hi.ctypes.data))
    assert (hi != keep[:E]).mean() > 0.2


def test_errors_are_reported_not_thrown(lib):
    from tec_mollm_b200 import _lib

    with pytest.raises(RuntimeError, match="bad argument"):
        _lib.call("tecgat_dropout_mask_host", C.c_uint64(0), 0, 10, 0, C.c_float(0.1), 5, None)
    info = (C.c_int64 * 12)()
    assert lib.tecgat_plan_info(None, info) != 0
    assert b"NULL" in lib.tecgat_last_error()


def test_python_surface_mirrors_the_reference():
    """Constructor / forward signatures of modules.py:319,340 and the parameter names checkpoints rely on."""
    from tec_mollm_b200 import GATv2Conv, SpatialEncoder, graph

    sig = inspect.signature(SpatialEncoder.__init__)
    assert list(sig.parameters)[:5] == ["self", "in_channels", "out_channels", "heads", "dropout"]
    assert sig.parameters["heads"].default == 2 and sig.parameters["dropout"].default == 0.1
    fsig = inspect.signature(SpatialEncoder.forward)
    assert list(fsig.parameters) == ["self", "x", "edge_index", "edge_weight"]
    assert fsig.parameters["edge_weight"].default is None
    enc = SpatialEncoder(22, 11, heads=2)
    assert enc.output_channels == 22
    shapes = {k: tuple(v.shape) for k, v in enc.state_dict().items()}
    assert shapes == {
        "gat_conv.att": (1, 2, 11), "gat_conv.bias": (22,),
        "gat_conv.lin_l.weight": (22, 22), "gat_conv.lin_l.bias": (22,),
        "gat_conv.lin_r.weight": (22, 22), "gat_conv.lin_r.bias": (22,),
    }
    assert sum(p.numel() for p in enc.parameters()) == 1056
    conv = GATv2Conv(22, 11, heads=2, dropout=0.1, concat=True, add_self_loops=True)  # the call at modules.py:329-336
    assert torch.count_nonzero(conv.bias) == 0
    with pytest.raises(NotImplementedError):
        GATv2Conv(22, 11, heads=2, concat=False)
    with pytest.raises(NotImplementedError):
        GATv2Conv(22, 11, edge_dim=3)
    for name in ("calculate_haversine_distance_matrix", "construct_binary_adjacency", "compute_degree_matrix",
                 "symmetrically_normalize_adjacency", "convert_to_pyg_and_save", "get_coordinates_from_data"):
        assert callable(getattr(graph, name))


def test_no_cpu_fallback():
    """A CPU tensor must raise: the product path never computes on the host."""
    from tec_mollm_b200 import SpatialEncoder

    enc = SpatialEncoder(6, 3, heads=2)
    x = torch.randn(2, 5, 6)
    ei = torch.tensor([[0, 1], [1, 0]])
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(x, ei)


def test_dense_graph_api_helpers_match_the_reference_golden():
    """The host-side mirrors of construct_binary_adjacency / symmetrically_normalize_adjacency / convert_to_pyg_and_save
    (dense-matrix API kept for compatibility) reproduce the reference's saved graph from the golden distances."""
    import tempfile
    from helpers import load_golden
    from tec_mollm_b200 import graph

    g = load_golden("graph_small150.npz")
    A = graph.construct_binary_adjacency(g["D"], float(g["thr"]))
    assert np.array_equal(np.diag(graph.compute_degree_matrix(A)), g["degree"])
    norm = graph.symmetrically_normalize_adjacency(A)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "graph_A.pt")
        graph.convert_to_pyg_and_save(norm, path)
        saved = torch.load(path)
    assert saved["edge_index"].dtype == torch.int64 and saved["edge_weight"].dtype == torch.float32
    assert np.array_equal(saved["edge_index"].numpy(), g["edge_index"])
    assert np.array_equal(saved["edge_weight"].numpy(), g["edge_weight"])
