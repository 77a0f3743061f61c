"""Host-side pieces of bench.py that run without a GPU: the algorithmic byte model, the workload description
and the NUMA-affinity helper (must restore the caller's affinity and never raise when NVML is absent)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bench_under_test"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_bytes_per_row_match_the_survey():
    b = _bench()
    fp32 = b.bytes_per_row(22, 2, 11, 4)  # 264 / 280 / 544 / 352 B per row = 1,440 in total
    assert fp32 == {"proj_fwd": 264, "edge_fwd": 280, "edge_bwd": 544, "proj_bwd": 352}
    bf16 = b.bytes_per_row(22, 2, 11, 2)
    assert bf16["proj_fwd"] == 22 * 4 + 2 * 22 * 2 and bf16["edge_bwd"] < fp32["edge_bwd"]


def test_gpu_local_cpus_restores_affinity_and_never_raises():
    b = _bench()
    before = os.sched_getaffinity(0)
    with b.gpu_local_cpus(0) as bound:
        assert bound in (True, False)  # no GPU / no NVML here: False, and no exception
    assert os.sched_getaffinity(0) == before


def test_workload_table_and_reference_arm_config():
    """Every --config names its BASELINE row, and the grids are the ones the graph goldens were generated on."""
    import numpy as np

    b = _bench()
    assert set(b.CONFIGS) == {"default", "dense300h4", "global64k"}
    lat, lon = b.grid_axes(b.CONFIGS["default"])
    assert lat.size == 41 and lon.size == 71 and lat[0] == 15.0 and lon[-1] == 140.0
    g = np.load(os.path.join(ROOT, "tests", "golden", "graph_cn150.npz"))
    assert np.array_equal(lat, g["lat"]) and np.array_equal(lon, g["lon"])
    glat, glon = b.grid_axes(b.CONFIGS["global64k"])
    assert glat.size * glon.size == 64800 and glat[0] == -89.5 and glon[-1] == 179.5

    class A:
        config, scaling, batch, dropout, autocast = "dense300h4", "strong", 64, 0.1, False

    c = b.workload_config(A, b.CONFIGS["dense300h4"], 8, 8)
    assert c["global_batch"] == 64 and "300 km" in c["workload"] and c["name"] == "dense300h4"
