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
