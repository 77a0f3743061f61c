#!/usr/bin/env python
"""bench.py -- GATv2 SpatialEncoder fwd+bwd throughput on B200 (metric of BASELINE.json), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME] [--scaling weak|strong]
                    [--batch B] [--autocast]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workloads (``--config``, named in ``config.workload``):
  default     BASELINE config 3's largest single-GPU point: F=22, H=2, C=11 on the 2911-node / 150 km graph (20,924 edges +
              2,911 self loops), B=128 x 48 = 6,144 snapshots PER GPU.
  dense300h4  BASELINE config 4: the same grid at 300 km (76,532 edges, max degree 38), H=4, C=11 (H*C = 44), B=32 per GPU.
  global64k   BASELINE config 5: global 1 x 1 degree cell-centred grid (64,800 nodes, 1,548,000 edges, max degree 486),
              H=2, C=11, B=4 per GPU.
Every workload builds its graph with the product's own haversine kernels (``graph.build_graph``), timed as the
``graph_build`` leg next to the reference's CPU functions.  fp32 contract, training mode with the reference's attention
dropout p=0.1, synthetic N(0,1) inputs.  A "step" is one forward + backward of the encoder over the rank's batch plus the
parameter-gradient all-reduce; edge-msgs/s = snapshots * E / time with E counting self loops.

  scaling : "weak" (default) -- B per GPU is fixed, the snapshot batch grows with N; "strong" -- ``--batch`` is the GLOBAL
            batch, sharded B/N per rank DDP-style.
  value   : inputs resident in HBM, timed with CUDA events on the launching stream, max over ranks.
  e2e     : the same step through the public module call with the step's x coming from PINNED HOST memory (H2D inside the
            timed region, double-buffered on a copy stream) and the flat parameter gradient read back to the host; y and dx
            stay on the device (what a training step does with them).
  roofline: dominant kernel (largest share of the step), algorithmic bytes of SURVEY.md 8(d) / its CUDA-event time (events
            recorded by the library between the phases of tecgat_forward / tecgat_backward on the launching stream).
  gpu_launches : kernels launched by libtecgat inside the timed region, counted by the library (tecgat_launch_count).
  cpu_baseline : the oracle (op-for-op restatement of PyG GATv2Conv -- torch_geometric is not installable here) timed on this
            box's host cores on a bounded sample; kind "port".
  --impl reference : the same CPU implementation as the reference arm (all host threads, bounded sample per step).
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_IN = 48
METRIC, UNIT = "gatv2_fwd_bwd_edge_msgs_per_s", "edge-msgs/s"

CONFIGS = {
    # name: grid axes (degrees), threshold, shape, default batch per GPU, CPU sample (snapshots)
    "default": dict(lat=(15.0, 55.0, 41), lon=(70.0, 140.0, 71), thr=150.0, F=22, H=2, C=11, batch=128, cpu_snapshots=96,
                    what="2911-node 150 km graph (BASELINE config 3, largest single-GPU point)"),
    "dense300h4": dict(lat=(15.0, 55.0, 41), lon=(70.0, 140.0, 71), thr=300.0, F=22, H=4, C=11, batch=32, cpu_snapshots=24,
                       what="2911-node 300 km graph, 4 heads (BASELINE config 4)"),
    "global64k": dict(lat=(-89.5, 89.5, 180), lon=(-179.5, 179.5, 360), thr=150.0, F=22, H=2, C=11, batch=4, cpu_snapshots=4,
                      what="global 1x1 degree grid, 64,800 nodes, 150 km (BASELINE config 5)"),
}


def grid_axes(cfg):
    return np.linspace(*cfg["lat"]), np.linspace(*cfg["lon"])


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def bytes_per_row(F, H, C, bp):
    """Algorithmic bytes per row and phase; bp = bytes of an xl/xr element."""
    HC = H * C
    return {
        "proj_fwd": F * 4 + 2 * HC * bp,
        "edge_fwd": 2 * HC * bp + HC * 4 + H * 8,
        "edge_bwd": HC * 4 + 2 * HC * bp + HC * 4 + H * 8 + 2 * HC * bp,
        "proj_bwd": 2 * HC * bp + F * 4 + F * 4,
    }


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of PyG GATv2Conv (what the reference executes on CPU) and of the reference's graph
# functions, on bounded samples
# ------------------------------------------------------------------------------------------------------------------
def cpu_graph(cfg, max_rows=None):
    """The reference's graph pipeline restated on CPU (oracle/graph_oracle.py: sklearn haversine + threshold + normalise).
    Returns (edge_index, seconds, rows evaluated); for grids too large for a dense (N, N) matrix a bounded row block is timed."""
    from oracle import graph_oracle as go

    lat, lon = grid_axes(cfg)
    n = lat.size * lon.size
    t0 = time.perf_counter()
    if n <= 4096:
        ei, _ = go.graph_edges_dense(lat, lon, cfg["thr"])
        return torch.from_numpy(np.asarray(ei)), time.perf_counter() - t0, n
    rows = min(n, max_rows or 2048)
    coords = go.node_coords_rad(lat, lon)
    t0 = time.perf_counter()
    go.graph_edges_blocked(coords, cfg["thr"], row_range=(0, rows))
    return None, time.perf_counter() - t0, rows


def cpu_fwd_bwd_time(cfg, ei, snapshots, repeats, warmup, threads, N=None):
    from oracle import gatv2_oracle as G

    torch.set_num_threads(threads)
    if N is None:
        N = int(np.prod([cfg["lat"][2], cfg["lon"][2]]))
    F, H, C = cfg["F"], cfg["H"], cfg["C"]
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(snapshots, N, F, generator=gen)
    gy = torch.randn(snapshots, N, H * C, generator=gen)
    params = G.init_params(F, C, H, seed=0)
    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        G.fwd_bwd(x, ei, params, H, C, gy)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    E = ei.size(1) - int((ei[0] == ei[1]).sum()) + N
    return times, snapshots * E


def cpu_sample_grid(cfg):
    """The grid the CPU legs run on: the workload's own grid when the reference's dense pipeline can hold it, otherwise a
    bounded sample -- the northern 30 latitude rows of the global grid (10,800 nodes incl. the polar rows; the whole grid takes
    the single-threaded reference pipeline ~6 minutes to build)."""
    lat, lon = grid_axes(cfg)
    if lat.size * lon.size <= 4096:
        return lat, lon, None
    return lat[-30:], lon, f"northern 30 of {lat.size} latitude rows"


def cpu_edge_index(cfg):
    """edge_index for the CPU legs without a GPU, from the reference restatement (row-blocked for the sampled global band, so no
    (N, N) matrix is formed)."""
    from oracle import graph_oracle as go

    lat, lon, sample = cpu_sample_grid(cfg)
    if sample is None:
        ei, _ = go.graph_edges_dense(lat, lon, cfg["thr"])
    else:
        ei, _ = go.graph_edges_blocked(go.node_coords_rad(lat, lon), cfg["thr"])
    return torch.from_numpy(np.asarray(ei)), lat.size * lon.size, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    ei, N, band = cpu_edge_index(cfg)
    snaps = cfg["cpu_snapshots"]
    times, edges = cpu_fwd_bwd_time(cfg, ei, snaps, args.steps, max(1, args.warmup), threads, N=N)
    total = sum(times)
    value = edges * len(times) / total
    sample = (f"{snaps} snapshots x {N} nodes{' (' + band + ')' if band else ''} per step (a bounded sample of the workload), fp32, "
              f"oracle port of PyG GATv2Conv fwd+bwd (autograd), {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": max(1, args.warmup), "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg, args.batch or cfg["batch"], args.gpus, note="CPU arm: bounded sample of the workload per step"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "samples_per_s": (snaps / L_IN) * len(times) / total,
    }
    emit(line)
    return 0


def workload_config(args, cfg, batch_per_gpu, world, note=None, edges=None):
    N = cfg["lat"][2] * cfg["lon"][2]
    out = {
        "workload": f"GATv2 SpatialEncoder fwd+bwd, F={cfg['F']} H={cfg['H']} C={cfg['C']}, {cfg['what']}, "
                    f"B={batch_per_gpu} x {L_IN} snapshots per GPU",
        "name": args.config, "nodes": N, "threshold_km": cfg["thr"], "batch_per_gpu": batch_per_gpu,
        "global_batch": batch_per_gpu * world if args.scaling == "weak" else (args.batch or cfg["batch"]),
        "snapshots_per_gpu": batch_per_gpu * L_IN, "snapshot_mode": "shared", "dropout_p": args.dropout, "training": True,
        "autocast_bf16": bool(args.autocast), "parallelism": f"dp{world} (snapshot-sharded, {args.scaling} scaling)",
        "l2": "inputs_larger_than_l2" if batch_per_gpu * L_IN * N * cfg["F"] * 4 > 126e6 else "working set below L2 (small batch)",
        "graph": "built by tec_mollm_b200.graph.build_graph (haversine kernels)",
    }
    if edges is not None:
        out["edges_incl_self_loops"] = edges
    if note:
        out["note"] = note
    return out


@contextlib.contextmanager
def gpu_local_cpus(gpu_index):
    """Temporarily bind the calling thread to the CPUs NVML reports as local to GPU ``gpu_index`` so that host buffers
    allocated inside land on that GPU's NUMA node (at 8 GPUs the H2D copies otherwise cross the socket link).  Yields whether
    the binding happened; the previous affinity is restored on exit (the CPU baseline leg uses every core)."""
    old, bound = None, False
    try:
        import pynvml

        old = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(gpu_index))
        bound = True
    except Exception:
        bound = False
    try:
        yield bound
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except Exception:
                pass


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def graph_build_leg(cfg, dev, with_cpu):
    """Builds the workload's graph with the product kernels and times it: CUDA events around the whole call (count pass,
    host guard-band re-check, fill pass) and host wall clock; nominal pairs = N*(N-1) ordered pairs."""
    from tec_mollm_b200 import graph

    lat, lon = grid_axes(cfg)
    graph.build_graph(lat, lon, cfg["thr"], device=dev)  # warm-up (context, allocator)
    torch.cuda.synchronize(dev)
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        ei, ew, stats = graph.build_graph(lat, lon, cfg["thr"], device=dev, return_stats=True)
    e1.record()
    torch.cuda.synchronize(dev)
    wall = (time.perf_counter() - t0) / reps
    ms = e0.elapsed_time(e1) / reps
    n = lat.size * lon.size
    pairs = n * (n - 1)
    leg = {"nodes": n, "edges": int(ei.size(1)), "ms": ms, "wall_ms": wall * 1e3, "pairs_per_s": pairs / (ms * 1e-3),
           "pairs": "nominal N*(N-1) ordered pairs (latitude-band skipping evaluates fewer)", "guard_band_pairs": stats["guard_band_pairs"],
           "kernel_ms": stats.get("kernel_ms"), "evaluated_pairs": stats.get("evaluated_pairs"), "dtype": "f64"}
    if stats.get("evaluated_pairs") and stats.get("kernel_ms"):
        # both passes evaluate the same pairs: count (classify) and fill (classify + write)
        leg["evaluated_pairs_per_s"] = 2 * stats["evaluated_pairs"] / (stats["kernel_ms"] * 1e-3)
        leg["count_kernel_ms"], leg["fill_kernel_ms"] = stats["count_kernel_ms"], stats["fill_kernel_ms"]
        leg["pairs"] = ("nominal N*(N-1) ordered pairs over the whole call; evaluated_pairs = what the latitude / longitude band "
                        "skipping leaves per pass (kernel_ms = the two edge kernels, CUDA events inside the library)")
    if with_cpu:
        _, secs, rows = cpu_graph(cfg)
        leg["cpu_baseline"] = {"value": rows * (n - 1) / secs, "unit": "pairs/s", "cores": 1, "kind": "port",
                               "sample": f"{rows} of {n} rows of the reference pipeline (sklearn haversine_distances + threshold + "
                                         f"scipy normalise, restated in oracle/graph_oracle.py), {secs * 1e3:.0f} ms"}
    return ei, leg


def run_gpu_arm(args):
    import torch.distributed as dist
    from tec_mollm_b200 import SpatialEncoder, _lib
    from tec_mollm_b200 import dist as tdist

    rank, world, local_rank = tdist.init_from_env("nccl")
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(1234 + rank)
    lib = _lib.lib()
    cfg = CONFIGS[args.config]
    F_IN, HEADS, C_OUT = cfg["F"], cfg["H"], cfg["C"]

    want_cpu = world == 1 and not args.no_cpu_baseline
    ei, graph_leg = graph_build_leg(cfg, dev, with_cpu=want_cpu)
    N_NODES = graph_leg["nodes"]

    gbatch = args.batch or cfg["batch"]
    if args.scaling == "strong":
        lo, hi = tdist.shard_range(gbatch, rank, world)
        B = hi - lo
        if B <= 0:
            raise RuntimeError(f"strong scaling: global batch {gbatch} < {world} ranks")
        b_max = -(-gbatch // world)
    else:
        B = b_max = gbatch
    S = B * L_IN
    enc = SpatialEncoder(F_IN, C_OUT, heads=HEADS, dropout=args.dropout, snapshot_mode="shared").to(dev).train()
    flat = tdist.FlatGradAllReduce(enc.parameters(), module=enc)
    if world > 1:  # identical parameters on every rank, as DDP does at construction (train.py:354)
        for p in enc.parameters():
            dist.broadcast(p.data, src=0)
    x = torch.randn(S, N_NODES, F_IN, device=dev).requires_grad_(True)
    gy = torch.randn(S, N_NODES, HEADS * C_OUT, device=dev)
    plan = enc.gat_conv.plan_for(ei, N_NODES)
    E = plan.num_edges
    edges_per_step = S * E
    ar_events = []

    def step(xin, time_ar=False):
        flat.zero_()
        xin.grad = None
        if args.autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = enc(xin, ei)
        else:
            y = enc(xin, ei)
        y.backward(gy)
        if time_ar and world > 1:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            flat.all_reduce_mean()
            b.record()
            ar_events.append((a, b))
        else:
            flat.all_reduce_mean()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    # ---- value: inputs resident in HBM ------------------------------------------------------------------------
    for _ in range(args.warmup):
        step(x)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.tecgat_phase_timing(1)
    ms4 = (ctypes.c_double * 4)()
    lib.tecgat_phase_times(ms4)  # clear
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = lib.tecgat_launch_count()
    ev0.record()
    for _ in range(args.steps):
        step(x, time_ar=True)
    ev1.record()
    barrier()
    launches = lib.tecgat_launch_count() - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    lib.tecgat_phase_times(ms4)
    lib.tecgat_phase_timing(0)
    clocks = sampler.stop()
    names = ("proj_fwd", "edge_fwd", "edge_bwd", "proj_bwd")
    phase_ms = {k: ms4[i] / args.steps for i, k in enumerate(names)}
    ar_ms = sum(a.elapsed_time(b) for a, b in ar_events) / max(1, len(ar_events)) if ar_events else 0.0

    # ---- multi-GPU correctness (outside the timed region): the all-reduced flat gradient is the mean of the ranks' own ------
    ar_check = None
    if world > 1:
        flat.zero_()
        x.grad = None
        enc(x, ei).backward(gy)
        local = flat.flat.clone()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        flat.all_reduce_mean()
        mean = torch.stack(gathered).double().mean(0)
        err = ((flat.flat.double() - mean).abs().max() / mean.abs().max()).item()
        ar_check = {"max_rel_err_vs_mean_of_rank_gradients": err, "ok": bool(err <= 1e-6)}
        if not ar_check["ok"]:
            raise RuntimeError(f"all-reduced gradient differs from the mean of the per-rank gradients: {err:.3e}")

    # ---- e2e: RAW features from pinned host memory every step (double-buffered), gradients read back -------------------
    # What the reference moves per step (train.py:58-65): x (B, L, N, 6) and the (B, L, 4) time features.  The public call is
    # the two drop-in modules, SpatioTemporalEmbedding (fused gather + concat) -> SpatialEncoder; its backward also reduces the
    # embedding-table gradients, so this step does MORE than the device-timed one.
    from tec_mollm_b200 import SpatioTemporalEmbedding

    C_RAW = 6
    D_EMB = F_IN - C_RAW
    emb = SpatioTemporalEmbedding(D_EMB, num_nodes=N_NODES).to(dev).train()
    with gpu_local_cpus(local_rank) as numa_bound:  # pinned pages land on the GPU's own NUMA node (first touch)
        x_host = [torch.randn(B, L_IN, N_NODES, C_RAW).pin_memory() for _ in range(2)]
        tf_host = [torch.stack([torch.randint(0, 12, (B, L_IN)), torch.randint(0, 366, (B, L_IN)), torch.randint(0, 13, (B, L_IN)),
                                torch.randint(0, 4, (B, L_IN))], dim=-1).float().pin_memory() for _ in range(2)]
    x_dev = [torch.empty(B, L_IN, N_NODES, C_RAW, device=dev) for _ in range(2)]
    tf_dev = [torch.empty(B, L_IN, 4, device=dev) for _ in range(2)]
    gy4 = gy.view(B, L_IN, N_NODES, HEADS * C_OUT)
    emb_params = list(emb.parameters())
    g_host = torch.empty(flat.flat.numel() + sum(p.numel() for p in emb_params)).pin_memory()
    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_step(i):
        flat.zero_()
        for p in emb_params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16) if args.autocast else contextlib.nullcontext():
            y = enc(emb(x_dev[i], tf_dev[i]), ei)      # (B, L, N, 22): snapshots in (b, l) order, no permute copy needed
        y.backward(gy4)
        flat.all_reduce_mean()

    def e2e_loop(n):
        with torch.cuda.stream(copy_stream):
            x_dev[0].copy_(x_host[0], non_blocking=True)
            tf_dev[0].copy_(tf_host[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(n):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(consumed[nxt])
                    x_dev[nxt].copy_(x_host[nxt], non_blocking=True)
                    tf_dev[nxt].copy_(tf_host[nxt], non_blocking=True)
                    ready[nxt].record(copy_stream)
            main.wait_event(ready[cur])
            e2e_step(cur)
            consumed[cur].record(main)
            nflat = flat.flat.numel()
            g_host[:nflat].copy_(flat.flat, non_blocking=True)
            off = nflat
            for p in emb_params:  # the embedding tables' gradients leave with the encoder's (47k floats)
                g_host[off:off + p.numel()].copy_(p.grad.view(-1), non_blocking=True)
                off += p.numel()
        main.synchronize()

    e2e_loop(min(2, args.warmup))
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_loop(args.steps)
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    h2d_per_rank = x_host[0].numel() * 4 + tf_host[0].numel() * 4

    # ---- max over ranks; totals over ranks -----------------------------------------------------------------------
    stats = torch.tensor([elapsed_ms, e2e_ms, ar_ms] + [phase_ms[k] for k in names], device=dev, dtype=torch.float64)
    tot = torch.tensor([float(S), float(B), float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms, ar_ms = stats[0].item(), stats[1].item(), stats[2].item()
    phase_ms = dict(zip(names, stats[3:].tolist()))
    total_snapshots, total_batch, total_launches = tot[0].item(), tot[1].item(), int(tot[2].item())

    if rank == 0:
        bp = 2 if args.autocast else 4
        bpr = bytes_per_row(F_IN, HEADS, C_OUT, bp)
        rows = S * N_NODES
        peak, peak_src = measured_peak_gbs()
        phases = {}
        for k, ms in phase_ms.items():
            gbs = rows * bpr[k] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            phases[k] = {"ms": ms, "bytes_per_row": bpr[k], "achieved_gbs": gbs, "frac": gbs / peak}
        dom = max(phase_ms, key=phase_ms.get)
        traffic, traffic_src = measured_traffic(dom, rows, args)
        step_ms = elapsed_ms / args.steps
        value = total_snapshots * E * args.steps / (elapsed_ms * 1e-3)
        e2e_value = total_snapshots * E * args.steps / (e2e_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "bf16" if args.autocast else "f32", "data": "synthetic",
            "config": workload_config(args, cfg, b_max, world, edges=E),
            "samples_per_s": total_batch * args.steps / (elapsed_ms * 1e-3),
            "roofline": {
                "kernel": dom, "bound": "hbm", "achieved": phases[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": phases[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": rows * bpr[dom], "share_of_step": phase_ms[dom] / step_ms,
            },
            "phases": phases,
            "whole_path_frac_of_roofline": (rows * sum(bpr.values()) / (step_ms * 1e-3) / 1e9) / peak,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(h2d_per_rank * total_snapshots / S), "d2h_bytes_per_step": world * g_host.numel() * 4,
                    "host_buffers_numa_local": bool(numa_bound),
                    "call": "SpatioTemporalEmbedding(x_raw, time_features) -> SpatialEncoder(.., edge_index) -> backward",
                    "note": "raw (B, L, N, 6) features + (B, L, 4) time indices cross H2D (what train.py:58-65 copies); the embedding "
                            "gather + concat runs on the device; y and dx stay on the device (training consumes them there); the "
                            "encoder's and the embedding tables' gradients are read back D2H"},
            "gpu_launches": total_launches,
            "gpu_launches_per_step_per_rank": launches / args.steps,
            "allreduce": {"ms_per_step": ar_ms, "bytes": flat.flat.numel() * 4, "share_of_step": ar_ms / step_ms if step_ms else 0.0,
                          "check": ar_check},
            "graph_build": graph_leg,
            "clocks": clocks,
        }
        if want_cpu:
            line["other_configs"] = other_configs(enc, flat, ei, args, dev, E, cfg, N_NODES, x, gy)
            threads = os.cpu_count() or 1
            snaps = cfg["cpu_snapshots"]
            if N_NODES <= 4096:
                ei_cpu, n_cpu, band = ei.cpu(), N_NODES, None
            else:  # bounded sample of the big grid: its northern band, edges from the product's builder (bit-identical to the reference's)
                lat_s, lon_s, band = cpu_sample_grid(cfg)
                from tec_mollm_b200 import graph as _graph
                ei_cpu, n_cpu = _graph.build_graph(lat_s, lon_s, cfg["thr"], device=dev)[0].cpu(), lat_s.size * lon_s.size
            times, edges = cpu_fwd_bwd_time(cfg, ei_cpu, snaps, 3, 1, threads, N=n_cpu)
            best = min(times)
            line["cpu_baseline"] = {
                "value": edges / best, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{snaps} snapshots x {n_cpu} nodes{' (' + band + ')' if band else ''}, fp32, oracle port of PyG GATv2Conv "
                          f"fwd+bwd, best of 3 ({best * 1e3:.0f} ms)",
            }
        emit(line)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    return 0


def other_configs(enc, flat, ei, args, dev, E, cfg, N, x, gy):
    """Secondary points of the same path, same run, untimed by the driver: the other precision contract at this batch, and the
    reference's training batch B = 2 (train.py:182) eager and through the opt-in CUDA-graph mode."""
    F_IN, HC = cfg["F"], cfg["H"] * cfg["C"]

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3

    def make(xin, gin, autocast):
        def fn():
            flat.zero_()
            xin.grad = None
            if autocast:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    y = enc(xin, ei)
            else:
                y = enc(xin, ei)
            y.backward(gin)
        return fn

    out = {}
    other = not args.autocast
    ms, _ = timed(make(x, gy, other), 5)
    out["bf16_autocast" if other else "fp32"] = {"batch_per_gpu": x.size(0) // L_IN, "ms_per_step": ms,
                                                  "edge_msgs_per_s": x.size(0) * E / (ms * 1e-3)}
    # "dropout 0 for parity, 0.1 for timing realism (report both)" -- the same step with the p = 0 kernels
    p_saved = enc.gat_conv.dropout
    try:
        enc.gat_conv.dropout = 0.0
        ms, _ = timed(make(x, gy, args.autocast), 5)
        out["dropout_0"] = {"batch_per_gpu": x.size(0) // L_IN, "ms_per_step": ms, "edge_msgs_per_s": x.size(0) * E / (ms * 1e-3),
                            "note": "training mode with attention dropout p = 0 (the headline runs the reference's p = 0.1)"}
    finally:
        enc.gat_conv.dropout = p_saved
    S2 = 2 * L_IN
    x2 = torch.randn(S2, N, F_IN, device=dev).requires_grad_(True)
    g2 = torch.randn(S2, N, HC, device=dev)
    ms, wall = timed(make(x2, g2, args.autocast), 50)
    out["batch_2"] = {"batch_per_gpu": 2, "ms_per_step": ms, "wall_ms_per_step": wall, "edge_msgs_per_s": S2 * E / (ms * 1e-3),
                      "samples_per_s": 2 / (ms * 1e-3), "mode": "eager",
                      "note": "the reference's training batch (train.py:182)"}
    try:  # opt-in CUDA-graph mode (SpatialEncoder.graphed): forward and backward are one replay each
        enc.gat_conv.fused_grad_accumulation = False
        f = enc.graphed(x2, ei)

        def gfn():
            x2.grad = None
            f(x2).backward(g2)

        ms, wall = timed(gfn, 50)
        out["batch_2_graphed"] = {"batch_per_gpu": 2, "ms_per_step": ms, "wall_ms_per_step": wall,
                                  "edge_msgs_per_s": S2 * E / (ms * 1e-3), "samples_per_s": 2 / (ms * 1e-3),
                                  "mode": "SpatialEncoder.graphed (CUDA graphs, training mode, dropout seed on the device)"}
    except Exception as exc:  # never lose the bench line to the secondary point
        out["batch_2_graphed"] = {"error": repr(exc)[:300]}
    finally:
        enc.gat_conv.fused_grad_accumulation = True
    return out


def measured_traffic(kernel, rows, args):
    """DRAM bytes of one launch of `kernel` from the newest committed ncu --set full capture (profiles/traffic_r0X.json):
    measured per row at the default workload (fp32, B=128 x 48 snapshots); other sizes scale it by rows, other dtypes and
    workloads report null."""
    if args.autocast or args.config != "default":
        return None, None
    for name in ("traffic_r02.json", "traffic_r01.json"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        try:
            with open(path) as f:
                t = json.load(f)
            per_row = t["kernels"][kernel]["dram_bytes_per_row"]
        except (KeyError, ValueError):
            continue
        src = "ncu --set full, dram read+write per launch, " + ("this workload" if rows == t["rows"] else f"per-row figure measured at {t['rows']} rows")
        return per_row * rows, f"{src} (profiles/{name}; captured separately, never under this run)"
    return None, None


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="default", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--batch", type=int, default=0, help="samples (x48 snapshots) per GPU (weak) or in total (strong); 0 = the config's default")
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--autocast", action="store_true", help="bf16-autocast contract instead of fp32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner there) and stray prints
    # are sent to stderr by pointing fd 1 at fd 2; emit() writes the line to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
