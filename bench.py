#!/usr/bin/env python
"""bench.py -- GATv2 SpatialEncoder fwd+bwd throughput on B200 (metric of BASELINE.json), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--autocast]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE config 3's largest single-GPU point -- default shape F=22, H=2, C=11 on the
2911-node / 150 km graph (20,924 edges + 2,911 self loops), B=128 x 48 = 6,144 snapshots PER GPU (weak scaling: the
snapshot batch is sharded DDP-style, the only collective is the 4.2 KB parameter-gradient all-reduce), fp32 contract,
training mode with the reference's attention dropout p=0.1, synthetic N(0,1) inputs.  A "step" is one forward + backward
of the encoder over the rank's batch; edge-msgs/s = snapshots * E / time with E counting self loops.

  value : inputs resident in HBM, timed with CUDA events on the launching stream, max over ranks.
  e2e   : same step through the public module call with the step's x coming from PINNED HOST memory (H2D inside the
          timed region, double-buffered on a copy stream) and the flat parameter gradient read back to the host.
  roofline : dominant kernel (largest share of the step), algorithmic bytes of SURVEY.md 8(d) / its CUDA-event time.
  cpu_baseline : the oracle (op-for-op restatement of PyG GATv2Conv -- torch_geometric is not installable here) timed on
          this box's host cores on a bounded sample (B=2 x 48 snapshots); kind "port".
  --impl reference : the same CPU implementation as the reference arm (all host threads, bounded sample per step).
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_IN, HEADS, C_OUT, L_IN, N_NODES = 22, 2, 11, 48, 2911
METRIC, UNIT = "gatv2_fwd_bwd_edge_msgs_per_s", "edge-msgs/s"


def load_graph_edges():
    g = np.load(os.path.join(ROOT, "tests", "golden", "graph_cn150.npz"))
    return torch.from_numpy(g["edge_index"])


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
# CPU arm: the oracle restatement of PyG GATv2Conv (what the reference executes on CPU), bounded sample
# ------------------------------------------------------------------------------------------------------------------
def cpu_fwd_bwd_time(batch, repeats, warmup, threads):
    from oracle import gatv2_oracle as G

    torch.set_num_threads(threads)
    ei = load_graph_edges()
    S = batch * L_IN
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(S, N_NODES, F_IN, generator=gen)
    gy = torch.randn(S, N_NODES, HEADS * C_OUT, generator=gen)
    params = G.init_params(F_IN, C_OUT, HEADS, seed=0)
    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        G.fwd_bwd(x, ei, params, HEADS, C_OUT, gy)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    E = ei.size(1) - int((ei[0] == ei[1]).sum()) + N_NODES
    return times, S * E


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    batch = 2
    times, edges = cpu_fwd_bwd_time(batch, args.steps, max(1, args.warmup), threads)
    total = sum(times)
    value = edges * len(times) / total
    sample = f"B={batch} x {L_IN} snapshots x {N_NODES} nodes per step, fp32, oracle port of PyG GATv2Conv fwd+bwd (autograd)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": max(1, args.warmup), "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, batch_override=batch, note="CPU arm: bounded sample of the workload per step"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "samples_per_s": batch * len(times) / total,
    }
    emit(line)
    return 0


def workload_config(args, batch_override=None, note=None):
    b = args.batch if batch_override is None else batch_override
    cfg = {
        "workload": f"GATv2 SpatialEncoder fwd+bwd, F={F_IN} H={HEADS} C={C_OUT}, N={N_NODES} nodes, 150 km graph "
                    f"(E=23835 incl. self loops), B={b} x {L_IN} snapshots per GPU (BASELINE config 3, largest single-GPU point)",
        "batch_per_gpu": b, "snapshots_per_gpu": b * L_IN, "snapshot_mode": "shared", "dropout_p": args.dropout,
        "training": True, "autocast_bf16": bool(args.autocast), "parallelism": f"dp{args.gpus} (snapshot-sharded)",
        "l2": "inputs_larger_than_l2",
    }
    if note:
        cfg["note"] = note
    return cfg


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
def run_gpu_arm(args):
    import torch.distributed as dist
    from tec_mollm_b200 import SpatialEncoder, gatv2
    from tec_mollm_b200 import dist as tdist

    rank, world, local_rank = tdist.init_from_env("nccl")
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(1234 + rank)

    B, S = args.batch, args.batch * L_IN
    ei = load_graph_edges().to(dev)
    enc = SpatialEncoder(F_IN, C_OUT, heads=HEADS, dropout=args.dropout).to(dev).train()
    flat = tdist.FlatGradAllReduce(enc.parameters())
    if world > 1:  # identical parameters on every rank, as DDP does at construction (train.py:354)
        for p in enc.parameters():
            dist.broadcast(p.data, src=0)
    x = torch.randn(S, N_NODES, F_IN, device=dev).requires_grad_(True)
    gy = torch.randn(S, N_NODES, HEADS * C_OUT, device=dev)
    plan = enc.gat_conv.plan_for(ei, N_NODES)
    E = plan.num_edges
    edges_per_step = S * E

    def step(xin):
        flat.zero_()
        xin.grad = None
        if args.autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = enc(xin, ei)
        else:
            y = enc(xin, ei)
        y.backward(gy)
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
    gatv2.PHASE_EVENTS = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step(x)
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    events, gatv2.PHASE_EVENTS = gatv2.PHASE_EVENTS, None
    clocks = sampler.stop()

    # per-phase device time from the events recorded inside the timed region
    phase_ms = {"proj_fwd": 0.0, "edge_fwd": 0.0, "edge_bwd": 0.0, "proj_bwd": 0.0}
    prev = None
    for name, ev in events:
        if name in phase_ms and prev is not None:
            phase_ms[name] += prev.elapsed_time(ev)
        prev = ev
    phase_ms = {k: v / args.steps for k, v in phase_ms.items()}

    # ---- e2e: x from pinned host memory every step (double-buffered), gradients read back ----------------------
    with gpu_local_cpus(local_rank) as numa_bound:  # pinned pages land on the GPU's own NUMA node (first touch)
        x_host = [torch.randn(S, N_NODES, F_IN).pin_memory() for _ in range(2)]
    x_dev = [torch.empty(S, N_NODES, F_IN, device=dev).requires_grad_(True) for _ in range(2)]
    g_host = torch.empty(flat.flat.numel()).pin_memory()
    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        with torch.cuda.stream(copy_stream):
            x_dev[0].data.copy_(x_host[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(n):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(consumed[nxt])
                    x_dev[nxt].data.copy_(x_host[nxt], non_blocking=True)
                    ready[nxt].record(copy_stream)
            main.wait_event(ready[cur])
            step(x_dev[cur])
            consumed[cur].record(main)
            g_host.copy_(flat.flat, non_blocking=True)
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

    # ---- max over ranks -----------------------------------------------------------------------------------------
    stats = torch.tensor([elapsed_ms, e2e_ms] + [phase_ms[k] for k in ("proj_fwd", "edge_fwd", "edge_bwd", "proj_bwd")],
                         device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = stats[0].item(), stats[1].item()
    phase_ms = dict(zip(("proj_fwd", "edge_fwd", "edge_bwd", "proj_bwd"), stats[2:].tolist()))

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
        value = world * edges_per_step * args.steps / (elapsed_ms * 1e-3)
        e2e_value = world * edges_per_step * args.steps / (e2e_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.autocast else "f32", "data": "synthetic", "config": workload_config(args),
            "samples_per_s": world * B * args.steps / (elapsed_ms * 1e-3),
            "roofline": {
                "kernel": dom, "bound": "hbm", "achieved": phases[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": phases[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": rows * bpr[dom], "share_of_step": phase_ms[dom] / step_ms,
            },
            "phases": phases,
            "whole_path_frac_of_roofline": (rows * sum(bpr.values()) / (step_ms * 1e-3) / 1e9) / peak,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": world * x_host[0].numel() * 4, "d2h_bytes_per_step": world * g_host.numel() * 4,
                    "host_buffers_numa_local": bool(numa_bound)},
            "gpu_launches": 6 * args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["other_configs"] = other_configs(enc, ei, gy, x, args, dev, E)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            times, edges = cpu_fwd_bwd_time(2, 3, 1, threads)
            best = min(times)
            line["cpu_baseline"] = {
                "value": edges / best, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"B=2 x {L_IN} snapshots x {N_NODES} nodes, fp32, oracle port of PyG GATv2Conv fwd+bwd, best of 3 "
                          f"({best * 1e3:.0f} ms)",
            }
        emit(line)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    return 0


def other_configs(enc, ei, gy, x, args, dev, E):
    """Secondary points of the same path, same run, untimed by the driver (BASELINE.json configs 1-2 run at B = 2 and config 2
    under bf16 autocast): the other precision contract at this batch, and the reference's training batch B = 2."""
    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    def make(xin, gin, autocast):
        def fn():
            xin.grad = None
            for p in enc.parameters():
                p.grad = None
            if autocast:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    y = enc(xin, ei)
            else:
                y = enc(xin, ei)
            y.backward(gin)
        return fn

    out = {}
    other = not args.autocast
    ms = timed(make(x, gy, other), 5)
    out["bf16_autocast" if other else "fp32"] = {"batch_per_gpu": args.batch, "ms_per_step": ms,
                                                  "edge_msgs_per_s": args.batch * L_IN * E / (ms * 1e-3)}
    S2 = 2 * L_IN
    x2 = torch.randn(S2, N_NODES, F_IN, device=dev).requires_grad_(True)
    g2 = torch.randn(S2, N_NODES, HEADS * C_OUT, device=dev)
    ms = timed(make(x2, g2, args.autocast), 50)
    out["batch_2"] = {"batch_per_gpu": 2, "ms_per_step": ms, "edge_msgs_per_s": S2 * E / (ms * 1e-3), "samples_per_s": 2 / (ms * 1e-3),
                      "note": "the reference's training batch (train.py:182): host-launch bound in eager mode, CUDA-graph capturable"}
    return out


def measured_traffic(kernel, rows, args):
    """DRAM bytes of one launch of `kernel` from the committed ncu --set full capture (profiles/traffic_r01.json): measured
    per row at the default workload (fp32, B=128 x 48 snapshots); other sizes scale it by rows, other dtypes report null."""
    path = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if args.autocast or not os.path.exists(path):
        return None, None
    try:
        with open(path) as f:
            t = json.load(f)
        per_row = t["kernels"][kernel]["dram_bytes_per_row"]
    except (KeyError, ValueError):
        return None, None
    src = "ncu --set full, dram read+write per launch, " + ("this workload" if rows == t["rows"] else f"per-row figure measured at {t['rows']} rows")
    return per_row * rows, src + " (profiles/r01_ncu_full_B128.csv)"


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
    ap.add_argument("--batch", type=int, default=128, help="samples (x48 snapshots) per GPU")
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
