#!/usr/bin/env python
"""Benchmark of the slam_recognition filter-pipeline hot path on B200 (contract: see the driver's prompt / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A step = one pass of the hot path (pyramid -> S1-S7 fused stack -> feature-point emit) over one batch of synthetic
1080p uint8 BGR frames (BASELINE config C3: 1920x1080, scale sqrt(2) -> 6 levels of 192x288, batch 64 per GPU; frames of
a video batch are independent, so N GPUs each process their own 64 frames: weak scaling, only the feature-point gather
crosses NVLink). Prints ONE JSON line on rank 0.

  value     frames/s, whole job, inputs resident in HBM, timed with CUDA events on the launching stream (max over ranks)
  e2e       same metric through LineEndPipeline.run_host (C-ABI silent_pipeline_run_host) with pinned HOST buffers:
            host->device copy of the frames and device->host copy of both result tensors + points inside the timed region
  roofline  algorithmic bytes per step (union crop read once + the two output tensors) / device time of the
            step, against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the bit-defined C oracle (a port: real TensorFlow-1 / the reference cannot run here) on the host cores
"""
import argparse
import concurrent.futures
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec (1080p RGB, full pyramid)"
FRAME_HW = (1080, 1920)
CENTER = (288, 192)
SCALE = 2 ** .5
CONFIG_ID = 3


def synthetic_frames(first, count):
    """uint8 uniform noise, RandomState(1000 * config + global_frame_index)."""
    out = np.empty((count,) + FRAME_HW + (3,), np.uint8)
    for i in range(count):
        out[i] = np.random.RandomState(1000 * CONFIG_ID + first + i).randint(0, 256, size=FRAME_HW + (3,))
    return out


# ---- CPU arm: the oracle port on the host cores --------------------------------------------------------------------------

def cpu_port_fps(budget_s, threads=None):
    """Frames/s of the bit-defined C oracle (pyramid + S1-S8), `threads` frames in flight (ctypes releases the GIL)."""
    from oracle import c_oracle
    import pysilent_b200.constant_convolutions as cc
    c_oracle.build()
    filters = dict(rgc=cc.midget_rgc(2), rgby=cc.rgby_3(2), stripe=cc.rgb_2d_stripe_tensors(),
                   blur=cc.blur_tensor(2, lengths=7), end=cc.rgb_2d_end_tensors())
    threads = threads or os.cpu_count() or 1
    frames = synthetic_frames(0, min(threads, 8))

    def one(i):
        pyr = c_oracle.from_image(frames[i % len(frames)], 3, CENTER, SCALE)
        return len(c_oracle.line_end_stack(pyr, filters)["points"])

    t0 = time.perf_counter()
    one(0)
    single = time.perf_counter() - t0
    rounds = max(1, int(budget_s / max(single * 1.5, 1e-3)))
    rounds = min(rounds, 8)
    with concurrent.futures.ThreadPoolExecutor(threads) as pool:
        t0 = time.perf_counter()
        done = 0
        for _ in range(rounds):
            done += len(list(pool.map(one, range(threads))))
        elapsed = time.perf_counter() - t0
    return dict(value=done / elapsed, unit="frames/s", cores=threads, kind="port",
                sample="%d frames of the same 1080p/L=6 workload through oracle/silent_oracle.c (pyramid + S1-S8), "
                       "%d threads x %d rounds, %.1f s; single-frame latency %.3f s" % (done, threads, rounds, elapsed,
                                                                                      single))


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    threads = os.cpu_count() or 1
    for _ in range(max(warmup, 0) and 1):
        cpu_port_fps(0.1, threads)
    res = cpu_port_fps(max(5.0, min(60.0, 2.0 * steps)), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 * threads / res["value"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3 sample: 1920x1080 uint8 BGR, scale sqrt2 -> 6 levels of 192x288, full stack; each "
                               "step = one frame per host thread", "frame": "1920x1080x3 uint8", "levels": 6},
        "cpu_baseline": res,
        "e2e": {"value": res["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference CPU path restated (oracle port): TensorFlow 1.x is not installable, the reference cannot run",
    }
    print(json.dumps(line), flush=True)


# ---- clocks ---------------------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (t, l) in self.lines if (t0 is None or t >= t0) and (t1 is None or t <= t1)] or \
               [l for (_, l) in self.lines]
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for row in rows:
            parts = [p.strip() for p in row.split(",")]
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- GPU arm ----------------------------------------------------------------------------------------------------------------

def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from pysilent_b200 import LineEndPipeline, _lib, distributed as sdist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = sdist.bind_to_device_numa_node(local_rank)   # page-locked buffers below land next to the GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    pipe = LineEndPipeline(zoom_ratio=SCALE)
    host_frames = torch.empty((B,) + FRAME_HW + (3,), dtype=torch.uint8).pin_memory()
    host_frames.numpy()[...] = synthetic_frames(rank * B, B)
    frames = host_frames.to(dev, non_blocking=True)
    plan = pipe.plan_for(frames)
    L, h, w = plan.levels, plan.h, plan.w
    n = B * L
    cap = 64 * n
    bufs = (torch.empty((n, h, w, 3), dtype=torch.float32, device=dev), torch.empty((n, h, w, 3), dtype=torch.float32,
            device=dev), torch.empty((cap, 4), dtype=torch.int64, device=dev), torch.zeros(1, dtype=torch.int64, device=dev))
    gather_cap = 16 * n
    # the path's only exchange: feature points to every rank over NCCL/NVLink, one all-gather per step on a side stream
    # (it overlaps the next step's kernels; the final barrier + synchronize below waits for the last one)
    gatherer = sdist.PointGather(gather_cap, L, dev) if world > 1 else None

    def step():
        pipe.run_frames(frames, out=bufs)
        return gatherer.submit(bufs[2], bufs[3], rank * B) if gatherer is not None else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _lib.lib().silent_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    slot = None
    for _ in range(args.steps):
        slot = step()
    if gatherer is not None and slot is not None:   # the last step's gather (side stream) belongs to the timed region
        torch.cuda.current_stream().wait_event(gatherer.done[slot])
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = _lib.lib().silent_launch_count() - launches0
    ms = ev0.elapsed_time(ev1) / args.steps
    count_points = int(bufs[3].item())

    # per-stage device time (separate, untimed pass with the library's event hooks)
    import ctypes
    _lib.check(_lib.lib().silent_plan_enable_timing(plan.handle, 1))
    stage = np.zeros((5, 5))
    for i in range(5):
        pipe.run_frames(frames, out=bufs)
        vals = [ctypes.c_float() for _ in range(5)]
        _lib.check(_lib.lib().silent_plan_stage_ms(plan.handle, *[ctypes.byref(v) for v in vals[:3]]))
        _lib.check(_lib.lib().silent_plan_stack_split_ms(plan.handle, ctypes.byref(vals[3]), ctypes.byref(vals[4])))
        stage[i] = [v.value for v in vals]
    _lib.check(_lib.lib().silent_plan_enable_timing(plan.handle, 0))
    stage_ms = stage[1:].mean(axis=0)   # pyramid, stack (a + b), emit, stack_a, stack_b

    # end to end through the host-buffer entry point
    orient_host = torch.empty((n, h, w, 3), dtype=torch.float32).pin_memory().numpy()
    line_end_host = torch.empty((n, h, w, 3), dtype=torch.float32).pin_memory().numpy()
    host_np = host_frames.numpy()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        res = pipe.run_host(host_np, orient_host, line_end_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = pipe.run_host(host_np, orient_host, line_end_host)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1000.0 / e2e_steps
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    if world > 1:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = (float(v) for v in t.tolist())
        total_launches = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(total_launches)
        launches = int(total_launches.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    alg_bytes = plan.algorithmic_bytes            # per frame, SURVEY 8(d): 13,240,584 B at 1080p / L=6 / uint8
    peak, peak_src = measured_peak()
    achieved = alg_bytes * B / (ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    fps = world * B / (ms * 1e-3)
    # per-kernel view: each kernel's own algorithmic bytes per frame = what it must read + write
    level_bytes = L * h * w * 4
    crop_bytes = alg_bytes - 2 * 3 * level_bytes
    kernel_alg = {"pyramid_pair_kernel": crop_bytes + 3 * level_bytes,           # union crop in, planar pyramid out
                  "stack_a_kernel": 3 * level_bytes + level_bytes,               # pyramid in, channel-sum plane out
                  "stack_b_kernel": level_bytes + 2 * 3 * level_bytes + level_bytes}   # plane in; orient, line_end, gray out
    kernel_ms = {"pyramid_pair_kernel": stage_ms[0], "stack_a_kernel": stage_ms[3], "stack_b_kernel": stage_ms[4]}
    kernels = {}
    for name, kb in kernel_alg.items():
        gbs = kb * B / (kernel_ms[name] * 1e-3) / 1e9 if kernel_ms[name] > 0 else 0.0
        kernels[name] = {"ms": float(kernel_ms[name]), "algorithmic_bytes_per_launch": int(kb * B), "achieved": gbs,
                         "frac": gbs / peak, "share_of_step": float(kernel_ms[name] / max(stage_ms[:3].sum(), 1e-9)),
                         "traffic": (traffic or {}).get("kernels", {}).get(name)}
    dominant = max(kernel_ms, key=kernel_ms.get)
    h2d = B * FRAME_HW[0] * FRAME_HW[1] * 3
    d2h = 2 * n * h * w * 3 * 4 + 8 + 32 * len(res.points)
    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "C3: 1920x1080 uint8 BGR frames, scale sqrt2 -> 6 levels of 192x288x3, full stack "
                               "(pyramid + center-surround + stripe/regulator + line-end + feature points), batch %d per "
                               "GPU" % B, "frame": "1920x1080x3 uint8", "levels": L, "batch_per_gpu": B,
                   "parallelism": "frame-sharded dp%d, NCCL point gather" % world,
                   "l2": "inputs (%.0f MB) and outputs (%.0f MB) per step exceed the 126 MB L2; no flush needed" % (
                       h2d / 1e6, d2h / 1e6),
                   "points_per_step": count_points},
        "clocks": clocks,
        "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "numa_node_rank0": numa_node,
                "api": "LineEndPipeline.run_host -> silent_pipeline_run_host, pinned host buffers"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic.get("dram_bytes_per_step") if traffic else None,
                     "kernel": "whole step per GPU: pyramid_pair_kernel + stack_a_kernel + stack_b_kernel + 4 emit "
                               "kernels (the step's algorithmic bytes over the step's device time; dominant kernel: %s)"
                               % dominant,
                     "algorithmic_bytes_per_step": alg_bytes * B, "peak_source": peak_src,
                     "traffic_source": traffic.get("source") if traffic else None,
                     "stage_ms": {"pyramid": stage_ms[0], "stack_fused": stage_ms[1], "emit": stage_ms[2]},
                     "kernels": kernels},
    }
    if args.batch1:
        one = frames[:1].contiguous()
        plan.reserve(B)
        small = tuple(t[: L] if t.dim() == 4 else t for t in bufs)
        for _ in range(5):
            pipe.run_frames(one, out=small)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            pipe.run_frames(one, out=small)
        b.record()
        torch.cuda.synchronize()
        line["batch1"] = {"workload": "C2: one 1080p frame per call (launch/latency bound)",
                          "ms_per_frame": a.elapsed_time(b) / 50, "frames_per_s": 50e3 / a.elapsed_time(b)}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_port_fps(args.cpu_seconds)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-batch1", dest="batch1", action="store_false")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
