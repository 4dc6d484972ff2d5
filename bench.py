#!/usr/bin/env python
"""Contract benchmark: instance-masks/sec of the unmold hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          (N>1: launched under torchrun)
    python bench.py --impl reference ...                   (the reference's CPU algorithm)

A step = one pass of the hot path (prologue -> class-tile gather -> fused mask expand)
over one batch of synthetic detections: BASELINE.json configs[1], 32 images of 1024x1024
with 100 instances each, per GPU (weak scaling: every rank processes its own batch; the
path has no exchange step, the final NCCL gather to rank 0 is timed separately as
`gather`).  `value` is device-timed with inputs resident in HBM; `e2e` goes through the
public NumPy-facing path with pinned host buffers, H2D of the inputs and D2H of the
masks inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "instance-masks/sec"
UNIT = "masks/s"
WORKLOAD = "BASELINE.json configs[1]: batch 32 images 1024x1024, 100 instances each, mask unmold"
BATCH, HW, N_INST, CLASSES = 32, (1024, 1024), 100, 81
SEED = 20260921


# ----------------------------------------------------------------------------- CPU legs
def _cpu_worker(args):
    """Unmold one synthetic image with the oracle (the reference's NumPy/SciPy algorithm).
    Runs in a spawned process; returns (n_masks, seconds of the unmold call alone)."""
    seed, hw, n_inst, classes = args
    import numpy as np

    import oracle
    from matterport_maskrcnn_with_tensorflow_serving_b200 import synth

    im = synth.make_batch(seed, 1, hw, n_inst, num_classes=classes)[0]
    det = im.detections.astype(np.float64)      # serve.py:131-136 hands float64 arrays over
    msk = im.mrcnn_mask.astype(np.float64)
    t0 = time.perf_counter()
    out = oracle.unmold_detections(det, msk, im.original_image_shape, im.image_shape, im.window)
    dt = time.perf_counter() - t0
    return int(out[0].shape[0]), dt


class CpuPool:
    """Process pool over host cores running the oracle on whole images."""

    def __init__(self, procs):
        import multiprocessing as mp

        self.procs = int(procs)
        self.ctx = mp.get_context("spawn")
        self.pool = self.ctx.Pool(self.procs)
        # spin the workers up (imports) outside any timed region
        self.pool.map(_cpu_worker, [(1, (64, 64), 2, 3)] * self.procs)

    def run(self, n_images, seed0):
        jobs = [(seed0 + i, HW, N_INST, CLASSES) for i in range(n_images)]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        masks = sum(r[0] for r in res)
        return masks, wall, [r[1] for r in res]

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def default_procs(cores):
    """Worker processes for the CPU legs.  The reference's path is memory-bound (one fresh
    H x W canvas per instance, then np.stack(axis=-1)).  Measured on this pool's 128-thread
    Xeon 8562Y+ host (masks/s at 16/32/64/128 processes: 1067 / 851 / 604 / 396), more
    processes only make it slower, so the CPU legs use 16."""
    return max(1, min(cores, 16))


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline_block(procs=None, images=None):
    """Bounded sample of the workload on the host: `images` images (100 masks each) over
    `procs` processes, plus one image alone for the single-core figure."""
    cores = host_cores()
    procs = procs or default_procs(cores)
    images = images or procs
    pool = CpuPool(procs)
    try:
        m1, w1, _ = pool.run(1, SEED + 1000)
        masks, wall, per = pool.run(images, SEED + 2000)
    finally:
        pool.close()
    return {
        "value": masks / wall, "unit": UNIT, "cores": procs, "kind": "port",
        "sample": f"{images} of the workload's 1024x1024x100-instance images "
                  f"({masks} masks) over {procs} processes, oracle NumPy/SciPy float64",
        "single_core_value": m1 / w1, "host_cores": cores, "cpu_model": cpu_model(),
    }


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (oracle port; the reference's
    code for this path is not vendored) on all the host threads it can use."""
    if rank != 0:
        return 0
    cores = host_cores()
    procs = args.cpu_procs or default_procs(cores)
    total_steps = args.steps + args.warmup
    images = max(4, min(procs, int(600 / max(total_steps, 1))))
    pool = CpuPool(procs)
    try:
        for w in range(args.warmup):
            pool.run(images, SEED + 10 * w)
        t0 = time.perf_counter()
        masks = 0
        for s in range(args.steps):
            m, _, _ = pool.run(images, SEED + 100 + s)
            masks += m
        wall = time.perf_counter() - t0
    finally:
        pool.close()
    value = masks / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_step": images, "instances_per_image": N_INST,
                   "num_classes": CLASSES, "layout": "[H,W,N] bool"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"each step = {images} images x {N_INST} instances over "
                                   f"{procs} processes (host has {cores} logical cores, "
                                   f"{cpu_model()})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "samples": len(sm),
                "reasons": sorted(reasons),
                "window": "identical untimed load (--load-ms) + the timed region, 50 ms period"}


# ----------------------------------------------------------------------------- GPU arm
def load_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu summary."""
    path = os.path.join(ROOT, "profiles", "mask_expand_ncu_summary.json")
    try:
        with open(path) as f:
            j = json.load(f)
        return j.get("dram_bytes_per_launch"), j.get("source")
    except (OSError, ValueError):
        return None, None


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    from matterport_maskrcnn_with_tensorflow_serving_b200 import sharding, synth
    from matterport_maskrcnn_with_tensorflow_serving_b200.engine import UnmoldEngine, make_geom

    cpu_block = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_block = cpu_baseline_block(args.cpu_procs or None)   # before CUDA is initialised here

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic inputs (seeded, per rank), resident in HBM and in pinned host memory
    base = synth.make_batch(SEED + rank, 4, HW, N_INST, num_classes=CLASSES)
    ims = [base[i % len(base)] for i in range(BATCH)]
    det_np = np.stack([im.detections for im in ims])              # [32,100,6] f32
    h_det = torch.from_numpy(det_np).pin_memory()
    h_msk = torch.empty((BATCH, N_INST, 28, 28, CLASSES), dtype=torch.float32).pin_memory()
    for i, im in enumerate(ims):
        h_msk[i].copy_(torch.from_numpy(im.mrcnn_mask))
    d_det = h_det.to(dev)
    d_msk = h_msk.to(dev)
    geoms = [make_geom(im.original_image_shape, im.image_shape, im.window) for im in ims]

    eng = UnmoldEngine(BATCH, N_INST, (28, 28), CLASSES, chunk_bytes=args.chunk_bytes,
                       ctas_per_sm=args.ctas_per_sm)
    eng.plan(geoms)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        eng.enqueue(d_det, d_msk, stream)
    torch.cuda.synchronize()
    counts = eng.d_counts[:BATCH].cpu().numpy()
    status = eng.d_status[:BATCH].cpu().numpy()
    assert int(status.max()) == 0, "synthetic inputs flagged invalid"
    masks_per_step = int(counts.sum())
    out_bytes = eng.canvas_bytes(counts)
    algo_bytes = out_bytes + masks_per_step * (28 * 28 * 4 + 24)   # SURVEY.md 8d per-instance figure

    # ---- timed region: exactly K steps, device-timed, expand kernel timed per launch
    sampler = ClockSampler(local_rank)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    import ctypes as C

    from matterport_maskrcnn_with_tensorflow_serving_b200 import _native as N
    lib = eng.lib
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    st = N.stream_ptr(stream)
    # clock sampling needs load for longer than a few ms: run the same steps untimed for
    # --load-ms first (clocks settle, nvidia-smi gets samples), keep sampling through the
    # timed region
    sampler.start()
    t_load = time.perf_counter()
    while (time.perf_counter() - t_load) * 1e3 < args.load_ms:
        for _ in range(10):
            eng.enqueue(d_det, d_msk, stream)
        torch.cuda.synchronize()
    barrier()
    ev0.record(stream)
    for s in range(args.steps):
        N.check(lib.mrx_unmold_prologue(P(d_det), N.MRX_F32, BATCH, N_INST, CLASSES, 28,
                                        P(eng.d_geom), P(eng.d_boxes), P(eng.d_class_ids),
                                        P(eng.d_scores), P(eng.d_src_index), P(eng.d_box_aux),
                                        P(eng.d_counts), P(eng.d_status), P(eng.d_job_counter),
                                        st), "prologue")
        N.check(lib.mrx_gather_tiles(P(d_msk), N.MRX_F32, BATCH, N_INST, 28, 28, CLASSES,
                                     P(eng.d_class_ids), P(eng.d_src_index), P(eng.d_counts),
                                     P(eng.d_tiles), st), "gather")
        kev[s][0].record(stream)
        N.check(lib.mrx_mask_expand(P(eng.d_tiles), P(eng.d_boxes), P(eng.d_box_aux),
                                    P(eng.d_counts), P(eng.d_geom),
                                    P(eng.d_canvas_off), P(eng.d_canvas), BATCH, N_INST, 28, 28,
                                    eng.chunk_bytes, eng.ctas_per_sm, P(eng.d_job_counter), st),
                "expand")
        kev[s][1].record(stream)
    ev1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    expand_ms = [a.elapsed_time(b) for a, b in kev]
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    barrier()
    max_ms = float(t.item())
    total_masks = torch.tensor([masks_per_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_masks, op=dist.ReduceOp.SUM)
    value = float(total_masks.item()) * args.steps / (max_ms * 1e-3)

    # ---- e2e: host buffers in, host masks out, every step (public API: StreamingUnmolder,
    # which overlaps the H2D of batch k+1 with the D2H of batch k's masks)
    from matterport_maskrcnn_with_tensorflow_serving_b200.engine import StreamingUnmolder

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    sm = StreamingUnmolder(eng, geoms)
    sm.submit(h_det, h_msk)                 # warm-up batch (allocations, first-touch)
    sm.wait(0)
    barrier()
    t0 = time.perf_counter()
    last = None
    for _ in range(e2e_steps):
        kk = sm.submit(h_det, h_msk)
        if last is not None:
            sm.wait(last)                   # batch k-1 is consumed while batch k is in flight
        last = kk
    h_counts, h_boxes, h_out = sm.wait(last)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = float(total_masks.item()) * e2e_steps / float(te.item())
    h2d = sm.h2d_bytes
    d2h = sm.d2h_bytes
    # sanity: the host copy really holds this step's masks
    assert int(h_counts.sum()) == masks_per_step and int(h_out[: 1 << 20].max()) <= 1

    # ---- gather-inclusive (N > 1): per-rank canvases to rank 0 over NCCL
    gather = None
    if world > 1:
        sizes = [int(out_bytes)] * world      # same geometry and counts on every rank
        local = eng.d_canvas[:int(out_bytes)]
        recv = sharding.gather_bytes_to_root(local, sizes, 0)     # warm-up (allocates)
        del recv
        barrier()
        g0 = torch.cuda.Event(enable_timing=True)
        g1 = torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        eng.enqueue(d_det, d_msk, stream)
        recv = sharding.gather_bytes_to_root(local, sizes, 0)
        g1.record(stream)
        torch.cuda.synchronize()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather = {"value": float(total_masks.item()) / (float(tg.item()) * 1e-3), "unit": UNIT,
                  "ms": float(tg.item()), "bytes_into_rank0": int(out_bytes) * (world - 1),
                  "note": "one step + NCCL gather of all canvases to rank 0 (NVLink ingress bound)"}
        del recv

    if rank == 0:
        peak, peak_src = load_peak()
        traffic, traffic_src = load_traffic()
        k_ms = float(np.mean(expand_ms))
        achieved = algo_bytes / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_gpu": BATCH, "instances_per_image": N_INST,
                       "num_classes": CLASSES, "mask_tile": "28x28 f32", "layout": "[H,W,N] bool",
                       "sharding": f"images over {world} rank(s), no data-path collective",
                       "l2": "per-step working set (813 MB in + 3.36 GB out per GPU) exceeds the "
                             "126 MB L2; no explicit flush",
                       "tile_buffer_bytes": eng.chunk_bytes or "auto", "seed": SEED},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "path": "engine.StreamingUnmolder: pinned host detections+mrcnn_mask -> H2D "
                            "(own stream, overlaps the previous batch's D2H) -> 3 kernels -> D2H "
                            "of counts, boxes and the [H,W,N] bool canvases; every batch's "
                            "masks are waited for on the host"},
            "gpu_launches": 3 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "kernel": "mask_expand_team_kernel",
                         "kernel_ms": k_ms, "algorithmic_bytes_per_launch": int(algo_bytes),
                         "peak_source": peak_src, "traffic_source": traffic_src},
            "cpu_baseline": cpu_block,
        }
        if gather is not None:
            line["gather"] = gather
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--load-ms", type=float, default=400.0,
                    help="untimed identical load before the timed region (clock sampling window)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--chunk-bytes", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-procs", type=int, default=0, help="worker processes of the CPU legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
