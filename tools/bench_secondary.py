"""Secondary workloads of BASELINE.json (configs[2], [3] per-GPU shard, [4]) and the mold step,
device-timed with CUDA events.  Not the contract bench (bench.py measures configs[1]); the
numbers go into profiles/README.md.  One JSON line per workload on stdout.

  python tools/bench_secondary.py [--iters 20] [--cpu]     (--cpu also times the oracle)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matterport_maskrcnn_with_tensorflow_serving_b200 import synth  # noqa: E402
from matterport_maskrcnn_with_tensorflow_serving_b200.engine import (  # noqa: E402
    AnchorGenerator, Molder, UnmoldEngine, make_geom)
from matterport_maskrcnn_with_tensorflow_serving_b200.model_configs import MaskRCNNServingConfig  # noqa: E402


def time_ms(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def unmold_case(name, batch, hw, n, classes, R, iters, base_images=4, seed=7, composite=False):
    base = synth.make_batch(seed, min(batch, base_images), hw, n, num_classes=classes, max_instances=R)
    ims = [base[i % len(base)] for i in range(batch)]
    d_det = torch.from_numpy(np.stack([im.detections for im in ims])).cuda()
    d_msk = torch.from_numpy(np.stack([im.mrcnn_mask for im in ims])).cuda()
    eng = UnmoldEngine(batch, R, (28, 28), classes)
    eng.plan([make_geom(im.original_image_shape, im.image_shape, im.window) for im in ims])
    eng.enqueue(d_det, d_msk)
    torch.cuda.synchronize()
    counts = eng.d_counts[:batch].cpu().numpy()
    masks = int(counts.sum())
    out_bytes = eng.canvas_bytes(counts)
    algo = out_bytes + masks * (28 * 28 * 4 + 24)
    step_ms, _ = time_ms(lambda: eng.enqueue(d_det, d_msk), iters)

    k_ms, _ = time_ms(lambda: eng.enqueue_expand(), iters)
    # extension layout: the expand kernel that writes bit-packed masks, and the pack kernel
    pk_ms, _ = time_ms(lambda: eng.enqueue_expand_packed(), iters)
    pack_ms, _ = time_ms(lambda: eng.pack_masks(), iters)
    pro_ms, _ = time_ms(lambda: eng.enqueue(d_det, d_msk, expand=False), iters)
    # COCO RLE from the tiles (count pass, one host read of the totals, write pass)
    rle_ms, _ = time_ms(lambda: eng.enqueue_rle(), max(3, iters // 4))
    d_runs, off = eng.enqueue_rle()
    rle_bytes = int(d_runs.numel()) * 4
    if composite:
        import random

        from matterport_maskrcnn_with_tensorflow_serving_b200 import visualize
        rng = np.random.default_rng(1)
        images = [torch.from_numpy(synth.synth_rgb_image(rng, *hw)).cuda() for _ in range(min(batch, 2))]
        images = [images[i % len(images)] for i in range(batch)]
        colors = visualize.random_colors(R, rng=random.Random(0))
        c_ms, _ = time_ms(lambda: visualize.composite_batch(eng, images, colors), max(3, iters // 4))
        print(json.dumps({"workload": name + " -> mask overlay (display_instances blend) on the device canvas",
                          "composite_ms_incl_staging": round(c_ms, 3),
                          "canvas_read_GBps": round(out_bytes / c_ms / 1e6, 1),
                          "note": "includes the device-side staging copies of the 32 input images"}),
              flush=True)
    print(json.dumps({"workload": name, "images": batch, "hw": list(hw), "masks": masks,
                      "canvas_GB": round(out_bytes / 1e9, 3), "step_ms": round(step_ms, 4),
                      "Mmasks_per_s": round(masks / step_ms / 1e3, 3),
                      "expand_ms": round(k_ms, 4),
                      "expand_algorithmic_GBps": round(algo / k_ms / 1e6, 1),
                      "prologue_plus_class_gather_ms": round(pro_ms, 4),
                      "expand_packed_ms": round(pk_ms, 4),
                      "expand_packed_Mmasks_per_s": round(masks / pk_ms / 1e3, 2),
                      "rle_ms_incl_host_read": round(rle_ms, 4), "rle_output_MB": round(rle_bytes / 1e6, 2),
                      "pack_kernel_ms": round(pack_ms, 4),
                      "pack_kernel_canvas_read_GBps": round(out_bytes / pack_ms / 1e6, 1)}), flush=True)
    del eng, d_det, d_msk
    torch.cuda.empty_cache()


def anchors_sweep(iters, cpu):
    import oracle
    gen = AnchorGenerator(MaskRCNNServingConfig)
    for s in [512, 640, 768, 896, 1000, 1024, 1280, 1536, 1792, 2000, 2048]:
        shape = (s, s, 3)
        n = gen.count(shape)
        out = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        med, best = time_ms(lambda: gen.generate_device(shape, out=out), iters)
        rec = {"workload": "configs[4] get_anchors", "size": s, "anchors": int(n),
               "gpu_us_per_call": round(med * 1e3, 2), "gpu_Ganchors_per_s": round(n / med / 1e6, 2),
               "gpu_write_GBps": round(n * 16 / med / 1e6, 1)}
        if cpu:
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                oracle.get_anchors(shape)
            cpu_ms = (time.perf_counter() - t0) / reps * 1e3
            rec["cpu_oracle_ms_per_call"] = round(cpu_ms, 3)
            rec["cpu_Manchors_per_s"] = round(n / cpu_ms / 1e3, 2)
        print(json.dumps(rec), flush=True)


def mold_cases(iters, cpu):
    import oracle
    rng = np.random.default_rng(2)
    m = Molder(MaskRCNNServingConfig)
    for hw in [(1024, 1024), (800, 1333), (2160, 3840)]:
        img = synth.synth_rgb_image(rng, *hw)
        d_img = torch.from_numpy(img).cuda()
        med, _ = time_ms(lambda: m.mold_device(d_img, np.float32), iters)
        rec = {"workload": "mold step (resize_image + mold_image, f32 out)", "hw": list(hw),
               "gpu_us_per_image": round(med * 1e3, 2),
               "algorithmic_GBps": round((3 * hw[0] * hw[1] + 12 * 1024 * 1024) / med / 1e6, 1)}
        if cpu:
            t0 = time.perf_counter()
            u8, *_ = oracle.resize_image(img, min_dim=800, max_dim=1024, min_scale=0, mode="square")
            oracle.mold_image(u8)
            rec["cpu_oracle_ms_per_image"] = round((time.perf_counter() - t0) * 1e3, 2)
        print(json.dumps(rec), flush=True)
    img = synth.synth_rgb_image(rng, 1080, 1920)
    d_img = torch.from_numpy(img).cuda()
    med, _ = time_ms(lambda: m.cv2_resize_device(d_img, (640, 640)), iters)
    print(json.dumps({"workload": "cv2.resize 1080x1920 -> 640x640 (u8, INTER_LINEAR)",
                      "gpu_us_per_image": round(med * 1e3, 2)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    unmold_case("configs[1] 32 x 1024x1024 x 100", 32, (1024, 1024), 100, 81, 100, args.iters, composite=True)
    unmold_case("configs[2] 64 x 800x1333 (HxW) x U{1..100}", 64, (800, 1333), (1, 100), 81, 100,
                args.iters, base_images=16)
    unmold_case("configs[3] per-GPU shard: 16 x 2160x3840 x 50", 16, (2160, 3840), 50, 81, 50,
                args.iters, base_images=2)
    anchors_sweep(args.iters, args.cpu)
    mold_cases(args.iters, args.cpu)


if __name__ == "__main__":
    main()
