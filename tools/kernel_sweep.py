"""Developer sweep: time the expand kernels on BASELINE configs[1] for several compiled
variants of the team kernel.  Needs the development build of the library:

    MRX_NVCC_FLAGS=-DMRX_DEV MRX_LIB_NAME=libmrx_dev.so python -m matterport_maskrcnn_with_tensorflow_serving_b200.build
    MRX_LIB=libmrx_dev.so python tools/kernel_sweep.py --variants 6x5x10w0 6x5x10w1 ...

Each variant runs in its own subprocess (the library reads MRX_EXPAND_TEAMS / MRX_EXPAND_FLAGS
at launch time in -DMRX_DEV builds only).  Not the contract bench (see bench.py).
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(args):
    import numpy as np
    import torch

    sys.path.insert(0, ROOT)
    import bench
    from matterport_maskrcnn_with_tensorflow_serving_b200.engine import UnmoldEngine, make_geom

    torch.cuda.set_device(0)
    gran = os.environ.get("MRX_L2_GRAN")
    if gran:      # experiment: cudaLimitMaxL2FetchGranularity (0x05), a hint in bytes
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        torch.zeros(1, device="cuda")
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(gran)))
        val = ctypes.c_size_t(0)
        rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
        print(json.dumps({"l2_fetch_granularity_set": int(gran), "rc": rc, "now": val.value}), flush=True)
    ims = bench.make_bench_images(0, args.batch)
    d_det = torch.from_numpy(np.stack([im.detections for im in ims])).cuda()
    d_msk = torch.from_numpy(np.stack([im.mrcnn_mask for im in ims])).cuda()
    eng = UnmoldEngine(args.batch, 100, (28, 28), 81)
    eng.plan([make_geom(im.original_image_shape, im.image_shape, im.window) for im in ims])
    for _ in range(3):
        eng.enqueue(d_det, d_msk)
    torch.cuda.synchronize()
    counts = eng.d_counts[:args.batch].cpu().numpy()
    nbytes = eng.canvas_bytes(counts) + int(counts.sum()) * 3160
    digest = bench._bytesum(torch, eng.d_canvas[:int(eng._offsets[args.batch])])

    def timeit(fn):
        ts = []
        for _ in range(args.iters):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    exp_med, exp_min = timeit(lambda: eng.enqueue_expand())
    step_med, _ = timeit(lambda: eng.enqueue(d_det, d_msk))
    pre_med, _ = timeit(lambda: eng.enqueue(d_det, d_msk, expand=False))
    pk_med, pk_min = timeit(lambda: eng.enqueue_expand_packed())
    pack_med, _ = timeit(lambda: eng.pack_masks())
    # mask overlay on the device canvas: kernel alone (inputs staged once), all four forms
    import random
    from matterport_maskrcnn_with_tensorflow_serving_b200 import synth, visualize
    rng = np.random.default_rng(0)
    img = torch.from_numpy(synth.synth_rgb_image(rng, 1024, 1024)).cuda()
    colors = visualize.random_colors(100, rng=random.Random(0))
    stage = visualize.CompositeStage(eng, [img] * args.batch, colors, 0.5)
    comp, _ = timeit(lambda: stage.run())
    comp = round(comp, 4)
    print(json.dumps({"variant": os.environ.get("MRX_EXPAND_TEAMS", "default"),
                      "flags": os.environ.get("MRX_EXPAND_FLAGS", ""),
                      "bits_warps": os.environ.get("MRX_BITS_WARPS", "default"),
                      "l2_gran": os.environ.get("MRX_L2_GRAN", ""),
                      "expand_ms": round(exp_med, 4), "expand_ms_min": round(exp_min, 4),
                      "expand_GBps": round(nbytes / exp_med / 1e6, 1), "step_ms": round(step_med, 4),
                      "prologue_gather_ms": round(pre_med, 4),
                      "expand_packed_ms": round(pk_med, 4), "expand_packed_ms_min": round(pk_min, 4),
                      "pack_kernel_ms": round(pack_med, 4), "composite_ms": comp,
                      "ones": digest}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--variants", nargs="*", default=["6x5x10w0", "6x5x10w1"])
    ap.add_argument("--flags", nargs="*", default=[""])
    ap.add_argument("--bits-warps", nargs="*", default=[""])
    ap.add_argument("--l2-gran", nargs="*", default=[""])
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args)
    for v in args.variants:
        for f in args.flags:
            for bw in args.bits_warps:
                for g in args.l2_gran:
                    env = dict(os.environ, MRX_EXPAND_TEAMS=v)
                    if f:
                        env["MRX_EXPAND_FLAGS"] = f
                    if bw:
                        env["MRX_BITS_WARPS"] = bw
                    if g:
                        env["MRX_L2_GRAN"] = g
                    subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--batch",
                                    str(args.batch), "--iters", str(args.iters)], env=env, check=False)


if __name__ == "__main__":
    main()
