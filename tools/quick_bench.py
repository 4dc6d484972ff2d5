"""Developer sweep: time mrx_mask_expand on BASELINE config 2 for several chunk sizes /
CTAs per SM.  Not the contract bench (see bench.py)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matterport_maskrcnn_with_tensorflow_serving_b200 import synth  # noqa: E402
from matterport_maskrcnn_with_tensorflow_serving_b200.engine import UnmoldEngine, make_geom  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--hw", type=int, nargs=2, default=[1024, 1024])
    ap.add_argument("--n", type=int, default=100)
    ap.add_argument("--classes", type=int, default=81)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--chunks", type=int, nargs="*", default=[51200])
    ap.add_argument("--ctas", type=int, nargs="*", default=[0])
    args = ap.parse_args()
    torch.cuda.set_device(0)
    # one image's worth of random input, replicated (generation is the slow part)
    base = synth.make_batch(123, min(args.batch, 4), tuple(args.hw), args.n, num_classes=args.classes)
    ims = [base[i % len(base)] for i in range(args.batch)]
    d_det = torch.from_numpy(np.stack([im.detections for im in ims])).cuda()
    d_msk = torch.from_numpy(np.stack([im.mrcnn_mask for im in ims])).cuda()
    geoms = [make_geom(im.original_image_shape, im.image_shape, im.window) for im in ims]
    for chunk in args.chunks:
        for ctas in args.ctas:
            eng = UnmoldEngine(args.batch, 100, (28, 28), args.classes, chunk_bytes=chunk, ctas_per_sm=ctas)
            eng.plan(geoms)
            eng.enqueue(d_det, d_msk)
            torch.cuda.synchronize()
            counts = eng.d_counts[:args.batch].cpu().numpy()
            nbytes = eng.canvas_bytes(counts)
            # full pipeline timing
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for _ in range(3):
                eng.enqueue(d_det, d_msk)
            ev[0].record()
            for _ in range(args.iters):
                eng.enqueue(d_det, d_msk)
            ev[1].record()
            torch.cuda.synchronize()
            t_all = ev[0].elapsed_time(ev[1]) / args.iters
            # expand-only timing
            ts = []
            for _ in range(args.iters):
                eng.d_job_counter.zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); eng.enqueue_expand(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t_exp = float(np.median(ts))
            print(json.dumps({"chunk": chunk, "ctas": ctas, "masks": int(counts.sum()),
                              "ms_all": round(t_all, 4), "ms_expand": round(t_exp, 4),
                              "GBps_expand": round(nbytes / t_exp / 1e6, 1),
                              "Mmasks_s_all": round(counts.sum() / t_all / 1e3, 3)}), flush=True)
            del eng
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
