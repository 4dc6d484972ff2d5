"""Time the bit-pack kernel on the configs[1] batch (device events) and the packed D2H."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matterport_maskrcnn_with_tensorflow_serving_b200 import synth
from matterport_maskrcnn_with_tensorflow_serving_b200.engine import UnmoldEngine, make_geom
base = synth.make_batch(7, 4, (1024, 1024), 100, num_classes=81)
ims = [base[i % 4] for i in range(32)]
d_det = torch.from_numpy(np.stack([im.detections for im in ims])).cuda()
d_msk = torch.from_numpy(np.stack([im.mrcnn_mask for im in ims])).cuda()
eng = UnmoldEngine(32, 100, (28, 28), 81)
eng.plan([make_geom(im.original_image_shape, im.image_shape, im.window) for im in ims])
eng.enqueue(d_det, d_msk)
for _ in range(3):
    d_packed, off = eng.pack_masks()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.pack_masks(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
h = torch.empty(int(off[-1]), dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
t0 = time.perf_counter(); h.copy_(d_packed[:int(off[-1])]); torch.cuda.synchronize(); t_d2h = time.perf_counter() - t0
print(json.dumps({"workload": "configs[1] batch -> bit-packed masks", "pack_ms": round(float(np.median(ts)), 4),
                  "canvas_read_GBps": round(3.3554432e9 / np.median(ts) / 1e6, 1),
                  "packed_bytes": int(off[-1]), "packed_d2h_ms": round(t_d2h * 1e3, 2)}))
