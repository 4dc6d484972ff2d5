import random, sys, os
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from matterport_maskrcnn_with_tensorflow_serving_b200 import synth, visualize
from matterport_maskrcnn_with_tensorflow_serving_b200.engine import UnmoldEngine, make_geom
base = synth.make_batch(7, 4, (1024, 1024), 100, num_classes=81)
ims = [base[i % 4] for i in range(32)]
d_det = torch.from_numpy(np.stack([im.detections for im in ims])).cuda()
d_msk = torch.from_numpy(np.stack([im.mrcnn_mask for im in ims])).cuda()
eng = UnmoldEngine(32, 100, (28, 28), 81)
eng.plan([make_geom(im.original_image_shape, im.image_shape, im.window) for im in ims])
eng.enqueue(d_det, d_msk)
rng = np.random.default_rng(1)
img = torch.from_numpy(synth.synth_rgb_image(rng, 1024, 1024)).cuda()
colors = visualize.random_colors(100, rng=random.Random(0))
for _ in range(3):
    visualize.composite_batch(eng, [img] * 32, colors)
torch.cuda.synchronize()
