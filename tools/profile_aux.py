"""Run each auxiliary kernel once (anchors, cv2 resize, mold) so that one ncu invocation can
capture them:  ncu --set full -k regex:'anchors|cv2_resize|mold_image' python tools/profile_aux.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matterport_maskrcnn_with_tensorflow_serving_b200 import synth  # noqa: E402
from matterport_maskrcnn_with_tensorflow_serving_b200.engine import AnchorGenerator, Molder  # noqa: E402
from matterport_maskrcnn_with_tensorflow_serving_b200.model_configs import MaskRCNNServingConfig  # noqa: E402

torch.cuda.set_device(0)
gen = AnchorGenerator(MaskRCNNServingConfig)
m = Molder(MaskRCNNServingConfig)
rng = np.random.default_rng(0)
img = torch.from_numpy(synth.synth_rgb_image(rng, 1080, 1920)).cuda()
for _ in range(2):
    gen.generate_device((1024, 1024, 3))
    small = m.cv2_resize_device(img, (640, 640))
    m.mold_device(img, np.float32)
torch.cuda.synchronize()
