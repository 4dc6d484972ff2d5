"""Shared helpers for the parity tests (oracle = checker, CUDA path = thing under test)."""
import numpy as np

import oracle
from matterport_maskrcnn_with_tensorflow_serving_b200 import synth

# Stated fp32 tolerance of the resized (pre-threshold) mask values: the reference
# interpolates in float64 (inputs are widened float32, serve.py:131-136); the device does
# two fp32 lerps (|err| <~ 3e-7 for values in [0,1]).
MASK_VALUE_ATOL = 1e-6


def oracle_unmold(im, dtype=np.float64, return_resized=False):
    return oracle.unmold_detections(
        im.detections.astype(dtype), im.mrcnn_mask.astype(dtype),
        im.original_image_shape, im.image_shape, im.window, return_resized=return_resized)


def item_of(im, dtype=np.float64):
    return (im.detections.astype(dtype), im.mrcnn_mask.astype(dtype),
            im.original_image_shape, im.image_shape, im.window)


def compare_masks(gpu_masks, ref_masks, resized, boxes):
    """Binary masks must agree everywhere the oracle's pre-threshold value is further than
    MASK_VALUE_ATOL from 0.5.  Returns (n_flips_outside_band, n_pixels_in_band)."""
    assert gpu_masks.shape == ref_masks.shape, (gpu_masks.shape, ref_masks.shape)
    assert gpu_masks.dtype == np.bool_
    diff = gpu_masks != ref_masks
    in_band = 0
    bad = 0
    if diff.any():
        for i, (y1, x1, y2, x2) in enumerate(boxes):
            d = diff[y1:y2, x1:x2, i]
            if d.any():
                near = np.abs(resized[i] - 0.5) <= MASK_VALUE_ATOL
                bad += int((d & ~near).sum())
        # differences outside any box are always errors
        outside = diff.copy()
        for i, (y1, x1, y2, x2) in enumerate(boxes):
            outside[y1:y2, x1:x2, i] = False
        bad += int(outside.sum())
    for r in resized:
        in_band += int((np.abs(r - 0.5) <= MASK_VALUE_ATOL).sum())
    return bad, in_band
