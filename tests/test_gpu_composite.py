"""GPU parity of the mask overlay of visualize.display_instances (serve.py:160-169; SURVEY.md
8f rank 2) against the oracle restatement of upstream's apply_mask loop: bit-exact uint8."""
import random

import numpy as np
import pytest

import oracle
from matterport_maskrcnn_with_tensorflow_serving_b200 import api_utils, synth, visualize
from matterport_maskrcnn_with_tensorflow_serving_b200.engine import UnmoldEngine, make_geom

from helpers import item_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw,n,alpha", [((96, 128), 12, 0.5), ((333, 517), 37, 0.5),
                                        ((64, 80), 5, 0.3), ((1024, 1024), 100, 0.5)])
def test_apply_masks_matches_oracle(cuda_device, hw, n, alpha):
    rng = np.random.default_rng(41)
    im = synth.make_batch(41, 1, hw, n, num_classes=5)[0]
    boxes, _, _, masks = api_utils.unmold_detections(*item_of(im))
    image = synth.synth_rgb_image(rng, *hw)
    colors = visualize.random_colors(boxes.shape[0], rng=random.Random(7))
    ref = oracle.composite_instances(image, boxes, masks, colors, alpha)
    got = visualize.apply_masks(image, boxes, masks, colors, alpha)
    assert got.dtype == np.uint8 and got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_all_zero_box_is_skipped(cuda_device):
    """upstream: `if not np.any(boxes[i]): continue` -- the instance is not blended."""
    rng = np.random.default_rng(42)
    image = synth.synth_rgb_image(rng, 40, 56)
    masks = rng.random((40, 56, 4)) > 0.4
    boxes = np.array([[0, 0, 40, 56], [0, 0, 0, 0], [3, 4, 30, 40], [0, 0, 40, 56]], dtype=np.int32)
    colors = visualize.random_colors(4, rng=random.Random(1))
    ref = oracle.composite_instances(image, boxes, masks, colors)
    got = visualize.apply_masks(image, boxes, masks, colors)
    assert np.array_equal(got, ref)
    assert not np.array_equal(got, oracle.composite_instances(
        image, np.array([[0, 0, 40, 56]] * 4), masks, colors))


def test_random_colors_values(cuda_device):
    ours = visualize.random_colors(9, rng=random.Random(3))
    ref = oracle.random_colors(9, rng=random.Random(3))
    assert ours == ref


def test_composite_on_device_canvas(cuda_device):
    """Overlay straight from the engine's canvas (no device -> host copy of the masks)."""
    import torch

    rng = np.random.default_rng(43)
    hw = (200, 264)
    ims = synth.make_batch(43, 3, hw, (0, 30), num_classes=6, max_instances=30)
    ims[1] = synth.make_image(rng, hw, 0, num_classes=6, max_instances=30)    # an empty image
    eng = UnmoldEngine(3, 30, (28, 28), 6)
    eng.plan([make_geom(im.original_image_shape, im.image_shape, im.window) for im in ims])
    d_det = torch.from_numpy(np.stack([im.detections for im in ims])).cuda()
    d_msk = torch.from_numpy(np.stack([im.mrcnn_mask for im in ims])).cuda()
    eng.enqueue(d_det, d_msk)
    images = [synth.synth_rgb_image(rng, *hw) for _ in ims]
    colors = [visualize.random_colors(30, rng=random.Random(10 + b)) for b in range(3)]
    outs = visualize.composite_batch(eng, images, colors)
    for b, im in enumerate(ims):
        boxes, _, _, masks = oracle.unmold_detections(
            im.detections, im.mrcnn_mask, im.original_image_shape, im.image_shape, im.window)
        got_b, _, _, got_masks = api_utils.unmold_detections(*item_of(im, np.float32))
        assert np.array_equal(got_b, boxes)
        # composite of the device masks; the oracle overlay is computed from the same masks
        ref = oracle.composite_instances(images[b], got_b, got_masks, colors[b]) \
            if got_b.shape[0] else images[b]
        assert np.array_equal(outs[b].cpu().numpy(), ref)


def test_unmold_overlay_batch(cuda_device):
    rng = np.random.default_rng(44)
    hw = (120, 160)
    ims = synth.make_batch(44, 2, hw, 9, num_classes=4, max_instances=12)
    images = [synth.synth_rgb_image(rng, *hw) for _ in ims]
    colors = visualize.random_colors(12, rng=random.Random(5))
    res = api_utils.unmold_overlay_batch([item_of(im, np.float32) for im in ims], images, colors)
    for (b, c, s, overlay), im, image in zip(res, ims, images):
        rb, rc, rs, rm = api_utils.unmold_detections(*item_of(im, np.float32))
        assert np.array_equal(b, rb) and np.array_equal(c, rc) and np.array_equal(s, rs)
        assert np.array_equal(overlay, oracle.composite_instances(image, rb, rm, colors))


def test_alpha_sweep_on_device_canvas(cuda_device):
    """Several alphas (including the degenerate 0 and 1) on the engine's canvas vs the oracle."""
    import torch

    rng = np.random.default_rng(45)
    hw = (256, 320)
    ims = synth.make_batch(45, 2, hw, 40, num_classes=6, max_instances=40)
    eng = UnmoldEngine(2, 40, (28, 28), 6)
    eng.plan([make_geom(im.original_image_shape, im.image_shape, im.window) for im in ims])
    d_det = torch.from_numpy(np.stack([im.detections for im in ims])).cuda()
    d_msk = torch.from_numpy(np.stack([im.mrcnn_mask for im in ims])).cuda()
    eng.enqueue(d_det, d_msk)
    images = [synth.synth_rgb_image(rng, *hw) for _ in ims]
    colors = visualize.random_colors(40, rng=random.Random(11))
    counts, boxes, _, _ = eng.fetch_meta()
    masks = eng.canvas_view(0, int(counts[0])).cpu().numpy().view(np.bool_)
    for alpha in (0.5, 0.3, 1.0, 0.0):
        ref = oracle.composite_instances(images[0], boxes[0, :int(counts[0])], masks, colors, alpha)
        got = visualize.composite_batch(eng, images, colors, alpha)[0].cpu().numpy()
        assert np.array_equal(got, ref), alpha


def test_out_of_range_colours(cuda_device):
    """Colours above 1 push values past 255 (uint32 working copy, wrapped by the final uint8
    cast): the per-pixel float64 blend must still match the oracle."""
    rng = np.random.default_rng(46)
    hw = (90, 120)
    im = synth.make_batch(46, 1, hw, 15, num_classes=3, max_instances=16)[0]
    boxes, _, _, masks = api_utils.unmold_detections(*item_of(im))
    image = synth.synth_rgb_image(rng, *hw)
    colors = [(1.7, 0.2, 3.1)] * boxes.shape[0]
    ref = oracle.composite_instances(image, boxes, masks, colors, 0.5)
    got = visualize.apply_masks(image, boxes, masks, colors, 0.5)
    assert np.array_equal(got, ref)
