"""CPU tests of the oracle (the checker itself): known answers derived from the published
upstream algorithm, agreement of `resize` with a hand-written formula, semantic edge cases,
and the committed golden vectors.  No GPU."""
import hashlib
import os

import numpy as np
import pytest

import oracle
from matterport_maskrcnn_with_tensorflow_serving_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------ anchors
def test_anchor_count_and_first_anchor_1024():
    a = oracle.get_anchors((1024, 1024, 3))
    assert a.shape == (261888, 4) and a.dtype == np.float32   # 3*(256^2+128^2+64^2+32^2+16^2)
    # level P2, cell (0,0), ratio 0.5: h = 32/sqrt(.5), w = 32*sqrt(.5), centred on (0,0)
    h, w = 32 / np.sqrt(0.5), 32 * np.sqrt(0.5)
    want = np.array([-h / 2, -w / 2, h / 2 - 1, w / 2 - 1]) / 1023
    np.testing.assert_allclose(a[0], want.astype(np.float32), rtol=0, atol=1e-9)


def test_anchor_order_level_y_x_ratio():
    cfg = oracle.OracleConfig
    a = oracle.generate_pyramid_anchors(cfg.RPN_ANCHOR_SCALES, cfg.RPN_ANCHOR_RATIOS,
                                        oracle.compute_backbone_shapes(cfg, (128, 192, 3)),
                                        cfg.BACKBONE_STRIDES, 1)
    R, W0 = 3, 192 // 4
    cy = (a[:, 0] + a[:, 2]) / 2
    cx = (a[:, 1] + a[:, 3]) / 2
    # ratio innermost: first three anchors share a centre; x next; then y
    assert np.allclose(cy[:R], 0) and np.allclose(cx[:R], 0)
    assert np.isclose(cx[R], 4) and np.isclose(cy[R], 0)
    assert np.isclose(cy[R * W0], 4) and np.isclose(cx[R * W0], 0)
    n0 = R * (128 // 4) * W0
    assert np.isclose(cx[n0 + R], 8)       # level P3 starts after all of P2, stride 8


def test_anchor_ceil_for_non_multiple_sizes():
    a = oracle.get_anchors((1030, 770, 3))
    hs = [int(np.ceil(1030 / s)) for s in (4, 8, 16, 32, 64)]
    ws = [int(np.ceil(770 / s)) for s in (4, 8, 16, 32, 64)]
    assert a.shape[0] == 3 * sum(h * w for h, w in zip(hs, ws))


def test_anchor_golden():
    g = np.load(os.path.join(GOLD, "anchors.npz"))
    for key in ["64x64", "256x320", "512x512", "1024x1024", "800x1344"]:
        h, w = [int(v) for v in key.split("x")]
        a = oracle.get_anchors((h, w, 3))
        assert a.shape[0] == int(g[key + "_count"])
        assert hashlib.sha256(a.tobytes()).digest() == g[key + "_sha256"].tobytes()
        np.testing.assert_array_equal(a[:6], g[key + "_head"])
        np.testing.assert_array_equal(a[-6:], g[key + "_tail"])


# ------------------------------------------------------------------ boxes
def test_norm_denorm_roundtrip_and_half_even():
    boxes = np.array([[0, 0, 10, 20], [5, 7, 1024, 1024]])
    n = oracle.norm_boxes(boxes, (1024, 1024))
    assert n.dtype == np.float32
    np.testing.assert_array_equal(oracle.denorm_boxes(n, (1024, 1024)), boxes)
    # np.around is round-half-to-even: 2.5 -> 2, 3.5 -> 4, (1.5 + 1) -> 2, (0.5 + 1) -> 2
    got = oracle.denorm_boxes(np.array([[2.5 / 10, 3.5 / 10, 1.5 / 10, 0.5 / 10]]), (11, 11))
    np.testing.assert_array_equal(got, [[2, 4, 2, 2]])


# ------------------------------------------------------------------ resize
@pytest.mark.parametrize("shape", [(28, 28), (100, 37), (5, 9), (300, 280), (1, 1), (2, 60)])
def test_resize_matches_explicit_formula(shape):
    rng = np.random.default_rng(0)
    m = rng.random((28, 28))
    np.testing.assert_allclose(oracle.resize(m, shape), oracle.resize_explicit(m, shape),
                               rtol=0, atol=1e-14)


def test_resize_identity_and_zero_border():
    rng = np.random.default_rng(1)
    m = rng.random((28, 28)) * 0.5 + 0.5          # all >= 0.5
    assert np.array_equal(oracle.resize(m, (28, 28)), m)
    big = oracle.resize(m, (280, 280))
    # zero-border bilinear attenuates the outer half source pixel: corner value = m[0,0]*(.55)^2
    assert np.isclose(big[0, 0], m[0, 0] * 0.55 * 0.55)
    assert big[0, 0] < 0.5 <= big[140, 140]


def test_resize_golden():
    g = np.load(os.path.join(GOLD, "resize.npz"))
    tile = g["tile"].astype(np.float64)
    for key in g.files:
        if key.startswith("out_"):
            bh, bw = [int(v) for v in key[4:].split("x")]
            np.testing.assert_array_equal(oracle.resize(tile, (bh, bw)), g[key])


# ------------------------------------------------------------------ unmold
def _run(im, dtype=np.float64):
    return oracle.unmold_detections(im.detections.astype(dtype), im.mrcnn_mask.astype(dtype),
                                    im.original_image_shape, im.image_shape, im.window)


def test_unmold_config1_plumbing():
    """BASELINE.json config 1: one 1024x1024 image, 10 synthetic detections, CPU only."""
    im = synth.make_batch(1, 1, (1024, 1024), 10)[0]
    boxes, class_ids, scores, masks = _run(im)
    assert boxes.shape == (10, 4) and boxes.dtype == np.int32
    assert class_ids.dtype == np.int32 and scores.shape == (10,)
    assert masks.shape == (1024, 1024, 10) and masks.dtype == np.bool_
    for i, (y1, x1, y2, x2) in enumerate(boxes):
        assert masks[:, :, i].sum() == masks[y1:y2, x1:x2, i].sum() > 0     # support inside box
    # scale 1, window = whole image: the pixel boxes survive the normalise/denormalise trip
    px = oracle.denorm_boxes(im.detections[:10, :4], (1024, 1024))
    np.testing.assert_array_equal(boxes, px)


def test_unmold_truncates_at_first_zero_class_and_drops_zero_area():
    rng = np.random.default_rng(5)
    im = synth.make_image(rng, (1024, 1024), 20, num_classes=7, zero_area_rows=(0, 7, 19))
    im.detections[15, 4] = 0.0
    boxes, class_ids, scores, masks = _run(im)
    assert boxes.shape[0] == 13 and masks.shape[2] == 13
    keep = [i for i in range(15) if i not in (0, 7)]
    np.testing.assert_array_equal(class_ids, im.detections[keep, 4].astype(np.int32))
    np.testing.assert_array_equal(scores, im.detections[keep, 5].astype(np.float64))


def test_unmold_empty_and_leading_dim():
    rng = np.random.default_rng(6)
    im = synth.make_image(rng, (64, 80), 0, num_classes=3)
    b, c, s, m = _run(im)
    assert b.shape == (0, 4) and m.shape == (64, 80, 0)
    im = synth.make_batch(3, 1, (120, 90), 6, num_classes=4)[0]
    a = _run(im)
    bb = oracle.unmold_detections(im.detections[None].astype(np.float64),
                                  im.mrcnn_mask[None].astype(np.float64),
                                  im.original_image_shape, im.image_shape, im.window)
    for x, y in zip(a, bb):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("name", ["unmold_small", "unmold_coco_shape"])
def test_unmold_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    b, c, s, m = oracle.unmold_detections(
        g["detections"].astype(np.float64), g["mrcnn_mask"].astype(np.float64),
        tuple(g["original_image_shape"]), tuple(g["image_shape"]), tuple(g["window"]))
    np.testing.assert_array_equal(b, g["boxes"])
    np.testing.assert_array_equal(c, g["class_ids"])
    np.testing.assert_array_equal(s, g["scores"])
    want = np.unpackbits(g["masks_packed"])[:m.size].reshape(tuple(g["masks_shape"])).astype(bool)
    np.testing.assert_array_equal(m, want)


# ------------------------------------------------------------------ mold
def test_resize_image_square_geometry():
    img = np.zeros((800, 1333, 3), dtype=np.uint8)
    out, window, scale, padding, crop = oracle.resize_image(img, min_dim=800, max_dim=1024,
                                                            min_scale=0, mode="square")
    assert out.shape == (1024, 1024, 3) and out.dtype == np.uint8
    assert window == (204, 0, 819, 1024) and np.isclose(scale, 1024 / 1333)
    assert padding == [(204, 205), (0, 0), (0, 0)] and crop is None
    assert synth.square_mold_geometry(800, 1333)[:2] == ((1024, 1024, 3), window)


def test_mold_image_and_meta():
    img = np.full((4, 4, 3), 200, dtype=np.uint8)
    m = oracle.mold_image(img)
    assert m.dtype == np.float64
    np.testing.assert_allclose(m[0, 0], [200 - 123.7, 200 - 116.8, 200 - 103.9])
    meta = oracle.compose_image_meta(0, (4, 4, 3), (8, 8, 3), (2, 2, 6, 6), 1, np.zeros(5, np.int32))
    assert meta.shape == (12 + 5,)


def test_preprocess_input_shapes():
    rng = np.random.default_rng(3)
    img = synth.synth_rgb_image(rng, 240, 320)
    molded, meta, anchors, window = oracle.preprocess_input(img, 640)
    assert molded.shape == (1024, 1024, 3) and molded.dtype == np.float64
    assert window == (112, 112, 912, 912)         # 640x640 -> scale 800/640 -> 800x800, centred in 1024
    assert meta.shape == (12 + 81,) and anchors.shape == (261888, 4)


def test_composite_known_answers():
    """apply_mask arithmetic (upstream visualize.py): float64 blend, truncating uint32 store,
    instances applied in order, all-zero boxes skipped."""
    image = np.full((2, 2, 3), 100, dtype=np.uint8)
    masks = np.zeros((2, 2, 3), dtype=bool)
    masks[0, 0, 0] = masks[0, 0, 2] = True      # pixel (0,0): instances 0 and 2
    masks[1, 1, 1] = True                       # pixel (1,1): instance 1, whose box is all zeros
    boxes = np.array([[0, 0, 2, 2], [0, 0, 0, 0], [0, 0, 1, 1]])
    colors = [(1.0, 0.0, 0.5), (1.0, 1.0, 1.0), (0.0, 1.0, 0.25)]
    out = oracle.composite_instances(image, boxes, masks, colors, alpha=0.5)
    assert out.dtype == np.uint8
    # instance 0: 100*0.5 + 0.5*c*255 -> (177.5, 50, 113.75) -> (177, 50, 113)
    # instance 2 on top: (177*0.5 + 0, 50*0.5 + 127.5, 113*0.5 + 31.875) -> (88, 152, 88)
    assert out[0, 0].tolist() == [88, 152, 88]
    assert out[1, 1].tolist() == [100, 100, 100]    # skipped instance
    assert out[0, 1].tolist() == [100, 100, 100]
    assert len(oracle.random_colors(7)) == 7 and oracle.random_colors(4)[0] in [
        (1.0, 0.0, 0.0), (0.5, 1.0, 0.0), (0.0, 1.0, 1.0), (0.5, 0.0, 1.0)]


def test_composite_golden():
    """The committed overlay of the unmold_small masks (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "unmold_small.npz"))
    c = np.load(os.path.join(GOLD, "composite_small.npz"))
    shape = tuple(int(v) for v in g["masks_shape"])
    masks = np.unpackbits(g["masks_packed"], count=int(np.prod(shape))).reshape(shape).astype(bool)
    colors = [tuple(row) for row in c["colors"]]
    out = oracle.composite_instances(c["image"], g["boxes"], masks, colors, alpha=0.5)
    assert np.array_equal(out, c["overlay"])
    assert (out != c["image"]).any()


def _rle_by_loop(mask):
    """Independent restatement of pycocotools' rleEncode: walk the pixels in column-major order
    and count (plain Python loop, small masks only)."""
    h, w = mask.shape
    counts, prev, run = [], 0, 0
    for x in range(w):
        for y in range(h):
            v = int(mask[y, x])
            if v != prev:
                counts.append(run)
                run, prev = 0, v
            run += 1
    counts.append(run)
    return counts


def test_rle_known_answers_and_roundtrip():
    """COCO 'uncompressed RLE': column-major runs, starting with zeros."""
    m = np.array([[0, 1], [1, 1]], dtype=bool)            # column-major: 0 1 | 1 1
    assert oracle.rle_encode(m)["counts"].tolist() == [1, 3]
    assert oracle.rle_encode(np.ones((3, 2), bool))["counts"].tolist() == [0, 6]
    assert oracle.rle_encode(np.zeros((3, 2), bool))["counts"].tolist() == [6]
    m = np.zeros((4, 3), bool)
    m[3, 0] = m[0, 1] = True                               # run crosses the column seam
    assert oracle.rle_encode(m)["counts"].tolist() == [3, 2, 7]
    rng = np.random.default_rng(12)
    for shape in [(1, 1), (5, 7), (16, 3), (9, 31)]:
        for p in (0.1, 0.5, 0.9):
            m = rng.random(shape) < p
            r = oracle.rle_encode(m)
            assert r["size"] == list(shape) and int(r["counts"].sum()) == m.size
            assert r["counts"].tolist() == _rle_by_loop(m)
            assert np.array_equal(oracle.rle_decode(r), m)
            assert (r["counts"][1:] > 0).all()             # only the first run may be empty
