"""GPU parity of get_anchors (serve.py:105) and the mold step (serve.py:83-107)."""
import numpy as np
import pytest

import oracle
from matterport_maskrcnn_with_tensorflow_serving_b200 import api_utils, serve, synth
from matterport_maskrcnn_with_tensorflow_serving_b200.engine import AnchorGenerator, Molder
from matterport_maskrcnn_with_tensorflow_serving_b200.model_configs import MaskRCNNServingConfig

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(512, 512), (640, 640), (1024, 1024), (2048, 2048),
                                (1000, 1000), (800, 1344), (1030, 770), (64, 64)])
def test_anchors_bit_exact(cuda_device, hw):
    """fp64 device arithmetic in numpy's operation order, one rounding to fp32: bit-exact."""
    shape = (hw[0], hw[1], 3)
    ref = oracle.get_anchors(shape)
    got = api_utils.get_anchors(shape)
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_anchors_known_answer(cuda_device):
    a = api_utils.get_anchors((1024, 1024, 3))
    assert a.shape == (261888, 4)
    np.testing.assert_allclose(a[0] * 1023, [-22.627417, -11.313708, 21.627417, 10.313708],
                               rtol=0, atol=2e-5)
    assert api_utils.get_anchors((1024, 1024, 3)) is a      # memoised by shape


def test_anchors_custom_config(cuda_device):
    class Cfg(MaskRCNNServingConfig):
        RPN_ANCHOR_SCALES = (8, 16, 32, 64, 128)
        RPN_ANCHOR_RATIOS = [0.25, 0.5, 1, 2, 4]
        RPN_ANCHOR_STRIDE = 2
        BACKBONE_STRIDES = [4, 8, 16, 32, 64]

    shape = (384, 512, 3)
    ref = oracle.get_anchors(shape, Cfg)
    got = AnchorGenerator(Cfg).get_anchors(shape)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("src,dst", [((480, 640), (640, 640)), ((1080, 1920), (640, 640)),
                                     ((1280, 1280), (640, 640)), ((333, 517), (1024, 1024)),
                                     ((100, 100), (37, 91)), ((640, 640), (640, 640)),
                                     ((7, 9), (64, 64))])
def test_cv2_resize_bit_exact(cuda_device, src, dst):
    """OpenCV's fixed-point INTER_LINEAR (and its exact-2x area shortcut) vs the real cv2."""
    import cv2
    import torch

    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(src[0], src[1], 3), dtype=np.uint8)
    ref = cv2.resize(img, (dst[1], dst[0]))
    m = Molder(MaskRCNNServingConfig)
    got = m.cv2_resize_device(torch.from_numpy(img).cuda(), dst).cpu().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("hw", [(1024, 1024), (800, 1333), (480, 640), (2160, 3840), (300, 200)])
def test_mold_image_bit_exact(cuda_device, hw):
    """resize_image(square) + mold_image: uint8 image after the truncating cast and the
    float64 / float32 molded tensors are bit-exact (device fp64 in scipy's op order)."""
    import torch

    rng = np.random.default_rng(2)
    img = synth.synth_rgb_image(rng, *hw)
    ref_u8, window, scale, padding, crop = oracle.resize_image(
        img, min_dim=800, max_dim=1024, min_scale=0, mode="square")
    ref_molded = oracle.mold_image(ref_u8)
    m = Molder(MaskRCNNServingConfig)
    d64, u8, win, sc, pad = m.mold_device(torch.from_numpy(img).cuda(), np.float64, want_u8=True)
    assert win == window and sc == scale and pad == padding
    assert np.array_equal(u8.cpu().numpy(), ref_u8)
    got64 = d64.cpu().numpy()
    assert got64.dtype == ref_molded.dtype == np.float64
    assert np.array_equal(got64, ref_molded)
    d32, _, _, _, _ = m.mold_device(torch.from_numpy(img).cuda(), np.float32)
    assert np.array_equal(d32.cpu().numpy(), ref_molded.astype(np.float32))


@pytest.mark.parametrize("img_size", [640, None])
def test_preprocess_input_matches_reference_flow(cuda_device, img_size):
    rng = np.random.default_rng(3)
    img = synth.synth_rgb_image(rng, 480, 640)
    ref_molded, ref_meta, ref_anchors, ref_window = oracle.preprocess_input(img, img_size)
    molded, meta, anchors, window = serve.preprocess_input(img, img_size)
    assert window == ref_window
    assert molded.dtype == ref_molded.dtype and np.array_equal(molded, ref_molded)
    assert np.array_equal(meta, ref_meta)
    assert np.array_equal(anchors.view(np.uint32), ref_anchors.view(np.uint32))


def test_do_inference_flow_lists_and_wire_bytes(cuda_device):
    """serve.py:141-154 end to end around an injected TF-Serving call: the model outputs arrive
    once as float lists (what `float_val` yields, serve.py:131-136) and once as serialized
    TensorProto messages (wire.py); both must equal the oracle's unmold of the same outputs."""
    from matterport_maskrcnn_with_tensorflow_serving_b200 import configs as cf
    from matterport_maskrcnn_with_tensorflow_serving_b200 import wire

    rng = np.random.default_rng(9)
    img = synth.synth_rgb_image(rng, 480, 640)
    molded, meta, anchors, window = serve.preprocess_input(img, cf.IMAGE_SIZE)
    im = synth.make_image(rng, img.shape[:2], 17, num_classes=cf.OUT_MASK_SHAPE[-1],
                          max_instances=cf.OUT_DETECTION_SHAPE[0], mold=(molded.shape, window))
    det32, msk32 = im.detections, im.mrcnn_mask
    ref = oracle.unmold_detections(det32.astype(np.float64), msk32.astype(np.float64),
                                   img.shape, molded.shape, window)
    seen = {}

    def predict_lists(molded_f32, meta_f32, anchors_f32):
        seen["dtypes"] = (molded_f32.dtype, meta_f32.dtype, anchors_f32.dtype)
        return det32.reshape(-1).tolist(), msk32.reshape(-1).tolist()

    def predict_wire(molded_f32, meta_f32, anchors_f32):
        return (wire.ndarray_to_tensor_proto_bytes(det32[None]),
                wire.ndarray_to_tensor_proto_bytes(msk32[None]))

    try:
        for fn in (predict_lists, predict_wire):
            serve.set_predict_fn(fn)
            b, c, s, m = serve.do_inference_unmolded(img)
            assert np.array_equal(b, ref[0]) and np.array_equal(c, ref[1])
            assert np.array_equal(s, ref[2]) and s.dtype == ref[2].dtype
            assert m.shape == ref[3].shape and m.dtype == np.bool_
            assert (m != ref[3]).sum() == 0      # seeded data: no pixel inside the 1e-6 band
        assert seen["dtypes"] == (np.float32, np.float32, np.float32)    # serve.py:117-119
    finally:
        serve.set_predict_fn(None)


def _fake_model(rng, imgs, n_inst):
    """Synthetic TF-Serving: per request image, detections + masks laid out for the molded
    geometry `preprocess_input` produces (looked up by the molded tensor's id in call order)."""
    from matterport_maskrcnn_with_tensorflow_serving_b200 import configs as cf

    outs = []
    for img in imgs:
        molded, meta, anchors, window = serve.preprocess_input(img, cf.IMAGE_SIZE, np.float32)
        im = synth.make_image(rng, img.shape[:2], n_inst, num_classes=cf.OUT_MASK_SHAPE[-1],
                              max_instances=cf.OUT_DETECTION_SHAPE[0],
                              mold=(molded.shape, window))
        outs.append((im, molded, meta, window))
    calls = {"k": 0, "inputs": []}

    def predict(molded_f32, meta_f32, anchors_f32):
        k = calls["k"]
        calls["k"] += 1
        calls["inputs"].append((molded_f32, meta_f32))
        im = outs[k % len(outs)][0]
        return im.detections.reshape(-1).tolist(), im.mrcnn_mask.reshape(-1).tolist()

    return outs, predict, calls


def test_do_inference_returns_the_saved_overlay_path(cuda_device, tmp_path):
    """serve.py:141-173: `do_inference(img)` returns `save_path` of media/mask-<uuid>.png whose
    pixels are the mask overlay of display_instances (oracle.composite_instances of the
    oracle's unmold), computed without the masks leaving the device."""
    import random

    import cv2

    from matterport_maskrcnn_with_tensorflow_serving_b200 import visualize

    rng = np.random.default_rng(19)
    img = synth.synth_rgb_image(rng, 300, 420)
    outs, predict, calls = _fake_model(rng, [img], 23)
    im, molded, meta, window = outs[0]
    rb, rc, rs, rm = oracle.unmold_detections(
        im.detections.astype(np.float64), im.mrcnn_mask.astype(np.float64), img.shape,
        molded.shape, window)
    colors = visualize.random_colors(100, rng=random.Random(5))
    ref = oracle.composite_instances(img, rb, rm, colors)
    serve.set_predict_fn(predict)
    try:
        path = serve.do_inference(img, colors=colors, media_dir=str(tmp_path))
    finally:
        serve.set_predict_fn(None)
    assert isinstance(path, str) and path.startswith(str(tmp_path))
    name = path[len(str(tmp_path)) + 1:]
    assert name.startswith("mask-") and name.endswith(".png") and len(name) == len("mask-.png") + 36
    got = cv2.cvtColor(cv2.imread(path, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_do_inference_batch_matches_single_calls(cuda_device, tmp_path):
    """Batched surface: images of different sizes through
    `do_inference_batch` give, per image, exactly the picture `do_inference` gives alone; the
    batched pre-processing sends the RPC the same tensors the single path sends."""
    import random

    import cv2

    from matterport_maskrcnn_with_tensorflow_serving_b200 import visualize

    rng = np.random.default_rng(29)
    imgs = [synth.synth_rgb_image(rng, 300, 420), synth.synth_rgb_image(rng, 640, 640),
            synth.synth_rgb_image(rng, 222, 150)]
    outs, predict, calls = _fake_model(rng, imgs, 11)
    colors = visualize.random_colors(100, rng=random.Random(6))
    serve.set_predict_fn(predict)
    try:
        calls["k"] = 0
        calls["inputs"].clear()
        singles = [serve.do_inference(img, colors=colors, media_dir=str(tmp_path)) for img in imgs]
        single_inputs = list(calls["inputs"])
        calls["k"] = 0
        calls["inputs"].clear()
        paths = serve.do_inference_batch(imgs, colors=colors, media_dir=str(tmp_path))
        batch_inputs = list(calls["inputs"])
    finally:
        serve.set_predict_fn(None)
    assert len(paths) == len(imgs) and len(set(paths + singles)) == 2 * len(imgs)
    for a, b in zip(singles, paths):
        assert np.array_equal(cv2.imread(a), cv2.imread(b))
    for (m1, meta1), (m2, meta2) in zip(single_inputs, batch_inputs):
        assert m1.dtype == m2.dtype == np.float32 and np.array_equal(m1, m2)
        assert np.array_equal(meta1, meta2)


@pytest.mark.parametrize("img_size", [640, None])
def test_preprocess_input_batch_matches_oracle(cuda_device, img_size):
    """One cv2.resize launch + one mold launch for the batch == the oracle per image."""
    rng = np.random.default_rng(31)
    if img_size is None:
        imgs = [synth.synth_rgb_image(rng, 480, 640) for _ in range(3)]
    else:
        imgs = [synth.synth_rgb_image(rng, 480, 640), synth.synth_rgb_image(rng, 1280, 1280),
                synth.synth_rgb_image(rng, 97, 333), synth.synth_rgb_image(rng, 640, 640)]
    molded, metas, anchors, windows = serve.preprocess_input_batch(imgs, img_size, np.float64)
    for b, img in enumerate(imgs):
        ref_molded, ref_meta, ref_anchors, ref_window = oracle.preprocess_input(img, img_size)
        assert windows[b] == ref_window
        assert molded[b].dtype == ref_molded.dtype and np.array_equal(molded[b], ref_molded)
        assert np.array_equal(metas[b], ref_meta)
        assert np.array_equal(anchors.view(np.uint32), ref_anchors.view(np.uint32))
    m32, _, _, _ = serve.preprocess_input_batch(imgs, img_size)      # wire dtype by default
    assert m32.dtype == np.float32 and np.array_equal(m32, molded.astype(np.float32))
