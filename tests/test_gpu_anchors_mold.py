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
            b, c, s, m = serve.do_inference(img)
            assert np.array_equal(b, ref[0]) and np.array_equal(c, ref[1])
            assert np.array_equal(s, ref[2]) and s.dtype == ref[2].dtype
            assert m.shape == ref[3].shape and m.dtype == np.bool_
            assert (m != ref[3]).sum() == 0      # seeded data: no pixel inside the 1e-6 band
        assert seen["dtypes"] == (np.float32, np.float32, np.float32)    # serve.py:117-119
    finally:
        serve.set_predict_fn(None)
