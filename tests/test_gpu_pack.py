"""Bit-packed mask transport (extension): the packed planes must equal
np.packbits of the masks unmold_detections returns, and unpacking must restore them exactly."""
import numpy as np
import pytest

from matterport_maskrcnn_with_tensorflow_serving_b200 import api_utils, synth

from helpers import item_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw,n,R", [((96, 128), 12, 16), ((75, 333), 37, 40), ((64, 21), 5, 8),
                                    ((40, 260), 0, 4), ((300, 517), 100, 100),
                                    ((1024, 1024), 100, 100)])
def test_packed_masks_equal_packbits(cuda_device, hw, n, R):
    im = synth.make_batch(61, 1, hw, n, num_classes=4, max_instances=R)[0]
    b, c, s, m = api_utils.unmold_detections(*item_of(im, np.float32))
    (pb, pc, ps, packed), = api_utils.unmold_detections_packed_batch([item_of(im, np.float32)])
    assert np.array_equal(pb, b) and np.array_equal(pc, c) and np.array_equal(ps, s)
    H, W = hw
    assert packed.dtype == np.uint8 and packed.shape == (b.shape[0], H, (W + 7) // 8)
    if b.shape[0]:
        ref = np.packbits(m.transpose(2, 0, 1), axis=-1)
        assert np.array_equal(packed, ref)
    restored = api_utils.unpack_masks(packed, W)
    assert restored.shape == m.shape and np.array_equal(restored, m)


def test_packed_batch_ragged(cuda_device):
    ims = synth.make_batch(62, 3, (120, 200), (0, 20), num_classes=3, max_instances=20)
    res = api_utils.unmold_detections_packed_batch([item_of(im, np.float32) for im in ims])
    for (b, c, s, packed), im in zip(res, ims):
        rb, rc, rs, rm = api_utils.unmold_detections(*item_of(im, np.float32))
        assert np.array_equal(b, rb)
        assert np.array_equal(api_utils.unpack_masks(packed, 200), rm)
