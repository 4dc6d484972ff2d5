"""COCO run-length masks computed on the device straight from the tiles (extension; SURVEY.md 8f
rank 4): every encoding must equal the oracle's RLE (pycocotools' uncompressed format) of the
bool mask `unmold_detections` returns for the same instance -- and therefore decode to it."""
import numpy as np
import pytest

import oracle
from matterport_maskrcnn_with_tensorflow_serving_b200 import api_utils, synth

from helpers import item_of

pytestmark = pytest.mark.gpu


def _check(ims, dtype=np.float32):
    items = [item_of(im, dtype) for im in ims]
    got = api_utils.unmold_detections_rle_batch(items)
    ref = api_utils.unmold_detections_batch(items)
    runs = 0
    for (b, c, s, rles), (rb, rc, rs, rm) in zip(got, ref):
        assert np.array_equal(b, rb) and np.array_equal(c, rc) and np.array_equal(s, rs)
        assert len(rles) == rb.shape[0]
        for i, rle in enumerate(rles):
            want = oracle.rle_encode(rm[:, :, i])
            assert rle["size"] == want["size"]
            assert rle["counts"].dtype == np.uint32
            assert np.array_equal(rle["counts"], want["counts"]), (i, rb[i])
            runs += len(rle["counts"])
    return runs


@pytest.mark.parametrize("hw,n,R,kw", [
    ((96, 128), 12, 16, {}),
    ((64, 96), 40, 40, dict(min_box=60, max_box_frac=1.0)),     # full-height / full-width boxes
    ((40, 56), 30, 32, dict(min_box=20, max_box_frac=1.0)),     # boxes touching every border
    ((150, 150), 30, 32, dict(min_box=1, max_box_frac=0.1)),    # boxes smaller than the tile
    ((75, 333), 37, 40, {}),                                    # widths that are no multiple of 32
    ((33, 1000), 7, 8, {}),
    ((17, 9), 3, 4, dict(min_box=1, max_box_frac=1.0)),
    ((64, 80), 0, 4, {}),                                       # nothing detected
])
def test_rle_equals_oracle_encoding(cuda_device, hw, n, R, kw):
    rng = np.random.default_rng(71)
    ims = [synth.make_image(rng, hw, n, num_classes=4, max_instances=R, **kw) for _ in range(3)]
    _check(ims)


def test_rle_touching_the_canvas_edges(cuda_device):
    """Hand-placed boxes: full canvas, full height at the left / right edge, bottom rows only,
    top rows only, last column only -- the seams between columns of the column-major order."""
    rng = np.random.default_rng(72)
    H, W = 48, 70
    boxes = [(0, 0, H, W), (0, 0, H, 9), (0, W - 11, H, W), (H - 7, 3, H, 40), (0, 5, 6, W),
             (0, W - 1, H, W), (10, 0, H, 1), (0, 20, H, 21), (H - 1, 0, H, W)]
    im = synth.make_image(rng, (H, W), len(boxes), num_classes=3, max_instances=12,
                          mold=((H, W, 3), (0, 0, H, W)))
    for i, (y1, x1, y2, x2) in enumerate(boxes):
        im.detections[i, :4] = synth._norm_boxes_f32(np.array([[y1, x1, y2, x2]], np.float64), (H, W))[0]
    b, c, s, m = api_utils.unmold_detections(*item_of(im, np.float32))
    assert [tuple(r) for r in b] == boxes
    _check([im])
    _check([im], np.float64)


def test_rle_full_size_batch(cuda_device):
    """BASELINE.json configs[1] shape, ragged second image; and the encodings are small."""
    ims = synth.make_batch(73, 1, (1024, 1024), 100) + synth.make_batch(74, 1, (1024, 1024), 37)
    runs = _check(ims)
    assert runs * 4 < 137 * 1024 * 1024 // 8      # far below even the bit-packed size
