"""Bit-packed mask transport (extension): the packed planes must equal
np.packbits of the masks unmold_detections returns, and unpacking must restore them exactly."""
import numpy as np
import pytest

from matterport_maskrcnn_with_tensorflow_serving_b200 import api_utils, synth

from helpers import item_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("direct", [True, False], ids=["expand_packed", "pack_kernel"])
@pytest.mark.parametrize("hw,n,R", [((96, 128), 12, 16), ((75, 333), 37, 40), ((64, 21), 5, 8),
                                    ((40, 260), 0, 4), ((300, 517), 100, 100),
                                    ((1024, 1024), 100, 100), ((800, 1333), 61, 100),
                                    ((33, 1000), 7, 8), ((17, 9), 3, 4),
                                    # pack kernel, N % 4 == 0, staged form: ragged last block,
                                    # tiny image, odd width, rows not 16-byte aligned; and the
                                    # direct form (224 slots do not fit the staging buffer) and
                                    # pack_bytes_kernel (226: N % 4 != 0 there)
                                    ((37, 1000), 100, 100), ((21, 36), 8, 8), ((50, 1333), 16, 16),
                                    ((30, 1100), 100, 100), ((44, 1333), 100, 100),
                                    ((120, 300), 224, 224), ((120, 300), 226, 226)])
def test_packed_masks_equal_packbits(cuda_device, hw, n, R, direct):
    """Both producers of the packed layout -- the expand kernel that writes bits directly
    (mrx_mask_expand_packed) and the pack kernel over a byte canvas (mrx_pack_masks) -- must give
    exactly np.packbits of the bool masks `unmold_detections` returns."""
    im = synth.make_batch(61, 1, hw, n, num_classes=4, max_instances=R)[0]
    b, c, s, m = api_utils.unmold_detections(*item_of(im, np.float32))
    (pb, pc, ps, packed), = api_utils.unmold_detections_packed_batch([item_of(im, np.float32)],
                                                                     direct=direct)
    assert np.array_equal(pb, b) and np.array_equal(pc, c) and np.array_equal(ps, s)
    H, W = hw
    assert packed.dtype == np.uint8 and packed.shape == (b.shape[0], H, (W + 7) // 8)
    if b.shape[0]:
        ref = np.packbits(m.transpose(2, 0, 1), axis=-1)
        assert np.array_equal(packed, ref)
    restored = api_utils.unpack_masks(packed, W)
    assert restored.shape == m.shape and np.array_equal(restored, m)


@pytest.mark.parametrize("direct", [True, False], ids=["expand_packed", "pack_kernel"])
def test_packed_batch_ragged(cuda_device, direct):
    ims = synth.make_batch(62, 3, (120, 200), (0, 20), num_classes=3, max_instances=20)
    res = api_utils.unmold_detections_packed_batch([item_of(im, np.float32) for im in ims],
                                                   direct=direct)
    for (b, c, s, packed), im in zip(res, ims):
        rb, rc, rs, rm = api_utils.unmold_detections(*item_of(im, np.float32))
        assert np.array_equal(b, rb)
        assert np.array_equal(api_utils.unpack_masks(packed, 200), rm)


@pytest.mark.parametrize("hw,n,R,batch", [((1024, 1024), 100, 100, 3), ((2160, 3840), 50, 50, 1),
                                          ((800, 1333), (1, 100), 100, 4),
                                          ((96, 160), (0, 16), 16, 9)])
def test_expand_packed_equals_byte_canvas_on_device(cuda_device, hw, n, R, batch):
    """Full-size shapes (BASELINE.json configs[1], [2], [3]): the directly packed output equals
    the bit-packing of the byte canvas the headline kernel writes, plane for plane, compared on
    the device (no oracle in the loop: two independent kernels, one arithmetic)."""
    import torch

    from matterport_maskrcnn_with_tensorflow_serving_b200.engine import UnmoldEngine, make_geom

    ims = synth.make_batch(63, batch, hw, n, num_classes=5, max_instances=R)
    eng = UnmoldEngine(batch, R, (28, 28), 5)
    eng.plan([make_geom(im.original_image_shape, im.image_shape, im.window) for im in ims])
    d_det = torch.from_numpy(np.stack([im.detections for im in ims])).cuda()
    d_msk = torch.from_numpy(np.stack([im.mrcnn_mask for im in ims])).cuda()
    eng.enqueue(d_det, d_msk)
    d_packed, off = eng.enqueue_expand_packed()
    direct = d_packed[:int(off[-1])].clone()
    d_packed.fill_(0xAA)
    d_packed2, off2 = eng.pack_masks()
    counts, boxes, cls, scores = eng.fetch_meta()
    H, W = hw
    wb = (W + 7) // 8
    weights = (2 ** torch.arange(7, -1, -1, device="cuda")).to(torch.int32)
    for b in range(batch):
        k = int(counts[b])
        if k == 0:
            continue
        lo = int(off[b])
        a = direct[lo:lo + k * H * wb]
        c = d_packed2[lo:lo + k * H * wb]
        assert torch.equal(a, c), f"image {b}: direct and pack kernel differ"
        # and both equal packbits of the canvas, computed with torch on the device for one plane
        i = k // 2
        plane = eng.canvas_view(b, k)[:, :, i].to(torch.int32)
        pad = wb * 8 - W
        if pad:
            plane = torch.nn.functional.pad(plane, (0, pad))
        ref = (plane.view(H, wb, 8) * weights).sum(-1).to(torch.uint8)
        assert torch.equal(a.view(k, H, wb)[i], ref)


def test_byte_canvas_after_a_canvasless_plan(cuda_device):
    """Regression: a packed / RLE call plans a new geometry without a byte canvas; the next
    byte-canvas call on the same cached engine must size the canvas for THAT geometry."""
    small = synth.make_batch(64, 1, (40, 56), 5, num_classes=4, max_instances=8)[0]
    big = synth.make_batch(65, 1, (300, 260), 8, num_classes=4, max_instances=8)[0]
    api_utils.unmold_detections(*item_of(small, np.float32))          # small canvas cached
    (pb, pc, ps, packed), = api_utils.unmold_detections_packed_batch([item_of(big, np.float32)])
    b, c, s, m = api_utils.unmold_detections(*item_of(big, np.float32))
    assert m.shape == (300, 260, b.shape[0])
    assert np.array_equal(packed, np.packbits(m.transpose(2, 0, 1), axis=-1))


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("upload", ["copy", "zero_copy"])
def test_streaming_unmolder_modes(cuda_device, packed, upload):
    """engine.StreamingUnmolder over three streams: every combination of output layout (byte
    canvas / bit-packed) and mask upload (copied, or read in place from pinned memory) returns, batch after batch, the bytes the plain engine produces."""
    import torch

    from matterport_maskrcnn_with_tensorflow_serving_b200.engine import (StreamingUnmolder, UnmoldEngine,
                                                                         make_geom)

    n_img, R, hw = 5, 12, (96, 136)
    batches = [synth.make_batch(700 + k, n_img, hw, (0, 12), num_classes=6, max_instances=R)
               for k in range(4)]
    geoms = [make_geom(im.original_image_shape, im.image_shape, im.window) for im in batches[0]]
    ref = UnmoldEngine(n_img, R, (28, 28), 6)
    ref.plan(geoms)
    want = []
    for ims in batches:
        d_det = torch.from_numpy(np.stack([im.detections for im in ims])).cuda()
        d_msk = torch.from_numpy(np.stack([im.mrcnn_mask for im in ims])).cuda()
        ref.enqueue(d_det, d_msk)
        if packed:
            d_out, off = ref.enqueue_expand_packed()
            total = int(off[-1])
        else:
            d_out, total = ref.d_canvas, int(ref._offsets[n_img])
        counts, boxes, _, _ = ref.fetch_meta()
        want.append((counts.copy(), d_out[:total].cpu().clone(), ref.packed_layout()[0] if packed else ref._offsets))
    eng = UnmoldEngine(n_img, R, (28, 28), 6)
    sm = StreamingUnmolder(eng, geoms, packed=packed, mask_upload=upload)
    pinned = [(torch.from_numpy(np.stack([im.detections for im in ims])).pin_memory(),
               torch.from_numpy(np.stack([im.mrcnn_mask for im in ims])).pin_memory()) for ims in batches]
    got = {}
    for k, (h_det, h_msk) in enumerate(pinned):
        sm.submit(h_det, h_msk)
        if k:
            c, b, o = sm.wait(k - 1)
            got[k - 1] = (c.clone(), o.clone())
    c, b, o = sm.wait(len(pinned) - 1)
    got[len(pinned) - 1] = (c.clone(), o.clone())
    H, W = hw
    for k in range(4):
        wc, wout, woff = want[k]
        gc, gout = got[k]
        assert np.array_equal(gc.numpy(), wc)
        for i in range(n_img):
            nb = int(wc[i]) * H * ((W + 7) // 8) if packed else H * W * int(wc[i])
            lo = int(woff[i])
            assert torch.equal(gout[lo:lo + nb], wout[lo:lo + nb]), (k, i)
