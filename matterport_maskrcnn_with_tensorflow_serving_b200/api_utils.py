"""Drop-in for the reference's `api.helpers.utils` (imported as `api_utils`,
/root/reference/serve.py:21): same function names, positional arguments, return tuples and
NumPy value semantics, computed by the sm_100a kernels in libmrx.so.

    get_anchors(image_shape)                                         serve.py:105
    unmold_detections(detections, mrcnn_mask, original_image_shape,
                      image_shape, window)                           serve.py:147-154
    load_img(path)                                                   serve.py:86

Plus batched entry points the reference lacks (it is hard-wired to one image per call,
serve.py:48): `unmold_detections_batch`.
"""
from __future__ import annotations

import numpy as np

from . import _native as N
from .engine import AnchorGenerator, UnmoldEngine, make_geom
from .model_configs import mconfig as _default_config

_state = {"config": _default_config, "anchors": None, "engines": {}}


def set_config(config):
    """Use another Matterport-style config object (attribute names of mrcnn/config.py)."""
    _state["config"] = config
    _state["anchors"] = None


def get_config():
    return _state["config"]


def load_img(path):
    """serve.py:85-86: file -> HxWx3 uint8 RGB ndarray (file IO; stays on the host)."""
    import cv2

    img = cv2.imread(path, cv2.IMREAD_COLOR)
    if img is None:
        raise FileNotFoundError(path)
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def get_anchors(image_shape):
    """[A,4] float32 normalised (y1,x1,y2,x2) FPN anchors for a molded image shape;
    memoised by shape like upstream MaskRCNN.get_anchors."""
    if _state["anchors"] is None:
        _state["anchors"] = AnchorGenerator(_state["config"])
    return _state["anchors"].get_anchors(image_shape)


def _squeeze_inputs(detections, mrcnn_mask):
    detections = np.asarray(detections)
    mrcnn_mask = np.asarray(mrcnn_mask)
    # serve.py:131-136 reshapes to (-1, *cf.OUT_*_SHAPE): accept the leading unit dim
    if detections.ndim == 3 and detections.shape[0] == 1:
        detections = detections[0]
    if mrcnn_mask.ndim == 5 and mrcnn_mask.shape[0] == 1:
        mrcnn_mask = mrcnn_mask[0]
    if detections.ndim != 2 or detections.shape[1] != 6:
        raise ValueError(f"detections must be [R,6], got {detections.shape}")
    if mrcnn_mask.ndim != 4 or mrcnn_mask.shape[0] != detections.shape[0]:
        raise ValueError(f"mrcnn_mask must be [R,mh,mw,C] with R={detections.shape[0]}, "
                         f"got {mrcnn_mask.shape}")
    if detections.dtype not in (np.float32, np.float64):
        detections = detections.astype(np.float64)
    if mrcnn_mask.dtype not in (np.float32, np.float64):
        mrcnn_mask = mrcnn_mask.astype(np.float64)
    # zero-copy views of a received message (wire.py) are read-only; torch wants writable memory
    if not mrcnn_mask.flags.writeable:
        mrcnn_mask = mrcnn_mask.copy()
    if not detections.flags.writeable:
        detections = detections.copy()
    return np.ascontiguousarray(detections), np.ascontiguousarray(mrcnn_mask)


def _engine_for(batch, R, mh, mw, Cc, det_dtype, mask_dtype):
    import torch

    key = (torch.cuda.current_device(), R, mh, mw, Cc, np.dtype(det_dtype).str,
           np.dtype(mask_dtype).str)
    eng = _state["engines"].get(key)
    if eng is None or eng.B < batch:
        eng = UnmoldEngine(max(batch, 1), R, (mh, mw), Cc, det_dtype, mask_dtype)
        _state["engines"][key] = eng
    return eng


def unmold_detections_batch(items):
    """items: sequence of (detections, mrcnn_mask, original_image_shape, image_shape,
    window) with equal R / mask shape / dtypes.  Returns a list of
    (boxes, class_ids, scores, masks) exactly as `unmold_detections` would per image."""
    import torch

    N.require_cuda()
    if len(items) == 0:
        return []
    dets, masks, geoms = [], [], []
    for det, msk, osh, ish, win in items:
        d, m = _squeeze_inputs(det, msk)
        dets.append(d)
        masks.append(m)
        geoms.append(make_geom(osh, ish, win))
    d0, m0 = dets[0], masks[0]
    for d, m in zip(dets, masks):
        if d.shape != d0.shape or m.shape != m0.shape or d.dtype != d0.dtype or \
                m.dtype != m0.dtype:
            raise ValueError("all images of a batch must share shapes and dtypes")
    R, (mh, mw, Cc) = d0.shape[0], m0.shape[1:]
    n = len(items)
    eng = _engine_for(n, R, mh, mw, Cc, d0.dtype, m0.dtype)
    eng.plan(geoms)
    dev = eng.device
    d_det = torch.from_numpy(np.stack(dets)).to(dev)
    d_msk = torch.from_numpy(np.stack(masks) if n > 1 else masks[0][None]).to(dev)
    eng.enqueue(d_det, d_msk)
    counts, boxes, class_ids, scores = eng.fetch_meta()
    out = []
    for b in range(n):
        k = int(counts[b])
        H, W = geoms[b][0], geoms[b][1]
        if k == 0:
            full = np.empty((H, W, 0))            # upstream: np.empty(shape[:2] + (0,))
        else:
            full = eng.canvas_view(b, k).cpu().numpy().view(np.bool_)
        out.append((boxes[b, :k].copy(), class_ids[b, :k].copy(), scores[b, :k].copy(), full))
    return out


def unpack_masks(packed, width):
    """Inverse of the packed transport: uint8 [N, H, ceil(W/8)] -> bool [H, W, N] (the array
    `unmold_detections` returns)."""
    if packed.shape[0] == 0:
        return np.empty((packed.shape[1], width, 0))
    return np.unpackbits(packed, axis=-1, count=width).transpose(1, 2, 0).astype(np.bool_)


def unmold_detections_packed_batch(items):
    """EXTENSION (not the reference layout): like `unmold_detections_batch` but the masks come
    back bit-packed, uint8 [N, H, ceil(W/8)] with packed[n, y] == np.packbits(masks[y, :, n]) --
    8x less device -> host traffic; `unpack_masks(packed, W)` restores the reference array."""
    import torch

    N.require_cuda()
    if len(items) == 0:
        return []
    dets, masks, geoms = [], [], []
    for det, msk, osh, ish, win in items:
        d, m = _squeeze_inputs(det, msk)
        dets.append(d)
        masks.append(m)
        geoms.append(make_geom(osh, ish, win))
    R, (mh, mw, Cc) = dets[0].shape[0], masks[0].shape[1:]
    n = len(items)
    eng = _engine_for(n, R, mh, mw, Cc, dets[0].dtype, masks[0].dtype)
    eng.plan(geoms)
    d_det = torch.from_numpy(np.stack(dets)).to(eng.device)
    d_msk = torch.from_numpy(np.stack(masks)).to(eng.device)
    eng.enqueue(d_det, d_msk)
    d_packed, off = eng.pack_masks()
    counts, boxes, class_ids, scores = eng.fetch_meta()
    out = []
    for b in range(n):
        k = int(counts[b])
        H, W = int(geoms[b][0]), int(geoms[b][1])
        wb = (W + 7) // 8
        pk = d_packed[int(off[b]):int(off[b]) + k * H * wb].cpu().numpy().reshape(k, H, wb)
        out.append((boxes[b, :k].copy(), class_ids[b, :k].copy(), scores[b, :k].copy(), pk))
    return out


def unmold_overlay_batch(items, images, colors=None, alpha=0.5):
    """`unmold_detections` followed by the mask overlay of `visualize.display_instances`
    (serve.py:147-169) without moving the masks to the host: the [H,W,N] canvases stay on
    the device, only boxes, class ids, scores and the blended uint8 images come back.

    items as for `unmold_detections_batch`; images: the original uint8 HxWx3 images;
    colors: RGB triples (shared list, or one list per image; default: `random_colors(R)`).
    Returns a list of (boxes, class_ids, scores, overlay_uint8)."""
    import torch

    from . import visualize

    N.require_cuda()
    if len(items) == 0:
        return []
    dets, masks, geoms = [], [], []
    for det, msk, osh, ish, win in items:
        d, m = _squeeze_inputs(det, msk)
        dets.append(d)
        masks.append(m)
        geoms.append(make_geom(osh, ish, win))
    R, (mh, mw, Cc) = dets[0].shape[0], masks[0].shape[1:]
    n = len(items)
    eng = _engine_for(n, R, mh, mw, Cc, dets[0].dtype, masks[0].dtype)
    eng.plan(geoms)
    d_det = torch.from_numpy(np.stack(dets)).to(eng.device)
    d_msk = torch.from_numpy(np.stack(masks)).to(eng.device)
    eng.enqueue(d_det, d_msk)
    if colors is None:
        colors = visualize.random_colors(R)
    overlays = visualize.composite_batch(eng, images, colors, alpha)
    counts, boxes, class_ids, scores = eng.fetch_meta()
    out = []
    for b in range(n):
        k = int(counts[b])
        out.append((boxes[b, :k].copy(), class_ids[b, :k].copy(), scores[b, :k].copy(),
                    overlays[b].cpu().numpy()))
    return out


def unmold_detections(detections, mrcnn_mask, original_image_shape, image_shape, window):
    """Reformat one image's detections from the molded image back to the original image.

    detections: [R, (y1, x1, y2, x2, class_id, score)] normalised coordinates
    mrcnn_mask: [R, mh, mw, num_classes]
    original_image_shape: (H, W, 3) before resizing     image_shape: molded shape
    window: (y1, x1, y2, x2) pixel box of the real image inside the molded image

    Returns boxes [N,4] int32 pixels, class_ids [N] int32, scores [N], masks [H,W,N] bool.
    """
    return unmold_detections_batch(
        [(detections, mrcnn_mask, original_image_shape, image_shape, window)])[0]
