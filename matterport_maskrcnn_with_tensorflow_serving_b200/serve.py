"""Request/response surface of the reference's serve.py, kept signature for signature:

    preprocess_input(img, img_size=640) -> (molded_image, image_meta, anchors, window)   serve.py:83-107
    grpc_inference(img) -> (mrcnn_detection, mrcnn_mask, molded_image, window)            serve.py:110-138
    do_inference(img) -> (final_rois, final_class_ids, final_scores, final_masks)         serve.py:141-173

The TensorFlow-Serving RPC itself (serve.py:26-80) is out of scope and is injected:
`set_predict_fn(fn)` installs `fn(molded_image_f32, image_meta_f32, anchors_f32) ->
(mrcnn_detection, mrcnn_mask)` -- float lists / arrays as `float_val` gives them, or the
`TensorProto` messages themselves (or their bytes), which are then decoded from the wire format
without a Python float per element (`wire.py`); without one `grpc_inference` raises.  Rendering the PNG
(`visualize.display_instances`, serve.py:160-169) is also out of scope, so `do_inference`
returns the unmolded tuple instead of a file path.
"""
from __future__ import annotations

import numpy as np

from . import api_utils
from . import configs as cf
from . import wire
from .engine import Molder

_predict_fn = None
_molder = None


def set_predict_fn(fn):
    global _predict_fn
    _predict_fn = fn


def compose_image_meta(image_id, original_image_shape, image_shape, window, scale,
                       active_class_ids):
    """serve.py:100-103 (upstream model.compose_image_meta): host-side list packing."""
    return np.array(
        [image_id] + list(original_image_shape) + list(image_shape) + list(window) +
        [scale] + list(active_class_ids))


def preprocess_input(img, img_size=640, molded_dtype=np.float64):
    """serve.py:83-107.  molded_dtype float64 reproduces the reference's return value
    (uint8 -> float32 -> minus float64 MEAN_PIXEL); float32 returns exactly what
    serve.py:117 sends on the wire (`molded_image.astype(np.float32)`)."""
    import torch

    global _molder
    mcf = api_utils.get_config()
    if isinstance(img, str):
        img = api_utils.load_img(img)
    img = np.asarray(img)
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise TypeError("preprocess_input expects an HxWx3 uint8 image")
    if _molder is None or _molder.config is not mcf:
        _molder = Molder(mcf)
    d_img = torch.from_numpy(np.ascontiguousarray(img)).to(_molder.device)
    if img_size is not None:
        d_img = _molder.cv2_resize_device(d_img, (img_size, img_size))
    img_shape = tuple(d_img.shape)
    d_molded, _, window, scale, _padding = _molder.mold_device(d_img, out_dtype=molded_dtype)
    molded_image = d_molded.cpu().numpy()
    image_meta = compose_image_meta(
        0, img_shape, molded_image.shape, window, scale,
        np.zeros([mcf.NUM_CLASSES], dtype=np.int32))
    anchors = api_utils.get_anchors(molded_image.shape)
    return molded_image, image_meta, anchors, window


def _is_tensor_proto(x):
    return isinstance(x, (bytes, bytearray, memoryview)) or hasattr(x, "SerializeToString")


def grpc_inference(img):
    """serve.py:110-138 with the RPC injected (see module docstring)."""
    if _predict_fn is None:
        raise RuntimeError("no TensorFlow-Serving client installed: call set_predict_fn()")
    molded_image, image_meta, anchors, window = preprocess_input(img, cf.IMAGE_SIZE)
    mrcnn_detection, mrcnn_mask = _predict_fn(
        molded_image.astype(np.float32), image_meta.astype(np.float32),
        anchors.astype(np.float32))
    if _is_tensor_proto(mrcnn_detection) and _is_tensor_proto(mrcnn_mask):
        # result.outputs[...] handed over as TensorProto messages (or their bytes): take the
        # values from the wire instead of through per-element Python floats (wire.py)
        mrcnn_detection, mrcnn_mask = wire.decode_predict_outputs(
            mrcnn_detection, mrcnn_mask, cf.OUT_DETECTION_SHAPE, cf.OUT_MASK_SHAPE)
        return mrcnn_detection, mrcnn_mask, molded_image, window
    # serve.py:131-136: float_val lists become float64 arrays with a leading -1 dim
    mrcnn_detection = np.array(mrcnn_detection).reshape((-1, *cf.OUT_DETECTION_SHAPE))
    mrcnn_mask = np.array(mrcnn_mask).reshape((-1, *cf.OUT_MASK_SHAPE))
    return mrcnn_detection, mrcnn_mask, molded_image, window


def do_inference(img):
    """serve.py:141-154; the visualisation tail (:156-173) is not reproduced."""
    mrcnn_detection, mrcnn_mask, molded_image, window = grpc_inference(img)
    return api_utils.unmold_detections(
        mrcnn_detection, mrcnn_mask, img.shape, molded_image.shape, window)
