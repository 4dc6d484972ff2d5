"""Response decode / request encode without the Python-list detour (SURVEY.md 8f rank 1;
serve.py:49-76, :131-136).  A message type with TensorProto's public field numbers is built
with protobuf at test time so that the bytes under test come from a real protobuf encoder."""
import time

import numpy as np
import pytest

from matterport_maskrcnn_with_tensorflow_serving_b200 import wire

pb = pytest.importorskip("google.protobuf")


def _tensor_class():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fdp = descriptor_pb2.FileDescriptorProto()
    fdp.name = "mini_tensor_%d.proto" % time.perf_counter_ns()
    fdp.package = "mini"
    fdp.syntax = "proto3"
    dim = fdp.message_type.add()
    dim.name = "Dim"
    f = dim.field.add()
    f.name, f.number, f.type, f.label = "size", 1, f.TYPE_INT64, f.LABEL_OPTIONAL
    shape = fdp.message_type.add()
    shape.name = "TensorShapeProto"
    f = shape.field.add()
    f.name, f.number, f.type, f.label, f.type_name = "dim", 2, f.TYPE_MESSAGE, f.LABEL_REPEATED, ".mini.Dim"
    t = fdp.message_type.add()
    t.name = "TensorProto"
    for name, num, typ, rep, tn in [("dtype", 1, "TYPE_INT32", False, None),
                                    ("tensor_shape", 2, "TYPE_MESSAGE", False, ".mini.TensorShapeProto"),
                                    ("tensor_content", 4, "TYPE_BYTES", False, None),
                                    ("float_val", 5, "TYPE_FLOAT", True, None),
                                    ("double_val", 6, "TYPE_DOUBLE", True, None)]:
        f = t.field.add()
        f.name, f.number, f.type = name, num, getattr(f, typ)
        f.label = f.LABEL_REPEATED if rep else f.LABEL_OPTIONAL
        if tn:
            f.type_name = tn
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("mini.TensorProto"))


@pytest.fixture(scope="module")
def T():
    return _tensor_class()


def test_float_val_decode_equals_the_reference_conversion(T):
    rng = np.random.default_rng(0)
    a = rng.random((1, 100, 6)).astype(np.float32)
    m = T()
    m.float_val.extend(a.reshape(-1).tolist())
    ref = np.array(m.float_val).reshape((-1, 100, 6))          # serve.py:131-133
    got = wire.tensor_proto_to_ndarray(m, dtype=np.float64).reshape((-1, 100, 6))
    assert ref.dtype == got.dtype == np.float64 and np.array_equal(ref, got)
    assert np.array_equal(wire.tensor_proto_to_ndarray(m.SerializeToString()), a.reshape(-1))


def test_tensor_content_shape_and_double(T):
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4) / 7
    m = T()
    m.dtype = wire.DT_FLOAT
    for s in a.shape:
        m.tensor_shape.dim.add().size = s
    m.tensor_content = a.tobytes()
    got = wire.tensor_proto_to_ndarray(m)
    assert got.shape == a.shape and got.dtype == np.float32 and np.array_equal(got, a)
    d = T()
    d.double_val.extend([0.1, 0.25, 1e-300])
    assert wire.tensor_proto_to_ndarray(d).tolist() == [0.1, 0.25, 1e-300]
    assert wire.tensor_proto_to_ndarray(T()).size == 0


def test_unpacked_float_val():
    """proto2-style unpacked repeated floats (one fixed32 per element) keep wire order."""
    vals = np.array([1.5, -2.25, 3.0], dtype="<f4")
    buf = b"".join(bytes([(5 << 3) | 5]) + v.tobytes() for v in vals)
    assert np.array_equal(wire.tensor_proto_to_ndarray(buf), vals)


def test_decode_predict_outputs_dtypes(T):
    rng = np.random.default_rng(1)
    det = rng.random((100, 6)).astype(np.float32)
    msk = rng.random((100, 28, 28, 3)).astype(np.float32)
    md, mm = T(), T()
    md.float_val.extend(det.reshape(-1).tolist())
    mm.float_val.extend(msk.reshape(-1).tolist())
    d, k = wire.decode_predict_outputs(md, mm, (100, 6), (100, 28, 28, 3))
    assert d.dtype == np.float64 and d.shape == (1, 100, 6)
    assert k.dtype == np.float32 and k.shape == (1, 100, 28, 28, 3)
    assert np.array_equal(d[0], det.astype(np.float64)) and np.array_equal(k[0], msk)


def test_request_encode_round_trip(T):
    a = np.random.default_rng(2).random((1, 64, 64, 3)).astype(np.float32)
    b = wire.ndarray_to_tensor_proto_bytes(a)
    m = T.FromString(b)                                         # a real protobuf parser reads it
    assert m.dtype == wire.DT_FLOAT and [d.size for d in m.tensor_shape.dim] == [1, 64, 64, 3]
    assert np.array_equal(np.frombuffer(m.tensor_content, np.float32).reshape(a.shape), a)
    assert np.array_equal(wire.tensor_proto_to_ndarray(b), a)
    with pytest.raises(ValueError):
        wire.ndarray_to_tensor_proto_bytes(np.zeros(3, np.int32))


def test_decode_is_much_faster_than_the_list_detour(T):
    a = np.random.default_rng(3).random(100 * 28 * 28 * 8, dtype=np.float32)
    m = T()
    m.float_val.extend(a.tolist())
    t0 = time.perf_counter()
    ref = np.array(m.float_val)
    t_ref = time.perf_counter() - t0
    t0 = time.perf_counter()
    got = wire.tensor_proto_to_ndarray(m)
    t_new = time.perf_counter() - t0
    assert np.array_equal(ref.astype(np.float32), got)
    assert t_new * 5 < t_ref, (t_new, t_ref)
