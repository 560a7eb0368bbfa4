"""Self-play record writer: the TFRecord framing and the tf.train.Example payload written by
libugeval are read back by an independent reader — framing checked with the oracle's own CRC-32C (pure C, not
the product's), the payload parsed by the protobuf runtime from an Example schema built here field by field
(TensorFlow itself is not available; field numbers are tensorflow/core/example/{example,feature}.proto's)."""
import struct

import numpy as np
import pytest
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory


def _example_class():
    fd = descriptor_pb2.FileDescriptorProto(name="ug_example.proto", package="ugt", syntax="proto3")

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    F = descriptor_pb2.FieldDescriptorProto
    m = msg("BytesList"); m.field.add(name="value", number=1, type=F.TYPE_BYTES, label=F.LABEL_REPEATED)
    m = msg("FloatList"); m.field.add(name="value", number=1, type=F.TYPE_FLOAT, label=F.LABEL_REPEATED)
    m = msg("Int64List"); m.field.add(name="value", number=1, type=F.TYPE_INT64, label=F.LABEL_REPEATED)
    m = msg("Feature")
    m.oneof_decl.add(name="kind")
    m.field.add(name="bytes_list", number=1, type=F.TYPE_MESSAGE, type_name=".ugt.BytesList", label=F.LABEL_OPTIONAL, oneof_index=0)
    m.field.add(name="float_list", number=2, type=F.TYPE_MESSAGE, type_name=".ugt.FloatList", label=F.LABEL_OPTIONAL, oneof_index=0)
    m.field.add(name="int64_list", number=3, type=F.TYPE_MESSAGE, type_name=".ugt.Int64List", label=F.LABEL_OPTIONAL, oneof_index=0)
    m = msg("Features")
    e = m.nested_type.add(name="FeatureEntry")
    e.options.map_entry = True
    e.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    e.field.add(name="value", number=2, type=F.TYPE_MESSAGE, type_name=".ugt.Feature", label=F.LABEL_OPTIONAL)
    m.field.add(name="feature", number=1, type=F.TYPE_MESSAGE, type_name=".ugt.Features.FeatureEntry", label=F.LABEL_REPEATED)
    m = msg("Example")
    m.field.add(name="features", number=1, type=F.TYPE_MESSAGE, type_name=".ugt.Features", label=F.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("ugt.Example"))


def _mask(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def read_records(path):
    """TensorFlow RecordReader framing, restated: u64 len | u32 masked crc(len) | data | u32 masked crc(data)"""
    import oracle
    crc = lambda b: oracle.lib().og_crc32c(b, len(b))  # noqa: E731
    buf = open(path, "rb").read()
    off, out = 0, []
    while off < len(buf):
        (n,) = struct.unpack_from("<Q", buf, off)
        (lc,) = struct.unpack_from("<I", buf, off + 8)
        assert lc == _mask(crc(buf[off:off + 8])), "length crc"
        data = buf[off + 12:off + 12 + n]
        (dc,) = struct.unpack_from("<I", buf, off + 12 + n)
        assert dc == _mask(crc(data)), "data crc"
        out.append(data)
        off += 16 + n
    assert off == len(buf)
    return out


def test_crc32c_known_answers():
    from unrealgo_b200 import _lib
    lib = _lib.load()
    # RFC 3720 B.4 check values
    assert lib.ug_crc32c(b"123456789", 9) == 0xE3069283
    assert lib.ug_crc32c(bytes(32), 32) == 0x8A9136AA
    assert lib.ug_crc32c(bytes([0xFF] * 32), 32) == 0x62A8AB43
    assert lib.ug_crc32c(bytes(range(32)), 32) == 0x46DD794E


def test_examples_round_trip(tmp_path):
    from unrealgo_b200 import DlTFRecordWriter
    Example = _example_class()
    rng = np.random.default_rng(3)
    path = str(tmp_path / "selfplay.tfrecord")
    w = DlTFRecordWriter(path)
    want = []
    for i in range(25):
        feature = rng.integers(0, 2, size=(17, 19, 19), dtype=np.int8)  # char[NUM_MAPS][19][19], 6137 bytes
        policy = rng.random(362).astype(np.float32)
        policy /= policy.sum()
        reward = 1.0 if i % 2 else -1.0
        w.WriteExample(feature, policy, reward)
        want.append((feature.tobytes(), policy, reward))
    w.Flush()
    assert w.close() == 25
    recs = read_records(path)
    assert len(recs) == 25
    for data, (fb, pol, rew) in zip(recs, want):
        ex = Example()
        ex.ParseFromString(data)
        feats = ex.features.feature
        assert sorted(feats.keys()) == ["feature", "policy", "reward"]  # DlTFRecordWriter.cc:63-69
        assert feats["feature"].WhichOneof("kind") == "bytes_list" and list(feats["feature"].bytes_list.value) == [fb]
        assert len(fb) == 6137
        assert np.array_equal(np.array(feats["policy"].float_list.value, np.float32), pol)
        assert list(feats["reward"].float_list.value) == [rew]
        # byte-exact against the protobuf runtime's own deterministic serialisation of the same message
        assert ex.SerializeToString(deterministic=True) == data


def test_named_keys_and_empty_file(tmp_path):
    from unrealgo_b200 import DlTFRecordWriter
    Example = _example_class()
    p0 = str(tmp_path / "empty.tfrecord")
    assert DlTFRecordWriter(p0).close() == 0
    assert read_records(p0) == []
    p1 = str(tmp_path / "named.tfrecord")
    w = DlTFRecordWriter(p1)
    w.WriteExample(np.arange(10, dtype=np.uint8), np.array([0.25, 0.75], np.float32), 0.5, names=("x", "pi", "z"))
    assert w.close() == 1
    ex = Example()
    ex.ParseFromString(read_records(p1)[0])
    assert sorted(ex.features.feature.keys()) == ["pi", "x", "z"]
    assert list(ex.features.feature["z"].float_list.value) == [0.5]
    with pytest.raises(OSError):
        DlTFRecordWriter(str(tmp_path / "no_such_dir" / "f.tfrecord"))
