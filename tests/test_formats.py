"""On-disk formats either side of LoadGraph/UpdateCheckPoint, exercised without a GPU through the host-only
ug_inspect_model entry (same C++ loader the evaluator uses) and the oracle's independent Python readers."""
import os
import shutil

import numpy as np
import pytest

from unrealgo_b200 import inspect_model, tfformat


def test_loader_reads_writer_and_agrees_with_oracle_reader(small_ckpt):
    from oracle import tf_reader
    meta, prefix = small_ckpt
    d = inspect_model(meta, prefix)
    assert d["blocks"] == 2 and d["checkpoint"]["filters"] == 64
    assert d["input"] == "feature_input" and d["input_dtype"] == tfformat.DT_BOOL
    assert d["policy_softmax"] and d["value_tanh"]
    assert d["saver_filename_tensor"] == "save/Const:0" and d["saver_restore_op"] == "save/restore_all"
    nodes, saver = tf_reader.read_graph(meta)
    assert saver == {"filename_tensor_name": "save/Const:0", "restore_op_name": "save/restore_all"}
    tensors = tf_reader.read_bundle(prefix)
    assert d["checkpoint"]["entries"] == len(tensors)
    assert d["checkpoint"]["parameters"] == sum(t.size for t in tensors.values())
    assert d["stem"]["kernel"] in tensors and tensors[d["stem"]["kernel"]].shape == (3, 3, 17, 64)
    # the walk orders the tower by execution: block k conv1, conv2
    kernels = [c["kernel"] for c in d["tower"]]
    assert kernels == [f"resnet/unit_res_{k}/sub{s}/conv{s}/conv/kernel" for k in range(2) for s in (1, 2)]


def test_bundle_roundtrip_bitexact(tmp_path):
    from oracle import tf_reader
    rng = np.random.default_rng(4)
    tensors = {f"scope_{i:03d}/w": rng.standard_normal((3, 3, i % 5 + 1, 4)).astype(np.float32) for i in range(150)}
    tensors["z/scalar_like"] = np.float32([3.5])
    prefix = str(tmp_path / "ck")
    tfformat.write_bundle(prefix, tensors, block_size=512)  # many data blocks + prefix compression
    back = tf_reader.read_bundle(prefix)
    assert set(back) == set(tensors)
    for k in tensors:
        assert back[k].dtype == np.float32 and np.array_equal(back[k], tensors[k])


def test_corrupted_checkpoint_is_rejected(small_ckpt, tmp_path):
    meta, prefix = small_ckpt
    for suf in (".index", ".data-00000-of-00001"):
        shutil.copy(prefix + suf, str(tmp_path / ("bad" + suf)))
    bad = str(tmp_path / "bad")
    data = bytearray(open(bad + ".data-00000-of-00001", "rb").read())
    data[1000] ^= 0x40
    open(bad + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(RuntimeError, match="crc32c"):
        inspect_model(meta, bad)
    with pytest.raises(RuntimeError, match="cannot read"):
        inspect_model(meta, str(tmp_path / "missing"))


def test_graph_errors_are_loud(tmp_path, small_ckpt):
    with pytest.raises(RuntimeError, match="Error reading graph definition"):
        inspect_model(str(tmp_path / "nope.meta"))
    junk = tmp_path / "junk.meta"
    junk.write_bytes(b"\xff\xff\xff\xff not a protobuf")
    with pytest.raises(RuntimeError):
        inspect_model(str(junk))
    with pytest.raises(RuntimeError, match=".meta or .pb"):
        inspect_model(str(tmp_path / "graph.txt"))
    meta, _ = small_ckpt
    with pytest.raises(RuntimeError, match="not in graph"):
        inspect_model(meta, None, "feature_input", ["no/such/policy", "resnet/tower_0/value_head/reward_predict"])
    with pytest.raises(RuntimeError, match="nn_input"):
        inspect_model(meta, None, "input", None)


def test_unrecognised_op_pattern_names_the_node(tmp_path):
    """a tower whose activation is not Relu must be refused, not silently mis-evaluated"""
    _, gd2 = tfformat.build_graph(1, 64)
    broken = gd2.replace(tfformat.pb_bytes(2, "Relu"), tfformat.pb_bytes(2, "Selu"), 1)
    path = tmp_path / "broken.meta"
    tfformat.write_metagraph(str(path), broken)
    with pytest.raises(RuntimeError, match="unsupported graph: .*'resnet/tower_0/"):
        inspect_model(str(path))


def test_graphdef_pb_and_float_placeholder(tmp_path):
    g, gd = tfformat.build_graph(1, 64, input_dtype=tfformat.DT_FLOAT, conv_bias=True)
    pb = tmp_path / "frozen.pb"
    tfformat.write_graphdef(str(pb), gd)
    d = inspect_model(str(pb))
    assert d["input_dtype"] == tfformat.DT_FLOAT and d["stem"]["bias"] == "resnet/init_conv/conv/bias"
    assert d["blocks"] == 1


def test_40_block_graph_walk():
    import tempfile
    d = tempfile.mkdtemp()
    g, gd = tfformat.build_graph(40, 256)
    tfformat.write_metagraph(os.path.join(d, "big.meta"), gd)
    info = inspect_model(os.path.join(d, "big.meta"))
    assert info["blocks"] == 40 and len(info["tower"]) == 80


def test_checkpoint_info_file_round_trip(tmp_path):
    """DlCheckPoint's three-line info file (funcapproximator/DlCheckPoint.cc:162-203) and the bundle-presence check
    (:130-137) that gate a checkpoint hot swap."""
    import ctypes as C

    from unrealgo_b200 import _lib, tfformat
    lib = _lib.load()
    path = str(tmp_path / "bestcheckpoint.info").encode()
    bufs = [C.create_string_buffer(512) for _ in range(3)]
    assert lib.ug_ckptinfo_read(path, *bufs, 512) == 1                      # missing file: outputs untouched
    assert lib.ug_ckptinfo_write(path, b"ckpt-1200", b"da39a3ee5e6b4b0d3255bfef95601890afd80709",
                                 b"http://host/ckpt-1200.zip", 1) == 0
    assert open(path, "rb").read() == (b"ckpt-1200\nda39a3ee5e6b4b0d3255bfef95601890afd80709\n"
                                       b"http://host/ckpt-1200.zip")      # no newline after the url
    assert lib.ug_ckptinfo_read(path, *bufs, 512) == 0
    assert [b.value for b in bufs] == [b"ckpt-1200", b"da39a3ee5e6b4b0d3255bfef95601890afd80709",
                                       b"http://host/ckpt-1200.zip"]
    assert lib.ug_ckptinfo_write(path, b"other", b"", b"", 0) == 1          # exists, no override
    assert lib.ug_ckptinfo_read(path, *bufs, 512) == 0 and bufs[0].value == b"ckpt-1200"
    short = str(tmp_path / "short.info")
    open(short, "w").write("only\ntwo")
    assert lib.ug_ckptinfo_read(short.encode(), *bufs, 512) == 1           # fewer than three lines: ignored
    # bundle presence
    meta, prefix = tfformat.write_synthetic_checkpoint(str(tmp_path / "ck"), 1, 64, seed=1)
    assert lib.ug_checkpoint_files_exist(prefix.encode()) == 1
    assert lib.ug_checkpoint_files_exist((prefix + "-nope").encode()) == 0
    assert lib.ug_checkpoint_files_exist(b"") == 0
