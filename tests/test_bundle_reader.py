"""The product's checkpoint-bundle reader (csrc/tfmodel.cpp TfBundle: the loader behind UpdateCheckPointForMetaGraph,
funcapproximator/DlTFNetworkEvaluator.cc:170-200) value for value, on the CPU: every float variable it returns must be
exactly what the writers stored — the foreign writer of tests/foreign_graphs.py (single 256 KB index block, many small
blocks, merged multi-shard bundles as Saver(sharded=True) leaves them) and the product's own — and it must agree with
the oracle's independent Python reader.  Damaged bundles are refused by name."""
import os
import shutil

import numpy as np
import pytest

from tests import foreign_graphs as FG
from unrealgo_b200 import bundle_read_float, inspect_model, tfformat


def _weights(seed=5, **kw):
    g, _ = FG.build(**kw)
    return FG.random_weights(g.variables, seed)


@pytest.mark.parametrize("shards,block", [(1, 262144), (1, 700), (2, 262144), (3, 900), (5, 262144)])
def test_reader_returns_what_the_foreign_writer_stored(tmp_path, shards, block):
    w = _weights(blocks=2, filters=64, bn="v1")
    prefix = str(tmp_path / "model.ckpt-7")
    FG.write_bundle(prefix, w, block_size=block, num_shards=shards)
    assert sorted(f for f in os.listdir(tmp_path) if ".data-" in f) == \
        ["model.ckpt-7.data-%05d-of-%05d" % (k, shards) for k in range(shards)]
    for key, want in w.items():
        got = bundle_read_float(prefix, key)
        assert got.shape == want.shape and got.dtype == np.float32 and np.array_equal(got, want), key
    scalar = bundle_read_float(prefix, "beta1_power")                                   # a scalar: shape ()
    assert scalar.shape == () and scalar == np.float32(0.9)
    with pytest.raises(RuntimeError, match="not DT_FLOAT"):
        bundle_read_float(prefix, "global_step")
    with pytest.raises(RuntimeError, match="not in checkpoint"):
        bundle_read_float(prefix, "resnet/no_such/kernel")


def test_sharded_bundle_loads_like_the_single_file_one(tmp_path, positions):
    """the whole loader (graph walk + every variable the plan needs) on a three-shard bundle"""
    one = FG.write(str(tmp_path / "one"), seed=3, bn="cond", filters=64)
    three = FG.write(str(tmp_path / "three"), seed=3, bn="cond", filters=64, num_shards=3)
    a, b = inspect_model(*one), inspect_model(*three)
    assert a["checkpoint"] == b["checkpoint"] and a["tower"] == b["tower"]
    for key in (a["stem"]["kernel"], a["tower"][-1]["mean"], a["value_fc2"]["bias"]):
        assert np.array_equal(bundle_read_float(one[1], key), bundle_read_float(three[1], key))
    # ... and so do the oracle's reader and both float oracles: same policy, same value
    from oracle import tf_reader
    from oracle.net_oracle import GraphOracle
    from oracle.net_oracle_np import NumpyGraphOracle
    r1, r3 = tf_reader.read_bundle(one[1]), tf_reader.read_bundle(three[1])
    assert sorted(r1) == sorted(r3) and all(np.array_equal(r1[k], r3[k]) for k in r1)
    _, planes = positions
    p1 = GraphOracle(*one).evaluate_reference(planes[:3])
    p3 = GraphOracle(*three).evaluate_reference(planes[:3])
    q3 = NumpyGraphOracle(*three).evaluate_reference(planes[:3])
    assert np.array_equal(p1[0], p3[0]) and np.array_equal(p1[1], p3[1])
    assert np.abs(p3[0] - q3[0]).max() <= 2e-6 and np.abs(p3[1] - q3[1]).max() <= 5e-6
    os.remove(three[1] + ".data-00002-of-00003")
    with pytest.raises(RuntimeError, match="data-00002-of-00003"):
        inspect_model(*three)


def test_reader_agrees_with_the_oracles_reader_and_the_product_writer(tmp_path):
    from oracle import tf_reader
    meta, prefix = tfformat.write_synthetic_checkpoint(str(tmp_path / "w"), 2, 64, seed=9)
    ref = tf_reader.read_bundle(prefix)
    assert len(ref) > 20
    for key, want in ref.items():
        if want.dtype == np.float32:
            assert np.array_equal(bundle_read_float(prefix, key), want), key
    fmeta, fprefix = FG.write(str(tmp_path / "f"), seed=4, bn="v3", conv_bias=True, small_blocks=True)
    for key, want in tf_reader.read_bundle(fprefix).items():
        if want.dtype == np.float32:
            assert np.array_equal(bundle_read_float(fprefix, key), want), key


def test_damaged_bundles_are_refused_by_name(tmp_path):
    w = _weights(blocks=1, filters=64, bn="v1")
    prefix = str(tmp_path / "ck")
    FG.write_bundle(prefix, w)
    key = sorted(k for k in w if k.endswith("/kernel"))[0]
    data = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data, "rb").read())
    # a flipped bit inside that tensor: its own crc32c catches it, other tensors still read
    # locate the tensor through the index the product reader itself walks: read once, find its bytes in the data file
    good = bundle_read_float(prefix, key)
    off = bytes(raw).find(good.tobytes())
    assert off >= 0
    raw[off + 5] ^= 0x10
    open(data, "wb").write(bytes(raw))
    with pytest.raises(RuntimeError, match="crc32c"):
        bundle_read_float(prefix, key)
    other = sorted(k for k in w if k.endswith("/kernel"))[-1]
    assert np.array_equal(bundle_read_float(prefix, other), w[other])
    # truncated data file
    open(data, "wb").write(bytes(raw[:off + 16]))
    with pytest.raises(RuntimeError, match="inconsistent size/offset"):
        bundle_read_float(prefix, key)
    # damaged index block: checksum
    shutil.copy(prefix + ".index", prefix + ".index.bak")
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[20] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(RuntimeError, match="checksum|malformed"):
        bundle_read_float(prefix, other)
    # not an SSTable at all
    open(prefix + ".index", "wb").write(b"x" * 100)
    with pytest.raises(RuntimeError, match="magic"):
        bundle_read_float(prefix, other)
    # a bundle that says it is big-endian is refused rather than misread
    FG.write_bundle(str(tmp_path / "be"), w, endianness=1)
    with pytest.raises(RuntimeError, match="big-endian"):
        bundle_read_float(str(tmp_path / "be"), other)
