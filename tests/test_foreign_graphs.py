"""The loader (csrc/tfmodel.cpp, through the host-only ug_inspect_model) against graphs and checkpoints it did not grow
up with: written by tests/foreign_graphs.py through Google's protobuf runtime and an SSTable writer that follows
TensorFlow's BundleWriter, in the idioms a TensorFlow-1.x training script leaves behind (SURVEY.md
8(c) "unknowns to resolve from the real .meta").  The two float oracles must read the same files and agree."""
import numpy as np
import pytest

from tests import foreign_graphs as FG
from tests import tfproto as P
from unrealgo_b200 import inspect_model

CASES = {
    "cond": dict(bn="cond"),                                   # tf.layers.batch_normalization(training=placeholder)
    "cond_plain_placeholder": dict(bn="cond", training_placeholder="plain"),
    "v3_conv_bias": dict(bn="v3", conv_bias=True),
    "resource_vars_float_input": dict(bn="v1", resource_vars=True, input_dtype=P.DT_FLOAT),
    "cond_v3_128_two_value_channels": dict(bn="cond_v3", filters=128, value_channels=2),
    "192_filters": dict(bn="v1", filters=192, blocks=1),
    "scale_false": dict(bn="cond", scale=False),
    "center_false": dict(bn="v1", center=False, conv_bias=True),
    "eps_1e-3_hidden_128": dict(bn="v1", bn_eps=1e-3, value_hidden=128, policy_channels=4),
    "many_index_blocks": dict(bn="v1", small_blocks=True),
}


@pytest.fixture(scope="module")
def written(tmp_path_factory):
    d = tmp_path_factory.mktemp("foreign")
    return {k: FG.write(str(d / k), seed=11 + i, **kw) for i, (k, kw) in enumerate(CASES.items())}


@pytest.mark.parametrize("case", list(CASES))
def test_loader_reads_foreign_graph_and_bundle(written, case):
    from oracle import tf_reader
    meta, prefix = written[case]
    kw = CASES[case]
    info = inspect_model(meta, prefix)
    filters = kw.get("filters", 64)
    assert info["blocks"] == kw.get("blocks", 2) and info["checkpoint"]["filters"] == filters
    assert info["input"] == "feature_input" and info["input_dtype"] == kw.get("input_dtype", P.DT_BOOL)
    assert info["policy_softmax"] and info["value_tanh"]
    assert abs(info["stem"]["eps"] - kw.get("bn_eps", 1e-5)) < 1e-9
    assert bool(info["stem"]["bias"]) == bool(kw.get("conv_bias", False))
    assert (info["stem"]["gamma"] == "") == (not kw.get("scale", True)) and \
           info["stem"]["gamma_const"] == (0 if kw.get("scale", True) else filters)
    assert (info["stem"]["beta"] == "") == (not kw.get("center", True))
    # execution order of the tower: block k conv a, conv b
    assert [c["kernel"] for c in info["tower"]] == [f"resnet/res_block_{k}/{ab}/conv2d/kernel"
                                                    for k in range(kw.get("blocks", 2)) for ab in "ab"]
    tensors = tf_reader.read_bundle(prefix)      # the oracle's reader on the same files
    used = {c[k] for c in [info["stem"], info["policy_conv"], info["value_conv"]] + info["tower"]
            for k in ("kernel", "bias", "gamma", "beta", "mean", "var") if c[k]}
    used |= {info[d][k] for d in ("policy_fc", "value_fc1", "value_fc2") for k in ("kernel", "bias")}
    assert used <= set(tensors) and info["checkpoint"]["variables_used"] == len(used)
    assert "beta1_power" in tensors and "beta1_power" not in used          # optimizer slots are ignored
    assert info["checkpoint"]["entries"] == len(tensors) + 1                # + global_step (int64, not a float entry)


@pytest.mark.parametrize("case", ["cond", "cond_v3_128_two_value_channels", "scale_false", "center_false",
                                  "resource_vars_float_input"])
def test_the_two_float_oracles_agree_on_foreign_graphs(written, case, positions):
    from oracle.net_oracle import GraphOracle
    from oracle.net_oracle_np import NumpyGraphOracle
    meta, prefix = written[case]
    _, planes = positions
    a = GraphOracle(meta, prefix).evaluate_reference(planes[:5])
    b = NumpyGraphOracle(meta, prefix).evaluate_reference(planes[:5])
    assert np.abs(a[0] - b[0]).max() <= 2e-6 and np.abs(a[1] - b[1]).max() <= 5e-6
    assert np.abs(b[0].sum(1) - 1).max() < 1e-12


def test_training_branch_is_never_taken(written, positions):
    """feeding training=True routes the data through the is_training=true FusedBatchNorm: the oracle refuses (it has
    no batch statistics to offer) — evidence that the Switch/Merge semantics are evaluated, not pattern-matched"""
    from oracle.net_oracle_np import NumpyGraphOracle
    meta, prefix = written["cond_plain_placeholder"]
    orc = NumpyGraphOracle(meta, prefix)
    with pytest.raises(KeyError, match="training"):
        orc.evaluate_reference(positions[1][:1])                      # a plain placeholder must be fed
    with pytest.raises(NotImplementedError, match="training-mode"):
        orc.evaluate_reference(positions[1][:1], feeds={"training": True})
    p, v = orc.evaluate_reference(positions[1][:1], feeds={"training": False})
    assert abs(p.sum() - 1) < 1e-12 and abs(v[0]) <= 1


@pytest.mark.parametrize("kw,msg", [
    (dict(bn="v1", data_format="NCHW"), "NCHW"),                       # in-graph NCHW: refuse, do not mis-evaluate
    (dict(bn="cond", data_format="NCHW"), "NCHW"),
])
def test_unsupported_foreign_graphs_fail_loudly(tmp_path, kw, msg):
    meta, prefix = FG.write(str(tmp_path / "bad"), **kw)
    with pytest.raises(RuntimeError, match=msg):
        inspect_model(meta, prefix)


def test_training_only_graph_is_refused(tmp_path):
    """a graph whose batch-norm exists only in training mode (no inference branch) is not an evaluator"""
    g, meta = FG.build(bn="v1")
    for n in meta.graph_def.node:
        if n.op == "FusedBatchNorm":
            n.attr["is_training"].b = True
    path = tmp_path / "train_only.meta"
    path.write_bytes(meta.SerializeToString())
    with pytest.raises(RuntimeError, match="training mode"):
        inspect_model(str(path))
