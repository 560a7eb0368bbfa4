"""A graph run through freeze_graph — every variable folded into a float Const — is what people hand to a C++ session:
the reference evaluates such a .pb straight after LoadPbGraph (funcapproximator/DlTFNetworkEvaluator.cc:85-111) and its
UpdateCheckPointForPbGraph (:203-226) merely fails on it.  CPU side: the loader takes the weights from the graph, both
float oracles evaluate the frozen graph exactly as they evaluate graph + checkpoint."""
import numpy as np
import pytest

from tests import foreign_graphs as FG
from unrealgo_b200 import inspect_model


def _pair(tmp_path, **kw):
    g, meta = FG.build(**kw)
    w = FG.random_weights(g.variables, 21)
    prefix = str(tmp_path / "model.ckpt-9")
    FG.write_bundle(prefix, w)
    mpath = tmp_path / "model.ckpt-9.meta"
    mpath.write_bytes(meta.SerializeToString())
    frozen = tmp_path / "frozen.pb"
    frozen.write_bytes(FG.freeze(meta, w).SerializeToString())
    return str(mpath), prefix, str(frozen), w


@pytest.mark.parametrize("kw", [dict(bn="v1"), dict(bn="v3", conv_bias=True), dict(bn="v1", resource_vars=True, filters=128),
                                dict(bn="cond", scale=False)],
                         ids=["v1", "v3-bias", "resource-128", "cond-noscale"])
def test_loader_takes_the_weights_of_a_frozen_graph_from_its_constants(tmp_path, kw):
    mpath, prefix, frozen, w = _pair(tmp_path, **kw)
    a = inspect_model(mpath, prefix)
    b = inspect_model(frozen)                      # no checkpoint: nothing to restore
    assert not a["frozen"] and a["const_weights"] == 0
    assert b["frozen"] and b["const_weights"] == b["checkpoint"]["variables_used"]
    # batch-norm scale / offset constants travel in the plan itself (as for scale=False graphs), the rest is counted
    convs = lambda d: [d["stem"]] + d["tower"] + [d["policy_conv"], d["value_conv"]]   # noqa: E731
    folded = sum(c["gamma_const"] + c["beta_const"] for c in convs(b)) - sum(c["gamma_const"] + c["beta_const"] for c in convs(a))
    assert folded > 0 and b["checkpoint"]["parameters"] + folded == a["checkpoint"]["parameters"]
    assert b["checkpoint"]["filters"] == a["checkpoint"]["filters"] == kw.get("filters", 64)
    assert b["blocks"] == a["blocks"] and b["stem"]["eps"] == a["stem"]["eps"]
    assert b["stem"]["kernel"].startswith("@const:") and b["stem"]["kernel"].endswith(a["stem"]["kernel"])
    # a checkpoint prefix beside a frozen graph is ignored, present or not
    strip = lambda c: {k: v for k, v in c.items() if k != "entries"}   # noqa: E731  ("entries" counts the bundle's too)
    assert strip(inspect_model(frozen, prefix)["checkpoint"]) == strip(b["checkpoint"])
    assert inspect_model(frozen, str(tmp_path / "no-such-prefix"))["checkpoint"] == b["checkpoint"]


def test_partly_frozen_graph_mixes_constants_and_checkpoint(tmp_path):
    """only the batch-norm statistics folded into constants: the rest still comes from the bundle"""
    g, meta = FG.build(bn="v1")
    w = FG.random_weights(g.variables, 5)
    prefix = str(tmp_path / "ck")
    FG.write_bundle(prefix, w)
    stats = {k: v for k, v in w.items() if k.endswith("moving_mean") or k.endswith("moving_variance")}
    part = tmp_path / "part.pb"
    part.write_bytes(FG.freeze(meta, stats).SerializeToString())
    info = inspect_model(str(part), prefix)
    assert not info["frozen"] and info["const_weights"] == len(stats)
    with pytest.raises(RuntimeError, match="cannot read|not in checkpoint"):
        inspect_model(str(part), str(tmp_path / "missing"))


def test_both_float_oracles_evaluate_the_frozen_graph_like_graph_plus_checkpoint(tmp_path, positions):
    from oracle.net_oracle import GraphOracle
    from oracle.net_oracle_np import NumpyGraphOracle
    mpath, prefix, frozen, _ = _pair(tmp_path, bn="v1")
    _, planes = positions
    want = GraphOracle(mpath, prefix).evaluate_reference(planes[:4])
    for cls in (GraphOracle, NumpyGraphOracle):
        got = cls(frozen, None).evaluate_reference(planes[:4])
        assert np.abs(got[0] - want[0]).max() <= 2e-6 and np.abs(got[1] - want[1]).max() <= 5e-6, cls.__name__
