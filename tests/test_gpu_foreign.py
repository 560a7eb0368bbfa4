"""The CUDA evaluator on foreign graphs and checkpoints (tests/foreign_graphs.py: protobuf-runtime GraphDefs in
TensorFlow-1.x idioms, BundleWriter-style index) against the float64 numpy oracle: tower widths 64 / 128 / 192, cond-style
batch-norm, FusedBatchNormV3, conv biases, Const scale/offset, resource variables, a 2-channel value head, a float
placeholder — what the real bootstrap-ckpt.meta may turn out to contain."""
import numpy as np
import pytest

from tests import foreign_graphs as FG
from tests import tfproto as P

pytestmark = pytest.mark.gpu

CASES = {
    "cond": dict(bn="cond"),
    "v3_conv_bias": dict(bn="v3", conv_bias=True),
    "resource_vars_float_input": dict(bn="v1", resource_vars=True, input_dtype=P.DT_FLOAT),
    "cond_v3_128_two_value_channels": dict(bn="cond_v3", filters=128, value_channels=2),
    "192_filters": dict(bn="v1", filters=192, blocks=1),
    "scale_false": dict(bn="cond", scale=False),
    "center_false_policy4": dict(bn="v1", center=False, conv_bias=True, policy_channels=4, value_hidden=128),
}


@pytest.mark.parametrize("case", list(CASES))
def test_foreign_checkpoint_parity(tmp_path, positions, case):
    from oracle.net_oracle_np import NumpyGraphOracle
    from unrealgo_b200 import PREC_FP16, PREC_FP32, DlTFNetworkEvaluator
    meta, prefix = FG.write(str(tmp_path / case), seed=3, **CASES[case])
    packed, planes = positions
    want_p, want_v = NumpyGraphOracle(meta, prefix).evaluate_reference(planes[:40])
    for prec, ptol, vtol in ((PREC_FP32, 1e-5, 1e-5), (PREC_FP16, 1e-3, 5e-3)):
        ev = DlTFNetworkEvaluator(meta, precision=prec, max_batch=64)
        assert ev.UpdateCheckPoint(prefix), ev._err()
        info = ev.info()
        assert info["filters"] == CASES[case].get("filters", 64)
        for n in (40, 3):   # large-batch and small-batch tower kernels
            p, v = ev.evaluate_packed(packed[:n])
            dp, dv = np.abs(p - want_p[:n]).max(), np.abs(v - want_v[:n]).max()
            assert dp <= ptol and dv <= vtol, (case, prec, n, dp, dv)
        ev.close()


def test_frozen_graph_evaluates_straight_after_load_graph(tmp_path, positions):
    """a graph run through freeze_graph (weights as Const nodes): the reference's session evaluates it after LoadPbGraph
    alone and its UpdateCheckPointForPbGraph merely fails (funcapproximator/DlTFNetworkEvaluator.cc:85-111, :203-226).
    Same here, and the answers are the bits the graph + checkpoint pair gives."""
    from oracle.net_oracle_np import NumpyGraphOracle
    from unrealgo_b200 import PREC_FP16, DlTFNetworkEvaluator
    g, meta = FG.build(bn="v3", conv_bias=True, filters=128)
    w = FG.random_weights(g.variables, 21)
    prefix = str(tmp_path / "model.ckpt-9")
    FG.write_bundle(prefix, w)
    mpath = tmp_path / "model.ckpt-9.meta"
    mpath.write_bytes(meta.SerializeToString())
    frozen = tmp_path / "frozen.pb"
    frozen.write_bytes(FG.freeze(meta, w).SerializeToString())
    packed, planes = positions
    want_p, want_v = NumpyGraphOracle(str(frozen), None).evaluate_reference(planes[:40])
    ev = DlTFNetworkEvaluator(str(frozen), precision=PREC_FP16, max_batch=64)
    assert ev.MetaGraphLoaded() and ev.info()["filters"] == 128          # weights are in: no UpdateCheckPoint needed
    ref = DlTFNetworkEvaluator(str(mpath), precision=PREC_FP16, max_batch=64)
    assert ref.UpdateCheckPoint(prefix), ref._err()
    for n in (40, 3):
        p, v = ev.evaluate_packed(packed[:n])
        assert np.abs(p - want_p[:n]).max() <= 1e-3 and np.abs(v - want_v[:n]).max() <= 5e-3
        p2, v2 = ref.evaluate_packed(packed[:n])
        assert np.array_equal(p, p2) and np.array_equal(v, v2)
    assert not ev.UpdateCheckPoint(prefix) and "frozen" in ev._err()     # nothing to restore; the evaluator stays usable
    p3, v3 = ev.evaluate_packed(packed[:3])
    assert np.array_equal(p3, p) and np.array_equal(v3, v)
    pol, val = np.zeros((3, 362)), np.zeros(3)
    assert ev.Evaluate(planes[:3], pol, val, 3) and np.array_equal(pol.astype(np.float32), p)
    ev.close()
    ref.close()
