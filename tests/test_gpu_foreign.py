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
