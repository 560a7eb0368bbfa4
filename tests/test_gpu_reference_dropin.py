"""The reference's own search/UctBoardEvaluator.cpp + DlConfig.cc, compiled unmodified in the build container against
the ugeval include overlay and linked with libugeval.so (oracle/Makefile `dropin` -> oracle/_ref/ref_dropin_eval),
run on the GPU: the answers must be the oracle's within tolerance and bit-identical to the Python mirror's."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_dropin_eval")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/ref_dropin_eval not prebuilt (needs the reference tree)")
def test_reference_uct_board_evaluator_runs_on_ugeval(small_ckpt, positions, tmp_path):
    from oracle.net_oracle import GraphOracle
    from unrealgo_b200 import DlTFNetworkEvaluator
    meta, prefix = small_ckpt
    _, planes = positions
    n = 16
    outs = "resnet/tower_0/policy_head/policy_predict:resnet/tower_0/value_head/reward_predict"
    (tmp_path / "dlcfg-19").write_text(f"# test config\nmeta_graph={meta}\ncheckpoint={prefix}\nnn_input=feature_input\n"
                                       f"nn_outputs={outs}\nvalue_transform=1\nreuse_searchtree=true\n")
    planes[:n].tofile(str(tmp_path / "planes.bin"))
    r = subprocess.run([EXE, "planes.bin", str(n)], cwd=str(tmp_path), capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stderr
    flags, pol, val = {}, np.zeros((n, 362)), np.zeros(n)
    for line in r.stdout.splitlines():
        t = line.split()
        if t[0] == "value":
            val[int(t[1])] = float(t[2])
        elif t[0] == "policy":
            pol[int(t[1])] = [float(x) for x in t[2:]]
        else:
            flags[t[0]] = int(t[1])
    assert flags == {"loaded_before": 0, "bad_path_throws": 1, "loaded_after": 1,
                     # the class driven directly, as funcapproximator/test/DlTFTestVariableBatchSize.cc:80-90 does:
                     "direct_loaded": 1, "direct_checkpoint": 1, "create_tensor": 1, "transform_chw_is_hwc": 1,
                     "transform_hwc_is_copy": 1, "transform_df_chw": 1, "direct_identical": 1}
    want_p, want_v = GraphOracle(meta, prefix).evaluate_reference(planes[:n], value_transform=1)
    assert np.abs(pol - want_p).max() <= 1e-3 and np.abs(val - want_v).max() <= 5e-3
    ev = DlTFNetworkEvaluator("", feature_input="feature_input", outputs=outs.split(":"), max_batch=512)
    got_p, got_v = np.zeros((n, 362)), np.zeros(n)
    try:
        ev.LoadGraph(meta)
        ev.UpdateCheckPoint(prefix)
        ev.set_value_transform(1)
        assert ev.Evaluate(planes[:n], got_p, got_v)
    finally:
        ev.close()
    assert np.array_equal(pol, got_p) and np.array_equal(val, got_v)
