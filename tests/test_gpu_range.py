"""fp16 range safety (VERDICT r1 "what's weak" 2): fp16 operands saturate at 65504 and the tower clamps rather than
producing inf, so a checkpoint outside that range would evaluate wrongly but plausibly.  ug_eval_load_checkpoint
calibrates every checkpoint (fp32 validation tower on built-in positions); by policy an fp16 evaluator rescales the hot
tensors by exact powers of two (default), falls back to bf16, refuses or ignores; the tower epilogues count clamped outputs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _hot_checkpoint(d, scale):
    """2x64 net whose stem batch-norm scale is multiplied by `scale`: stem outputs of the order of `scale`"""
    from unrealgo_b200 import tfformat
    os.makedirs(d, exist_ok=True)
    g, gd = tfformat.build_graph(2, 64)
    w = tfformat.random_weights(g.variables, 11)
    w["resnet/init_conv/bn/gamma"] = w["resnet/init_conv/bn/gamma"] * np.float32(scale)
    meta, prefix = os.path.join(d, "hot.meta"), os.path.join(d, "hot")
    tfformat.write_metagraph(meta, gd)
    tfformat.write_bundle(prefix, w)
    return meta, prefix


def test_calibration_of_an_ordinary_checkpoint(small_ckpt, positions):
    from unrealgo_b200 import PREC_FP16, DlTFNetworkEvaluator
    ev = DlTFNetworkEvaluator(small_ckpt[0], max_batch=64)
    assert ev.UpdateCheckPoint(small_ckpt[1])
    info = ev.info()
    assert info["precision"] == PREC_FP16 and info["fp16_headroom"] > 8
    assert 0.1 < info["calib_max_activation"] < 1000 and 0 < info["calib_max_weight"] < 100
    ev.evaluate_packed(positions[0][:40])
    ev.evaluate_packed(positions[0][:3])
    assert ev.info()["saturation_events"] == 0
    ev.close()


def test_checkpoint_outside_fp16_range(tmp_path, small_ckpt, positions, capfd):
    from oracle.net_oracle_np import NumpyGraphOracle
    from unrealgo_b200 import PREC_BF16, PREC_FP16, DlTFNetworkEvaluator
    packed, planes = positions
    meta, prefix = _hot_checkpoint(str(tmp_path / "hot"), 3e4)
    want_p, want_v = NumpyGraphOracle(meta, prefix).evaluate_reference(planes[:40])
    # default policy (auto): the hot tensors are stored times 2^-k — exact rescaling folded into weights and biases — so
    # the evaluator stays on fp16 operands AND inside the fp16 tolerance
    ev = DlTFNetworkEvaluator(meta, max_batch=64)
    assert ev.UpdateCheckPoint(prefix)
    info = ev.info()
    assert info["precision"] == PREC_FP16 and info["calib_max_activation"] > 65504 / 8
    assert info["fp16_stream_shift"] > 0 and info["fp16_headroom"] >= 16
    assert "rescaled by powers of two" in capfd.readouterr().err
    for n in (40, 3):                                            # both tower kernels
        p, v = ev.evaluate_packed(packed[:n])
        dp, dv = np.abs(p - want_p[:n]).max(), np.abs(v - want_v[:n]).max()
        assert dp <= 1e-3 and dv <= 5e-3, (n, dp, dv)
    pol, val = np.zeros((5, 362)), np.zeros(5)
    assert ev.Evaluate(planes[:5], pol, val, 5)                  # reference-facing entry, same numbers
    assert np.array_equal(pol.astype(np.float32), ev.evaluate_packed(packed[:5])[0])
    assert ev.info()["saturation_events"] == 0
    # an ordinary checkpoint afterwards is not rescaled
    assert ev.UpdateCheckPoint(small_ckpt[1]) and ev.info()["fp16_stream_shift"] == 0
    # policy "bf16": no rescaling, bf16 operands (coarser, but the right evaluator), and back to fp16 afterwards
    ev.set_fp16_range_policy("bf16")
    assert ev.UpdateCheckPoint(prefix) and ev.info()["precision"] == PREC_BF16
    assert "bf16 operands instead" in capfd.readouterr().err
    p, v = ev.evaluate_packed(packed[:20])
    assert np.isfinite(p).all() and np.abs(p.sum(1) - 1).max() < 1e-5
    assert (p.argmax(1) == want_p[:20].argmax(1)).mean() >= 0.8
    assert ev.UpdateCheckPoint(small_ckpt[1]) and ev.info()["precision"] == PREC_FP16
    # refuse
    ev.set_fp16_range_policy("refuse")
    assert not ev.UpdateCheckPoint(prefix) and "headroom" in ev._err() and "refused" in ev._err()
    assert ev.info()["precision"] == PREC_FP16                   # the previous network is still loaded and usable
    p0, _ = ev.evaluate_packed(packed[:4])
    assert np.isfinite(p0).all()
    # ignore: fp16 as packed — the epilogues then report what they had to clamp
    ev.set_fp16_range_policy("ignore")
    assert ev.UpdateCheckPoint(prefix) and ev.info()["precision"] == PREC_FP16 and ev.info()["fp16_stream_shift"] == 0
    ev.evaluate_packed(packed[:40])     # large-batch kernel
    s1 = ev.info()["saturation_events"]
    ev.evaluate_packed(packed[:2])      # small-batch kernel
    s2 = ev.info()["saturation_events"]
    assert s1 > 0 and s2 > s1
    ev.close()


def test_hot_block_internal_tensor_is_rescaled_on_its_own(tmp_path, positions):
    """only the first conv of block 1 runs hot: its block shift, not the stream's"""
    from oracle.net_oracle_np import NumpyGraphOracle
    from unrealgo_b200 import PREC_FP16, DlTFNetworkEvaluator, tfformat
    d = str(tmp_path / "hotblock")
    os.makedirs(d)
    g, gd = tfformat.build_graph(2, 64)
    w = tfformat.random_weights(g.variables, 12)
    w["resnet/unit_res_1/sub1/conv1/bn/gamma"] *= np.float32(2e4)
    w["resnet/unit_res_1/sub2/conv2/conv/kernel"] *= np.float32(1 / 2e4)   # the block's output stays ordinary
    meta, prefix = os.path.join(d, "m.meta"), os.path.join(d, "m")
    tfformat.write_metagraph(meta, gd)
    tfformat.write_bundle(prefix, w)
    packed, planes = positions
    ev = DlTFNetworkEvaluator(meta, max_batch=64)
    assert ev.UpdateCheckPoint(prefix)
    info = ev.info()
    assert info["precision"] == PREC_FP16 and info["fp16_stream_shift"] == 0 and info["fp16_max_block_shift"] > 0
    want_p, want_v = NumpyGraphOracle(meta, prefix).evaluate_reference(planes[:30])
    for n in (30, 4):
        p, v = ev.evaluate_packed(packed[:n])
        assert np.abs(p - want_p[:n]).max() <= 1e-3 and np.abs(v - want_v[:n]).max() <= 5e-3
    assert ev.info()["saturation_events"] == 0
    ev.close()
