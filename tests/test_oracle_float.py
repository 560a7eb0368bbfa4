"""The floating-point oracle is two independent readings of TensorFlow's op definitions — oracle/net_oracle.py (torch,
fp32) and oracle/net_oracle_np.py (numpy, float64, explicit tap sums and dataflow semantics).  Neither is TensorFlow:
parity against the reference's real evaluator stays unpinned until tools/tf_dump_reference.py has been run where
TensorFlow and the bootstrap checkpoint exist.  What this file pins is that the two readings agree, so a
misunderstanding of an op would have to be made twice to go unnoticed."""
import numpy as np
import pytest


@pytest.mark.parametrize("kw", [dict(), dict(conv_bias=True, input_dtype=1), dict(value_channels=2, policy_channels=4),
                                dict(bn_eps=1e-3, value_hidden=64)])
def test_torch_and_numpy_oracles_agree(tmp_path, positions, kw):
    from oracle.net_oracle import GraphOracle
    from oracle.net_oracle_np import NumpyGraphOracle
    from unrealgo_b200 import tfformat
    ckpt = tfformat.write_synthetic_checkpoint(str(tmp_path / "ck"), 3, 64, seed=5, **kw)
    _, planes = positions
    a = GraphOracle(*ckpt).evaluate_reference(planes[:8])
    b = NumpyGraphOracle(*ckpt).evaluate_reference(planes[:8])
    assert np.abs(a[0] - b[0]).max() <= 2e-6 and np.abs(a[1] - b[1]).max() <= 5e-6
    for vt in (1, 2):
        av = GraphOracle(*ckpt).evaluate_reference(planes[:2], value_transform=vt)[1]
        bv = NumpyGraphOracle(*ckpt).evaluate_reference(planes[:2], value_transform=vt)[1]
        assert np.abs(av - bv).max() <= 5e-6


def test_intermediate_tensors_agree(small_ckpt, positions):
    """layer by layer, not only at the outputs (errors could cancel in a softmax)"""
    import oracle
    from oracle.net_oracle import GraphOracle
    from oracle.net_oracle_np import NumpyGraphOracle
    _, planes = positions
    nhwc = oracle.transform_feature(planes[:3], 0, True)
    names = ["resnet/tower_0/init_conv/Relu", "resnet/tower_0/unit_res_0/sub1/conv1/Relu", "resnet/tower_0/unit_res_1/Relu",
             "resnet/tower_0/policy_head/dense/BiasAdd", "resnet/tower_0/value_head/dense1/Relu"]
    _, _, ta = GraphOracle(*small_ckpt).run(nhwc, fetch=names)
    _, _, tb = NumpyGraphOracle(*small_ckpt).run(nhwc, fetch=names)
    for n in names:
        scale = np.abs(tb[n]).max()
        assert np.abs(ta[n] - tb[n]).max() <= 2e-6 * max(scale, 1.0), n


def test_numpy_oracle_conv_against_a_hand_computed_case():
    """SAME padding, NHWC x HWIO, tap order: a 3x3 board-corner case small enough to write down"""
    from oracle.net_oracle_np import NumpyGraphOracle
    x = np.zeros((1, 3, 3, 1))
    x[0, 0, 0, 0], x[0, 1, 2, 0] = 2.0, 5.0
    w = np.arange(9, dtype=np.float64).reshape(3, 3, 1, 1)      # w[dy][dx]
    y = NumpyGraphOracle._conv2d(x, w, {"padding": {"s": b"SAME"}, "strides": {"list_i": [1, 1, 1, 1]}})[0, :, :, 0]
    # out[y][x] = sum_{dy,dx} in[y+dy-1][x+dx-1] * w[dy][dx]
    want = np.zeros((3, 3))
    for oy in range(3):
        for ox in range(3):
            for dy in range(3):
                for dx in range(3):
                    iy, ix = oy + dy - 1, ox + dx - 1
                    if 0 <= iy < 3 and 0 <= ix < 3:
                        want[oy, ox] += x[0, iy, ix, 0] * w[dy, dx, 0, 0]
    assert np.array_equal(y, want) and y[0, 0] == 2 * 4 and y[1, 1] == 2 * 0 + 5 * 5
