"""BASELINE.json configs[1]: the 20-block x 256-filter net at batch 256, through the C ABI, against the oracle on
the same 256 positions — plus size-independent properties at that size (softmax sums, value range,
batch-composition independence, determinism)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def flagship(tmp_path_factory):
    from unrealgo_b200 import tfformat
    d = str(tmp_path_factory.mktemp("flagship"))
    return tfformat.write_synthetic_checkpoint(d, 20, 256, seed=20260119)


def test_flagship_batch256_parity_and_properties(flagship):
    import oracle
    from oracle.net_oracle import GraphOracle
    from unrealgo_b200 import DlTFNetworkEvaluator
    packed, planes = oracle.synthetic_positions(256, seed=20260119)
    ev = DlTFNetworkEvaluator(flagship[0], max_batch=256)  # default arithmetic: fp16 operands, fp32 accumulate
    assert ev.UpdateCheckPoint(flagship[1])
    info = ev.info()
    assert (info["blocks"], info["filters"]) == (20, 256)
    assert info["flops_per_position"] == 17063658984.0
    pol = np.zeros((256, 362))
    val = np.zeros(256)
    assert ev.Evaluate(planes, pol, val, 256)
    want_p, want_v = GraphOracle(*flagship).evaluate_reference(planes)
    dp, dv = np.abs(pol - want_p).max(), np.abs(val - want_v).max()
    print(f"flagship 20x256 batch 256: max|dpolicy|={dp:.3e} max|dvalue|={dv:.3e} "
          f"mean|dpolicy|={np.abs(pol - want_p).mean():.3e} max p={want_p.max():.3f}")
    assert dp <= 1e-3 and dv <= 5e-3
    # properties
    assert np.abs(pol.sum(1) - 1).max() < 1e-5 and (pol >= 0).all()
    assert (np.abs(val) <= 1).all()
    p2, v2 = ev.evaluate_packed(packed)
    assert np.array_equal(p2.astype(np.float64), pol) and np.array_equal(v2.astype(np.float64), val)  # same planes
    # a position's result does not depend on its batch neighbours or on the batch size (tail tiles)
    for lo, hi in [(0, 1), (5, 12), (100, 117)]:
        ps, vs = ev.evaluate_packed(packed[lo:hi])
        assert np.array_equal(ps, p2[lo:hi]) and np.array_equal(vs, v2[lo:hi])
    for ctas in (1, 2):  # both tile shapes compute identical sums in identical order per output element
        ev.set_conv_ctas(ctas)
        pc, vc = ev.evaluate_packed(packed)
        assert np.array_equal(pc, p2) and np.array_equal(vc, v2)
    ev.close()
    # bf16-operand mode on the same inputs: documented looser bound (8-bit mantissa through 41 convolutions)
    from unrealgo_b200 import PREC_BF16
    evb = DlTFNetworkEvaluator(flagship[0], max_batch=256, precision=PREC_BF16)
    assert evb.UpdateCheckPoint(flagship[1])
    pb, vb = evb.evaluate_packed(packed)
    dpb, dvb = np.abs(pb - want_p).max(), np.abs(vb - want_v).max()
    print(f"flagship bf16 mode: max|dpolicy|={dpb:.3e} max|dvalue|={dvb:.3e} mean|dpolicy|={np.abs(pb - want_p).mean():.3e}")
    assert dpb <= 1e-2 and dvb <= 6e-2 and np.abs(pb - want_p).mean() <= 1e-5
    evb.close()
