"""K3/K4 for small batches (conv3x3_sb.cu): 128-row x NT-column tiles on single CTAs (cta_group::1) or 256-row x
NT-column tiles on CTA pairs (cta_group::2, each SM stages half of the weights) — the tower kernel for the batch sizes
the reference's own engine issues (1 .. MAX_BATCHES = 16, funcapproximator/DlTFNetworkEvaluator.h:15).

Unit parity of the raw kernel against an fp32 torch convolution of the same 16-bit operands (same harness and
tolerances as conv3x3_v2's), then the whole evaluator at n in {1, 7, 16, 24}: within the north-star tolerance of the
oracle, and bit-identical to what the large-batch kernel (conv3x3_v2) returns for the same positions — a position's
result must not depend on which kernel its batch size selected."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALL = [(r, s) for r in range(3) for s in range(3)]
SINGLE, PAIR = 256, 512   # added to the tile width: force single-CTA / CTA-pair tiles (otherwise chosen from the batch)
DIRECT = 1024             # raw entry only: per-thread global stores even for tiles of 64+ columns (else TMA stores)


def _run(*a, **k):
    from tests.gpu_debug_conv_v2 import run
    msg = run(*a, impl="sb", **k)
    assert " bad 0/" in msg and "nonfinite 0" in msg, msg


@pytest.mark.parametrize("nt", [16, SINGLE + 32, PAIR + 32])
@pytest.mark.parametrize("tap", ALL)
def test_single_tap_exact_padding(tap, nt):
    # n=3: 1083 rows = 8 full 128-row tiles + a partial one; tiles straddle image boundaries
    _run(3, 64, 64, [tap], 0, False, False, nt=nt)


@pytest.mark.parametrize("f16", [True, False])
@pytest.mark.parametrize("n,cin,cout,residual,out_lo,nt", [
    (1, 64, 64, False, False, 0),
    (1, 256, 256, True, True, 16),
    (1, 256, 256, True, True, 32),
    (5, 256, 256, True, True, 64),      # conv2 of a block: hi+lo residual in, hi+lo out
    (16, 256, 256, True, True, 128),
    (4, 128, 128, True, False, 0),      # one residual tensor (residual split off)
    (7, 32, 256, False, True, 0),       # stem: 32 stored input channels (SWIZZLE_64B), opens the hi/lo stream
    (9, 192, 192, False, False, 64),
    (24, 256, 256, False, False, 0),    # more CTAs than SMs at every tile width: two waves
    # CTA pairs: an odd number of 128-row tiles (the last pair's second tile lies wholly past M), every width
    (1, 256, 256, True, True, PAIR + 32),
    (5, 256, 256, True, True, PAIR + 64),
    (16, 256, 256, True, True, PAIR + 128),
    (2, 128, 128, True, False, PAIR + 128),
    (7, 32, 256, False, True, PAIR),
    (9, 192, 192, False, False, PAIR + 64),
    (24, 256, 256, False, False, PAIR),
    (16, 256, 256, True, True, SINGLE + 128),
    (3, 64, 64, False, False, SINGLE),
    # wide tiles with per-thread stores (the default for them is staged TMA stores)
    (16, 256, 256, True, True, SINGLE + 128 + DIRECT),
    (5, 256, 256, True, True, PAIR + 64 + DIRECT),
    (9, 192, 192, False, False, 64 + DIRECT),
    (3, 128, 128, True, False, PAIR + 128),   # 9 tiles: the last pair's second tile is past M, TMA stores skip it
])
def test_conv_sb_matches_fp32(n, cin, cout, residual, out_lo, nt, f16):
    _run(n, cin, cout, ALL, 0, residual, out_lo, f16=f16, seed=n + cin + cout, split_res=out_lo, nt=nt)


@pytest.mark.parametrize("mode", [SINGLE, PAIR])
def test_rows_past_the_batch_do_not_leak(mode):
    _run(2, 128, 128, ALL, 0, True, False, n_alloc=4, split_res=False, nt=mode)


@pytest.mark.parametrize("mode", [SINGLE, PAIR])
def test_no_relu_and_programmatic_launch(mode):
    _run(3, 64, 128, ALL, 0, False, False, relu=False, pdl=1, iters=3, nt=mode)


def _ev(ckpt, precision, max_batch=64):
    from unrealgo_b200 import DlTFNetworkEvaluator
    ev = DlTFNetworkEvaluator(ckpt[0], precision=precision, max_batch=max_batch)
    assert ev.UpdateCheckPoint(ckpt[1])
    return ev


@pytest.mark.parametrize("precision", [2, 0], ids=["fp16", "bf16"])
@pytest.mark.parametrize("which", ["small", "wide"])
def test_small_batches_through_the_evaluator(small_ckpt, wide_ckpt, positions, which, precision):
    from oracle.net_oracle import GraphOracle
    ckpt = small_ckpt if which == "small" else wide_ckpt
    packed, planes = positions
    ev = _ev(ckpt, precision)
    assert ev.info()["small_batch_max"] >= 16
    orc = GraphOracle(*ckpt)
    pol_tol, val_tol = (1e-3, 5e-3) if precision == 2 else (5e-3, 3e-2)   # as tests/test_gpu_eval.py
    for n in (1, 7, 16, 24):
        want_p, want_v = orc.evaluate_reference(planes[:n])
        ev.set_small_batch(24)
        p, v = ev.evaluate_packed(packed[:n])
        assert np.abs(p - want_p).max() <= pol_tol and np.abs(v - want_v).max() <= val_tol, (n, np.abs(p - want_p).max())
        pol, val = np.zeros((n, 362)), np.zeros(n)
        assert ev.Evaluate(planes[:n], pol, val, n)     # the reference-facing entry takes the same path
        assert np.array_equal(pol.astype(np.float32), p) and np.array_equal(val.astype(np.float32), v)
        # every tile width, on single CTAs or CTA pairs, computes the same sums in the same order
        for nt in (16, 32, 64, 128, SINGLE + 32, SINGLE + 128, PAIR + 32, PAIR + 64, PAIR + 128):
            ev.set_small_batch(24, nt)
            pn, vn = ev.evaluate_packed(packed[:n])
            assert np.array_equal(pn, p) and np.array_equal(vn, v), (n, nt)
        ev.set_small_batch(0)                           # the large-batch kernel on the same positions
        p2, v2 = ev.evaluate_packed(packed[:n])
        assert np.array_equal(p2, p) and np.array_equal(v2, v), (n, np.abs(p2 - p).max(), np.abs(v2 - v).max())
    ev.close()


def test_small_batch_intermediate_layers_match_the_large_batch_kernel(wide_ckpt, positions):
    packed, _ = positions
    ev = _ev(wide_ckpt, 2)
    n = 5
    for layer in range(0, 7):
        ev.set_debug_stop(layer)
        ev.set_small_batch(24)
        ev.evaluate_packed(packed[:n])
        a = ev.debug_layer(layer, n)
        ev.set_small_batch(0)
        ev.evaluate_packed(packed[:n])
        b = ev.debug_layer(layer, n)
        assert np.array_equal(a, b), (layer, np.abs(a - b).max())
    ev.close()
