"""K3/K4 (second generation) unit parity: conv3x3_v2 — shared-memory-resident input tile, filter taps as
row-shifted descriptors with disable-output-lane masks, TMA-staged epilogue — against an fp32 torch
convolution of the same 16-bit-rounded operands.  Single-tap weight patterns isolate the shifted-descriptor
and edge-mask logic (a wrong mask only shows on board edges, a wrong shift everywhere)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALL = [(r, s) for r in range(3) for s in range(3)]


def _run(*a, **k):
    from tests.gpu_debug_conv_v2 import run
    msg = run(*a, **k)
    assert " bad 0/" in msg and "nonfinite 0" in msg, msg


@pytest.mark.parametrize("tap", ALL)
def test_single_tap_exact_padding(tap):
    # n=3: 1083 rows = 4 full 256-row tiles + a partial one; tiles straddle image boundaries
    _run(3, 64, 64, [tap], 0, False, False)


@pytest.mark.parametrize("f16", [True, False])
@pytest.mark.parametrize("n,cin,cout,residual,out_lo", [
    (1, 64, 64, False, False),
    (5, 256, 256, True, True),      # conv2 of a block: hi+lo residual in, hi+lo out
    (4, 128, 128, True, False),     # one residual tensor (residual split off)
    (7, 32, 256, False, True),      # stem: 32 stored input channels (SWIZZLE_64B), opens the hi/lo stream
    (9, 192, 192, False, False),
    (16, 256, 256, False, False),
])
def test_conv_v2_matches_fp32(n, cin, cout, residual, out_lo, f16):
    _run(n, cin, cout, ALL, 0, residual, out_lo, f16=f16, seed=n + cin + cout, split_res=out_lo)


def test_rows_past_the_batch_do_not_leak():
    # buffers sized for 4 images, batch of 2: the rows of images 2..3 are read as halo by the last tile and
    # must be masked out
    _run(2, 128, 128, ALL, 0, True, False, n_alloc=4, split_res=False)


def test_no_relu():
    _run(3, 64, 128, ALL, 0, False, False, relu=False)
