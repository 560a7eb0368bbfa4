"""K3/K4 unit parity: one tcgen05 implicit-GEMM 3x3 conv launch vs an fp32 torch convolution of the same
bf16-rounded operands (floating-point kernel => torch fp32 reference; tolerance = one bf16 rounding of the
output plus fp32 accumulation-order noise)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _conv_case(n, cin, cout, ck, ctas, residual, relu, seed):
    import torch
    import torch.nn.functional as F
    from unrealgo_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 19, 19, cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5).to(torch.bfloat16)  # HWIO
    bias = torch.randn(cout, generator=g) * 0.1
    res = torch.randn(n, 19, 19, cout, generator=g).to(torch.bfloat16) if residual else None
    # packed weights [cout][tap*cin + ci]
    wp = w.permute(3, 0, 1, 2).reshape(cout, 9 * cin).contiguous()
    dx, dw, db = x.cuda(), wp.cuda(), bias.cuda()
    dres = res.cuda() if residual else None
    dout = torch.full((n, 19, 19, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    rc = lib.ug_k_conv3x3_bf16(0, dx.data_ptr(), dw.data_ptr(), db.data_ptr(), dres.data_ptr() if residual else None,
                               dout.data_ptr(), n, cin, cout, 1 if relu else 0, ck, ctas)
    assert rc == 0, (rc, lib.ug_last_error(None))
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    ref = ref + bias
    if residual:
        ref = ref + res.float()
    if relu:
        ref = torch.relu(ref)
    got = dout.float().cpu()
    assert torch.isfinite(got).all(), f"non-finite outputs: {(~torch.isfinite(got)).sum().item()}"
    err = (got - ref).abs()
    tol = 2.0 ** -7 * ref.abs().clamp(min=1.0)
    bad = err > tol
    if bad.any():
        idx = bad.nonzero()[:8].tolist()
        raise AssertionError(f"max err {err.max().item():.4g}; {bad.sum().item()}/{bad.numel()} outside tol; "
                             f"first bad (n,y,x,c): {idx}")


@pytest.mark.parametrize("ctas", [1, 2])
@pytest.mark.parametrize("n,cin,cout,ck,residual,relu", [
    (1, 64, 64, 64, False, True),
    (3, 64, 64, 64, True, True),
    (2, 256, 256, 64, True, True),
    (16, 256, 256, 64, False, False),
    (5, 32, 64, 32, False, True),     # stem shape: 32 stored input channels, SWIZZLE_64B
    (7, 32, 256, 32, False, True),
    (4, 128, 128, 64, True, True),
])
def test_conv3x3_matches_fp32(n, cin, cout, ck, residual, relu, ctas):
    _conv_case(n, cin, cout, ck, ctas, residual, relu, seed=n * 1000 + cin + cout + ctas)
