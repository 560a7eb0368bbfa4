"""Bring-up diagnostics for the tcgen05 conv kernel (run on the GPU box, not a pytest).  For each variant it
runs a one-hot-weight convolution (out[..., co] = in[y+r-1, x+s-1, ci]) so that a wrong tap order, a missing
halo shift or a swizzle/descriptor mismatch shows up as a recognisable pattern instead of noise."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from unrealgo_b200 import _lib  # noqa: E402

lib = _lib.load()


def run(n, cin, cout, ck, ctas, w_hwio, x, bias=None, res=None, relu=0):
    wp = w_hwio.permute(3, 0, 1, 2).reshape(cout, 9 * cin).contiguous().to(torch.bfloat16).cuda()
    dx = x.to(torch.bfloat16).cuda()
    db = (bias if bias is not None else torch.zeros(cout)).float().cuda()
    dres = res.to(torch.bfloat16).cuda() if res is not None else None
    out = torch.full((n, 19, 19, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    rc = lib.ug_k_conv3x3_bf16(0, dx.data_ptr(), wp.data_ptr(), db.data_ptr(), dres.data_ptr() if res is not None else None,
                               out.data_ptr(), n, cin, cout, relu, ck, ctas)
    if rc != 0:
        print("   launch rc", rc, lib.ug_last_error(None))
        return None
    torch.cuda.synchronize()
    return out.float().cpu()


def ref_conv(x, w):
    return F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)


def diagnose(n, cin, cout, ck, ctas):
    print(f"== n={n} cin={cin} cout={cout} ck={ck} ctas={ctas}")
    g = torch.Generator().manual_seed(1)
    x = torch.randint(-4, 5, (n, 19, 19, cin), generator=g).float()  # small ints: exact in bf16 and fp32
    # one-hot: output channel co copies input channel (co*7)%cin at tap co%9
    w = torch.zeros(3, 3, cin, cout)
    for co in range(cout):
        t = co % 9
        w[t // 3, t % 3, (co * 7) % cin, co] = 1.0
    got = run(n, cin, cout, ck, ctas, w, x)
    if got is None:
        return False
    want = ref_conv(x, w)
    nan = (~torch.isfinite(got)).sum().item()
    bad = (got != want)
    print(f"   one-hot: nan={nan} mismatches={bad.sum().item()}/{bad.numel()}")
    if bad.any():
        per_co = bad.sum(dim=(0, 1, 2))
        print("   bad per output channel (first 18):", per_co[:18].tolist())
        per_px = bad.sum(dim=3)[0]
        print("   bad per pixel of image 0 (rows 0,1,9,18):", per_px[0].tolist(), per_px[1].tolist(), per_px[9].tolist(), per_px[18].tolist())
        # try alternative hypotheses
        for name, alt in [("taps transposed", w.permute(1, 0, 2, 3)), ("taps flipped", w.flip(0, 1))]:
            print(f"   hypothesis {name}: mismatches={(got != ref_conv(x, alt)).sum().item()}")
        idx = bad.nonzero()[:5].tolist()
        for i in idx:
            print("   sample", i, "got", got[tuple(i)].item(), "want", want[tuple(i)].item())
    # random weights + bias + residual + relu
    wr = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    xr = torch.randn(n, 19, 19, cin, generator=g)
    bias = torch.randn(cout, generator=g)
    res = torch.randn(n, 19, 19, cout, generator=g)
    got = run(n, cin, cout, ck, ctas, wr, xr, bias, res, 1)
    want = torch.relu(ref_conv(xr.to(torch.bfloat16), wr.to(torch.bfloat16)) + bias + res.to(torch.bfloat16).float())
    err = (got - want).abs()
    print(f"   random: max err {err.max().item():.4g} (ref max {want.abs().max().item():.3g}) nan={(~torch.isfinite(got)).sum().item()}")
    return not bad.any() and err.max().item() < 0.05


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), lib.ug_version())
    ok = True
    for cfg in [(1, 64, 64, 64, 1), (3, 64, 64, 64, 1), (2, 256, 256, 64, 1), (3, 32, 64, 32, 1),
                (1, 64, 64, 64, 2), (3, 256, 256, 64, 2), (3, 32, 64, 32, 2), (40, 256, 256, 64, 2)]:
        try:
            ok = diagnose(*cfg) and ok
        except Exception as e:  # a trapped kernel poisons the context: stop
            print("   EXCEPTION", repr(e))
            ok = False
            break
    print("ALL OK" if ok else "FAILURES")
