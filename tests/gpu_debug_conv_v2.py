"""Bring-up harness for conv3x3_v2 (run on the GPU box): isolates the three hardware assumptions —
(1) pipeline + epilogue (centre tap only), (2) row-shifted SWIZZLE_128B descriptors (single off-centre taps),
(3) disable-output-lane masks (errors confined to board edges if wrong) — then full random weights, both
descriptor modes, and a timing line."""
import ctypes as C
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from unrealgo_b200 import _lib  # noqa: E402

lib = None


def run(n, cin, cout, taps, desc_mode, residual, out_lo, f16=True, relu=True, seed=1, iters=0, n_alloc=None, prof=False, split_res=True,
        impl="v2", nt=0, pdl=0, sb_prof=None):
    """impl "v2": conv3x3_v2 (ug_k_conv3x3_v2); "sb": the small-batch kernel (ug_k_conv3x3_sb, nt = tile width, 0 = auto)"""
    global lib
    if lib is None:
        lib = _lib.load()
    dt = torch.float16 if f16 else torch.bfloat16
    g = torch.Generator().manual_seed(seed)
    n_alloc = n_alloc or n
    x = torch.randn(n_alloc, 19, 19, cin, generator=g).to(dt)
    w = torch.zeros(3, 3, cin, cout)
    for (r, s) in taps:
        w[r, s] = torch.randn(cin, cout, generator=g) * (2.0 / (len(taps) * cin)) ** 0.5
    w = w.to(dt)
    bias = torch.randn(cout, generator=g) * 0.1
    res_hi = res_lo = None
    if residual:
        rf = torch.randn(n_alloc, 19, 19, cout, generator=g)
        res_hi = rf.to(dt)
        res_lo = (rf - res_hi.float()).to(dt) if split_res else None
    wp = w.permute(3, 0, 1, 2).reshape(cout, 9 * cin).contiguous()
    dx, dw, db = x.cuda(), wp.cuda(), bias.cuda()
    drh = res_hi.cuda() if residual else None
    drl = res_lo.cuda() if res_lo is not None else None
    doh = torch.full((n_alloc, 19, 19, cout), float("nan"), dtype=dt, device="cuda")
    dol = torch.full((n_alloc, 19, 19, cout), float("nan"), dtype=dt, device="cuda") if out_lo else None
    ms = C.c_float(0)
    dprof = torch.zeros(74, 64, dtype=torch.int64, device="cuda") if prof else None
    if impl == "sb":
        rc = lib.ug_k_conv3x3_sb(0, dx.data_ptr(), dw.data_ptr(), db.data_ptr(), drh.data_ptr() if residual else None,
                                 drl.data_ptr() if drl is not None else None, doh.data_ptr(),
                                 dol.data_ptr() if out_lo else None, n, n_alloc, cin, cout, 1 if relu else 0,
                                 1 if f16 else 0, nt, pdl, iters, C.byref(ms), sb_prof.data_ptr() if sb_prof is not None else None)
    else:
        rc = lib.ug_k_conv3x3_v2(0, dx.data_ptr(), dw.data_ptr(), db.data_ptr(), drh.data_ptr() if residual else None,
                                 drl.data_ptr() if drl is not None else None, doh.data_ptr(), dol.data_ptr() if out_lo else None,
                                 n, n_alloc, cin, cout, 1 if relu else 0, 1 if f16 else 0, desc_mode, iters, C.byref(ms),
                                 dprof.data_ptr() if prof else None)
    if rc != 0:
        return f"rc={rc} {lib.ug_last_error(None)}"
    torch.cuda.synchronize()
    ref = F.conv2d(x[:n].float().permute(0, 3, 1, 2), w.float().permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1) + bias
    if residual:
        ref = ref + res_hi[:n].float() + (res_lo[:n].float() if res_lo is not None else 0)
    if relu:
        ref = torch.relu(ref)
    got = doh[:n].float().cpu()
    if out_lo:
        got = got + dol[:n].float().cpu()
    nonfinite = (~torch.isfinite(got)).sum().item()
    err = (torch.nan_to_num(got, nan=1e9) - ref).abs()
    # one 16-bit rounding of the output (hi only) or of the remainder (hi + lo)
    rel = (2.0 ** -9 if f16 else 2.0 ** -7) if not out_lo else (2.0 ** -14 if f16 else 2.0 ** -13)
    tol = rel * ref.abs().clamp(min=1.0)
    bad = err > tol
    msg = f"max err {err.max().item():.3g} bad {bad.sum().item()}/{bad.numel()} nonfinite {nonfinite}"
    if bad.any():
        b = bad.any(dim=3)  # [n,19,19]
        edge = torch.zeros(19, 19, dtype=torch.bool)
        edge[0, :] = edge[18, :] = edge[:, 0] = edge[:, 18] = True
        msg += f" bad pixels {b.sum().item()} (on edge {(b & edge).sum().item()}, interior {(b & ~edge).sum().item()})"
        msg += f" first {bad.nonzero()[:3].tolist()}"
    if prof:
        p = dprof.cpu().double()
        names = ["mma_total", "w_acc_empty", "w_mask", "w_a_full", "w_b_full", "tiles", "", "", "epi_total", "w_acc_full",
                 "w_res", "w_store"]
        msg += "\n    prof(mean cycles/cluster): " + " ".join(f"{nm}={p[:, i].mean().item():.0f}" for i, nm in enumerate(names) if nm)
        msg += f"\n    prof(max mma_total)={p[:, 0].max().item():.0f} min={p[:, 0].min().item():.0f}"
        # the raw entry runs 1 + iters launches on the same counters: bins accumulate over all of them
        bins = p[:, 16:52].mean(0).reshape(4, 9)
        msg += "\n    B-wait by (chunk j, tap step t), mean cycles over launches: " + " | ".join(
            " ".join(f"{x:.0f}" for x in row) for row in bins.tolist()) + f"  first-tile total {p[:, 60].mean().item():.0f}"
    if iters:
        fl = 2.0 * n * 361 * 9 * cin * cout
        msg += f" | {ms.value * 1e3:.1f} us/launch {fl / ms.value / 1e9:.0f} TFLOP/s"
    return msg


if __name__ == "__main__":
    ALL = [(r, s) for r in range(3) for s in range(3)]
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "bringup"):
        for dm in (0,):
            print(f"== desc_mode {dm}", flush=True)
            print("centre only  n=3 64->64 :", run(3, 64, 64, [(1, 1)], dm, False, False), flush=True)
            for tap in [(1, 0), (1, 2), (0, 1), (2, 1), (0, 0), (2, 2)]:
                print(f"tap {tap} n=3 64->64 :", run(3, 64, 64, [tap], dm, False, False), flush=True)
            print("all taps n=3 64->64     :", run(3, 64, 64, ALL, dm, False, False), flush=True)
            print("all taps n=5 256->256 res+lo:", run(5, 256, 256, ALL, dm, True, True), flush=True)
            print("stem     n=7 32->256 lo :", run(7, 32, 256, ALL, dm, False, True), flush=True)
            print("n=2 of alloc 4, 128->128 res:", run(2, 128, 128, ALL, dm, True, False, n_alloc=4), flush=True)
    if which == "prof":
        print("stem :", run(256, 32, 256, ALL, 0, False, True), flush=True)
        print("conv1:", run(256, 256, 256, ALL, 0, False, False), flush=True)
        print("conv2:", run(256, 256, 256, ALL, 0, True, True), flush=True)
    if which == "power":
        # sustained (power-capped) launch time of one tower conv, with and without the weight traffic
        import subprocess
        for dm, name in ((8, "normal"), (8 | 16, "no B reloads (wrong results, timing only)")):
            for rep in range(3):
                msg = run(256, 256, 256, ALL, dm, True, True, iters=12000 if rep == 2 else 3000)
            smi = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"],
                                 capture_output=True, text=True).stdout.strip()
            print(f"{name}: {msg.split('|')[-1].strip()}  (after the loop: {smi})", flush=True)
    if which == "probe":
        for dm in (0, 2, 4, 6):
            print(f"probe desc_mode={dm} conv1:", run(256, 256, 256, ALL, dm, False, False, iters=20, prof=True), flush=True)
    if which == "sbprof":
        # per-CTA time stamps of a chain of identical launches: where do the ~10 us per layer go at n = 1?
        import numpy as np
        # nt + 256: single-CTA tiles, nt + 512: CTA-pair tiles
        for n, nt in ((1, 256 + 32), (1, 512 + 32), (1, 16), (16, 256 + 128), (16, 512 + 128)):
            iters = 12
            prof = torch.zeros(iters, 148, 16, dtype=torch.int64, device="cuda")
            msg = run(n, 256, 256, ALL, 0, True, True, impl="sb", nt=nt, pdl=1, iters=iters, sb_prof=prof)
            p = prof.cpu().numpy().astype(np.float64)
            m_tiles = (n * 361 + 127) // 128
            ctas = ((m_tiles + 1) // 2 * 2 if nt & 512 else m_tiles) * (256 // (nt & 255))
            gx = (m_tiles + 1) // 2 * 2 if nt & 512 else m_tiles
            # CTA pairs: the MMA-issue stamps exist in the leader (even blockIdx.x) only
            sel = [i for i in range(ctas) if not (nt & 512) or (i % gx) % 2 == 0]
            p = p[2:, sel]   # skip the first launches
            clk = lambda a, b: (p[:, :, b] - p[:, :, a])
            ghz = 1.9
            names = [("prologue (start -> roles)", 0, 1), ("pdl_wait", 1, 2), ("wait -> first A chunk", 2, 3),
                     ("first -> last A chunk", 3, 4), ("last A chunk -> MMAs issued", 4, 5), ("MMAs issued -> acc_full", 5, 6),
                     ("acc_full -> epilogue done", 6, 7), ("epilogue done -> exit", 7, 10), ("whole CTA", 0, 10)]
            print(f"n={n} nt={nt} ctas={ctas}: {msg.split('|')[-1].strip()}")
            for nm, a, b in names:
                d = clk(a, b) / ghz / 1e3
                print(f"    {nm:<34s} mean {d.mean():6.2f} us  min {d.min():6.2f}  max {d.max():6.2f}")
            g0, g1 = p[:, :, 8], p[:, :, 9]
            print(f"    global timer: launch span (first start -> last end) {((g1.max(1) - g0.min(1))).mean() / 1e3:.2f} us, "
                  f"launch-to-launch (start_i+1 - start_i) {np.diff(g0.min(1)).mean() / 1e3:.2f} us, "
                  f"end_i -> first wait-return_i+1 n/a, overlap (end_i - start_i+1) {(g1.max(1)[:-1] - g0.min(1)[1:]).mean() / 1e3:.2f} us")
    if which == "sb":
        print("sb centre  n=3 64->64 nt32 :", run(3, 64, 64, [(1, 1)], 0, False, False, impl="sb", nt=32), flush=True)
        for tap in [(1, 0), (0, 1), (2, 2)]:
            print(f"sb tap {tap} n=3 64->64 nt16:", run(3, 64, 64, [tap], 0, False, False, impl="sb", nt=16), flush=True)
        for nt in (16, 32, 64, 128):
            print(f"sb all taps n=5 256->256 res+lo nt={nt}:", run(5, 256, 256, ALL, 0, True, True, impl="sb", nt=nt), flush=True)
        print("sb stem n=7 32->256 lo auto:", run(7, 32, 256, ALL, 0, False, True, impl="sb"), flush=True)
        for n in (1, 2, 4, 8, 16, 32):
            for nt in (16, 32, 64, 128):
                print(f"sb time conv2 n={n} nt={nt} pdl:", run(n, 256, 256, ALL, 0, True, True, impl="sb", nt=nt, pdl=1, iters=200).split("|")[-1], flush=True)
            print(f"v2 time conv2 n={n} pdl      :", run(n, 256, 256, ALL, 8, True, True, iters=200).split("|")[-1], flush=True)
    if which in ("all", "time"):
        dm = int(sys.argv[2]) if len(sys.argv) > 2 else 0
        print("time conv1 n=256 256->256         :", run(256, 256, 256, ALL, dm, False, False, iters=20, prof=True), flush=True)
        print("time conv2 n=256 256->256 res+lo  :", run(256, 256, 256, ALL, dm, True, True, iters=20, prof=True), flush=True)
        print("time stem  n=256 32->256 lo       :", run(256, 32, 256, ALL, dm, False, True, iters=20, prof=True), flush=True)
        print("time conv1 +PDL                   :", run(256, 256, 256, ALL, dm | 8, False, False, iters=20), flush=True)
        print("time conv2 +PDL                   :", run(256, 256, 256, ALL, dm | 8, True, True, iters=20), flush=True)
        print("time conv2 bf16 n=256             :", run(256, 256, 256, ALL, dm, True, True, f16=False, iters=20), flush=True)
