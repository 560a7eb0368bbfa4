"""Prints max/mean policy and value error of each arithmetic mode against the oracle on synthetic nets
(run on the GPU box; feeds the tolerances quoted in DESIGN.md)."""
import sys
import tempfile

import numpy as np

sys.path.insert(0, ".")
import oracle  # noqa: E402
from oracle.net_oracle import GraphOracle  # noqa: E402
from unrealgo_b200 import PREC_BF16, PREC_FP16, PREC_FP32, DlTFNetworkEvaluator, tfformat  # noqa: E402

packed, planes = oracle.synthetic_positions(256, seed=20260119)
for blocks, filters, n, seed in [(2, 64, 64, 11), (3, 256, 64, 13), (10, 128, 128, 3), (20, 256, 256, 20260119), (20, 256, 256, 7)]:
    d = tempfile.mkdtemp()
    meta, prefix = tfformat.write_synthetic_checkpoint(d, blocks, filters, seed=seed)
    want_p, want_v = GraphOracle(meta, prefix).evaluate_reference(planes[:n])
    for name, prec, split in [("bf16", PREC_BF16, 1), ("bf16-nosplit", PREC_BF16, 0), ("fp16", PREC_FP16, 1),
                              ("fp16-nosplit", PREC_FP16, 0), ("fp32", PREC_FP32, 1)]:
        ev = DlTFNetworkEvaluator(meta, precision=prec, max_batch=n)
        assert ev.UpdateCheckPoint(prefix)
        ev.set_residual_split(split)
        p, v = ev.evaluate_packed(packed[:n])
        dp, dv = np.abs(p - want_p), np.abs(v - want_v)
        print(f"{blocks}x{filters} seed {seed} n={n} {name}: max|dp|={dp.max():.3e} mean|dp|={dp.mean():.3e} "
              f"p99.9|dp|={np.quantile(dp, 0.999):.3e} max|dv|={dv.max():.3e} mean|dv|={dv.mean():.3e} "
              f"(max p {want_p.max():.3f}, |v| max {np.abs(want_v).max():.3f})", flush=True)
        ev.close()
