"""Latency of one evaluation at the batch sizes the reference's engine issues (1 .. 16) and a little beyond, on the
20x256 net: the small-batch tower kernel against the large-batch one, device time (CUDA events inside the library)
and end to end through DlTFNetworkEvaluator::Evaluate with host buffers.  Run on the GPU box:
    python tests/gpu_small_batch_probe.py [blocks filters] [--pairs]   (--pairs: single-CTA against CTA-pair tiles)"""
import json
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
from unrealgo_b200 import DlTFNetworkEvaluator, synth, tfformat  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    blocks = int(args[0]) if len(args) > 0 else 20
    filters = int(args[1]) if len(args) > 1 else 256
    d = tempfile.mkdtemp(prefix="ug_sb_")
    meta, prefix = tfformat.write_synthetic_checkpoint(d, blocks, filters, seed=20260119)
    ev = DlTFNetworkEvaluator(meta, max_batch=128)
    assert ev.UpdateCheckPoint(prefix)
    packed = synth.synthetic_packed_positions(128, seed=3)
    planes = synth.packed_to_planes(packed)
    out = {}
    sizes = (1, 2, 3, 4, 6, 8, 12, 16, 20, 24, 26) if "--pairs" in sys.argv else (1, 2, 4, 8, 16, 24, 32, 48, 64, 96, 128)
    # nt + 256: single-CTA tiles, nt + 512: CTA-pair tiles (ug_eval_set_small_batch)
    variants = (("v2", 0, 0), ("sb", 128, 0), ("single", 128, 256), ("pair", 128, 512), ("pair32", 128, 512 + 32),
                ("pair64", 128, 512 + 64), ("pair128", 128, 512 + 128)) if "--pairs" in sys.argv else \
               (("v2", 0, 0), ("sb", 128, 0), ("sb16", 128, 16), ("sb32", 128, 32), ("sb64", 128, 64), ("sb128", 128, 128))
    if "--auto" in sys.argv:
        sizes, variants = (1, 2, 4, 6, 8, 12, 14, 16, 20, 24, 26), (("v2", 0, 0), ("sb", 128, 0))
    for n in sizes:
        row = {}
        for name, sb, nt in variants:
            ev.set_small_batch(sb, nt)
            br = ev.bench(packed[:n], 50)
            pol, val = np.zeros((n, 362)), np.zeros(n)
            for _ in range(5):
                ev.Evaluate(planes[:n], pol, val, n)
            t0 = time.perf_counter()
            for _ in range(50):
                ev.Evaluate(planes[:n], pol, val, n)
            e2e = (time.perf_counter() - t0) / 50
            row[name] = {"device_us": round(br["ms_total"] / 50 * 1e3, 1), "tower_us": round(br["ms_tower"] * 1e3, 1),
                         "heads_us": round(br["ms_heads"] * 1e3, 1), "e2e_us": round(e2e * 1e6, 1),
                         "positions_per_s_e2e": round(n / e2e)}
        out[n] = row
        print(n, json.dumps(row), flush=True)
    ev.close()


if __name__ == "__main__":
    main()
