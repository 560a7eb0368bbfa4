#!/usr/bin/env python
"""Compare libugeval with what the reference's TensorFlow evaluator returned — the GPU half of the parity harness.

Input: the .npz written by tools/tf_dump_reference.py (planes + TensorFlow's policy/value for them) and the same
.meta / checkpoint.  Runs the product through its reference-facing entry (DlTFNetworkEvaluator::Evaluate ->
ug_eval_run_planes: host int8 planes in, host doubles out) in each arithmetic mode, prints max-abs errors, and exits
non-zero when a mode misses its tolerance (BASELINE.json north_star: policy <= 1e-3, value <= 5e-3 for the 16-bit
tensor-core build, tighter for the fp32 validation build).  Needs a B200; no TensorFlow.

    python tools/compare_with_tf.py --dump tf_reference.npz --meta bootstrap-ckpt.meta --ckpt bootstrap-ckpt
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

TOL = {"fp32": (1e-5, 1e-5), "fp16": (1e-3, 5e-3), "bf16": (1e-3, 5e-3)}   # (policy, value) max-abs


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--dump", required=True)
    ap.add_argument("--meta", required=True)
    ap.add_argument("--ckpt", required=True)
    ap.add_argument("--modes", default="fp32,fp16,bf16")
    ap.add_argument("--required", default="fp32,fp16", help="modes whose failure makes the exit code non-zero")
    ap.add_argument("--batch", type=int, default=16)
    args = ap.parse_args()

    from unrealgo_b200 import PREC_BF16, PREC_FP16, PREC_FP32, DlTFNetworkEvaluator
    prec = {"fp32": PREC_FP32, "fp16": PREC_FP16, "bf16": PREC_BF16}
    d = np.load(args.dump)
    info = json.loads(str(d["info"]))
    planes, want_p = d["planes"], d["policy"].astype(np.float64)
    want_v = d["value_transformed"].astype(np.float64)
    n = planes.shape[0]
    print(f"reference dump: producer={info.get('producer')} version={info.get('tensorflow_version')} positions={n} "
          f"input={info.get('input')} ({info.get('input_dtype')}) value_transform={info.get('value_transform')}")
    if info.get("producer") != "tensorflow":
        print("NOTE: this dump was NOT produced by TensorFlow (stand-in): it exercises the harness, it pins nothing.")
    failed = []
    for mode in args.modes.split(","):
        ev = DlTFNetworkEvaluator(args.meta, feature_input=info.get("input"), outputs=info.get("outputs"),
                                  precision=prec[mode], max_batch=max(args.batch, 16))
        if not ev.UpdateCheckPoint(args.ckpt):
            print(f"{mode}: UpdateCheckPoint failed: {ev._err()}")
            failed.append(mode)
            continue
        ev.set_value_transform(int(info.get("value_transform", 0)))
        pol, val = np.zeros((n, 362)), np.zeros(n)
        for lo in range(0, n, args.batch):
            hi = min(n, lo + args.batch)
            assert ev.Evaluate(planes[lo:hi], pol[lo:hi], val[lo:hi], hi - lo), ev._err()
        dp, dv = np.abs(pol - want_p).max(), np.abs(val - want_v).max()
        top1 = float((pol.argmax(1) == want_p.argmax(1)).mean())
        ok = dp <= TOL[mode][0] and dv <= TOL[mode][1]
        print(f"{mode}: max|dpolicy|={dp:.3e} (tol {TOL[mode][0]:g})  max|dvalue|={dv:.3e} (tol {TOL[mode][1]:g})  "
              f"mean|dpolicy|={np.abs(pol - want_p).mean():.3e}  top-1 agreement={top1:.3f}  {'OK' if ok else 'FAIL'}  "
              f"net={ev.info()['blocks']}x{ev.info()['filters']}")
        if not ok:
            failed.append(mode)
        ev.close()
    bad = [m for m in failed if m in args.required.split(",")]
    if bad:
        print("FAILED:", ",".join(bad))
        return 1
    print("parity within tolerance" + ("" if info.get("producer") == "tensorflow" else " (stand-in dump: harness check only)"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
