"""Probe (not a test): the reference's own search engine (oracle/_ref/ref_search) on libugeval with the flagship
20-block x 256-filter network — what the unmodified engine reaches with its one-leaf-per-call hand-off — next to the
product's self-play driver on one game.  Usage: python tests/gpu_reference_search_probe.py [playouts]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from unrealgo_b200 import tfformat
    playouts = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    d = tempfile.mkdtemp(prefix="refsearch_")
    meta, prefix = tfformat.write_synthetic_checkpoint(os.path.join(d, "flagship"), 20, 256, seed=5)
    outs = "resnet/tower_0/policy_head/policy_predict:resnet/tower_0/value_head/reward_predict"
    with open(os.path.join(d, "dlcfg-19"), "w") as f:
        f.write(f"meta_graph={meta}\ncheckpoint={prefix}\nnn_input=feature_input\nnn_outputs={outs}\n"
                "value_transform=0\nreuse_searchtree=false\n")
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_search")
    for threads in (1, 16):  # 16 = MAX_BATCHES, the most the engine's EvalBuffer holds (DlTFNetworkEvaluator.h:15)
        try:
            r = subprocess.run([exe, str(playouts * (4 if threads > 1 else 1)), "2", str(threads)], cwd=d,
                               capture_output=True, text=True, timeout=90)
            print(f"reference engine on libugeval (20x256), {threads} search thread(s):",
                  [ln for ln in r.stdout.splitlines() if ln.startswith("total")], "rc", r.returncode)
        except subprocess.TimeoutExpired:
            print(f"reference engine with {threads} search thread(s): no result within 90 s")
    # the same engine with the leaf-queue patch: every search thread owns a ticket
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_search_queue")
    for threads in (1, 16, 64, 128, 256):
        try:
            r = subprocess.run([exe, str(playouts * max(1, threads // 2)), "2", str(threads)], cwd=d,
                               capture_output=True, text=True, timeout=60)
            print(f"reference engine + ugeval leaf queue, {threads} search thread(s):",
                  [ln for ln in r.stdout.splitlines() if ln.startswith("total")], "rc", r.returncode)
        except subprocess.TimeoutExpired:
            print(f"queue-patched engine with {threads} search thread(s): no result within 60 s")


if __name__ == "__main__":
    main()
