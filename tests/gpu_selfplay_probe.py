"""Self-play leg alone on one device (diagnostics): python tests/gpu_selfplay_probe.py <device> [games] [moves] [threads]"""
import os
import sys
import time

sys.path.insert(0, ".")
from unrealgo_b200 import DlTFNetworkEvaluator, selfplay, tfformat  # noqa: E402

dev = int(sys.argv[1])
games = int(sys.argv[2]) if len(sys.argv) > 2 else 768
moves = int(sys.argv[3]) if len(sys.argv) > 3 else 1
threads = int(sys.argv[4]) if len(sys.argv) > 4 else 10
d = f"/tmp/ug_probe_ckpt_{dev}"
meta, prefix = tfformat.write_synthetic_checkpoint(d, 20, 256, seed=20260119)
ev = DlTFNetworkEvaluator(meta, device=dev, max_batch=262)
assert ev.UpdateCheckPoint(prefix), ev._err()
t0 = time.time()
st, _ = selfplay(ev, games=games, threads=threads, playouts_per_move=1600, max_moves=moves, max_batch=262,
                 timeout_us=300, reuse_tree=True, random_opening_moves=1, seed=1000 + dev)
print(f"device {dev}: {st['playouts_per_s']:.0f} playouts/s mean_batch {st['mean_batch']:.1f} in {time.time() - t0:.1f}s", flush=True)
ev.close()
