"""Diagnostics: two evaluators (own workspaces, streams, queues) driving the same GPU from one process, to see
whether interleaving two towers hides the per-layer tails.  python tests/gpu_selfplay_dual_probe.py [games_each]"""
import sys
import threading
import time

sys.path.insert(0, ".")
from unrealgo_b200 import DlTFNetworkEvaluator, selfplay, tfformat  # noqa: E402

games = int(sys.argv[1]) if len(sys.argv) > 1 else 512
meta, prefix = tfformat.write_synthetic_checkpoint("/tmp/ug_dual_ckpt", 20, 256, seed=20260119)
evs = []
for i in range(2):
    ev = DlTFNetworkEvaluator(meta, device=0, max_batch=262)
    assert ev.UpdateCheckPoint(prefix), ev._err()
    evs.append(ev)
out = [None, None]


def run(i):
    out[i], _ = selfplay(evs[i], games=games, threads=6, playouts_per_move=1600, max_moves=1, max_batch=262,
                         timeout_us=300, reuse_tree=True, random_opening_moves=1, seed=2000 + i)


t0 = time.time()
ths = [threading.Thread(target=run, args=(i,)) for i in range(2)]
for t in ths:
    t.start()
for t in ths:
    t.join()
dt = time.time() - t0
tot = sum(o["playouts"] for o in out)
print(f"two evaluators on one GPU: {tot / dt:.0f} playouts/s total ({[round(o['playouts_per_s']) for o in out]}, "
      f"mean batches {[round(o['mean_batch'], 1) for o in out]}) in {dt:.1f}s", flush=True)
