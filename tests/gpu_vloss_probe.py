"""Diagnostics: few games with many virtual-loss leaves in flight on the 20x256 net (python tests/gpu_vloss_probe.py)."""
import sys

sys.path.insert(0, ".")
from unrealgo_b200 import DlTFNetworkEvaluator, selfplay, tfformat  # noqa: E402

meta, prefix = tfformat.write_synthetic_checkpoint("/tmp/ug_vl_ckpt", 20, 256, seed=20260119)
ev = DlTFNetworkEvaluator(meta, max_batch=262)
assert ev.UpdateCheckPoint(prefix), ev._err()
for games, k in ((8, 48), (1, 64), (1, 1)):
    st, _ = selfplay(ev, games=games, threads=min(games, 4), playouts_per_move=1600, max_moves=3, max_batch=262,
                     timeout_us=300, random_opening_moves=1, seed=3, leaves_per_game=k)
    print(f"20x256: {games} game(s) x {k} leaves in flight: {st['playouts_per_s']:.0f} playouts/s, "
          f"mean batch {st['mean_batch']:.1f}", flush=True)
ev.close()
