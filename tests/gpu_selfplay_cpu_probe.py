"""Self-play on one GPU with the host budget of an 8-GPU box (4 CPUs per rank): where does the host side stand?
    python tests/gpu_selfplay_cpu_probe.py [cpus=4] [games=768] [prelude=80]
Prints playouts/s, mean batch, search-thread busy fraction, host us per playout and the queue's bookkeeping, with and
without transparent huge pages for the trees."""
import json
import os
import sys
import tempfile

sys.path.insert(0, ".")
from unrealgo_b200 import DlTFNetworkEvaluator, selfplay, tfformat  # noqa: E402


def main():
    cpus = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    games = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    prelude = int(sys.argv[3]) if len(sys.argv) > 3 else 80
    os.sched_setaffinity(0, set(sorted(os.sched_getaffinity(0))[:cpus]))
    d = tempfile.mkdtemp(prefix="ug_spprobe_")
    meta, prefix = tfformat.write_synthetic_checkpoint(d, 20, 256, seed=20260119)
    ev = DlTFNetworkEvaluator(meta, max_batch=262)
    assert ev.UpdateCheckPoint(prefix)
    for thp in ("1", "0"):
        os.environ["UGEVAL_SP_THP"] = thp
        for threads in (cpus, max(cpus - 1, 1)):
            st, _ = selfplay(ev, games=games, threads=threads, playouts_per_move=1600, max_moves=1, max_batch=262,
                             timeout_us=300, reuse_tree=True, prelude_moves=prelude, seed=5)
            keep = ("playouts_per_s", "mean_batch", "flush_timeouts", "batches", "full_batches", "wave_cut_batches",
                    "search_busy_fraction", "host_us_per_playout", "max_queue_depth", "gather_sleeps", "seconds")
            print(json.dumps({"cpus": cpus, "threads": threads, "thp": thp, **{k: round(st[k], 3) for k in keep}}), flush=True)
    ev.close()


if __name__ == "__main__":
    main()
