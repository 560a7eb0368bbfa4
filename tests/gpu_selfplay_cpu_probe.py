"""Self-play on one GPU with the host budget of an 8-GPU box (4 CPUs per rank): where does the host side stand, and how
much concurrency (games x leaves in flight against batch size) does the pipeline need?
    python tests/gpu_selfplay_cpu_probe.py [cpus=4] [prelude=80] [variants: games:threads:max_batch[:timeout_us[:moves]] ...]
One leaf per game is in flight, so throughput <= games / cycle time (queue wait + two batches ahead + own batch + host
processing): Little's law, not a saturated resource, caps the rate when the cycle time grows."""
import json
import os
import sys
import tempfile

sys.path.insert(0, ".")
from unrealgo_b200 import DlTFNetworkEvaluator, selfplay, tfformat  # noqa: E402


def main():
    cpus = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    prelude = int(sys.argv[2]) if len(sys.argv) > 2 else 80
    variants = sys.argv[3:] or ["768:4:262", "768:4:209", "768:4:157", "1024:4:262", "768:3:209"]
    os.sched_setaffinity(0, set(sorted(os.sched_getaffinity(0))[:cpus]))
    d = tempfile.mkdtemp(prefix="ug_spprobe_")
    meta, prefix = tfformat.write_synthetic_checkpoint(d, 20, 256, seed=20260119)
    ev = DlTFNetworkEvaluator(meta, max_batch=262)
    assert ev.UpdateCheckPoint(prefix)
    for v in variants:
        f = [int(x) for x in v.split(":")]
        games, threads, mb = f[0], f[1], f[2]
        to = f[3] if len(f) > 3 else 300
        moves = f[4] if len(f) > 4 else 1
        st, _ = selfplay(ev, games=games, threads=threads, playouts_per_move=1600, max_moves=moves, max_batch=mb,
                         timeout_us=to, reuse_tree=True, prelude_moves=prelude, seed=5)
        keep = ("playouts_per_s", "mean_batch", "flush_timeouts", "batches", "full_batches", "wave_cut_batches",
                "search_busy_fraction", "gpu_busy_fraction", "host_us_per_playout", "max_queue_depth", "gather_sleeps", "seconds")
        print(json.dumps({"cpus": cpus, "games": games, "threads": threads, "max_batch": mb, "timeout_us": to,
                          **{k: round(st[k], 3) for k in keep}}), flush=True)
    ev.close()


if __name__ == "__main__":
    main()
