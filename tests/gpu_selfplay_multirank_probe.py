"""Self-play on every GPU of a box at once (torchrun, one rank per GPU, each pinned to its slice of the host CPUs, as
bench.py does): per-rank playouts/s, device busy fraction, mean batch and host cost for a few variants of
games x search threads x batch cap — what limits an 8-GPU box with 4 CPUs per GPU?
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tests/gpu_selfplay_multirank_probe.py \\
        [prelude=80] [variants: games:threads:max_batch ...]"""
import json
import os
import sys
import tempfile

sys.path.insert(0, ".")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from unrealgo_b200 import DlTFNetworkEvaluator, selfplay, tfformat  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    prelude = int(sys.argv[1]) if len(sys.argv) > 1 else 80
    variants = sys.argv[2:] or ["768:4:262", "768:3:262", "1024:4:262", "1536:4:262"]
    torch.cuda.set_device(local)
    cpus = sorted(os.sched_getaffinity(0))
    per = max(len(cpus) // world, 1)
    os.sched_setaffinity(0, cpus[local * per:(local + 1) * per])
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = tempfile.mkdtemp(prefix=f"ug_mr_{rank}_")
    meta, prefix = tfformat.write_synthetic_checkpoint(d, 20, 256, seed=20260119)
    ev = DlTFNetworkEvaluator(meta, device=local, max_batch=262)
    assert ev.UpdateCheckPoint(prefix)
    for v in variants:
        games, threads, mb = (int(x) for x in v.split(":"))
        if world > 1:
            dist.barrier()
        st, _ = selfplay(ev, games=games, threads=threads, playouts_per_move=1600, max_moves=1, max_batch=mb,
                         timeout_us=300, reuse_tree=True, prelude_moves=prelude, seed=1000 + rank)
        row = [st["playouts_per_s"], st["gpu_busy_fraction"], st["mean_batch"], st["search_busy_fraction"],
               st["host_us_per_playout"], float(st["flush_timeouts"]), st["seconds"], float(st["playouts"]),
               float(st["idle_launches"]), st["idle_gap_seconds"], float(st["batches"])]
        t = torch.tensor(row, dtype=torch.float64, device="cuda")
        rows = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(rows, t)
        else:
            rows = [t]
        if rank == 0:
            rows = [r.tolist() for r in rows]
            total = sum(r[7] for r in rows) / max(r[6] for r in rows)
            print(json.dumps({"variant": v, "prelude": prelude, "whole_job_playouts_per_s": round(total),
                              "per_rank_playouts_per_s": [round(r[0]) for r in rows],
                              "gpu_busy": [round(r[1], 3) for r in rows], "mean_batch": [round(r[2], 1) for r in rows],
                              "search_busy": [round(r[3], 2) for r in rows],
                              "host_us_per_playout": [round(r[4], 1) for r in rows],
                              "flush_timeouts": [int(r[5]) for r in rows], "batches": [int(r[10]) for r in rows],
                              "idle_launches": [int(r[8]) for r in rows],
                              "idle_gap_s": [round(r[9], 2) for r in rows], "seconds": [round(r[6], 1) for r in rows]}), flush=True)
    ev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
