"""Multi-GPU plumbing of the evaluation path: games/positions shard by rank with no data-path collective
(a playout touches one game's tree and one network evaluation; the only cross-rank traffic is the
barrier and the reduction of the counters that bench.py reports).  Backend-agnostic: NCCL on the GPU box, gloo in
the CPU tests."""
import torch
import torch.distributed as dist


def world():
    """(world_size, rank) of the default process group, (1, 0) when torch.distributed is not initialised"""
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def owned_games(total_games, world_size, rank):
    """GPU g owns games {g, g+G, g+2G, ...}; returns the list of global game indices"""
    if not (0 <= rank < world_size):
        raise ValueError("rank outside the world")
    return list(range(rank, total_games, world_size))


def game_seed(base_seed, global_game):
    """per-game RNG seed that does not depend on how many GPUs share the job"""
    return (int(base_seed) * 1000003 + int(global_game)) & 0x7FFFFFFFFFFFFFFF


def reduce_throughput(units, seconds, device="cpu"):
    """whole-job throughput pieces: (sum of units over ranks, max of seconds over ranks).  Every rank must call."""
    ws, _ = world()
    if ws == 1:
        return float(units), float(seconds)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    s = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dist.all_reduce(s, op=dist.ReduceOp.MAX)
    return float(u.item()), float(s.item())
