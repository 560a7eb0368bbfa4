"""N>1 host logic on CPU: two gloo ranks shard games disjointly and reduce counters the way bench.py does
(sum of units, max of time); no collective touches the data path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unrealgo_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world_size, port, total_games, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        assert shard.world() == (world_size, rank)
        games = shard.owned_games(total_games, world_size, rank)
        seeds = [shard.game_seed(7, g) for g in games]
        # pretend each game produced (index + 1) playouts and rank r took (1 + r) seconds
        units, secs = shard.reduce_throughput(sum(g + 1 for g in games), 1.0 + rank)
        gathered = [None] * world_size
        dist.all_gather_object(gathered, (games, seeds))
        if rank == 0:
            out.put((units, secs, gathered))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size,total_games", [(2, 11), (2, 512), (3, 10)])
def test_games_shard_disjointly_and_counters_reduce(world_size, total_games):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world_size, port, total_games, out)) for r in range(world_size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    units, secs, gathered = out.get()
    all_games = sorted(g for games, _ in gathered for g in games)
    assert all_games == list(range(total_games))                       # every game owned exactly once
    assert units == sum(range(1, total_games + 1)) and secs == float(world_size)  # SUM of units, MAX of time
    all_seeds = [s for _, seeds in gathered for s in seeds]
    assert len(set(all_seeds)) == total_games                          # seeds are per game, not per rank
    # a game's seed does not depend on the world size
    assert shard.game_seed(7, 5) == [s for games, seeds in gathered for g, s in zip(games, seeds) if g == 5][0]


def test_single_process_is_a_world_of_one():
    assert shard.world() == (1, 0)
    assert shard.reduce_throughput(10, 2.5) == (10.0, 2.5)
    assert shard.owned_games(5, 1, 0) == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        shard.owned_games(5, 2, 2)
