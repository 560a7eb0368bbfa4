"""The self-play driver's host logic on the CPU (no GPU): csrc/selfplay.cpp linked against a stub leaf queue whose
answers depend only on the submitted position (tests/cpp/selfplay_stub.cpp)."""
import os
import subprocess

import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("spstub") / "selfplay_stub")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I/usr/local/cuda/include",
                    os.path.join(ROOT, "tests", "cpp", "selfplay_stub.cpp"),
                    os.path.join(ROOT, "unrealgo_b200", "csrc", "selfplay.cpp"), "-o", out],
                   check=True, capture_output=True)
    return out


def _run(exe, games, threads, playouts, moves, reuse, leaves, *more):
    r = subprocess.run([exe] + [str(x) for x in (games, threads, playouts, moves, reuse, leaves) + more],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.splitlines()
    t = lines[0].split()
    stats = {t[i]: int(t[i + 1]) for i in range(0, len(t), 2)}
    games_moves = [[int(x) for x in ln.split(":")[1].split()] for ln in lines[1:]]
    return stats, games_moves


@pytest.mark.parametrize("reuse", [1, 0])
def test_search_is_a_function_of_the_seed_not_of_the_threads(exe, reuse):
    s1, g1 = _run(exe, 12, 1, 120, 10, reuse, 1)
    s4, g4 = _run(exe, 12, 4, 120, 10, reuse, 1)
    assert g1 == g4                                   # one leaf per game: sequential PUCT per game, any threading
    assert s1["playouts"] == s4["playouts"] == 12 * 10 * 120 and s1["moves"] == 12 * 10 and s1["games_finished"] == 12
    for g in g1:                                      # legality judged by the oracle's board
        b = oracle.Board()
        for mv in g:
            assert b.is_legal(mv)
            b.play(mv)
    assert len({tuple(g) for g in g1}) > 1


def test_tree_reuse_changes_visit_counts_not_legality(exe):
    _, with_reuse = _run(exe, 6, 2, 150, 8, 1, 1)
    _, without = _run(exe, 6, 2, 150, 8, 0, 1)
    assert with_reuse != without                      # carried-over statistics shift some choices
    assert [g[:2] for g in with_reuse] == [g[:2] for g in without]  # the two random opening moves come from the seed


def test_virtual_loss_leaves_keep_the_books(exe):
    # the stub answers in submission order, so even the multi-leaf search is reproducible here
    s, g = _run(exe, 3, 1, 200, 6, 1, 16)
    s2, g2 = _run(exe, 3, 1, 200, 6, 1, 16)
    assert s["playouts"] == 3 * 6 * 200 and s["moves"] == 18 and g == g2
    for game in g:
        b = oracle.Board()
        for mv in game:
            assert b.is_legal(mv)
            b.play(mv)


def test_prelude_moves_start_the_search_in_the_middle_game(exe):
    """the benchmark's mid-game sample: 150 random legal moves, then 3 searched moves per game"""
    s1, g1 = _run(exe, 8, 1, 60, 3, 1, 1, 0, 0, 150)
    s3, g3 = _run(exe, 8, 3, 60, 3, 1, 1, 0, 0, 150)
    assert g1 == g3 and s1["playouts"] == 8 * 3 * 60 and s1["moves"] == 8 * 3
    for g in g1:
        assert len(g) == 153
        b = oracle.Board()
        for mv in g:
            assert b.is_legal(mv)
            b.play(mv)
        assert sum(1 for mv in g[:150] if mv == 361) == 0   # no pass while other moves exist
    assert len({tuple(g[:150]) for g in g1}) == 8
