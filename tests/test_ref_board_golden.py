"""tests/golden/ref_board.npz holds answers of the REFERENCE's own GoBoard (go/GoBoard.cpp compiled by oracle/Makefile;
generator: tests/golden/make_ref_board_golden.py).  It pins, bit for bit:
  * the oracle's board restatement and its CollectFeatures planes,
  * the product's host board (csrc/goboard.h) — colours via the planes, legal moves, generated move lists, Tromp-Taylor score,
and, where oracle/_ref/libugref.so exists (the build container), the live reference board against the file."""
import os

import numpy as np
import pytest

import oracle
from unrealgo_b200 import ProductBoard

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_board.npz"))
UNDO, SET_TO_PLAY = -2, -3
N_GAMES = int(G["op_game"].max()) + 1


def _samples_of(g):
    return {int(G["s_after"][i]): i for i in np.nonzero(G["s_game"] == g)[0]}


def _ops_of(g):
    sel = np.nonzero(G["op_game"] == g)[0]
    return [(int(G["ops"][i]), int(G["op_color"][i])) for i in sel]


def _apply(b, op, col):
    if op == UNDO:
        b.undo()
    elif op == SET_TO_PLAY:
        b.set_to_play(col)
    else:
        b.play(op, col)


def _check_board(b, i, what):
    assert np.array_equal(b.colors(), G["s_colors"][i]), what
    assert b.to_play == G["s_to_play"][i] and b.last_move() == G["s_last"][i], what
    for c in (0, 1):
        want = np.unpackbits(G["s_legal"][i, c])[:362].astype(bool)
        got = np.array([b.is_legal(p, c) for p in range(362)])
        assert np.array_equal(got, want), (what, c, np.nonzero(got != want)[0])
    if G["s_has_planes"][i]:
        want = np.unpackbits(G["s_planes"][i])[:17 * 361].reshape(17, 19, 19).astype(np.int8)
        assert np.array_equal(b.collect_features(), want), what
    if hasattr(b, "score"):
        assert b.score(0.0) == G["s_score"][i, 0] and b.score(7.5) == G["s_score"][i, 1], what


def _replay(make_board):
    checked = 0
    for g in range(N_GAMES):
        b = make_board()
        samples = _samples_of(g)
        if 0 in samples:
            _check_board(b, samples[0], (g, 0))
            checked += 1
        for n, (op, col) in enumerate(_ops_of(g), start=1):
            _apply(b, op, col)
            if n in samples:
                _check_board(b, samples[n], (g, n))
                checked += 1
    return checked


def test_fixture_shape():
    assert N_GAMES >= 27 and len(G["s_game"]) > 500 and (G["s_ko"] >= 0).sum() >= 5
    assert (G["ops"] == UNDO).sum() > 20 and (G["ops"] == SET_TO_PLAY).sum() >= 4 and (G["ops"] == 361).sum() > 50
    # out-of-turn moves exist in the mixed games only
    assert G["s_has_planes"].sum() > 350


def test_oracle_board_and_planes_match_the_reference_board():
    assert _replay(oracle.Board) == len(G["s_game"])


def test_live_reference_board_reproduces_the_fixture():
    try:
        oracle.RefBoard()
    except RuntimeError:
        pytest.skip("oracle/_ref/libugref.so not built here (needs the reference tree)")
    assert _replay(oracle.RefBoard) == len(G["s_game"])


def test_simple_eye_rule_of_the_fixture_is_the_two_block_rule():
    """the generated lists drop exactly the legal points whose neighbours are all own stones of at most two blocks
    (go/GoEyeUtil.h:133-176); a weaker 'all neighbours own' rule would drop more — make sure the file can tell"""
    dropped = weaker = 0
    for i in range(len(G["s_game"])):
        tp = int(G["s_to_play"][i])
        legal = np.unpackbits(G["s_legal"][i, tp])[:361].astype(bool)
        gen = np.unpackbits(G["s_gen"][i])[:361].astype(bool)
        if not gen.any() and not np.unpackbits(G["s_gen"][i])[361]:
            continue  # two passes: nothing generated at all
        assert not (gen & ~legal).any()
        c = np.full((21, 21), 3, np.int8)
        c[1:20, 1:20] = G["s_colors"][i].reshape(19, 19)
        nb = np.stack([c[:-2, 1:-1], c[2:, 1:-1], c[1:-1, :-2], c[1:-1, 2:]])
        all_own = ((nb == tp) | (nb == 3)).all(0).reshape(-1)
        assert not (legal & ~gen & ~all_own).any()      # whatever was dropped has only own stones around it
        dropped += int((legal & ~gen).sum())
        weaker += int((legal & all_own).sum())
    assert dropped > 100 and weaker > dropped + 10      # some all-own points are NOT eyes and must be generated


def test_product_host_board_matches_the_reference_board():
    """csrc/goboard.h (the self-play driver's incremental board) on the games that alternate colours, which is all it
    supports: legal moves for the side to move, planes and side to move at every sampled position"""
    checked = games = 0
    for g in range(N_GAMES):
        ops = _ops_of(g)
        if SET_TO_PLAY in [op for op, _ in ops]:
            ops = ops[:[op for op, _ in ops].index(SET_TO_PLAY)]
        colours = [col for _, col in ops]
        if UNDO in [op for op, _ in ops] or colours != [k & 1 for k in range(len(ops))]:
            continue  # a mixed game
        games += 1
        pb = ProductBoard()
        samples = _samples_of(g)
        for n, (op, _) in enumerate([(None, None)] + ops):
            if op is not None:
                pb.play(op)
            if n in samples:
                i = samples[n]
                assert pb.to_play == G["s_to_play"][i]
                want = np.unpackbits(G["s_legal"][i, pb.to_play])[:362].astype(bool)
                got = np.array([pb.legal(p) for p in range(362)])
                assert np.array_equal(got, want), (g, n, np.nonzero(got != want)[0])
                if G["s_has_planes"][i]:
                    planes = np.unpackbits(G["s_planes"][i])[:17 * 361].reshape(17, 19, 19).astype(np.int8)
                    assert np.array_equal(pb.planes(), planes), (g, n)
                moves, mask = pb.generate()   # GenerateLegalMoves: eyes of the side to move dropped, empty after two passes
                want_gen = np.nonzero(np.unpackbits(G["s_gen"][i])[:362])[0]
                assert np.array_equal(moves.astype(np.int64), want_gen), (g, n)
                assert np.array_equal(np.nonzero(np.unpackbits(mask.view(np.uint8), bitorder="little")[:362])[0], want_gen)
                assert pb.score(0.0) == G["s_score"][i, 0] and pb.score(7.5) == G["s_score"][i, 1], (g, n)
                checked += 1
    assert games >= 21 and checked > 350


def _template_fixture():
    F = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_collect_features.npz"))
    planes = np.unpackbits(F["planes"], axis=1)[:, :17 * 361].reshape(-1, 17, 19, 19).astype(np.int8)
    games, at = [], 0
    for n in F["game_len"].tolist():
        games.append(F["moves"][at:at + n].tolist())
        at += n
    return games, planes


def test_collect_features_template_golden():
    """tests/golden/ref_collect_features.npz: planes written by the reference's own CollectFeatures template (generator:
    make_ref_board_golden.py, `ref_search features`) after every move of 21 games, 2966 positions.  The oracle's
    restatement and the product's host board reproduce every one."""
    games, planes = _template_fixture()
    assert len(planes) == sum(len(g) + 1 for g in games) == 2966
    at = 0
    for g in games:
        ob, pb = oracle.Board(), ProductBoard()
        for n in range(len(g) + 1):
            if n > 0:
                ob.play(g[n - 1])
                pb.play(g[n - 1])
            assert np.array_equal(ob.collect_features(), planes[at + n]), (at, n)
            assert np.array_equal(pb.planes(), planes[at + n]), (at, n)
        at += len(g) + 1


REF_SEARCH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_search")


@pytest.mark.skipif(not os.path.exists(REF_SEARCH), reason="oracle/_ref/ref_search not built (needs the reference tree)")
def test_the_real_collect_features_template_on_every_position(tmp_path):
    """`ref_search features` calls the reference's own GoUctGlobalSearchState::CollectFeatures (the template, compiled
    with the whole search stack) on the search thread's state after every move of the alternating games: the fixture's
    planes (made by oracle/ref_board_shim.cc driving the real board), the oracle's restatement and the product's host
    board must all equal it — at every position, not only the sampled ones"""
    import subprocess
    games = []
    for g in range(N_GAMES):
        ops = _ops_of(g)
        if SET_TO_PLAY in [op for op, _ in ops]:
            ops = ops[:[op for op, _ in ops].index(SET_TO_PLAY)]
        if UNDO in [op for op, _ in ops] or [c for _, c in ops] != [k & 1 for k in range(len(ops))]:
            continue
        games.append((g, [op for op, _ in ops]))
    (tmp_path / "games.txt").write_text("".join(" ".join(map(str, ops)) + "\n" for _, ops in games))
    r = subprocess.run([REF_SEARCH, "features", "games.txt", "out.bin"], cwd=str(tmp_path), capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    real = np.fromfile(str(tmp_path / "out.bin"), np.int8).reshape(-1, 17, 19, 19)
    assert len(real) == sum(len(ops) + 1 for _, ops in games) > 2500
    at = sampled = 0
    for g, ops in games:
        samples = _samples_of(g)
        ob, pb = oracle.Board(), ProductBoard()
        for n in range(len(ops) + 1):
            if n > 0:
                ob.play(ops[n - 1])
                pb.play(ops[n - 1])
            assert np.array_equal(ob.collect_features(), real[at + n]), (g, n)
            assert np.array_equal(pb.planes(), real[at + n]), (g, n)
            if n in samples and G["s_has_planes"][samples[n]]:
                want = np.unpackbits(G["s_planes"][samples[n]])[:17 * 361].reshape(17, 19, 19).astype(np.int8)
                assert np.array_equal(want, real[at + n]), (g, n)
                sampled += 1
        at += len(ops) + 1
    assert sampled > 350


def test_live_fuzz_product_board_against_the_reference_board():
    """where libugref.so exists: 30 more random games to the end (moves drawn from the reference's GenerateLegalMoves
    list, so eyes are never filled and games finish by two passes) — after every move the product board's move list
    equals the reference's, every fourth move also planes, Tromp-Taylor score and the oracle's planes"""
    try:
        oracle.RefBoard()
    except RuntimeError:
        pytest.skip("oracle/_ref/libugref.so not built here (needs the reference tree)")
    rng = np.random.default_rng(77)
    positions = finished = 0
    for g in range(30):
        rb, pb, ob = oracle.RefBoard(), ProductBoard(), oracle.Board()
        for n in range(500):
            want = rb.generate()
            got, _ = pb.generate()
            assert np.array_equal(got.astype(np.int64), want.astype(np.int64)), (g, n)
            if len(want) == 0:
                finished += 1
                break
            if n % 4 == 0:
                planes = rb.collect_features()
                assert np.array_equal(pb.planes(), planes) and np.array_equal(ob.collect_features(), planes), (g, n)
                assert pb.score(7.5) == rb.score(7.5), (g, n)
            # mostly a random non-pass move; the pass (last entry) when it is the only one or now and then
            mv = int(want[-1]) if len(want) == 1 or rng.random() < 0.01 else int(rng.choice(want[:-1]))
            assert not rb.play(mv)
            pb.play(mv)
            ob.play(mv)
            positions += 1
    assert positions > 8000 and finished >= 25
