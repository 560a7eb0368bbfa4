"""Generates tests/golden/ref_board.npz from the REFERENCE's own GoBoard.  Run from the repo root IN THE BUILD
CONTAINER (needs /root/reference; `make -C oracle` compiles go/GoBoard.cpp and friends into oracle/_ref/libugref.so):

    python tests/golden/make_ref_board_golden.py

Every number in the file comes from oracle.RefBoard, i.e. from the reference's compiled board code: seeded random
games (moves drawn from the reference's IsLegal, 3 % passes), and at sampled move numbers the colours, side to move,
ko point, GetLastMove, the legal-move mask of both colours, the Tromp-Taylor score (GoBoardUtil::TrompTaylorScore,
komi 0 and 7.5), the move list of GenerateLegalMoves (simple eyes excluded, nothing after two passes), and the 17 planes produced by driving that board the way
CollectFeatures does (oracle/ref_board_shim.cc).  Games 0..IN_TURN-1 alternate colours like self-play; the rest mix in
out-of-turn moves and undos (board checks only — CollectFeatures replays undone moves with the side to move, so planes
after an out-of-turn move are not a meaningful contract).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
IN_TURN, MIXED = 20, 6
SAMPLE_EVERY = 9
UNDO = -2  # op code in the move list


def legal_mask(b, color):
    return np.array([b.is_legal(p, color) for p in range(362)], np.bool_)


def main():
    rng = np.random.default_rng(20261020)
    ops, op_color, op_game = [], [], []
    s_game, s_after, s_colors, s_to_play, s_ko, s_last, s_legal, s_planes, s_has_planes, s_score, s_gen = ([] for _ in range(11))

    def sample(g, n_ops, b, planes):
        s_game.append(g); s_after.append(n_ops)
        s_colors.append(b.colors()); s_to_play.append(b.to_play); s_ko.append(b.ko_point()); s_last.append(b.last_move())
        s_legal.append(np.stack([np.packbits(legal_mask(b, c)) for c in (0, 1)]))
        s_has_planes.append(planes)
        gen = np.zeros(362, np.bool_); gen[b.generate()] = True  # GenerateLegalMoves for the side to move
        s_gen.append(np.packbits(gen))
        s_score.append([b.score(0.0), b.score(7.5)])  # GoBoardUtil::TrompTaylorScore, komi 0 and 7.5
        s_planes.append(np.packbits(b.collect_features().reshape(-1)) if planes else np.zeros(768, np.uint8))

    for g in range(IN_TURN + MIXED):
        mixed = g >= IN_TURN
        length = [0, 1, 5, 7, 8, 9][g] if g < 6 else int(rng.integers(40, 400))
        b = oracle.RefBoard()
        n_ops = 0
        sample(g, 0, b, not mixed)
        for _ in range(length):
            if mixed and b.move_number > 0 and rng.random() < 0.05:
                b.undo()
                ops.append(UNDO); op_color.append(-1); op_game.append(g)
            else:
                col = b.to_play if (not mixed or rng.random() > 0.08) else int(rng.integers(0, 2))
                legal = np.nonzero(legal_mask(b, col)[:361])[0]
                mv = 361 if (len(legal) == 0 or rng.random() < 0.03) else int(rng.choice(legal))
                assert not b.play(mv, col)
                ops.append(mv); op_color.append(col); op_game.append(g)
            n_ops += 1
            if n_ops % SAMPLE_EVERY == 0 or n_ops == length or b.ko_point() >= 0:  # every ko position is sampled
                sample(g, n_ops, b, not mixed)
        if not mixed and g % 4 == 3:  # GetLastMove null rule: side to move forced to the colour that just played
            b.set_to_play(1 - b.to_play)
            ops.append(-3); op_color.append(b.to_play); op_game.append(g)  # op -3 = SetToPlay(op_color)
            sample(g, n_ops + 1, b, True)
    # a hand-made ko fight, sampled after every move: take, forbidden retake, threats elsewhere, retake, fill
    def pt(col, row):
        return (row - 1) * 19 + (col - 1)
    g = IN_TURN + MIXED
    b = oracle.RefBoard()
    sample(g, 0, b, True)
    fight = [pt(4, 4), pt(5, 4), pt(3, 3), pt(6, 3), pt(4, 2), pt(5, 2), pt(5, 3), pt(4, 3),   # White takes the ko
             pt(16, 16), pt(16, 15), pt(5, 3), pt(10, 10), pt(4, 3),                            # Black retakes, fills
             361, pt(3, 16), 361, 361]                                                        # passes; two in a row end it
    for n, mv in enumerate(fight):
        assert b.is_legal(mv), (n, mv)
        assert not b.play(mv)
        ops.append(mv); op_color.append(1 - b.to_play); op_game.append(g)
        sample(g, n + 1, b, True)
    np.savez_compressed(os.path.join(HERE, "ref_board.npz"),
                        ops=np.array(ops, np.int16), op_color=np.array(op_color, np.int8),
                        op_game=np.array(op_game, np.int16),
                        s_game=np.array(s_game, np.int16), s_after=np.array(s_after, np.int32),
                        s_colors=np.stack(s_colors), s_to_play=np.array(s_to_play, np.int8),
                        s_ko=np.array(s_ko, np.int16), s_last=np.array(s_last, np.int16),
                        s_legal=np.stack(s_legal), s_planes=np.stack(s_planes),
                        s_has_planes=np.array(s_has_planes, np.bool_), s_score=np.array(s_score, np.float32),
                        s_gen=np.stack(s_gen))
    print("ref_board.npz:", len(ops), "ops,", len(s_game), "samples,", int(np.sum(np.array(s_ko) >= 0)), "with a ko point,",
          os.path.getsize(os.path.join(HERE, "ref_board.npz")), "bytes")


def collect_features_golden():
    """tests/golden/ref_collect_features.npz: the planes of the reference's own CollectFeatures TEMPLATE
    (GoUctGlobalSearchState::CollectFeatures, gouct/GoUctGlobalSearch.h:160-206, compiled with the whole search stack into
    oracle/_ref/ref_search — `ref_search features`) after every move of the alternating games of ref_board.npz."""
    import subprocess
    import tempfile
    G = np.load(os.path.join(HERE, "ref_board.npz"))
    games = []
    for g in range(int(G["op_game"].max()) + 1):
        sel = np.nonzero(G["op_game"] == g)[0]
        ops, cols = G["ops"][sel].tolist(), G["op_color"][sel].tolist()
        if -3 in ops:
            cut = ops.index(-3)
            ops, cols = ops[:cut], cols[:cut]
        if UNDO in ops or cols != [k & 1 for k in range(len(ops))]:
            continue
        games.append(ops)
    d = tempfile.mkdtemp(prefix="refplanes_")
    with open(os.path.join(d, "games.txt"), "w") as f:
        f.write("".join(" ".join(map(str, ops)) + "\n" for ops in games))
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_search"), "features", "games.txt", "out.bin"], cwd=d,
                   check=True, capture_output=True)
    planes = np.fromfile(os.path.join(d, "out.bin"), np.int8).reshape(-1, 17 * 361)
    assert len(planes) == sum(len(ops) + 1 for ops in games) and set(np.unique(planes).tolist()) <= {0, 1}
    np.savez_compressed(os.path.join(HERE, "ref_collect_features.npz"),
                        moves=np.array([m for ops in games for m in ops], np.int16),
                        game_len=np.array([len(ops) for ops in games], np.int32),
                        planes=np.packbits(planes.astype(np.uint8), axis=1))
    print("ref_collect_features.npz:", len(games), "games,", len(planes), "positions,",
          os.path.getsize(os.path.join(HERE, "ref_collect_features.npz")), "bytes")


if __name__ == "__main__":
    if "--planes-only" not in sys.argv:
        main()
    collect_features_golden()
