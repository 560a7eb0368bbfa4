"""Host side of the path, no GPU: the self-play driver's incremental board (csrc/goboard.h) against the
oracle's restatement of GoBoard + CollectFeatures' undo/replay walk.  Integer work => bit-exact."""
import numpy as np

import oracle
from unrealgo_b200 import ProductBoard


def _random_game(seed, length, pass_prob):
    rng = np.random.default_rng(seed)
    ob, pb = oracle.Board(), ProductBoard()
    assert np.array_equal(pb.pack(), ob.pack())
    for step in range(length):
        # legality agrees on every point (simple ko, suicide, captures)
        if step % 7 == 0:
            for p in range(361):
                assert pb.legal(p) == ob.is_legal(p), (seed, step, p)
        if rng.random() < pass_prob:
            mv = 361
        else:
            mv = 361
            for _ in range(60):
                p = int(rng.integers(0, 361))
                if ob.is_legal(p):
                    mv = p
                    break
        ob.play(mv)
        pb.play(mv)
        assert pb.to_play == ob.to_play
        # the ring of packed positions == the masks the reference's undo/replay walk would visit
        assert np.array_equal(pb.pack(), ob.pack()), (seed, step)
    return ob, pb


def test_incremental_history_matches_undo_walk():
    for seed, length, pp in [(1, 40, 0.0), (2, 300, 0.02), (3, 450, 0.05), (4, 12, 0.3), (5, 600, 0.01)]:
        _random_game(seed, length, pp)


def test_generate_moves_excludes_eyes_and_illegal():
    ob, pb = _random_game(9, 260, 0.0)
    moves, mask = pb.generate()
    assert moves[-1] == 361 and len(set(moves.tolist())) == len(moves)
    for m in moves[:-1]:
        assert ob.is_legal(int(m))
        assert (mask[m >> 5] >> (m & 31)) & 1
    legal_pts = {p for p in range(361) if ob.is_legal(p)}
    dropped = legal_pts - set(int(m) for m in moves[:-1])
    colors = ob.colors().reshape(19, 19)
    me = ob.to_play
    for p in dropped:  # every legal point that was not generated is a simple eye: all neighbours are own stones
        y, x = divmod(p, 19)
        for yy, xx in ((y - 1, x), (y + 1, x), (y, x - 1), (y, x + 1)):
            if 0 <= yy < 19 and 0 <= xx < 19:
                assert colors[yy, xx] == me


def test_two_passes_end_move_generation():
    pb = ProductBoard()
    pb.play(361)
    assert len(pb.generate()[0]) == 362 - 0  # empty board: every point + pass
    pb.play(361)
    assert len(pb.generate()[0]) == 0


def _tromp_taylor(colors, komi):
    """independent restatement of GoBoardUtil::TrompTaylorScore (go/GoBoardUtil.h:626-680) on a 19x19 colour array"""
    c = colors.reshape(19, 19)
    score = -komi + float((c == 0).sum()) - float((c == 1).sum())
    seen = np.zeros((19, 19), bool)
    for y in range(19):
        for x in range(19):
            if c[y, x] != 2 or seen[y, x]:
                continue
            stack, size, adj = [(y, x)], 0, set()
            seen[y, x] = True
            while stack:
                yy, xx = stack.pop()
                size += 1
                for ny, nx in ((yy - 1, xx), (yy + 1, xx), (yy, xx - 1), (yy, xx + 1)):
                    if 0 <= ny < 19 and 0 <= nx < 19:
                        if c[ny, nx] == 2:
                            if not seen[ny, nx]:
                                seen[ny, nx] = True
                                stack.append((ny, nx))
                        else:
                            adj.add(int(c[ny, nx]))
            if adj == {0}:
                score += size
            elif adj == {1}:
                score -= size
    return score


def test_score_and_host_planes_match_the_oracle():
    assert ProductBoard().score(7.5) == -7.5  # empty board: one region touching neither colour
    for seed, length, pp in [(21, 30, 0.0), (22, 250, 0.02), (23, 500, 0.01), (24, 5, 0.4)]:
        ob, pb = _random_game(seed, length, pp)
        assert pb.score(7.5) == _tromp_taylor(ob.colors(), 7.5)
        assert pb.score(0.5) == _tromp_taylor(ob.colors(), 0.5)
        # the planes written to self-play records == CollectFeatures' undo/replay walk
        assert np.array_equal(pb.planes(), ob.collect_features())


def test_pack_planes_is_the_inverse_of_collect_features():
    """ug_pack_planes: from CollectFeatures' char[17][19][19] back to the 784-byte packed leaf the queue takes — checked
    on the planes of the reference's own CollectFeatures template (tests/golden/ref_collect_features.npz) against the
    packed history the self-play driver's board keeps for the same positions"""
    import os
    from unrealgo_b200 import pack_planes
    F = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_collect_features.npz"))
    planes = np.unpackbits(F["planes"], axis=1)[:, :17 * 361].reshape(-1, 17, 19, 19).astype(np.int8)
    got = pack_planes(planes)
    want, at = [], 0
    for n in F["game_len"].tolist():
        pb = ProductBoard()
        want.append(pb.pack())
        for mv in F["moves"][at:at + n].tolist():
            pb.play(mv)
            want.append(pb.pack())
        at += n
    assert np.array_equal(got, np.stack(want))
    # any non-zero byte is a stone, like (bool)char in TransformFeature
    assert np.array_equal(pack_planes(planes[100:110] * 5), got[100:110])
