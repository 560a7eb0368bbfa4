"""Pins the oracle (CPU restatement) before it is trusted as the checker: hand-derived known answers for the
encoder contract, the committed golden vectors, and the reference's own ArrayUtil
compiled into oracle/_ref."""
import json
import os

import numpy as np

import oracle

G = os.path.join(os.path.dirname(__file__), "golden")


def pt(col, row):
    return (row - 1) * 19 + (col - 1)


def test_empty_board_black_to_move():
    f = oracle.Board().collect_features()
    assert f.shape == (17, 19, 19) and f.dtype == np.int8
    assert (f[:16] == 0).all() and (f[16] == 1).all()  # plane 16 = 1 - cur, black = 0


def test_empty_board_white_to_move():
    b = oracle.Board()
    b.set_to_play(oracle.WHITE)
    f = b.collect_features()
    assert (f == 0).all()


def test_one_stone_and_history_repeat():
    b = oracle.Board()
    b.play(pt(4, 4))  # black D4; white to move
    f = b.collect_features()
    assert b.to_play == oracle.WHITE and (f[16] == 0).all()
    assert f[0].sum() == 0            # side to move (white) has no stones
    assert f[1].sum() == 1 and f[1][3][3] == 1   # opponent stone, index [row-1][col-1]
    # one move undone -> empty board, and the empty position repeats for the rest of the history
    assert (f[2:16] == 0).all()
    assert b.move_number == 1 and b.colors()[pt(4, 4)] == oracle.BLACK  # replayed


def test_cur_is_fixed_to_leaf_side_to_move():
    b = oracle.Board()
    b.play(pt(4, 4)); b.play(pt(16, 16)); b.play(pt(10, 10))  # B, W, B ; white to move
    f = b.collect_features()
    # k=0: white(cur)={Q16}, black(opp)={D4,K10}; k=1: cur {Q16} opp {D4}; k=2: cur {} opp {D4}; k>=3 empty
    assert f[0].sum() == 1 and f[1].sum() == 2
    assert f[2].sum() == 1 and f[3].sum() == 1
    assert f[4].sum() == 0 and f[5].sum() == 1
    assert (f[6:16] == 0).all()


def test_pass_consumes_a_history_slot():
    b = oracle.Board()
    b.play(pt(4, 4)); b.play(oracle.PASS)  # B D4, W pass; black to move
    f = b.collect_features()
    assert (f[16] == 1).all()
    assert f[0].sum() == 1 and f[2].sum() == 1  # same stones before/after the pass
    assert f[4].sum() == 0                       # after undoing D4


def test_capture_and_undo_restores_stone():
    b = oracle.Board()
    for mv in [pt(1, 1), pt(2, 1), pt(3, 1), pt(10, 10), pt(2, 2)]:
        b.play(mv)
    assert b.colors()[pt(2, 1)] == oracle.EMPTY  # captured
    f = b.collect_features()  # white to move
    assert f[0].sum() == 1 and f[1].sum() == 3   # leaf: W {K10}, B {A1,C1,B2}
    assert f[2].sum() == 2 and f[2][0][1] == 1   # one move earlier the white B1 stone is back
    assert b.colors()[pt(2, 1)] == oracle.EMPTY and b.move_number == 5


def test_simple_ko_is_illegal_then_legal():
    b = oracle.Board()
    for mv in [pt(4, 4), pt(5, 4), pt(3, 3), pt(6, 3), pt(4, 2), pt(5, 2), pt(5, 3), pt(4, 3)]:
        b.play(mv)
    assert b.colors()[pt(5, 3)] == oracle.EMPTY          # black stone captured in the ko
    assert not b.is_legal(pt(5, 3))                      # immediate recapture forbidden (SIMPLEKO)
    b.play(pt(16, 16)); b.play(pt(16, 4))
    assert b.is_legal(pt(5, 3))


def test_last_move_null_rule_stops_undo():
    b = oracle.Board()
    for mv in [pt(4, 4), pt(16, 16), pt(10, 10)]:
        b.play(mv)
    b.set_to_play(oracle.BLACK)  # last entry colour (black) != opp(to_play)=white -> GO_NULLMOVE
    assert b.last_move() == -1
    f = b.collect_features()
    for k in range(8):  # nothing undone: all eight steps show the same position
        assert np.array_equal(f[2 * k], f[0]) and np.array_equal(f[2 * k + 1], f[1])
    assert f[0].sum() == 2 and f[1].sum() == 1 and (f[16] == 1).all()


def test_golden_kat_file():
    z = np.load(os.path.join(G, "encoder_kat.npz"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    cases = mg.kat_cases()
    for i, name in enumerate(z["names"]):
        b = cases[str(name)]
        assert np.array_equal(b.collect_features(), z["planes"][i]), name
        assert np.array_equal(b.pack(), z["packed"][i]), name


def test_golden_positions_regenerate():
    z = np.load(os.path.join(G, "positions_64.npz"))
    packed, planes = oracle.synthetic_positions(64, seed=20260119)
    assert np.array_equal(packed, z["packed"]) and np.array_equal(planes, z["planes"])


def test_packed_masks_agree_with_planes():
    """ug_packed_pos (by colour) and the cur/opp planes describe the same histories"""
    z = np.load(os.path.join(G, "positions_64.npz"))
    packed, planes = z["packed"], z["planes"]
    for i in range(packed.shape[0]):
        tp = int(packed[i, 192])
        assert (planes[i, 16] == 1 - tp).all()
        for k in range(8):
            for colour in (0, 1):
                words = packed[i, colour * 96 + k * 12: colour * 96 + (k + 1) * 12]
                bits = ((words[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)[:361].astype(np.int8)
                plane = 2 * k + (0 if colour == tp else 1)
                assert np.array_equal(bits, planes[i, plane].reshape(-1))


def test_transform_feature_matches_reference_get_offset():
    cases = json.load(open(os.path.join(G, "get_offset.json")))
    for c in cases[:400]:
        dims, idx = c["dims"], c["idx"]
        flat = ((idx[0] * dims[1] + idx[1]) * dims[2] + idx[2]) * dims[3] + idx[3]
        assert flat == c["offset"]
    rng = np.random.default_rng(0)
    planes = rng.integers(0, 2, size=(3, 17, 19, 19)).astype(np.int8)
    hwc = oracle.transform_feature(planes, 0, True)
    assert np.array_equal(hwc, planes.transpose(0, 2, 3, 1))   # flat = ((b*19+row)*19+col)*17+plane
    assert np.array_equal(oracle.transform_feature(planes, 1, True), planes)
    assert np.array_equal(oracle.transform_feature(hwc, 1, True, hwc_input=True), planes)
    live = oracle.ref_lib()
    if live is not None:  # the compiled reference, when present in this container
        import ctypes as C
        for c in cases[:50]:
            assert live.ref_get_offset((C.c_int * 4)(*c["dims"]), 4, (C.c_int * 4)(*c["idx"])) == c["offset"]


def test_point2index_roundtrip():
    L = oracle.lib()
    seen = set()
    for row in range(1, 20):
        for col in range(1, 20):
            idx = L.og_point2index(20 * row + col)  # GoPointUtil::Pt
            assert idx == (row - 1) * 19 + (col - 1)
            seen.add(idx)
    go_pass = 19 * 19 + 3 * 20 + 1
    assert L.og_point2index(go_pass) == 361
    assert seen == set(range(361))


def test_value_transform_and_update_prior():
    v = np.array([0.25, -0.5, 1.0])
    assert np.array_equal(oracle.value_transform(v, 0), v)
    assert np.array_equal(oracle.value_transform(v, 1), -v)
    assert np.array_equal(oracle.value_transform(v, 2), 1 - v)
    p = np.linspace(0.001, 0.01, 362)
    ch = np.array([0, 5, 361])
    pr = oracle.update_prior(p, ch)
    assert abs(pr.sum() - 1) < 1e-12 and np.allclose(pr, p[ch] / p[ch].sum())
