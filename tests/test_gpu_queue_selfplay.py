"""Q1 (leaf-batching queue) and the self-play driver on the GPU.

Queue property the reference violates (SURVEY.md 3.1 defect 1, A.6.7): every submitted leaf gets exactly its
own result — under many producer threads, random arrival, flush-on-timeout and a checkpoint swap mid-stream.
Self-play: per-game search is sequential PUCT, so move sequences are a pure function of the weights — they must
not depend on how many games/threads share the GPU (batch composition), which checks queue + evaluator +
driver end to end."""
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ev(ckpt, max_batch=128):
    from unrealgo_b200 import DlTFNetworkEvaluator
    ev = DlTFNetworkEvaluator(ckpt[0], max_batch=max_batch)
    assert ev.UpdateCheckPoint(ckpt[1])
    return ev


def test_queue_every_leaf_gets_its_own_result(small_ckpt):
    import oracle
    from unrealgo_b200 import LeafQueue
    packed, _ = oracle.synthetic_positions(400, seed=5)
    ev = _ev(small_ckpt)
    want_p, want_v = ev.evaluate_packed(packed[:128])
    for lo in range(128, 400, 128):
        p, v = ev.evaluate_packed(packed[lo:lo + 128])
        want_p, want_v = np.concatenate([want_p, p]), np.concatenate([want_v, v])
    q = LeafQueue(ev, max_batch=64, timeout_us=300)
    errors = []
    n_threads = 32

    def producer(tid):
        rng = np.random.default_rng(tid)
        try:
            for rep in range(40):
                idx = int(rng.integers(0, 400))
                if rng.random() < 0.3:
                    time.sleep(float(rng.random()) * 2e-3)  # random arrival -> partial batches, timeouts
                t = q.submit(packed[idx])
                pol, val = q.wait(t)
                if not (np.array_equal(pol, want_p[idx]) and val == want_v[idx]):
                    errors.append((tid, rep, idx))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=producer, args=(i,)) for i in range(n_threads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    st = q.stats()
    assert not errors, errors[:5]
    assert st["leaves"] == n_threads * 40 and st["batches"] >= st["leaves"] / 64
    q.close()
    ev.close()


def test_queue_checkpoint_swap_midstream(small_ckpt, small_ckpt_b):
    import oracle
    from unrealgo_b200 import LeafQueue
    packed, _ = oracle.synthetic_positions(64, seed=6)
    ev_a, ev_b = _ev(small_ckpt), _ev(small_ckpt_b)
    pa, va = ev_a.evaluate_packed(packed)
    pb, vb = ev_b.evaluate_packed(packed)
    ev_b.close()
    q = LeafQueue(ev_a, max_batch=32, timeout_us=200)
    stop = threading.Event()
    bad = []

    def producer(tid):
        rng = np.random.default_rng(100 + tid)
        while not stop.is_set():
            idx = int(rng.integers(0, 64))
            pol, val = q.wait(q.submit(packed[idx]))
            is_a = np.array_equal(pol, pa[idx]) and val == va[idx]
            is_b = np.array_equal(pol, pb[idx]) and val == vb[idx]
            if not (is_a or is_b):  # never a mixture or a stale neighbour's result
                bad.append((tid, idx))

    threads = [threading.Thread(target=producer, args=(i,)) for i in range(8)]
    for t in threads:
        t.start()
    time.sleep(0.2)
    assert q.swap_checkpoint(small_ckpt_b[1])
    time.sleep(0.2)
    stop.set()
    for t in threads:
        t.join()
    pol, val = q.wait(q.submit(packed[3]))
    assert np.array_equal(pol, pb[3]) and val == vb[3]  # the new weights are live
    assert not bad, bad[:5]
    q.close()
    ev_a.close()


def test_selfplay_is_deterministic_and_batch_independent(small_ckpt):
    from unrealgo_b200 import selfplay
    ev = _ev(small_ckpt, max_batch=64)
    kw = dict(playouts_per_move=24, max_moves=14, random_opening_moves=3, seed=7, record_moves=True, max_batch=64)
    st1, g1 = selfplay(ev, games=6, threads=1, **kw)
    st2, g2 = selfplay(ev, games=24, threads=4, **kw)
    assert g2[:6] == g1                       # same seeds -> same games, whatever shares the batch
    assert len({tuple(g) for g in g2}) > 1    # the random openings do diversify
    assert st1["moves"] == 6 * 14 and st2["moves"] == 24 * 14
    assert st2["playouts"] == st2["evaluations"] + st2["terminal_leaves"] == 24 * 14 * 24
    assert st2["mean_batch"] > st1["mean_batch"]
    # without tree reuse each move searches from scratch: still deterministic, generally different visit counts
    st3, g3 = selfplay(ev, games=6, threads=2, reuse_tree=False, **kw)
    st4, g4 = selfplay(ev, games=6, threads=1, reuse_tree=False, **kw)
    assert g3 == g4
    ev.close()


def test_selfplay_moves_are_legal_and_follow_visit_counts(small_ckpt):
    import oracle
    from unrealgo_b200 import selfplay
    ev = _ev(small_ckpt, max_batch=64)
    _, games = selfplay(ev, games=4, threads=2, playouts_per_move=16, max_moves=40, random_opening_moves=6, seed=3,
                        record_moves=True, max_batch=64)
    for g in games:
        b = oracle.Board()
        for mv in g:
            assert b.is_legal(mv), (g, mv)   # legality judged by the oracle's board, not the driver's
            b.play(mv)
    ev.close()


def test_selfplay_records_follow_write_tfrecord(small_ckpt, tmp_path):
    """Recorded games (UctDeepPlayer::WriteTFRecord, search/UctDeepPlayer.cpp:362-388): one example per move up to
    the last non-pass move, newest first; planes = CollectFeatures of the position after the move (oracle board),
    policy = root visit counts of the search that chose it, reward = +-1 alternating, sign from Tromp-Taylor."""
    import oracle
    from tests.test_board_host import _tromp_taylor
    from tests.test_tfrecord import _example_class, read_records
    from unrealgo_b200 import selfplay
    Example = _example_class()
    ev = _ev(small_ckpt, max_batch=64)
    playouts = 20
    st, games = selfplay(ev, games=5, threads=2, playouts_per_move=playouts, max_moves=17, random_opening_moves=2,
                         seed=11, record_moves=True, max_batch=64, reuse_tree=False, record_dir=str(tmp_path),
                         game_index_base=40)
    ev.close()
    assert st["games_finished"] == 5
    for gi, moves in enumerate(games):
        recs = read_records(str(tmp_path / f"selfplay_{40 + gi}.tfrecord"))
        last = len(moves) - 1
        while last >= 0 and moves[last] == 361:
            last -= 1
        assert len(recs) == last + 1
        b = oracle.Board()
        planes_after, colors_after = [], None
        for mv in moves[:last + 1]:
            b.play(mv)
            planes_after.append(b.collect_features())
        colors_after = b.colors()
        score = _tromp_taylor(colors_after, 7.5)
        mover_is_black = last % 2 == 0
        reward = 1.0 if (score > 0) == mover_is_black else -1.0
        for k, data in enumerate(recs):           # record k describes move index last - k
            i = last - k
            ex = Example()
            ex.ParseFromString(data)
            f = ex.features.feature
            got = np.frombuffer(f["feature"].bytes_list.value[0], np.int8).reshape(17, 19, 19)
            assert np.array_equal(got, planes_after[i]), (gi, i)
            pol = np.array(f["policy"].float_list.value, np.float32)
            assert pol.shape == (362,) and (pol >= 0).all() and np.all(pol == np.round(pol))   # raw visit counts
            # without tree reuse the root children share exactly playouts - 1 visits (the first playout expands the root)
            assert pol.sum() == playouts - 1
            if i >= 2:  # after the random opening the move played is the most visited child (first wins ties)
                assert pol[moves[i]] == pol.max()
            assert list(f["reward"].float_list.value) == [reward if k % 2 == 0 else -reward]


def test_selfplay_virtual_loss_leaves_fill_batches_from_few_games(small_ckpt):
    """leaves_per_game > 1: the reference's multi-thread search (virtual-loss counters on the descent path) lets two
    games keep dozens of leaves in flight.  Every playout is still counted exactly once, every move is legal, the
    virtual-loss counters net out (a second run from the same state plays on), and the batches are far larger than
    the two leaves a one-leaf-per-game search can offer."""
    import oracle
    from unrealgo_b200 import selfplay
    ev = _ev(small_ckpt, max_batch=64)
    kw = dict(games=2, threads=2, playouts_per_move=96, max_moves=8, random_opening_moves=2, seed=5,
              record_moves=True, max_batch=64, timeout_us=200)
    st1, g1 = selfplay(ev, leaves_per_game=1, **kw)
    st16, g16 = selfplay(ev, leaves_per_game=16, **kw)
    ev.close()
    for st in (st1, st16):
        assert st["moves"] == 2 * 8 and st["playouts"] == 2 * 8 * 96
        assert st["playouts"] == st["evaluations"] + st["terminal_leaves"]
    assert st1["mean_batch"] <= 2.0 and st16["mean_batch"] >= 8.0
    for g in g16:
        b = oracle.Board()
        for mv in g:
            assert b.is_legal(mv), (g, mv)
            b.play(mv)


def test_compact_priors_are_the_legal_entries_in_move_order(small_ckpt):
    """SURVEY 8(f) row 1, second half: the queue ships #legal priors per leaf instead of 362 floats.  Same numbers as the
    dense rows at the legal indices, in ascending move order (pass last); a leaf without a mask still gets 362."""
    import oracle
    from unrealgo_b200 import LeafQueue
    packed, _ = oracle.synthetic_positions(48, seed=8)
    rng = np.random.default_rng(3)
    legal = np.zeros((48, 12), np.uint32)
    idx = []
    for i in range(48):
        k = int(rng.integers(1, 362))
        moves = np.sort(rng.permutation(362)[:k])
        idx.append(moves)
        np.bitwise_or.at(legal[i], moves >> 5, (np.uint32(1) << (moves & 31).astype(np.uint32)))
    ev = _ev(small_ckpt)
    dense_p, dense_v = ev.evaluate_packed(packed, legal)
    plain_p, _ = ev.evaluate_packed(packed[:1])
    q = LeafQueue(ev, max_batch=32, timeout_us=300)
    q.set_compact_priors(True)
    tickets = [q.submit(packed[i], legal[i]) for i in range(48)] + [q.submit(packed[0])]
    for i in range(48):
        pol, val = q.wait(tickets[i])
        k = len(idx[i])
        assert np.array_equal(pol[:k], dense_p[i][idx[i]]) and val == dense_v[i], i
        assert abs(float(pol[:k].sum()) - 1) < 1e-5
    pol, _ = q.wait(tickets[48])                       # no mask: the dense softmax row
    assert np.array_equal(pol, plain_p[0])
    st = q.stats_ex()
    assert st["leaves"] == 49 and st["failed_batches"] == 0
    q.close()
    ev.close()


def test_waiting_on_a_consumed_ticket_returns_instead_of_hanging(small_ckpt):
    import oracle
    from unrealgo_b200 import LeafQueue
    packed, _ = oracle.synthetic_positions(2, seed=8)
    ev = _ev(small_ckpt)
    q = LeafQueue(ev, max_batch=8, timeout_us=100)
    t = q.submit(packed[0])
    q.wait(t)
    with pytest.raises(RuntimeError):
        q.wait(t)                                      # consumed: reported, not waited for
    q.close()
    ev.close()


def test_failed_batch_fails_its_own_leaves_only(small_ckpt):
    """a queue created before any checkpoint is loaded: the leaves of that batch come back failed (outputs untouched),
    and once weights are there the same queue evaluates — a failure belongs to a batch, not to the queue"""
    import oracle
    from unrealgo_b200 import DlTFNetworkEvaluator, LeafQueue
    packed, _ = oracle.synthetic_positions(4, seed=8)
    ev = DlTFNetworkEvaluator(small_ckpt[0], max_batch=64)        # graph only
    q = LeafQueue(ev, max_batch=8, timeout_us=100)
    t = q.submit(packed[0])
    with pytest.raises(RuntimeError):
        q.wait(t)
    assert q.swap_checkpoint(small_ckpt[1])
    pol, val = q.wait(q.submit(packed[1]))
    want_p, want_v = ev.evaluate_packed(packed[1:2])
    assert np.array_equal(pol, want_p[0]) and val == want_v[0]
    st = q.stats_ex()
    assert st["failed_batches"] == 1 and st["leaves"] == 2
    q.close()
    ev.close()
