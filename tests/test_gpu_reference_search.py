"""The reference's OWN search engine on the product library.  oracle/_ref/ref_search is built in the build container
(oracle/Makefile `refsearch`) from 93 unmodified reference sources — search/UctSearch.cpp with its NetworkEvalThread
and ExpandAndBackup, search/UctBoardEvaluator.cpp, gouct/, go/, board/ ... — with unrealgo_b200/host/overlay ahead of
the reference root on the include path and libugeval.so where TensorFlow was.  Here it searches on the GPU."""
import os
import subprocess

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_search")
OUTS = "resnet/tower_0/policy_head/policy_predict:resnet/tower_0/value_head/reward_predict"


def run_ref_search(cwd, meta, prefix, playouts, moves, timeout=300):
    with open(os.path.join(cwd, "dlcfg-19"), "w") as f:
        f.write(f"meta_graph={meta}\ncheckpoint={prefix}\nnn_input=feature_input\nnn_outputs={OUTS}\n"
                "value_transform=0\nreuse_searchtree=false\n")
    r = subprocess.run([EXE, str(playouts), str(moves)], cwd=cwd, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = {"moves": [], "games": []}
    for line in r.stdout.splitlines():
        t = line.split()
        if t and t[0] == "move":
            out["moves"].append(int(t[3]))
            out["games"].append(float(t[5]))
        elif t and t[0] == "root":
            out["root"] = {int(x.split(":")[0]): float(x.split(":")[1]) for x in t[1:]}
        elif t and t[0] == "total":
            out["playouts_per_s"] = float(t[6])
    return out


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/ref_search not prebuilt (needs the reference tree)")
def test_reference_search_engine_runs_on_ugeval(small_ckpt, tmp_path):
    meta, prefix = small_ckpt
    out = run_ref_search(str(tmp_path), meta, prefix, playouts=300, moves=4)
    assert len(out["moves"]) == 4 and all(g >= 300 for g in out["games"])
    b = oracle.Board()
    for mv in out["moves"]:                      # the engine's chosen moves are legal on the oracle's board
        assert b.is_legal(mv)
        b.play(mv)
    visits = out["root"]
    assert sum(visits.values()) >= 299 and len(visits) > 1 and out["playouts_per_s"] > 50


@pytest.mark.parametrize("playouts,moves,reuse", [(160, 10, False), (200, 12, True)])
@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/ref_search not prebuilt (needs the reference tree)")
def test_reference_engine_and_product_driver_play_the_same_game_on_the_gpu(small_ckpt, tmp_path, playouts, moves, reuse):
    """Both on the CUDA evaluator: the reference's engine through UctBoardEvaluator -> ug_eval_run_planes (host planes
    in, doubles out, its own UpdatePrior in double), the product's self-play driver through the leaf queue
    (packed positions, legal-mask renormalisation in the head kernel, floats).  With the reference's hand-off made
    leaf-accurate (tests/cpp/ugeval_sync_preload.c) the two searches must choose the same moves, from a fresh tree per
    move and with the played move's subtree carried over."""
    from unrealgo_b200 import DlTFNetworkEvaluator, selfplay
    meta, prefix = small_ckpt
    ev = DlTFNetworkEvaluator("", feature_input="feature_input", outputs=OUTS.split(":"), max_batch=64)
    try:
        ev.LoadGraph(meta)
        ev.UpdateCheckPoint(prefix)
        _, games = selfplay(ev, games=1, threads=1, playouts_per_move=playouts, max_moves=moves, reuse_tree=reuse,
                            random_opening_moves=0, record_moves=True, max_batch=64, leaves_per_game=1)
    finally:
        ev.close()
    so = str(tmp_path / "sync_preload.so")
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "cpp", "ugeval_sync_preload.c"),
                    "-ldl"], check=True)
    with open(os.path.join(str(tmp_path), "dlcfg-19"), "w") as f:
        f.write(f"meta_graph={meta}\ncheckpoint={prefix}\nnn_input=feature_input\nnn_outputs={OUTS}\n"
                f"value_transform=0\nreuse_searchtree={'true' if reuse else 'false'}\n")
    ref_moves = None
    for attempt in range(3):
        e = dict(os.environ, LD_PRELOAD=so, UGSTUB_SETTLE_US=str(250 * (attempt + 1)))
        r = subprocess.run([EXE, str(playouts), str(moves)], cwd=str(tmp_path), env=e, capture_output=True, text=True,
                           timeout=120)
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
        ref_moves = [int(ln.split()[3]) for ln in r.stdout.splitlines() if ln.startswith("move ")]
        if ref_moves == games[0]:
            break
    assert ref_moves == games[0]


EXE_QUEUE = os.path.join(ROOT, "oracle", "_ref", "ref_search_queue")


@pytest.mark.parametrize("playouts,moves,reuse", [(160, 10, False), (200, 12, True)])
@pytest.mark.skipif(not os.path.exists(EXE_QUEUE), reason="oracle/_ref/ref_search_queue not prebuilt")
def test_reference_engine_with_the_leaf_queue_patch(small_ckpt, tmp_path, playouts, moves, reuse):
    """INTEGRATION.md section 2 applied for real: unrealgo_b200/host/patches/uctsearch_leaf_queue.patch puts
    ug_queue_submit/ug_queue_wait where the reference's ExpandAndBackup handed its leaf to NetworkEvalThread.  Every
    leaf now gets its own evaluation, so no synchronising shim is needed: the patched engine and the product's
    self-play driver, both on the CUDA evaluator through the queue, play the same game."""
    from unrealgo_b200 import DlTFNetworkEvaluator, selfplay
    meta, prefix = small_ckpt
    ev = DlTFNetworkEvaluator("", feature_input="feature_input", outputs=OUTS.split(":"), max_batch=64)
    try:
        ev.LoadGraph(meta)
        ev.UpdateCheckPoint(prefix)
        _, games = selfplay(ev, games=1, threads=1, playouts_per_move=playouts, max_moves=moves, reuse_tree=reuse,
                            random_opening_moves=0, record_moves=True, max_batch=64, leaves_per_game=1)
    finally:
        ev.close()
    with open(os.path.join(str(tmp_path), "dlcfg-19"), "w") as f:
        f.write(f"meta_graph={meta}\ncheckpoint={prefix}\nnn_input=feature_input\nnn_outputs={OUTS}\n"
                f"value_transform=0\nreuse_searchtree={'true' if reuse else 'false'}\n")
    r = subprocess.run([EXE_QUEUE, str(playouts), str(moves)], cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    assert [int(ln.split()[3]) for ln in r.stdout.splitlines() if ln.startswith("move ")] == games[0]
