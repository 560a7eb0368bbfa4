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
