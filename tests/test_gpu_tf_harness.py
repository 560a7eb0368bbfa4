"""tools/compare_with_tf.py end to end on the GPU, fed by a stand-in dump (tests/make_standin_tf_dump.py: the numpy
float64 oracle in place of TensorFlow, which is not available offline) on a foreign-idiom checkpoint standing in for
the real bootstrap checkpoint.  Proves the harness runs and fails when it should; TensorFlow parity itself stays
"unverified" until tools/tf_dump_reference.py has produced a real dump (README.md)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_compare_with_tf_runs_end_to_end(tmp_path):
    from tests import foreign_graphs as FG
    from tests import make_standin_tf_dump
    meta, prefix = FG.write(str(tmp_path / "standin"), name="bootstrap-ckpt", bn="cond", blocks=3, filters=128, seed=21)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "positions_64.npz"))
    dump = make_standin_tf_dump.write(str(tmp_path / "dump.npz"), meta, prefix, gold["planes"][:32], gold["packed"][:32],
                                      value_transform=1)
    cmd = [sys.executable, os.path.join(ROOT, "tools", "compare_with_tf.py"), "--dump", dump, "--meta", meta, "--ckpt", prefix]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "fp32:" in r.stdout and "fp16:" in r.stdout and "stand-in" in r.stdout
    # a dump that does not belong to the checkpoint must fail the comparison
    d = dict(np.load(dump))
    d["policy"] = np.roll(d["policy"], 1, axis=1)
    bad = str(tmp_path / "bad.npz")
    np.savez_compressed(bad, **d)
    r = subprocess.run(cmd[:3] + [bad] + cmd[4:], capture_output=True, text=True, timeout=600)
    assert r.returncode == 1 and "FAILED" in r.stdout, r.stdout + r.stderr
