"""The C++ host shim (unrealgo_b200/host/*.h) driven like the reference's own call sites: compiled here with
g++ -std=c++11 against libugeval.so, run on the GPU, compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_uct_board_evaluator_cpp(small_ckpt, positions, tmp_path):
    from oracle.net_oracle import GraphOracle
    meta, prefix = small_ckpt
    _, planes = positions
    n = 16
    exe = str(tmp_path / "test_host_shim")
    libdir = os.path.join(ROOT, "unrealgo_b200", "lib")
    subprocess.run(["g++", "-std=c++11", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_host_shim.cpp"),
                    "-L" + libdir, "-lugeval", "-Wl,-rpath," + libdir], check=True)
    (tmp_path / "dlcfg-19").write_text(
        f"# test config\nmeta_graph={meta}\ncheckpoint={prefix}\nnn_input=feature_input\n"
        "nn_outputs=resnet/tower_0/policy_head/policy_predict:resnet/tower_0/value_head/reward_predict\n"
        "value_transform=1\nreuse_searchtree=true\n")
    planes[:n].tofile(str(tmp_path / "planes.bin"))
    r = subprocess.run([exe, "planes.bin", str(n)], cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    flags = {}
    pol = np.zeros((n, 362))
    val = np.zeros(n)
    for line in r.stdout.splitlines():
        t = line.split()
        if t[0] == "value":
            val[int(t[1])] = float(t[2])
        elif t[0] == "policy":
            pol[int(t[1])] = [float(x) for x in t[2:]]
        elif t[0] == "registered":
            flags["registered"], flags["identical"] = int(t[1]), int(t[3])
        else:
            flags[t[0]] = int(t[1])
    assert flags == {"loaded_before": 0, "bad_path_throws": 1, "untouched": 1, "loaded_after": 1,
                     "registered": 1, "identical": 1}
    assert "Running model failed" in r.stderr  # the reference logs and carries on (.cc:315-317)
    want_p, want_v = GraphOracle(meta, prefix).evaluate_reference(planes[:n], value_transform=1)
    assert np.abs(pol - want_p).max() <= 1e-3 and np.abs(val - want_v).max() <= 5e-3
