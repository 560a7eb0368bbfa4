"""CPU-side checks of the TensorFlow parity harness (tools/): TensorFlow itself is not available here, so the script's
TF-free parts are exercised directly — the dlcfg reader, the inference-sub-graph description (on a GraphDef built with
the protobuf runtime, the same message API TensorFlow hands the script), the dump layout through the stand-in writer —
and the script must say clearly what it needs when TensorFlow is missing."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_dump_script_reads_dlcfg_like_the_reference(tmp_path):
    mod = _load("tf_dump_reference")
    p = tmp_path / "dlcfg-19"
    p.write_text("# comment\nnn_input=feature_input\nnn_outputs=a/policy:b/value\nvalue_transform=1  # flip\n")
    cfg = mod.read_dlcfg(str(p))
    assert cfg == {"nn_input": "feature_input", "nn_outputs": "a/policy:b/value", "value_transform": "1"}
    ref = open(os.path.join(ROOT, "tests", "golden", "dlcfg_cases.json")).read()
    assert "nn_outputs" in ref   # the compiled reference parser's answers live there (tests/test_dlconfig.py)


def test_subgraph_description_answers_the_surveys_unknowns():
    from tests import foreign_graphs as FG
    mod = _load("tf_dump_reference")
    _, meta = FG.build(bn="cond", blocks=2, filters=64, bn_eps=1e-3, conv_bias=True)
    info = mod.describe_inference_subgraph(meta.graph_def, "feature_input", mod.DEFAULT_OUTPUTS)
    assert info["ops"]["Conv2D"] == 1 + 4 + 2 and info["ops"]["MatMul"] == 3 and info["ops"]["Merge"] == 7
    assert info["batch_norm_epsilons"] == [0.001] and info["data_formats"] == ["NHWC"]
    assert info["batch_norm_is_training_values"] == [False, True]          # both branches of the cond are reachable
    assert info["placeholders"] == {"feature_input": "Placeholder", "training": "PlaceholderWithDefault"}
    assert "save/restore_all" not in info["ops"] and info["nodes_on_path"] < len(meta.graph_def.node)


def test_dump_script_side_feeds():
    mod = _load("tf_dump_reference")
    assert mod.parse_feeds(["training=false", "keep_prob:0=1.0", "k=3"]) == {"training:0": False, "keep_prob:0": 1.0, "k:0": 3}
    assert mod.parse_feeds([]) == {} and mod.parse_feeds(None) == {}
    import pytest
    with pytest.raises(SystemExit):
        mod.parse_feeds(["training"])


def test_dump_script_needs_tensorflow_and_says_so(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "tf_dump_reference.py"), "--meta", "x.meta", "--ckpt", "x",
                        "--positions", "p.npz", "--out", str(tmp_path / "o.npz")], capture_output=True, text=True)
    assert r.returncode != 0 and "tensorflow" in (r.stderr + r.stdout).lower()


def test_standin_dump_has_the_layout_compare_expects(tmp_path, small_ckpt, positions):
    from tests import make_standin_tf_dump
    packed, planes = positions
    out = make_standin_tf_dump.write(str(tmp_path / "d.npz"), small_ckpt[0], small_ckpt[1], planes[:4], packed[:4], 2)
    d = np.load(out)
    info = json.loads(str(d["info"]))
    assert info["producer"] == "oracle-standin" and info["value_transform"] == 2
    assert d["policy"].shape == (4, 362) and d["value"].shape == (4,) and d["planes"].dtype == np.int8
    assert np.allclose(d["value_transformed"], 1 - d["value"].astype(np.float64), atol=1e-7)
    src = open(os.path.join(ROOT, "tools", "compare_with_tf.py")).read()
    for key in ("planes", "policy", "value_transformed", "info"):
        assert f'd["{key}"]' in src
    assert "import oracle" not in src and "from oracle" not in src        # the product-side tool never touches the oracle
