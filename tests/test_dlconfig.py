"""dlcfg parsing: the product's parser (libugeval, host-only code) must answer exactly like
the reference's own DlConfig class — golden answers were produced by the reference source compiled into
oracle/_ref (tests/golden/make_golden.py); when oracle/_ref is present it is also queried live."""
import json
import os

import pytest

from unrealgo_b200 import DlConfig

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = json.load(open(os.path.join(G, "dlcfg_cases.json")))


@pytest.mark.parametrize("name", sorted(k for k in CASES if not k.startswith("__")))
def test_dlcfg_matches_reference(name, tmp_path):
    case = CASES[name]
    path = tmp_path / "dlcfg-19"
    path.write_text(case["text"])
    cfg = DlConfig(str(path))
    a = case["answers"]
    for kd, want in a["get"].items():
        k, d = kd.split("|")
        assert cfg.get(k, d) == want, (name, k, d)
    assert cfg.get_network_input() == a["network_input"]
    assert cfg.get_metagraph() == a["metagraph"]
    assert cfg.default_checkpoint() == a["checkpoint"]
    assert cfg.get_network_outputs() == a["network_outputs"]
    assert cfg.get_value_transform() == a["value_transform"]
    assert int(cfg.reuse_search_tree()) == a["reuse_search_tree"]


def test_missing_file_gives_defaults():
    cfg = DlConfig("/nonexistent/dlcfg-19")
    want = CASES["__missing_file__"]
    assert cfg.get_network_input() == want["network_input"] == "input"
    assert cfg.get_value_transform() == want["value_transform"] == 0
    assert int(cfg.reuse_search_tree()) == want["reuse_search_tree"] == 0
    assert cfg.get_metagraph() == "bootstrap-ckpt.meta" and cfg.default_checkpoint() == "bootstrap-ckpt"


def test_reference_config_keys():
    """the keys the drop-in must honour, on the reference's shipped config (config/dlcfg-19:14-19)"""
    a = CASES["reference_dlcfg19"]["answers"]
    assert a["network_input"] == "feature_input"
    assert a["network_outputs"] == ["resnet/tower_0/policy_head/policy_predict",
                                    "resnet/tower_0/value_head/reward_predict"]
    assert a["value_transform"] == 0 and a["reuse_search_tree"] == 1
