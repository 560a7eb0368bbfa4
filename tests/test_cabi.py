"""The C-ABI library loads on a GPU-less box and exports every symbol include/ugeval.h declares; the product
path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ugeval.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ug_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    from unrealgo_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ugeval.h but not exported"
    assert set(_lib.SIGNATURES) == set(names), set(_lib.SIGNATURES) ^ set(names)


def test_struct_layout_matches_header():
    from unrealgo_b200 import _lib
    assert C.sizeof(_lib.PackedPos) == 784 == _lib.PACKED_POS_BYTES
    assert _lib.PackedPos.white.offset == 384 and _lib.PackedPos.to_play.offset == 768


def test_library_is_sm100a_only():
    import subprocess
    from unrealgo_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from unrealgo_b200 import DlTFNetworkEvaluator
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        DlTFNetworkEvaluator("")


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under unrealgo_b200/ may reference it"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "unrealgo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "ugo_oracle" not in src, f


def test_wave_aligned_batch_sizes():
    """ug_eval_efficient_batch's rule (host arithmetic): a wave of the tower kernel on a 148-SM part is 74 CTA pairs x
    256 pixel rows = 18944 rows = 52.47 positions; a partial batch is cut back to whole waves."""
    from unrealgo_b200 import _lib
    f = _lib.load().ug_efficient_batch_for_sms
    assert [f(148, n) for n in (1, 30, 52)] == [1, 30, 52]          # below / at one wave: unchanged
    assert f(148, 53) == 52 and f(148, 103) == 52 and f(148, 104) == 104 and f(148, 105) == 104
    assert f(148, 209) == 209 and f(148, 213) == 209 and f(148, 256) == 209 and f(148, 261) == 209 and f(148, 262) == 262
    bounds = [w * 18944 // 361 for w in range(1, 13)]               # 52, 104, 157, 209, 262, ...
    for n in range(1, 600):
        want = max([b for b in bounds if b <= n], default=n)
        assert f(148, n) == want, n
    assert f(0, 17) == 17 and f(148, 0) == 0
