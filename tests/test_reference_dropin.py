"""Drop-in boundary, checked against the reference's own call site (no GPU needed): search/UctBoardEvaluator.cpp is
compiled UNMODIFIED from /root/reference with unrealgo_b200/host/overlay ahead of the reference root on the include
path; everything it then needs from the evaluator must be C-ABI symbols declared in include/ugeval.h, and nothing from
TensorFlow.  Skipped where the reference tree is absent (the GPU box); there the prebuilt oracle/_ref/ref_dropin_eval
is exercised by tests/test_gpu_reference_dropin.py."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DEFS = ["-DGO_DEFINE_MAX_SIZE=19", "-DUSE_NNEVALTHREAD", "-DCHECK_EVAL_RESULT", "-DFORCE_SINGLE_THREAD",
        "-DUSE_DIRECT_VISITCOUNT"]  # CMakeLists.txt:12-16

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "search")), reason="reference tree not present")


def _inc():
    return ["-I" + os.path.join(ROOT, "unrealgo_b200", "host", "overlay"), "-I" + os.path.join(ROOT, "oracle", "ref_shim"),
            "-I" + REF] + ["-I" + os.path.join(REF, d) for d in ("go", "lib", "platform", "board", "search", "config")]


@needs_ref
def test_reference_call_site_compiles_against_the_overlay_and_binds_only_the_c_abi(tmp_path):
    obj = str(tmp_path / "UctBoardEvaluator.o")
    subprocess.run(["g++", "-std=c++11", "-O1", "-w", "-include", "string", "-c"] + DEFS + _inc() +
                   [os.path.join(REF, "search", "UctBoardEvaluator.cpp"), "-o", obj], check=True)
    und = subprocess.run(["nm", "-C", "-u", obj], capture_output=True, text=True, check=True).stdout
    syms = [ln.split()[-1] for ln in und.splitlines() if ln.strip()]
    ug = sorted(s for s in syms if s.startswith("ug_"))
    assert ug == ["ug_eval_create", "ug_eval_destroy", "ug_eval_load_checkpoint", "ug_eval_load_graph", "ug_eval_ready",
                  "ug_eval_run_planes", "ug_eval_run_planes_hwc",   # _hwc: Evaluate() under DF_CHW (see the header)
                  "ug_eval_set_io", "ug_eval_set_value_transform", "ug_last_error"]
    header = open(os.path.join(ROOT, "include", "ugeval.h")).read()
    for s in ug:
        assert re.search(r"\b%s\s*\(" % s, header), s          # every one is a declared entry point
    assert not [s for s in und.splitlines() if "tensorflow" in s or "TF_" in s]
    # the object's own definitions are the reference class's members, i.e. it really is the reference's code
    defined = subprocess.run(["nm", "-C", "--defined-only", obj], capture_output=True, text=True, check=True).stdout
    for member in ("UctBoardEvaluator::EvaluateState", "UctBoardEvaluator::LoadGraph",
                   "UctBoardEvaluator::UpdateCheckPoint", "UctBoardEvaluator::GraphLoaded"):
        assert member in defined


@needs_ref
def test_dropin_binary_builds_and_has_no_cpu_fallback(tmp_path):
    if not os.path.exists(os.path.join(ROOT, "unrealgo_b200", "lib", "libugeval.so")):
        pytest.skip("libugeval.so not built")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "dropin"], check=True, capture_output=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_dropin_eval")
    assert os.path.exists(exe)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run itself is tests/test_gpu_reference_dropin.py")
    (tmp_path / "planes.bin").write_bytes(bytes(17 * 361))
    r = subprocess.run([exe, "planes.bin", "1"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr   # the reference's call path fails loudly without a GPU


@needs_ref
def test_reference_record_writer_header_over_ug_tfrecord(tmp_path):
    """funcapproximator/DlTFRecordWriter.h is used UNMODIFIED (it only forward-declares TensorFlow types);
    unrealgo_b200/host/DlTFRecordWriter.cc supplies the members in place of the reference's TensorFlow-based .cc"""
    import numpy as np
    from tests.test_tfrecord import _example_class, read_records
    libdir = os.path.join(ROOT, "unrealgo_b200", "lib")
    if not os.path.exists(os.path.join(libdir, "libugeval.so")):
        pytest.skip("libugeval.so not built")
    exe = str(tmp_path / "ref_recordwriter")
    subprocess.run(["g++", "-std=c++11", "-O1", "-include", "string", "-I" + REF,
                    os.path.join(ROOT, "tests", "cpp", "ref_recordwriter_main.cc"),
                    os.path.join(ROOT, "unrealgo_b200", "host", "DlTFRecordWriter.cc"),
                    "-L" + libdir, "-lugeval", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    path = str(tmp_path / "out.tfrecord")
    r = subprocess.run([exe, path, "7"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "bad_write -1" in r.stdout, r.stdout + r.stderr
    recs = read_records(path)
    assert len(recs) == 7
    Example = _example_class()
    s = 12345
    for i, raw in enumerate(recs):
        ex = Example.FromString(raw)
        feat = np.empty(6137, np.uint8)
        pol = np.empty(362, np.float32)
        for k in range(6137):
            s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
            feat[k] = (s >> 24) & 1
        for k in range(362):
            s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
            pol[k] = np.float32((s >> 8) / 16777216.0)
        kf, kp, kr = ("feature", "policy", "reward") if i % 2 == 0 else ("f", "p", "r")
        f = ex.features.feature
        assert sorted(f.keys()) == sorted([kf, kp, kr])
        assert f[kf].bytes_list.value[0] == feat.tobytes()
        assert np.array_equal(np.array(f[kp].float_list.value, np.float32), pol)
        assert list(f[kr].float_list.value) == [float(i % 3 - 1)]


@needs_ref
def test_leaf_queue_patch_applies_to_the_reference_tree(tmp_path):
    """unrealgo_b200/host/patches/uctsearch_leaf_queue.patch, `patch -p1` from the reference root:
    applies without fuzz or rejects, everything it adds sits behind UGEVAL_LEAF_QUEUE, and what it adds calls only
    declared C-ABI entry points"""
    import shutil
    os.makedirs(tmp_path / "search")
    shutil.copy(os.path.join(REF, "search", "UctSearch.cpp"), tmp_path / "search" / "UctSearch.cpp")
    patch = os.path.join(ROOT, "unrealgo_b200", "host", "patches", "uctsearch_leaf_queue.patch")
    r = subprocess.run(["patch", "-p1", "-i", patch], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0 and "fuzz" not in r.stdout and "rej" not in r.stdout, r.stdout + r.stderr
    patched = open(tmp_path / "search" / "UctSearch.cpp").read()
    original = open(os.path.join(REF, "search", "UctSearch.cpp")).read()
    assert patched.count("UGEVAL_LEAF_QUEUE") >= 5 and "UGEVAL_LEAF_QUEUE" not in original
    header = open(os.path.join(ROOT, "include", "ugeval.h")).read()
    added = [ln[1:] for ln in open(patch).read().splitlines() if ln.startswith("+") and not ln.startswith("+++")]
    used = set(re.findall(r"\b(ug_[a-z_]+)\s*\(", "\n".join(added)))
    assert used == {"ug_pack_planes", "ug_queue_submit", "ug_queue_wait", "ug_queue_create", "ug_queue_swap_checkpoint",
                    "ug_eval_set_value_transform"}   # dlcfg value_transform: Evaluate() read it per call, the queue bypasses it
    for s in used:
        assert re.search(r"\b%s\s*\(" % s, header), s
