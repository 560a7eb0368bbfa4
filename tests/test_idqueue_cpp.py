"""Lock-free id queue (csrc/idqueue.h) under contention, on the CPU: compiled from tests/cpp/test_idqueue.cpp."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_idqueue_stress(tmp_path):
    exe = str(tmp_path / "test_idqueue")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(ROOT, "tests", "cpp", "test_idqueue.cpp"),
                    "-o", exe], check=True, capture_output=True)
    for producers, per in ((16, 100000), (3, 300000)):
        r = subprocess.run([exe, str(producers), str(per)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
