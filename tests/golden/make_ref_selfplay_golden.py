"""Generates tests/golden/ref_selfplay_14.tfrecords.gz with the REFERENCE's own self-play player.  Run from the repo
root IN THE BUILD CONTAINER after `python -c "import __graft_entry__ as g; g.build()"` (needs /root/reference):

    python tests/golden/make_ref_selfplay_golden.py

oracle/_ref/ref_selfplay is the reference's unmodified search/UctDeepPlayer.cpp + search engine (oracle/Makefile); it
plays 14 moves of self-play — 4000 playouts per move (MAX_SEARCH_ITERATIONS, search/UctDeepPlayer.cpp:15), tree reuse
on — against the stand-in network of tests/cpp/ugeval_stub.c (policy/value = a hash of the 17 planes, answered only
once the leaf's planes have settled: UGSTUB_SYNC) and writes the game with its own WriteTFRecord.  The file holds, per
move from the last back to the first: the 17 planes CollectFeatures produced, the root's visit counts of that move's
search, and the alternating Tromp-Taylor reward (komi 7.5)."""
import gzip
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
MOVES = 14


def main():
    d = tempfile.mkdtemp(prefix="refselfplay_")
    stub = os.path.join(d, "ugeval_stub.so")
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-o", stub, os.path.join(ROOT, "tests", "cpp", "ugeval_stub.c")],
                   check=True)
    with open(os.path.join(d, "dlcfg-19"), "w") as f:
        f.write("meta_graph=stub.meta\ncheckpoint=stub\nvalue_transform=0\nreuse_searchtree=true\n")
    out = os.path.join(d, "game.tfrecords")
    env = dict(os.environ, LD_PRELOAD=stub, UGSTUB_SYNC="1", UGSTUB_SETTLE_US="200")
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_selfplay"), str(MOVES), out], cwd=d, env=env,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    print("moves:", " ".join(ln.split()[3] for ln in r.stdout.splitlines() if ln.startswith("move ")))
    data = open(out, "rb").read()
    with open(os.path.join(ROOT, "tests", "golden", f"ref_selfplay_{MOVES}.tfrecords.gz"), "wb") as f:
        f.write(gzip.compress(data, 9, mtime=0))
    print(len(data), "bytes of records written")


if __name__ == "__main__":
    sys.exit(main())
