"""Generates tests/golden/*.npz and dlcfg_cases.json.  Run from the repo root IN THE BUILD CONTAINER:

    python tests/golden/make_golden.py

* encoder_kat.npz   — hand-constructed histories (empty board, one stone, capture, ko, short history, pass,
                      forced to-play) with the oracle's CollectFeatures planes and packed masks.  The reference
                      has no vectors for this path; these are self-made known-answer cases whose
                      expected properties are ALSO asserted independently in tests/test_oracle_golden.py.
* positions_64.npz  — first 64 positions of the seeded SURVEY 8(d) generator (packed + planes).
* dlcfg_cases.json  — config-file texts with the answers of the REFERENCE's own DlConfig class, compiled from
                      /root/reference into oracle/_ref/libugref.so (funcapproximator/DlConfig.cc, lib/StringUtil.cc).
* get_offset.json   — UnrealGo::ArrayUtil::GetOffset answers from the same compiled reference code.
"""
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pt(col, row):  # 1-based GTP-ish coordinates -> point index
    return (row - 1) * 19 + (col - 1)


def kat_cases():
    cases = {}
    b = oracle.Board()
    cases["empty_black"] = b
    b = oracle.Board(); b.set_to_play(oracle.WHITE)
    cases["empty_white_forced"] = b
    b = oracle.Board(); b.play(pt(4, 4))
    cases["one_stone"] = b
    b = oracle.Board(); b.play(oracle.PASS)
    cases["one_pass"] = b
    # capture: white stone at (2,1) surrounded by black (1,1),(3,1),(2,2)
    b = oracle.Board()
    for mv in [pt(1, 1), pt(2, 1), pt(3, 1), pt(10, 10), pt(2, 2)]:
        b.play(mv)
    cases["capture_corner"] = b
    # ko: classic shape, black captures at k, white may not retake immediately
    b = oracle.Board()
    seq = [pt(4, 4), pt(5, 4), pt(3, 3), pt(6, 3), pt(4, 2), pt(5, 2), pt(5, 3), pt(4, 3)]  # W takes the ko at (4,3)
    for mv in seq:
        b.play(mv)
    cases["ko_taken"] = b
    b = oracle.Board()
    for i in range(5):
        b.play(pt(3 + 2 * i, 16))
    cases["short_history_5"] = b
    b = oracle.Board()
    for i in range(12):
        b.play(pt(2 + i, 3 + (i % 2)))
    cases["long_history_12"] = b
    b = oracle.Board()
    for mv in [pt(4, 4), pt(16, 16), oracle.PASS, pt(4, 16), oracle.PASS, oracle.PASS, pt(16, 4)]:
        b.play(mv)
    cases["passes_inside"] = b
    # last stack entry's colour != opp(to_play): GetLastMove() is GO_NULLMOVE, nothing is undone
    b = oracle.Board()
    for mv in [pt(4, 4), pt(16, 16), pt(10, 10)]:
        b.play(mv)
    b.set_to_play(oracle.BLACK)
    cases["to_play_forced_mismatch"] = b
    return cases


def main():
    cases = kat_cases()
    names = sorted(cases)
    np.savez_compressed(os.path.join(HERE, "encoder_kat.npz"),
                        names=np.array(names),
                        packed=np.stack([cases[k].pack() for k in names]),
                        planes=np.stack([cases[k].collect_features() for k in names]),
                        to_play=np.array([cases[k].to_play for k in names], np.int8),
                        colors=np.stack([cases[k].colors() for k in names]))
    packed, planes = oracle.synthetic_positions(64, seed=20260119)
    np.savez_compressed(os.path.join(HERE, "positions_64.npz"), packed=packed, planes=planes)

    ref = oracle.ref_lib()
    assert ref is not None, "oracle/_ref/libugref.so missing: run make -C oracle with /root/reference present"
    texts = {
        "reference_dlcfg19": open("/root/reference/config/dlcfg-19").read(),
        "empty": "",
        "comments_and_spaces": "# comment\n   # indented comment\n  nn_input =  my_input  \nnn_outputs=a/b:c/d:e\nvalue_transform = 2\nreuse_searchtree=false\n",
        "duplicates_first_wins": "checkpoint=first\ncheckpoint=second\nvalue_transform=1\nvalue_transform=0\n",
        "no_equals_and_tabs": "justakey\n\tnn_input=\ttabbed\nmeta_graph=a=b=c\nreuse_searchtree=False\n",
        "value_transform_other": "value_transform=7\nreuse_searchtree=anything\nnn_outputs=only_one\n",
        "empty_values": "nn_input=\nmeta_graph=\nvalue_transform=0\n",
    }
    keys = ["meta_graph", "checkpoint", "nn_input", "nn_outputs", "value_transform", "reuse_searchtree", "host",
            "justakey", "", "checkevalsubpath"]
    out = {}
    for name, text in texts.items():
        with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
            f.write(text)
            path = f.name
        c = ref.ref_dlcfg_open(path.encode())
        buf = C.create_string_buffer(4096)
        ans = {"get": {}}
        for k in keys:
            for d in ("", "DEFAULT"):
                ref.ref_dlcfg_get(c, k.encode(), d.encode(), buf, len(buf))
                ans["get"][f"{k}|{d}"] = buf.value.decode()
        ref.ref_dlcfg_network_input(c, buf, len(buf)); ans["network_input"] = buf.value.decode()
        ref.ref_dlcfg_metagraph(c, buf, len(buf)); ans["metagraph"] = buf.value.decode()
        ref.ref_dlcfg_checkpoint(c, buf, len(buf)); ans["checkpoint"] = buf.value.decode()
        ob = C.create_string_buffer(512 * 8)
        n = ref.ref_dlcfg_network_outputs(c, ob, 512, 8)
        ans["network_outputs"] = [ob.raw[i * 512:(i + 1) * 512].split(b"\0", 1)[0].decode() for i in range(n)]
        ans["value_transform"] = ref.ref_dlcfg_value_transform(c)
        ans["reuse_search_tree"] = ref.ref_dlcfg_reuse_search_tree(c)
        ref.ref_dlcfg_close(c)
        os.unlink(path)
        out[name] = {"text": text, "answers": ans}
    # missing file
    c = ref.ref_dlcfg_open(b"/nonexistent/dlcfg-19")
    buf = C.create_string_buffer(4096)
    ref.ref_dlcfg_network_input(c, buf, len(buf))
    out["__missing_file__"] = {"network_input": buf.value.decode(), "value_transform": ref.ref_dlcfg_value_transform(c),
                               "reuse_search_tree": ref.ref_dlcfg_reuse_search_tree(c)}
    ref.ref_dlcfg_close(c)
    json.dump(out, open(os.path.join(HERE, "dlcfg_cases.json"), "w"), indent=1, sort_keys=True)

    rng = np.random.default_rng(1)
    cases = []
    for _ in range(200):
        n = int(rng.integers(1, 17))
        for dims in ([n, 19, 19, 17], [n, 17, 19, 19]):
            idx = [int(rng.integers(0, d)) for d in dims]
            d_arr = (C.c_int * 4)(*dims)
            i_arr = (C.c_int * 4)(*idx)
            cases.append({"dims": dims, "idx": idx, "offset": ref.ref_get_offset(d_arr, 4, i_arr)})
    json.dump(cases, open(os.path.join(HERE, "get_offset.json"), "w"))
    print("golden files written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
