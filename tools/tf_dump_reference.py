#!/usr/bin/env python
"""Dump what the REFERENCE's TensorFlow evaluator returns for committed positions — the TF half of the parity harness.

Runs wherever a TensorFlow with `tf.compat.v1` (any 1.x, or 2.x in v1-compat mode) and the bootstrap checkpoint (the
git-LFS object bootstrap.tar.gz of unrealgo/unrealgo) are available; neither exists in this repository's build
container, which is why floating-point parity against the reference is marked "unverified" in README.md until this
script has been run.  It does exactly what tensorflow::DlTFNetworkEvaluator<T> does:

  LoadMetaGraph        funcapproximator/DlTFNetworkEvaluator.cc:115-167   ReadBinaryProto(MetaGraphDef), Session::Create
  UpdateCheckPoint     .cc:170-200   Session::Run({filename_tensor_name: prefix}, {}, {restore_op_name})
  Evaluate             .cc:287-318   TransformFeature (DF_HWC: [b][i][j][d] = feature[b][d][i][j], T = placeholder dtype),
                                     Session::Run({nn_input: tensor}, nn_outputs), float -> double, value_transform
and writes planes + outputs (+ a description of the inference sub-graph: op histogram, batch-norm epsilons, data
formats, placeholder dtype — the unknowns SURVEY.md 8(c) lists) to an .npz that tools/compare_with_tf.py checks
libugeval against.

    python tools/tf_dump_reference.py --meta bootstrap-ckpt.meta --ckpt bootstrap-ckpt \\
        --positions tests/golden/positions_64.npz --out tf_reference.npz [--dlcfg config/dlcfg-19] [--fetch NODE ...]
        [--feed training=false ...]     (a plain `Placeholder` training switch has to be fed; the engine itself feeds
                                         nothing but nn_input, so a graph that needs this cannot run in the reference)
"""
import argparse
import collections
import json
import sys

import numpy as np

DEFAULT_INPUT = "feature_input"                                   # DlTFNetworkEvaluator.cc:33
DEFAULT_OUTPUTS = ["resnet/tower_0/policy_head/policy_predict",   # .cc:34-35; config/dlcfg-19:16-17
                   "resnet/tower_0/value_head/reward_predict"]


def read_dlcfg(path):
    """key=value lines, '#' comments (funcapproximator/DlConfig.cc:20-41)"""
    cfg = {}
    for line in open(path):
        line = line.split("#", 1)[0].strip()
        if "=" in line:
            k, v = line.split("=", 1)
            cfg[k.strip()] = v.strip()
    return cfg


def parse_feeds(items):
    """--feed NAME=VALUE ... -> {tensor name: python scalar}; true/false -> bool, otherwise int, then float"""
    feeds = {}
    for item in items or []:
        if "=" not in item:
            raise SystemExit(f"--feed expects NAME=VALUE, got {item!r}")
        name, value = item.split("=", 1)
        low = value.strip().lower()
        if low in ("true", "false"):
            v = low == "true"
        else:
            try:
                v = int(low)
            except ValueError:
                v = float(low)
        feeds[name if ":" in name else name + ":0"] = v
    return feeds


def describe_inference_subgraph(graph_def, input_name, outputs):
    """walk back from the outputs to the input placeholder: what the evaluator actually runs"""
    nodes = {n.name: n for n in graph_def.node}
    seen, stack = set(), list(outputs)
    while stack:
        name = stack.pop().lstrip("^").split(":")[0]
        if name in seen or name not in nodes:
            continue
        seen.add(name)
        stack.extend(nodes[name].input)
    ops = collections.Counter(nodes[n].op for n in seen)
    eps = sorted({round(float(nodes[n].attr["epsilon"].f), 10) for n in seen if nodes[n].op.startswith("FusedBatchNorm")})
    fmts = sorted({nodes[n].attr["data_format"].s.decode() for n in seen if "data_format" in nodes[n].attr})
    train = sorted({bool(nodes[n].attr["is_training"].b) for n in seen if nodes[n].op.startswith("FusedBatchNorm")})
    placeholders = {n: nodes[n].op for n in seen if nodes[n].op.startswith("Placeholder")}
    conv_shapes = []
    return {"ops": dict(ops), "batch_norm_epsilons": eps, "data_formats": fmts, "batch_norm_is_training_values": train,
            "placeholders": placeholders, "nodes_on_path": len(seen), "conv_shapes": conv_shapes}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--meta", required=True)
    ap.add_argument("--ckpt", required=True, help="checkpoint prefix (<p>.index + <p>.data-00000-of-00001)")
    ap.add_argument("--positions", required=True, help=".npz with `planes` int8 [n][17][19][19] (and optionally `packed`)")
    ap.add_argument("--out", required=True)
    ap.add_argument("--dlcfg", help="take nn_input / nn_outputs / value_transform from this dlcfg file")
    ap.add_argument("--batch", type=int, default=16, help="positions per Session::Run (MAX_BATCHES in the reference)")
    ap.add_argument("--fetch", nargs="*", default=[], help="extra tensors to dump (localises a mismatch)")
    ap.add_argument("--feed", nargs="*", default=[], help="NAME=VALUE scalars fed beside nn_input (e.g. training=false)")
    args = ap.parse_args()
    extra_feeds = parse_feeds(args.feed)

    import tensorflow as tf
    tf1 = tf.compat.v1
    tf1.disable_eager_execution()

    input_name, outputs, vt = DEFAULT_INPUT, list(DEFAULT_OUTPUTS), 0
    if args.dlcfg:
        cfg = read_dlcfg(args.dlcfg)
        input_name = cfg.get("nn_input", input_name)
        if "nn_outputs" in cfg:
            outputs = cfg["nn_outputs"].split(":")          # DlConfig.cc:214-219 splits at ':'
        vt = {"0": 0, "1": 1, "2": 2}.get(cfg.get("value_transform", "0"), 0)   # DlConfig.cc:228-246

    meta_graph = tf1.MetaGraphDef()
    with open(args.meta, "rb") as f:
        meta_graph.ParseFromString(f.read())                 # ReadBinaryProto, .cc:130
    graph = tf.Graph()
    with graph.as_default():
        tf.import_graph_def(meta_graph.graph_def, name="")   # Session::Create(graph_def), .cc:156
    sess = tf1.Session(graph=graph)
    saver = meta_graph.saver_def
    # restore exactly as the reference does (.cc:184-193): feed the filename tensor, run the restore op
    sess.run(saver.restore_op_name, feed_dict={saver.filename_tensor_name: args.ckpt})

    data = np.load(args.positions)
    planes = np.ascontiguousarray(data["planes"], np.int8)
    n = planes.shape[0]
    x = graph.get_tensor_by_name(input_name + ":0")
    # TransformFeature, DF_HWC (.cc:366-389): data[((b*19+i)*19+j)*17+d] = (T)feature[b][d][i][j]
    nhwc = np.transpose(planes, (0, 2, 3, 1))
    feed_arr = (nhwc != 0) if x.dtype == tf.bool else nhwc.astype(x.dtype.as_numpy_dtype)
    fetches = [graph.get_tensor_by_name(o + ":0") for o in outputs] + \
              [graph.get_tensor_by_name(f if ":" in f else f + ":0") for f in args.fetch]
    pol, val, extra = [], [], [[] for _ in args.fetch]
    info = describe_inference_subgraph(meta_graph.graph_def, input_name, outputs)
    unfed = [p for p, op in info["placeholders"].items() if op == "Placeholder" and p != input_name and p + ":0" not in extra_feeds]
    if unfed:
        print(f"note: the inference sub-graph has plain placeholders besides {input_name}: {unfed} — feed them with --feed "
              "(the reference engine cannot: it feeds nn_input only, DlTFNetworkEvaluator.cc:295)", file=sys.stderr)
    side = {graph.get_tensor_by_name(k): v for k, v in extra_feeds.items()}
    for lo in range(0, n, args.batch):
        out = sess.run(fetches, feed_dict={x: feed_arr[lo:lo + args.batch], **side})
        pol.append(np.asarray(out[0], np.float32).reshape(-1, 362))
        val.append(np.asarray(out[1], np.float32).reshape(-1))
        for k in range(len(args.fetch)):
            extra[k].append(np.asarray(out[2 + k]))
    pol, val = np.concatenate(pol), np.concatenate(val)
    v64 = val.astype(np.float64)                              # DlTensorUtil::GetValue widening, then .cc:300-310
    v64 = -v64 if vt == 1 else (1 - v64 if vt == 2 else v64)
    info.update({"producer": "tensorflow", "extra_feeds": {k: v for k, v in extra_feeds.items()}, "tensorflow_version": tf.__version__, "input": input_name, "outputs": outputs,
                 "input_dtype": x.dtype.name, "value_transform": vt, "positions": n, "batch": args.batch,
                 "policy_row_sums_minmax": [float(pol.sum(1).min()), float(pol.sum(1).max())],
                 "value_minmax": [float(val.min()), float(val.max())]})
    save = {"planes": planes, "policy": pol, "value": val, "value_transformed": v64, "info": np.array(json.dumps(info))}
    if "packed" in data.files:
        save["packed"] = data["packed"]
    for name, parts in zip(args.fetch, extra):
        save["fetch:" + name] = np.concatenate(parts)
    np.savez_compressed(args.out, **save)
    print(json.dumps(info, indent=1))
    print(f"wrote {args.out}: {n} positions")


if __name__ == "__main__":
    sys.exit(main())
