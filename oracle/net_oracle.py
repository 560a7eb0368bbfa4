"""TEST INFRASTRUCTURE — fp32 CPU evaluation of the network graph, the floating-point half of the oracle.

The reference runs the graph inside tensorflow::Session::Run (funcapproximator/DlTFNetworkEvaluator.cc:295);
TensorFlow (un-vendored third-party dependency, no version pinned: README.md:40) is not available here, so this
module *interprets* the GraphDef node by node with torch fp32 CPU ops that restate the published semantics of
each TensorFlow op (Conv2D NHWC/SAME, FusedBatchNorm inference: (x-mean)*rsqrt(var+eps)*gamma+beta, BiasAdd,
Add, Relu, Reshape, MatMul, Softmax, Tanh, Cast, Identity, Switch/Merge with dead-branch semantics).  It is independent of the product's C++ pattern
matcher: it evaluates whatever graph the .meta describes, starting from the two nn_outputs nodes.

"restatement, not TensorFlow": parity against the real TF evaluator + bootstrap checkpoint is UNPINNED until
both are available.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import tf_reader


def _split(name):
    """'node:1' -> ('node', 1); 'node' -> ('node', 0); control inputs ('^node') are filtered by the caller"""
    if name.startswith("^"):
        name = name[1:]
    head, sep, tail = name.rpartition(":")
    return (head, int(tail)) if sep and tail.isdigit() else (name, 0)


def _strip(name):
    return _split(name)[0]


class _Dead(Exception):
    """the untaken side of a Switch was asked for (TensorFlow's dead tensor): Merge catches it and takes the other input"""


class GraphOracle:
    def __init__(self, graph_path, ckpt_prefix, input_name="feature_input",
                 outputs=("resnet/tower_0/policy_head/policy_predict", "resnet/tower_0/value_head/reward_predict")):
        self.nodes, self.saver = tf_reader.read_graph(graph_path)
        # ckpt_prefix None: a frozen graph (every variable folded into a Const node), nothing to restore
        self.vars = {k: torch.from_numpy(v) for k, v in tf_reader.read_bundle(ckpt_prefix).items()} if ckpt_prefix else {}
        self.input_name = input_name
        self.outputs = list(outputs)

    def _data_inputs(self, node):
        return [i for i in node["inputs"] if not i.startswith("^")]

    def _eval(self, name, cache):
        name, idx = _split(name)
        if name in cache:
            return cache[name]
        node = self.nodes[name]
        op = node["op"]
        ins = self._data_inputs(node)
        a = node["attrs"]
        ev = lambda k: self._eval(ins[k], cache)
        if op == "Switch":  # (data, pred) -> output 0 when pred is false, output 1 when true; never cached (two outputs)
            taken = bool(torch.as_tensor(ev(1)).reshape(-1)[0])
            if idx != int(taken):
                raise _Dead(name)
            return ev(0)
        if op == "Merge":
            out = None
            for k in range(len(ins)):
                try:
                    out = ev(k)
                    break
                except _Dead:
                    continue
            if out is None:
                raise _Dead(name)
        elif op in ("VariableV2", "Variable", "VarHandleOp"):
            out = self.vars[name]
        elif op in ("Identity", "StopGradient", "ReadVariableOp", "PlaceholderWithDefault", "Snapshot"):
            out = ev(0)
        elif op == "Const":
            out = a["value"]["tensor"]
            out = torch.from_numpy(np.array(out)) if isinstance(out, np.ndarray) else out
        elif op == "Cast":
            out = ev(0).to(torch.float32)
        elif op == "Conv2D":
            assert a.get("data_format", {}).get("s", b"NHWC") == b"NHWC"
            assert a["padding"]["s"] == b"SAME" and all(s == 1 for s in a["strides"]["list_i"])
            x, w = ev(0), ev(1)  # NHWC, HWIO
            kh = w.shape[0]
            y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=kh // 2)
            out = y.permute(0, 2, 3, 1).contiguous()
        elif op in ("FusedBatchNorm", "FusedBatchNormV2", "FusedBatchNormV3"):
            assert idx == 0, "only the normalised output is on the inference path"
            assert not a.get("is_training", {}).get("b", True), "inference graph expected"
            x, g, b, m, v = (ev(k) for k in range(5))
            eps = a.get("epsilon", {}).get("f", 1e-4)
            out = (x - m) * torch.rsqrt(v + eps) * g + b
        elif op in ("BiasAdd", "Add", "AddV2"):
            out = ev(0) + ev(1)
        elif op == "Relu":
            out = torch.relu(ev(0))
        elif op == "Reshape":
            shape = [int(s) for s in np.asarray(ev(1)).reshape(-1)]
            out = ev(0).reshape(shape)
        elif op == "MatMul":
            assert not a.get("transpose_a", {}).get("b", False) and not a.get("transpose_b", {}).get("b", False)
            out = ev(0) @ ev(1)
        elif op == "Softmax":
            out = torch.softmax(ev(0), dim=-1)
        elif op == "Tanh":
            out = torch.tanh(ev(0))
        else:
            raise NotImplementedError(f"oracle: op {op} at {name}")
        cache[name] = out
        return out

    @torch.no_grad()
    def run(self, nhwc, fetch=None, feeds=None):
        """nhwc: bool/int8/float array [n,19,19,17] (what TransformFeature hands to TF).
        Returns (policy float32 [n,362], value float32 [n])."""
        x = torch.from_numpy(np.ascontiguousarray(nhwc)).to(torch.float32)
        cache = {self.input_name: x}
        for k, v in (feeds or {}).items():
            cache[_strip(k)] = torch.as_tensor(v)
        pol = self._eval(self.outputs[0], cache)
        val = self._eval(self.outputs[1], cache)
        if fetch:
            return pol.numpy(), val.reshape(-1).numpy(), {k: cache[_strip(k)].numpy() for k in fetch}
        return pol.numpy(), val.reshape(-1).numpy()

    def evaluate_reference(self, planes_chw, value_transform=0):
        """DlTFNetworkEvaluator<bool>::Evaluate restated end to end (.cc:287-318): TransformFeature(DF_HWC, bool),
        Session::Run, float->double GetValue, value_transform in double."""
        from . import transform_feature, value_transform as vt_fn
        nhwc = transform_feature(planes_chw, data_format=0, as_bool=True)
        pol, val = self.run(nhwc)
        return pol.astype(np.float64), vt_fn(val.astype(np.float64), value_transform)
