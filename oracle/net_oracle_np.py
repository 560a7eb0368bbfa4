"""TEST INFRASTRUCTURE — the floating-point oracle's SECOND reading: a torch-free float64 dataflow interpreter.

oracle/net_oracle.py evaluates the graph with torch fp32 ops; this module evaluates the same GraphDef with plain numpy
in float64, written from TensorFlow's op definitions (tensorflow/core/ops: nn_ops.cc Conv2D / BiasAdd / FusedBatchNorm
/ Relu / Softmax, math_ops.cc MatMul / Add / Tanh / Cast, array_ops.cc Reshape / Identity / Placeholder,
control_flow_ops.cc Switch / Merge) and sharing no arithmetic with it: the convolution is an explicit sum over filter
taps of padded slices (SAME padding for stride 1 puts floor((k-1)/2) zeros in front, TensorFlow's rule), batch-norm is
the inference formula spelled out, Switch/Merge follow TensorFlow's dead-tensor semantics (so graphs exported with a
`training` placeholder and tf.cond are evaluated, not pattern-matched).  tests/test_oracle_float.py requires the two
readings to agree on every graph shape the loader accepts; a misreading of TensorFlow would have to be made twice,
independently, to pass.

The reference runs the graph inside tensorflow::Session::Run (funcapproximator/DlTFNetworkEvaluator.cc:295);
TensorFlow and the bootstrap checkpoint are not available offline: parity against the real evaluator stays UNPINNED
until tools/tf_dump_reference.py has been run where they are (README.md).
"""
import numpy as np

from . import tf_reader

DEAD = object()  # TensorFlow's "dead tensor": the untaken side of a Switch


def _split(name):
    """'^ctl' -> (None, 0); 'node:1' -> ('node', 1); 'node' -> ('node', 0)"""
    if name.startswith("^"):
        return None, 0
    head, sep, tail = name.rpartition(":")
    if sep and tail.isdigit():
        return head, int(tail)
    return name, 0


class NumpyGraphOracle:
    def __init__(self, graph_path, ckpt_prefix, input_name="feature_input",
                 outputs=("resnet/tower_0/policy_head/policy_predict", "resnet/tower_0/value_head/reward_predict")):
        self.nodes, self.saver = tf_reader.read_graph(graph_path)
        # ckpt_prefix None: a frozen graph (every variable folded into a Const node), nothing to restore
        self.vars = {k: v.astype(np.float64) for k, v in tf_reader.read_bundle(ckpt_prefix).items()} if ckpt_prefix else {}
        self.input_name = input_name
        self.outputs = list(outputs)

    # ---- ops: every function returns a tuple of outputs
    @staticmethod
    def _conv2d(x, w, attrs):
        fmt = attrs.get("data_format", {}).get("s", b"NHWC")
        if fmt != b"NHWC":
            raise NotImplementedError(f"Conv2D data_format {fmt!r}")
        if attrs["padding"]["s"] != b"SAME":
            raise NotImplementedError("Conv2D padding")
        if any(s != 1 for s in attrs.get("strides", {}).get("list_i", [1])) or \
           any(d != 1 for d in attrs.get("dilations", {}).get("list_i", [1])):
            raise NotImplementedError("Conv2D stride/dilation")
        kh, kw, cin, cout = w.shape
        n, h, wd, c = x.shape
        assert c == cin, (x.shape, w.shape)
        ph, pw = (kh - 1) // 2, (kw - 1) // 2  # zeros in front; the remainder goes behind
        xp = np.zeros((n, h + kh - 1, wd + kw - 1, c), np.float64)
        xp[:, ph:ph + h, pw:pw + wd, :] = x
        out = np.zeros((n, h, wd, cout), np.float64)
        for dy in range(kh):
            for dx in range(kw):
                out += xp[:, dy:dy + h, dx:dx + wd, :] @ w[dy, dx]
        return out

    def _op(self, node, ins):
        op, a = node["op"], node["attrs"]
        if op in ("Identity", "StopGradient", "Snapshot", "ReadVariableOp", "PreventGradient"):
            return (ins[0],)
        if op == "Cast":
            return (np.asarray(ins[0], np.float64),)
        if op == "Conv2D":
            return (self._conv2d(ins[0], ins[1], a),)
        if op == "BiasAdd":
            if a.get("data_format", {}).get("s", b"NHWC") != b"NHWC":
                raise NotImplementedError("BiasAdd data_format")
            return (ins[0] + ins[1].reshape((1,) * (ins[0].ndim - 1) + (-1,)),)
        if op in ("Add", "AddV2"):
            return (ins[0] + ins[1],)
        if op in ("FusedBatchNorm", "FusedBatchNormV2", "FusedBatchNormV3"):
            if a.get("is_training", {}).get("b", True):
                raise NotImplementedError(f"{node['name']}: training-mode batch-norm on the inference path")
            if a.get("data_format", {}).get("s", b"NHWC") != b"NHWC":
                raise NotImplementedError("FusedBatchNorm data_format")
            x, scale, offset, mean, var = ins[:5]
            eps = a.get("epsilon", {}).get("f", 1e-4)  # the op's default attribute value
            inv = 1.0 / np.sqrt(var + np.float64(np.float32(eps)))
            return ((x - mean) * (inv * scale) + offset, mean, var, mean, var)
        if op == "Relu":
            return (np.maximum(ins[0], 0.0),)
        if op == "Tanh":
            return (np.tanh(ins[0]),)
        if op == "Softmax":
            z = ins[0] - ins[0].max(axis=-1, keepdims=True)
            e = np.exp(z)
            return (e / e.sum(axis=-1, keepdims=True),)
        if op == "Reshape":
            shape = [int(v) for v in np.asarray(ins[1]).reshape(-1)]
            return (np.reshape(ins[0], shape),)   # row-major, like TensorFlow
        if op == "Squeeze":
            dims = a.get("squeeze_dims", {}).get("list_i", [])
            return (np.squeeze(ins[0], axis=tuple(dims) if dims else None),)
        if op == "ExpandDims":
            return (np.expand_dims(ins[0], int(np.asarray(ins[1]).reshape(-1)[0])),)
        if op == "MatMul":
            x = ins[0].T if a.get("transpose_a", {}).get("b", False) else ins[0]
            w = ins[1].T if a.get("transpose_b", {}).get("b", False) else ins[1]
            return (x @ w,)
        raise NotImplementedError(f"numpy oracle: op {op} at {node['name']}")

    def _value(self, ref, cache, feeds):
        name, idx = _split(ref)
        key = (name, idx)
        if key in cache:
            return cache[key]
        if name in feeds:
            cache[key] = feeds[name]
            return cache[key]
        node = self.nodes[name]
        op, a = node["op"], node["attrs"]
        data_in = [i for i in node["inputs"] if not i.startswith("^")]
        if op in ("VariableV2", "Variable", "VarHandleOp"):
            outs = (self.vars[name],)
        elif op == "Const":
            outs = (np.asarray(a["value"]["tensor"]),)
        elif op == "Placeholder":
            raise KeyError(f"placeholder '{name}' is not fed")
        elif op == "PlaceholderWithDefault":
            outs = (self._value(data_in[0], cache, feeds),)
        elif op == "Switch":
            data = self._value(data_in[0], cache, feeds)
            pred = self._value(data_in[1], cache, feeds)
            if data is DEAD or pred is DEAD:
                outs = (DEAD, DEAD)
            else:
                taken = bool(np.asarray(pred).reshape(-1)[0])
                outs = (DEAD, data) if taken else (data, DEAD)   # output 0 = false side, output 1 = true side
        elif op == "Merge":
            live = [v for v in (self._value(i, cache, feeds) for i in data_in) if v is not DEAD]
            outs = (live[0] if live else DEAD,)
        else:
            ins = [self._value(i, cache, feeds) for i in data_in]
            outs = (DEAD,) * 6 if any(v is DEAD for v in ins) else self._op(node, ins)
        for k, v in enumerate(outs):
            cache[(name, k)] = v
        return cache[key]

    def run(self, nhwc, feeds=None, fetch=None):
        """nhwc: bool/int8/float array [n,19,19,17]; feeds: extra placeholder values (e.g. {'training': False}).
        Returns (policy float64 [n,362], value float64 [n]) (+ dict of fetched tensors)."""
        f = {self.input_name: np.ascontiguousarray(nhwc)}
        for k, v in (feeds or {}).items():
            f[_split(k)[0]] = np.asarray(v)
        cache = {}
        pol = self._value(self.outputs[0], cache, f)
        val = self._value(self.outputs[1], cache, f)
        if pol is DEAD or val is DEAD:
            raise RuntimeError("an output is dead: wrong feed for the training switch?")
        if fetch:
            return pol, val.reshape(-1), {k: self._value(k, cache, f) for k in fetch}
        return pol, val.reshape(-1)

    def evaluate_reference(self, planes_chw, value_transform=0, feeds=None):
        """DlTFNetworkEvaluator<bool>::Evaluate (.cc:287-318) in float64: TransformFeature(DF_HWC), Session::Run,
        value_transform."""
        nhwc = np.ascontiguousarray(np.transpose(np.asarray(planes_chw) != 0, (0, 2, 3, 1)))  # [b][i][j][d], .cc:366-389
        pol, val = self.run(nhwc, feeds)
        if value_transform == 1:
            val = -val
        elif value_transform == 2:
            val = 1 - val
        return pol, val
