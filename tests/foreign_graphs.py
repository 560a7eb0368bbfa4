"""TEST INFRASTRUCTURE — graphs and checkpoints the product's loader did not grow up with.

Every parity test of round 1 loaded files written by unrealgo_b200/tfformat.py: one author's reading of TensorFlow on
both sides.  This module writes the artefacts a TensorFlow-1.x training script would leave, through Google's protobuf
runtime (tests/tfproto.py) and an SSTable writer of its own that follows LevelDB's TableBuilder as TensorFlow's
BundleWriter drives it (one 256 KB data block, restart interval 16, shortest-separator index keys, non-float entries
such as global_step beside the weights).  Graph idioms covered (SURVEY.md 8(c) "unknowns to resolve from the real
.meta"):

  bn="cond"      tf.layers.batch_normalization(training=<PlaceholderWithDefault>, fused=True): both branches under
                 cond/ with Switch/Merge, the training-mode FusedBatchNorm beside the inference one
  bn="v3"        FusedBatchNormV3 (TF >= 1.15 exporters), bn="v1" FusedBatchNorm
  scale=False / center=False   gamma / beta are Const nodes, not variables
  conv_bias      Conv2D -> BiasAdd -> batch-norm
  resource_vars  VarHandleOp + ReadVariableOp instead of VariableV2 + Identity
  input_dtype    bool (what the engine feeds, search/UctBoardEvaluator.h:25) or float
                 (funcapproximator/test/DlTFTestVariableBatchSize.cc:81)
  data_format    "NCHW" in-graph must be refused by the loader, loudly
  filters, value_channels, policy_channels, value_hidden, bn_eps, scope names
"""
import os
import struct

import numpy as np

from tests import tfproto as P


# ----------------------------------------------------------------------------------------------- graph
class G:
    def __init__(self, resource_vars=False):
        self.nodes = []
        self.variables = {}   # checkpoint key -> shape
        self.resource_vars = resource_vars

    def node(self, name, op, inputs=(), **attrs):
        n = P.NodeDef(name=name, op=op, input=list(inputs))
        for k, v in attrs.items():
            a = n.attr[k.lstrip("_") if k.startswith("__") else k]
            if isinstance(v, bool):
                a.b = v
            elif isinstance(v, int):
                a.i = v
            elif isinstance(v, float):
                a.f = v
            elif isinstance(v, bytes):
                a.s = v
            elif isinstance(v, tuple) and v[0] == "type":
                a.type = v[1]
            elif isinstance(v, tuple) and v[0] == "ints":
                a.list.i.extend(v[1])
            elif isinstance(v, tuple) and v[0] == "shape":
                for d in v[1]:
                    a.shape.dim.add().size = d
            elif isinstance(v, P.TensorProto):
                a.tensor.CopyFrom(v)
            else:
                raise TypeError((k, v))
        self.nodes.append(n)
        return name

    @staticmethod
    def const_tensor(values, dtype=P.DT_FLOAT, splat=False):
        t = P.TensorProto(dtype=dtype)
        arr = np.asarray(values)
        for d in arr.shape:
            t.tensor_shape.dim.add().size = int(d)
        if splat:   # TensorFlow stores a constant filled with one value as a single float_val
            t.float_val.append(float(arr.reshape(-1)[0]))
        elif dtype == P.DT_FLOAT:
            t.tensor_content = arr.astype("<f4").tobytes()
        elif dtype == P.DT_INT32:
            t.tensor_content = arr.astype("<i4").tobytes()
        elif dtype == P.DT_BOOL:
            t.bool_val.extend(bool(x) for x in arr.reshape(-1))
        return t

    def variable(self, name, shape):
        """returns the tensor name a consumer reads"""
        self.variables[name] = tuple(shape)
        if self.resource_vars:
            self.node(name, "VarHandleOp", (), dtype=("type", P.DT_FLOAT), shape=("shape", shape), shared_name=name.encode(),
                      container=b"")
            self.node(name + "/Initializer/random_normal", "Const", (), dtype=("type", P.DT_FLOAT),
                      value=self.const_tensor(np.zeros(1, np.float32)))
            self.node(name + "/Assign", "AssignVariableOp", (name, name + "/Initializer/random_normal"),
                      dtype=("type", P.DT_FLOAT))
            return self.node(name + "/Read/ReadVariableOp", "ReadVariableOp", (name,), dtype=("type", P.DT_FLOAT))
        self.node(name, "VariableV2", (), dtype=("type", P.DT_FLOAT), shape=("shape", shape), container=b"", shared_name=b"")
        self.node(name + "/Initializer/random_normal", "Const", (), dtype=("type", P.DT_FLOAT),
                  value=self.const_tensor(np.zeros(1, np.float32)))
        self.node(name + "/Assign", "Assign", (name, name + "/Initializer/random_normal"), T=("type", P.DT_FLOAT),
                  use_locking=True, validate_shape=True)
        return self.node(name + "/read", "Identity", (name,), T=("type", P.DT_FLOAT),
                         _class=b"loc:@" + name.encode())


def _batch_norm(g, scope, vscope, x, c, eps, bn, training, scale, center, data_format):
    """tf.layers.batch_normalization(x, training=..., fused=True, scale=..., center=..., epsilon=eps) as TF 1.x emits it"""
    op = {"v1": "FusedBatchNorm", "v3": "FusedBatchNormV3", "cond": "FusedBatchNorm", "cond_v3": "FusedBatchNormV3"}[bn]
    if scale:
        gamma = g.variable(f"{vscope}/gamma", (c,))
    else:
        gamma = g.node(f"{scope}/Const", "Const", (), dtype=("type", P.DT_FLOAT),
                       value=g.const_tensor(np.ones(c, np.float32), splat=True))
    if center:
        beta = g.variable(f"{vscope}/beta", (c,))
    else:
        beta = g.node(f"{scope}/Const_1", "Const", (), dtype=("type", P.DT_FLOAT),
                      value=g.const_tensor(np.zeros(c, np.float32), splat=True))
    mean = g.variable(f"{vscope}/moving_mean", (c,))
    var = g.variable(f"{vscope}/moving_variance", (c,))
    extra = {"U": ("type", P.DT_FLOAT)} if op.endswith("V3") else {}
    if not bn.startswith("cond"):
        return g.node(f"{scope}/{op}", op, (x, gamma, beta, mean, var), T=("type", P.DT_FLOAT), epsilon=float(eps),
                      data_format=data_format.encode(), is_training=False, **extra)
    c_ = f"{scope}/cond"
    g.node(f"{c_}/Switch", "Switch", (training, training), T=("type", P.DT_BOOL))
    g.node(f"{c_}/switch_t", "Identity", (f"{c_}/Switch:1",), T=("type", P.DT_BOOL))
    g.node(f"{c_}/switch_f", "Identity", (f"{c_}/Switch",), T=("type", P.DT_BOOL))
    pred = g.node(f"{c_}/pred_id", "Identity", (training,), T=("type", P.DT_BOOL))
    # training branch: batch statistics (empty mean/variance inputs), is_training=true
    g.node(f"{c_}/Const", "Const", (f"^{c_}/switch_t",), dtype=("type", P.DT_FLOAT), value=g.const_tensor(np.zeros(0, np.float32)))
    g.node(f"{c_}/Const_1", "Const", (f"^{c_}/switch_t",), dtype=("type", P.DT_FLOAT), value=g.const_tensor(np.zeros(0, np.float32)))
    t = f"{c_}/{op}"
    for k, src in enumerate((x, gamma, beta)):
        g.node(f"{t}/Switch" + (f"_{k}" if k else ""), "Switch", (src, pred), T=("type", P.DT_FLOAT))
    g.node(t, op, (f"{t}/Switch:1", f"{t}/Switch_1:1", f"{t}/Switch_2:1", f"{c_}/Const", f"{c_}/Const_1"),
           T=("type", P.DT_FLOAT), epsilon=float(eps), data_format=data_format.encode(), is_training=True, **extra)
    # inference branch: moving statistics, is_training=false
    f = f"{c_}/{op}_1"
    for k, src in enumerate((x, gamma, beta, mean, var)):
        g.node(f"{f}/Switch" + (f"_{k}" if k else ""), "Switch", (src, pred), T=("type", P.DT_FLOAT))
    g.node(f, op, (f"{f}/Switch", f"{f}/Switch_1", f"{f}/Switch_2", f"{f}/Switch_3", f"{f}/Switch_4"),
           T=("type", P.DT_FLOAT), epsilon=float(eps), data_format=data_format.encode(), is_training=False, **extra)
    out = g.node(f"{c_}/Merge", "Merge", (f, t), T=("type", P.DT_FLOAT), N=2)
    # the moving-average update clutter a training graph carries (never on the inference path)
    g.node(f"{c_}/Merge_1", "Merge", (f"{f}:1", f"{t}:1"), T=("type", P.DT_FLOAT), N=2)
    g.node(f"{scope}/cond_1/Switch", "Switch", (training, training), T=("type", P.DT_BOOL))
    g.node(f"{scope}/AssignMovingAvg/decay", "Const", (), dtype=("type", P.DT_FLOAT),
           value=g.const_tensor(np.float32([0.01]).reshape(()), splat=True))
    return out


def _conv_bn(g, scope, vscope, x, cin, cout, k, o, relu=True):
    w = g.variable(f"{vscope}/conv2d/kernel", (k, k, cin, cout))
    strides = [1, 1, 1, 1]
    y = g.node(f"{scope}/conv2d/Conv2D", "Conv2D", (x, w), T=("type", P.DT_FLOAT), strides=("ints", strides),
               padding=b"SAME", data_format=o["data_format"].encode(), dilations=("ints", [1, 1, 1, 1]),
               use_cudnn_on_gpu=True)
    if o["conv_bias"]:
        b = g.variable(f"{vscope}/conv2d/bias", (cout,))
        y = g.node(f"{scope}/conv2d/BiasAdd", "BiasAdd", (y, b), T=("type", P.DT_FLOAT), data_format=o["data_format"].encode())
    y = _batch_norm(g, f"{scope}/batch_normalization", f"{vscope}/batch_normalization", y, cout, o["bn_eps"], o["bn"],
                    o["training"], o["scale"], o["center"], o["data_format"])
    if relu:
        y = g.node(f"{scope}/Relu", "Relu", (y,), T=("type", P.DT_FLOAT))
    return y


def _dense(g, scope, vscope, x, cin, cout):
    w = g.variable(f"{vscope}/kernel", (cin, cout))
    b = g.variable(f"{vscope}/bias", (cout,))
    y = g.node(f"{scope}/MatMul", "MatMul", (x, w), T=("type", P.DT_FLOAT), transpose_a=False, transpose_b=False)
    return g.node(f"{scope}/BiasAdd", "BiasAdd", (y, b), T=("type", P.DT_FLOAT), data_format=b"NHWC")


def build(blocks=2, filters=64, bn="cond", conv_bias=False, scale=True, center=True, resource_vars=False,
          input_dtype=P.DT_BOOL, data_format="NHWC", policy_channels=2, value_channels=1, value_hidden=256, bn_eps=1e-5,
          input_name="feature_input", policy_name="resnet/tower_0/policy_head/policy_predict",
          value_name="resnet/tower_0/value_head/reward_predict", training_placeholder="with_default"):
    """returns (G, MetaGraphDef message)"""
    g = G(resource_vars)
    t = "resnet/tower_0"
    o = {"bn": bn, "conv_bias": conv_bias, "scale": scale, "center": center, "bn_eps": bn_eps, "data_format": data_format,
         "training": None}
    if bn.startswith("cond"):
        if training_placeholder == "with_default":   # the reference feeds only nn_input (.cc:295): the switch must default
            g.node("training/input", "Const", (), dtype=("type", P.DT_BOOL), value=g.const_tensor(np.array(False), P.DT_BOOL))
            o["training"] = g.node("training", "PlaceholderWithDefault", ("training/input",), dtype=("type", P.DT_BOOL),
                                   shape=("shape", []))
        else:
            o["training"] = g.node("training", "Placeholder", (), dtype=("type", P.DT_BOOL), shape=("shape", []))
    x = g.node(input_name, "Placeholder", (), dtype=("type", input_dtype), shape=("shape", [-1, 19, 19, 17]))
    if input_dtype != P.DT_FLOAT:
        x = g.node(f"{t}/ToFloat", "Cast", (x,), SrcT=("type", input_dtype), DstT=("type", P.DT_FLOAT))
    x = _conv_bn(g, f"{t}/init_conv", "resnet/init_conv", x, 17, filters, 3, o)
    for k in range(blocks):
        u, uv = f"{t}/res_block_{k}", f"resnet/res_block_{k}"
        y = _conv_bn(g, f"{u}/a", f"{uv}/a", x, filters, filters, 3, o)
        y = _conv_bn(g, f"{u}/b", f"{uv}/b", y, filters, filters, 3, o, relu=False)
        y = g.node(f"{u}/add", "Add", (x, y), T=("type", P.DT_FLOAT))       # skip first, branch second
        x = g.node(f"{u}/out", "Relu", (y,), T=("type", P.DT_FLOAT))
    p = _conv_bn(g, f"{t}/policy_head/conv", "resnet/policy_head/conv", x, filters, policy_channels, 1, o)
    shp = g.node(f"{t}/policy_head/Reshape/shape", "Const", (), dtype=("type", P.DT_INT32),
                 value=g.const_tensor(np.int32([-1, 361 * policy_channels]), P.DT_INT32))
    p = g.node(f"{t}/policy_head/Reshape", "Reshape", (p, shp), T=("type", P.DT_FLOAT), Tshape=("type", P.DT_INT32))
    p = _dense(g, f"{t}/policy_head/dense", "resnet/policy_head/dense", p, 361 * policy_channels, 362)
    g.node(policy_name, "Softmax", (p,), T=("type", P.DT_FLOAT))
    v = _conv_bn(g, f"{t}/value_head/conv", "resnet/value_head/conv", x, filters, value_channels, 1, o)
    shv = g.node(f"{t}/value_head/Reshape/shape", "Const", (), dtype=("type", P.DT_INT32),
                 value=g.const_tensor(np.int32([-1, 361 * value_channels]), P.DT_INT32))
    v = g.node(f"{t}/value_head/Reshape", "Reshape", (v, shv), T=("type", P.DT_FLOAT), Tshape=("type", P.DT_INT32))
    v = _dense(g, f"{t}/value_head/dense", "resnet/value_head/dense", v, 361 * value_channels, value_hidden)
    v = g.node(f"{t}/value_head/dense/Relu", "Relu", (v,), T=("type", P.DT_FLOAT))
    v = _dense(g, f"{t}/value_head/dense_1", "resnet/value_head/dense_1", v, value_hidden, 1)
    v = g.node(f"{t}/value_head/Tanh", "Tanh", (v,), T=("type", P.DT_FLOAT))
    g.node(value_name, "Identity", (v,), T=("type", P.DT_FLOAT))
    # what a training script leaves beside the inference path
    g.node("global_step", "VariableV2", (), dtype=("type", P.DT_INT64), shape=("shape", []), container=b"", shared_name=b"")
    g.node("save/Const", "Const", (), dtype=("type", P.DT_STRING), value=P.TensorProto(dtype=P.DT_STRING, string_val=[b"model"]))
    g.node("save/RestoreV2", "RestoreV2", ("save/Const",))
    g.node("save/restore_all", "NoOp", ("^save/RestoreV2",))
    g.node("gradients/Shape", "Const", (), dtype=("type", P.DT_INT32), value=g.const_tensor(np.int32([0]), P.DT_INT32))
    g.node("train/Adam", "NoOp", ())
    meta = P.MetaGraphDef()
    meta.meta_info_def.tensorflow_version = "1.12.0"
    meta.meta_info_def.tensorflow_git_version = "v1.12.0-0-ga6d8ffae09"
    meta.graph_def.node.extend(g.nodes)
    meta.graph_def.versions.producer = 27
    meta.saver_def.filename_tensor_name = "save/Const:0"
    meta.saver_def.save_tensor_name = "save/control_dependency:0"
    meta.saver_def.restore_op_name = "save/restore_all"
    meta.saver_def.max_to_keep = 5
    meta.saver_def.keep_checkpoint_every_n_hours = 10000.0
    meta.saver_def.version = 2
    meta.collection_def["trainable_variables"].bytes_list.value.extend(k.encode() for k in sorted(g.variables))
    return g, meta


# ----------------------------------------------------------------------------------------------- bundle
def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _crc(data):
    import oracle
    c = oracle.lib().og_crc32c(bytes(data), len(data))   # the oracle's own C implementation (a third CRC-32C)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _shortest_separator(a, b):
    """leveldb BytewiseComparator::FindShortestSeparator"""
    n = min(len(a), len(b))
    i = 0
    while i < n and a[i] == b[i]:
        i += 1
    if i < n and a[i] < 0xFF and a[i] + 1 < b[i]:
        return a[:i] + bytes([a[i] + 1])
    return a


def _short_successor(a):
    for i, ch in enumerate(a):
        if ch != 0xFF:
            return a[:i] + bytes([ch + 1])
    return a


class _Block:
    def __init__(self, restart_interval):
        self.buf, self.restarts, self.n, self.last, self.ri = bytearray(), [0], 0, b"", restart_interval

    def add(self, key, value):
        shared = 0
        if self.n and self.n % self.ri == 0:
            self.restarts.append(len(self.buf))
        elif self.n:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last, self.n = key, self.n + 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def write_bundle(prefix, tensors, extra_entries=True, block_size=262144, num_shards=1, endianness=0):
    """TensorFlow's BundleWriter: data file = tensors in key order, back to back; index = SSTable of BundleEntryProto.
    num_shards > 1: what Saver(sharded=True) leaves after MergeBundles — one merged index, the tensors spread over
    <prefix>.data-0000k-of-0000N, every entry naming its shard"""
    items = {k: np.ascontiguousarray(v, "<f4") for k, v in tensors.items()}
    other = {}
    if extra_entries:
        other["global_step"] = (P.DT_INT64, (), struct.pack("<q", 123456))
        other["beta1_power"] = (P.DT_FLOAT, (), struct.pack("<f", 0.9))   # optimizer slot: a float the net does not use
    keys = sorted(list(items) + list(other))
    entries = []
    files = [open(prefix + ".data-%05d-of-%05d" % (k, num_shards), "wb") for k in range(num_shards)]
    for i, k in enumerate(keys):
        if k in items:
            dtype, shape, raw = P.DT_FLOAT, items[k].shape, items[k].tobytes()
        else:
            dtype, shape, raw = other[k]
        shard = (i * 7) % num_shards
        df = files[shard]
        e = P.BundleEntryProto(dtype=dtype, offset=df.tell(), size=len(raw), crc32c=_crc(raw))
        if shard:
            e.shard_id = shard
        for d in shape:
            e.shape.dim.add().size = int(d)
        df.write(raw)
        entries.append((k.encode(), e.SerializeToString()))
    for df in files:
        df.close()
    header = P.BundleHeaderProto(num_shards=num_shards, endianness=endianness)   # 0 = little, 1 = big
    header.version.producer = 1
    rows = [(b"", header.SerializeToString())] + entries
    with open(prefix + ".index", "wb") as f:
        def emit(contents):
            off = f.tell()
            f.write(contents + b"\x00" + struct.pack("<I", _crc(contents + b"\x00")))
            return _varint(off) + _varint(len(contents))
        index = _Block(1)
        blk = _Block(16)
        pending = None   # (last key of the finished block, its handle)
        for key, val in rows:
            if pending is not None:
                index.add(_shortest_separator(pending[0], key), pending[1])
                pending = None
            blk.add(key, val)
            if len(blk.buf) >= block_size:
                pending = (key, emit(blk.finish()))
                blk = _Block(16)
        if blk.n:
            pending = (blk.last, emit(blk.finish()))
        if pending is not None:
            index.add(_short_successor(pending[0]), pending[1])
        meta_h = emit(_Block(16).finish())
        index_h = emit(index.finish())
        footer = meta_h + index_h
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57))


def random_weights(variables, seed):
    """trained-looking statistics: He-normal kernels, gamma around 1 (small on a block's second conv), spread-out
    moving statistics"""
    rng = np.random.default_rng(seed)
    out = {}
    for name in sorted(variables):
        shape = variables[name]
        if name.endswith("/kernel"):
            out[name] = (rng.standard_normal(shape) * np.sqrt(2.0 / np.prod(shape[:-1]))).astype(np.float32)
        elif name.endswith("/b/batch_normalization/gamma"):
            out[name] = rng.uniform(0.15, 0.35, shape).astype(np.float32)
        elif name.endswith("/gamma"):
            out[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif name.endswith("/moving_variance"):
            out[name] = rng.uniform(0.3, 2.0, shape).astype(np.float32)
        else:
            out[name] = (rng.standard_normal(shape) * 0.1).astype(np.float32)
    return out


def write(directory, name="model.ckpt-123456", seed=7, small_blocks=False, num_shards=1, **kw):
    """-> (meta path, checkpoint prefix)"""
    os.makedirs(directory, exist_ok=True)
    g, meta = build(**kw)
    mpath = os.path.join(directory, name + ".meta")
    with open(mpath, "wb") as f:
        f.write(meta.SerializeToString())
    prefix = os.path.join(directory, name)
    write_bundle(prefix, random_weights(g.variables, seed), block_size=700 if small_blocks else 262144,
                 num_shards=num_shards)
    return mpath, prefix


def freeze(meta, weights):
    """what tensorflow.python.tools.freeze_graph does to the inference graph: every variable becomes a float Const of
    the same name holding its checkpoint value (the `read` Identity / ReadVariableOp consumers stay).  -> GraphDef"""
    gd = P.GraphDef()
    gd.CopyFrom(meta.graph_def)
    for n in gd.node:
        if n.op in ("VariableV2", "Variable", "VarHandleOp") and n.name in weights:
            w = np.ascontiguousarray(weights[n.name], "<f4")
            n.op = "Const"
            n.ClearField("attr")
            n.attr["dtype"].type = P.DT_FLOAT
            t = n.attr["value"].tensor
            t.dtype = P.DT_FLOAT
            for d in w.shape:
                t.tensor_shape.dim.add().size = int(d)
            if w.size == 1:
                t.float_val.append(float(w.reshape(-1)[0]))   # TensorFlow stores single values this way
            else:
                t.tensor_content = w.tobytes()
        elif n.op == "ReadVariableOp":
            n.op = "Identity"
            n.ClearField("attr")
            n.attr["T"].type = P.DT_FLOAT
    return gd
