"""Writers for the two on-disk artefacts the reference hands to TensorFlow — without TensorFlow.

* ``write_bundle``  : TF V2 checkpoint bundle (``<prefix>.index`` LevelDB-format table of BundleEntryProto +
                      ``<prefix>.data-00000-of-00001`` raw little-endian tensors), the format restored by the Saver
                      op at funcapproximator/DlTFNetworkEvaluator.cc:170-200.
* ``write_metagraph``: a MetaGraphDef (``.meta``, read at DlTFNetworkEvaluator.cc:130) holding an AlphaGo-Zero
                      shaped inference graph in TF-1.x idiom (VariableV2 + ``/read`` Identity, Conv2D,
                      FusedBatchNorm, Relu, Add, MatMul, BiasAdd, Softmax, Tanh) plus the kind of training-side
                      clutter (save/, gradients/) a real export carries, which loaders must ignore.

The real ``bootstrap-ckpt`` is a git-LFS object that is not available offline, so benchmarks and tests run on
checkpoints produced here (random-init, seeded).  The C++ loader (csrc/tfmodel.cpp) and the independent Python
reader in oracle/ must both round-trip what this module writes.
"""
import struct

import numpy as np

DT_FLOAT, DT_INT32, DT_STRING, DT_BOOL = 1, 3, 7, 10


# ----------------------------------------------------------------------------- protobuf wire helpers
def _varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _tag(num, wire):
    return _varint((num << 3) | wire)


def pb_varint(num, v):
    return _tag(num, 0) + _varint(v)


def pb_bytes(num, b):
    if isinstance(b, str):
        b = b.encode()
    return _tag(num, 2) + _varint(len(b)) + b


def pb_f32(num, f):
    return _tag(num, 5) + struct.pack("<f", f)


def pb_fixed32(num, v):
    return _tag(num, 5) + struct.pack("<I", v)


# ----------------------------------------------------------------------------- crc32c
_CRC_TABLE = None


# bench.py's reference arm sets this to False: that process must not map the product library at all, and it writes its
# checkpoint with tensor_crc=False so that only the few KB of index blocks go through the table loop below
NATIVE_CRC = True


def crc32c(data):
    """CRC-32C (Castagnoli).  Uses libugeval's exported implementation when built, else a table loop."""
    if NATIVE_CRC:
        try:
            from . import _lib
            lib = _lib.load()
            buf = bytes(data)
            return lib.ug_crc32c(buf, len(buf))
        except (ImportError, OSError, AttributeError):
            pass
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    c = 0xFFFFFFFF
    for b in bytes(data):
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def crc_mask(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- SSTable / bundle
class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b""
        self.ri = restart_interval

    def add(self, key, value):
        shared = 0
        if self.count % self.ri == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value))
        self.buf += key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self):
        out = bytes(self.buf)
        for r in self.restarts:
            out += struct.pack("<I", r)
        return out + struct.pack("<I", len(self.restarts))


def _write_block(f, contents):
    """returns the BlockHandle (offset, size) of a block written with its 5-byte trailer"""
    off = f.tell()
    f.write(contents)
    trailer_type = b"\x00"  # kNoCompression (BundleWriter sets table::kNoCompression)
    f.write(trailer_type + struct.pack("<I", crc_mask(crc32c(contents + trailer_type))))
    return off, len(contents)


def _handle(off, size):
    return _varint(off) + _varint(size)


def write_bundle(prefix, tensors, block_size=4096, tensor_crc=True):
    """tensors: dict name -> np.ndarray (float32).  Keys are stored sorted, as TensorFlow does.  tensor_crc=False omits
    BundleEntryProto.crc32c (an optional field: readers then skip the per-tensor check)."""
    names = sorted(tensors)
    entries = []
    with open(prefix + ".data-00000-of-00001", "wb") as df:
        for name in names:
            arr = np.ascontiguousarray(tensors[name], dtype="<f4")
            raw = arr.tobytes()
            off = df.tell()
            df.write(raw)
            shape = b"".join(pb_bytes(2, pb_varint(1, int(d))) for d in arr.shape)
            e = pb_varint(1, DT_FLOAT) + pb_bytes(2, shape)
            if off:
                e += pb_varint(4, off)
            e += pb_varint(5, len(raw))
            if tensor_crc:
                e += pb_fixed32(6, crc_mask(crc32c(raw)))
            entries.append((name.encode(), e))
    # BundleHeaderProto{num_shards=1, endianness=LITTLE(0, omitted), version{producer=1}}
    header = pb_varint(1, 1) + pb_bytes(3, pb_varint(1, 1))
    with open(prefix + ".index", "wb") as f:
        index = _BlockBuilder(restart_interval=1)
        blk = _BlockBuilder()
        last_key = b""

        def flush():
            nonlocal blk
            if blk.count:
                off, size = _write_block(f, blk.finish())
                index.add(last_key, _handle(off, size))
                blk = _BlockBuilder()

        for key, val in [(b"", header)] + entries:
            blk.add(key, val)
            last_key = key
            if len(blk.buf) >= block_size:
                flush()
        flush()
        meta_off, meta_size = _write_block(f, _BlockBuilder().finish())
        idx_off, idx_size = _write_block(f, index.finish())
        footer = _handle(meta_off, meta_size) + _handle(idx_off, idx_size)
        footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
        f.write(footer)


# ----------------------------------------------------------------------------- GraphDef pieces
def attr_s(s):
    return pb_bytes(2, s)


def attr_i(i):
    return pb_varint(3, i)


def attr_f(x):
    return pb_f32(4, x)


def attr_b(b):
    return pb_varint(5, 1 if b else 0)


def attr_type(t):
    return pb_varint(6, t)


def attr_list_i(vals):
    return pb_bytes(1, pb_bytes(3, b"".join(_varint(v) for v in vals)))


def attr_shape(dims):
    return pb_bytes(7, b"".join(pb_bytes(2, pb_varint(1, d)) for d in dims))


def attr_tensor_int32(vals):
    shape = pb_bytes(2, pb_varint(1, len(vals)))
    t = pb_varint(1, DT_INT32) + pb_bytes(2, shape) + pb_bytes(4, np.asarray(vals, "<i4").tobytes())
    return pb_bytes(8, t)


def attr_tensor_string(s):
    t = pb_varint(1, DT_STRING) + pb_bytes(2, b"") + pb_bytes(8, s)
    return pb_bytes(8, t)


def node(name, op, inputs=(), attrs=None):
    b = pb_bytes(1, name) + pb_bytes(2, op)
    for i in inputs:
        b += pb_bytes(3, i)
    for k in sorted(attrs or {}):
        b += pb_bytes(5, pb_bytes(1, k) + pb_bytes(2, attrs[k]))
    return pb_bytes(1, b)  # GraphDef.node = 1


class GraphBuilder:
    def __init__(self):
        self.nodes = []
        self.variables = {}  # name -> shape

    def add(self, name, op, inputs=(), attrs=None):
        self.nodes.append(node(name, op, inputs, attrs))
        return name

    def variable(self, name, shape):
        self.variables[name] = tuple(shape)
        self.add(name, "VariableV2", (), {"dtype": attr_type(DT_FLOAT), "shape": attr_shape(shape),
                                          "container": attr_s(""), "shared_name": attr_s("")})
        # initializer / assign clutter a real export carries
        self.add(name + "/Initializer/zeros", "Const", (), {"dtype": attr_type(DT_FLOAT)})
        self.add(name + "/Assign", "Assign", (name, name + "/Initializer/zeros"),
                 {"T": attr_type(DT_FLOAT), "use_locking": attr_b(True), "validate_shape": attr_b(True)})
        return self.add(name + "/read", "Identity", (name,), {"T": attr_type(DT_FLOAT)})

    def conv_bn_relu(self, scope, var_scope, x, cin, cout, k, eps, relu=True, conv_bias=False):
        w = self.variable(f"{var_scope}/conv/kernel", (k, k, cin, cout))
        y = self.add(f"{scope}/conv/Conv2D", "Conv2D", (x, w), {
            "T": attr_type(DT_FLOAT), "strides": attr_list_i([1, 1, 1, 1]), "padding": attr_s("SAME"),
            "data_format": attr_s("NHWC"), "dilations": attr_list_i([1, 1, 1, 1]), "use_cudnn_on_gpu": attr_b(True)})
        if conv_bias:
            bvar = self.variable(f"{var_scope}/conv/bias", (cout,))
            y = self.add(f"{scope}/conv/BiasAdd", "BiasAdd", (y, bvar),
                         {"T": attr_type(DT_FLOAT), "data_format": attr_s("NHWC")})
        g = self.variable(f"{var_scope}/bn/gamma", (cout,))
        b = self.variable(f"{var_scope}/bn/beta", (cout,))
        m = self.variable(f"{var_scope}/bn/moving_mean", (cout,))
        v = self.variable(f"{var_scope}/bn/moving_variance", (cout,))
        y = self.add(f"{scope}/bn/FusedBatchNorm", "FusedBatchNorm", (y, g, b, m, v), {
            "T": attr_type(DT_FLOAT), "epsilon": attr_f(eps), "data_format": attr_s("NHWC"),
            "is_training": attr_b(False)})
        if relu:
            y = self.add(f"{scope}/Relu", "Relu", (y,), {"T": attr_type(DT_FLOAT)})
        return y

    def dense(self, scope, var_scope, x, cin, cout):
        w = self.variable(f"{var_scope}/kernel", (cin, cout))
        b = self.variable(f"{var_scope}/bias", (cout,))
        y = self.add(f"{scope}/MatMul", "MatMul", (x, w), {"T": attr_type(DT_FLOAT), "transpose_a": attr_b(False),
                                                           "transpose_b": attr_b(False)})
        return self.add(f"{scope}/BiasAdd", "BiasAdd", (y, b), {"T": attr_type(DT_FLOAT), "data_format": attr_s("NHWC")})


def build_graph(blocks, filters, policy_channels=2, value_channels=1, value_hidden=256, bn_eps=1e-5,
                input_dtype=DT_BOOL, input_name="feature_input",
                policy_name="resnet/tower_0/policy_head/policy_predict",
                value_name="resnet/tower_0/value_head/reward_predict", conv_bias=False):
    """AGZ-shaped inference graph.  Returns (GraphBuilder, serialized GraphDef bytes)."""
    g = GraphBuilder()
    t = "resnet/tower_0"
    x = g.add(input_name, "Placeholder", (), {"dtype": attr_type(input_dtype), "shape": attr_shape([-1, 19, 19, 17])})
    if input_dtype != DT_FLOAT:
        x = g.add(f"{t}/Cast", "Cast", (x,), {"SrcT": attr_type(input_dtype), "DstT": attr_type(DT_FLOAT)})
    x = g.conv_bn_relu(f"{t}/init_conv", "resnet/init_conv", x, 17, filters, 3, bn_eps, conv_bias=conv_bias)
    for k in range(blocks):
        u, uv = f"{t}/unit_res_{k}", f"resnet/unit_res_{k}"
        y = g.conv_bn_relu(f"{u}/sub1/conv1", f"{uv}/sub1/conv1", x, filters, filters, 3, bn_eps, conv_bias=conv_bias)
        y = g.conv_bn_relu(f"{u}/sub2/conv2", f"{uv}/sub2/conv2", y, filters, filters, 3, bn_eps, relu=False,
                           conv_bias=conv_bias)
        y = g.add(f"{u}/add", "Add", (y, x), {"T": attr_type(DT_FLOAT)})
        x = g.add(f"{u}/Relu", "Relu", (y,), {"T": attr_type(DT_FLOAT)})
    # policy head
    p = g.conv_bn_relu(f"{t}/policy_head/conv1x1", "resnet/policy_head/conv1x1", x, filters, policy_channels, 1, bn_eps)
    shp = g.add(f"{t}/policy_head/flatten/shape", "Const", (), {"dtype": attr_type(DT_INT32),
                                                                "value": attr_tensor_int32([-1, 361 * policy_channels])})
    p = g.add(f"{t}/policy_head/flatten/Reshape", "Reshape", (p, shp), {"T": attr_type(DT_FLOAT), "Tshape": attr_type(DT_INT32)})
    p = g.dense(f"{t}/policy_head/dense", "resnet/policy_head/dense", p, 361 * policy_channels, 362)
    g.add(policy_name, "Softmax", (p,), {"T": attr_type(DT_FLOAT)})
    # value head
    v = g.conv_bn_relu(f"{t}/value_head/conv1x1", "resnet/value_head/conv1x1", x, filters, value_channels, 1, bn_eps)
    shv = g.add(f"{t}/value_head/flatten/shape", "Const", (), {"dtype": attr_type(DT_INT32),
                                                               "value": attr_tensor_int32([-1, 361 * value_channels])})
    v = g.add(f"{t}/value_head/flatten/Reshape", "Reshape", (v, shv), {"T": attr_type(DT_FLOAT), "Tshape": attr_type(DT_INT32)})
    v = g.dense(f"{t}/value_head/dense1", "resnet/value_head/dense1", v, 361 * value_channels, value_hidden)
    v = g.add(f"{t}/value_head/dense1/Relu", "Relu", (v,), {"T": attr_type(DT_FLOAT)})
    v = g.dense(f"{t}/value_head/dense2", "resnet/value_head/dense2", v, value_hidden, 1)
    g.add(value_name, "Tanh", (v,), {"T": attr_type(DT_FLOAT)})
    # saver + training clutter that the loader must never need
    g.add("save/Const", "Const", (), {"dtype": attr_type(DT_STRING), "value": attr_tensor_string(b"model")})
    g.add("save/RestoreV2", "RestoreV2", ("save/Const",), {})
    g.add("save/restore_all", "NoOp", ("^save/RestoreV2",), {})
    g.add("gradients/Fill", "Fill", (), {"T": attr_type(DT_FLOAT)})
    g.add("input_producer/Const", "Const", (), {"dtype": attr_type(DT_STRING)})
    return g, b"".join(g.nodes)


def write_metagraph(path, graph_def_bytes, filename_tensor="save/Const:0", restore_op="save/restore_all"):
    meta_info = pb_bytes(1, pb_bytes(1, "") + pb_bytes(5, "1.6.0"))  # MetaInfoDef{meta_graph_version, tf version}
    saver = pb_bytes(1, filename_tensor) + pb_bytes(2, "save/control_dependency:0") + pb_bytes(3, restore_op)
    with open(path, "wb") as f:
        f.write(meta_info + pb_bytes(2, graph_def_bytes) + pb_bytes(3, saver))


def write_graphdef(path, graph_def_bytes):
    with open(path, "wb") as f:
        f.write(graph_def_bytes)


# ----------------------------------------------------------------------------- synthetic weights
def random_weights(variables, seed):
    """He-normal conv/dense kernels; gamma~U(.5,1.5), beta,mean~N(0,.1), var~U(.5,1.5)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name in sorted(variables):
        shape = variables[name]
        if name.endswith("/kernel"):
            fan_in = int(np.prod(shape[:-1]))
            out[name] = (rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        elif name.endswith("sub2/conv2/bn/gamma"):
            # keep the residual stream from doubling its variance every block (20-40 blocks of random init)
            out[name] = rng.uniform(0.15, 0.35, shape).astype(np.float32)
        elif name.endswith("/gamma") or name.endswith("/moving_variance"):
            out[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        else:  # beta, moving_mean, bias
            out[name] = (rng.standard_normal(shape) * 0.1).astype(np.float32)
    return out


def write_synthetic_checkpoint(directory, blocks, filters, seed=20260119, name="bootstrap-ckpt", tensor_crc=True,
                               **graph_kw):
    """Writes <directory>/<name>.meta/.index/.data-00000-of-00001; returns (meta_path, ckpt_prefix)."""
    import os
    os.makedirs(directory, exist_ok=True)
    g, gd = build_graph(blocks, filters, **graph_kw)
    meta = os.path.join(directory, name + ".meta")
    prefix = os.path.join(directory, name)
    write_metagraph(meta, gd)
    write_bundle(prefix, random_weights(g.variables, seed), tensor_crc=tensor_crc)
    return meta, prefix
