"""TEST INFRASTRUCTURE — independent Python readers for MetaGraphDef / GraphDef and TF V2 bundles.

Written separately from the product's C++ loader (unrealgo_b200/csrc/tfmodel.cpp) so that the two can be
checked against each other.  Formats: SURVEY.md A.5 (general knowledge of TensorFlow 1.x; the reference only
hands paths to TF: funcapproximator/DlTFNetworkEvaluator.cc:130, :176-193).  Only tests/, smoke() and
bench.py's cpu_baseline leg may import this.
"""
import struct

import numpy as np


def read_varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def fields(buf):
    """yield (field_number, wire_type, value) over one protobuf message; bytes fields as memoryview slices"""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = read_varint(buf, pos)
        num, wire = key >> 3, key & 7
        if wire == 0:
            v, pos = read_varint(buf, pos)
        elif wire == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wire == 2:
            ln, pos = read_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wire == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wire}")
        yield num, wire, v


def _signed(v, bits=64):
    return v - (1 << bits) if v >> (bits - 1) else v


def parse_tensor(buf):
    dtype, dims, content, ints, floats, strs, bools = 0, [], None, [], [], [], []
    for num, wire, v in fields(buf):
        if num == 1:
            dtype = v
        elif num == 2:
            for n2, _, v2 in fields(v):
                if n2 == 2:
                    size = 0
                    for n3, _, v3 in fields(v2):
                        if n3 == 1:
                            size = _signed(v3)
                    dims.append(size)
        elif num == 4:
            content = v
        elif num == 5:
            floats += list(struct.unpack(f"<{len(v) // 4}f", v)) if wire == 2 else [struct.unpack("<f", struct.pack("<I", v))[0]]
        elif num == 7:
            if wire == 2:
                p = 0
                while p < len(v):
                    x, p = read_varint(v, p)
                    ints.append(_signed(x))
            else:
                ints.append(_signed(v))
        elif num == 8:
            strs.append(v)
        elif num == 11:  # bool_val (packed or not)
            bools += [bool(b) for b in v] if wire == 2 else [bool(v)]
    if dtype == 10:
        arr = np.frombuffer(content, np.bool_) if content is not None else np.asarray(bools, np.bool_)
        count = int(np.prod(dims)) if dims else 1
        if arr.size == 0 and count == 1:
            arr = np.zeros(1, np.bool_)   # proto3 omits a false scalar
        return arr.reshape(dims) if dims else arr.reshape(())
    if dtype == 3:
        arr = np.frombuffer(content, "<i4") if content is not None else np.asarray(ints, np.int32)
    elif dtype == 1:
        arr = np.frombuffer(content, "<f4") if content is not None else np.asarray(floats, np.float32)
    elif dtype == 7:
        return strs
    else:
        return None
    count = int(np.prod(dims)) if dims else 1
    if arr.size == 1 and count > 1:
        arr = np.repeat(arr, count)
    return arr.reshape(dims) if dims else arr.reshape(())


def parse_attr(buf):
    a = {}
    for num, wire, v in fields(buf):
        if num == 1:  # list
            li = []
            for n2, w2, v2 in fields(v):
                if n2 == 3:
                    if w2 == 2:
                        p = 0
                        while p < len(v2):
                            x, p = read_varint(v2, p)
                            li.append(_signed(x))
                    else:
                        li.append(_signed(v2))
            a["list_i"] = li
        elif num == 2:
            a["s"] = v
        elif num == 3:
            a["i"] = _signed(v)
        elif num == 4:
            a["f"] = struct.unpack("<f", struct.pack("<I", v))[0]
        elif num == 5:
            a["b"] = bool(v)
        elif num == 6:
            a["type"] = v
        elif num == 8:
            a["tensor"] = parse_tensor(v)
    return a


def parse_graph_def(buf):
    nodes = {}
    order = []
    for num, _, v in fields(buf):
        if num != 1:
            continue
        node = {"name": "", "op": "", "inputs": [], "attrs": {}}
        for n2, _, v2 in fields(v):
            if n2 == 1:
                node["name"] = v2.decode()
            elif n2 == 2:
                node["op"] = v2.decode()
            elif n2 == 3:
                node["inputs"].append(v2.decode())
            elif n2 == 5:
                key, val = None, {}
                for n3, _, v3 in fields(v2):
                    if n3 == 1:
                        key = v3.decode()
                    elif n3 == 2:
                        val = parse_attr(v3)
                node["attrs"][key] = val
        nodes[node["name"]] = node
        order.append(node["name"])
    return nodes, order


def read_graph(path):
    """returns (nodes dict, saver dict) for a .meta (MetaGraphDef) or .pb (GraphDef) file"""
    buf = open(path, "rb").read()
    if path.endswith(".meta"):
        graph, saver = None, {}
        for num, _, v in fields(buf):
            if num == 2:
                graph = v
            elif num == 3:
                for n2, _, v2 in fields(v):
                    if n2 == 1:
                        saver["filename_tensor_name"] = v2.decode()
                    elif n2 == 3:
                        saver["restore_op_name"] = v2.decode()
        if graph is None:
            raise ValueError("no graph_def in MetaGraphDef")
        return parse_graph_def(graph)[0], saver
    return parse_graph_def(buf)[0], {"filename_tensor_name": "save/Const:0", "restore_op_name": "save/restore_all"}


# ----------------------------------------------------------------------------- bundle
def _crc32c(data):
    from . import lib
    return lib().og_crc32c(data, len(data))


def _mask(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _block(buf, off, size, verify=True):
    contents = buf[off:off + size]
    if buf[off + size] != 0:
        raise ValueError("compressed block")
    if verify:
        stored = struct.unpack_from("<I", buf, off + size + 1)[0]
        if stored != _mask(_crc32c(buf[off:off + size + 1])):
            raise ValueError("block crc mismatch")
    nrestarts = struct.unpack_from("<I", contents, size - 4)[0]
    end = size - 4 - 4 * nrestarts
    pos = 0
    key = b""
    while pos < end:
        shared, pos = read_varint(contents, pos)
        non_shared, pos = read_varint(contents, pos)
        vlen, pos = read_varint(contents, pos)
        key = key[:shared] + contents[pos:pos + non_shared]
        pos += non_shared
        yield key, contents[pos:pos + vlen]
        pos += vlen


def read_bundle(prefix, verify=True):
    """returns dict name -> np.float32 array"""
    idx = open(prefix + ".index", "rb").read()
    if struct.unpack_from("<Q", idx, len(idx) - 8)[0] != 0xDB4775248B80FB57:
        raise ValueError("bad table magic")
    footer = idx[-48:]
    pos = 0
    _, pos = read_varint(footer, pos)
    _, pos = read_varint(footer, pos)
    ioff, pos = read_varint(footer, pos)
    isize, pos = read_varint(footer, pos)
    rows = []
    num_shards = 1
    for _, handle in _block(idx, ioff, isize, verify):
        boff, p = read_varint(handle, 0)
        bsize, p = read_varint(handle, p)
        for key, val in _block(idx, boff, bsize, verify):
            if not key:   # BundleHeaderProto: num_shards = 1
                for num, _, v in fields(val):
                    if num == 1:
                        num_shards = v
                continue
            rows.append((key, val))
    # <prefix>.data-0000k-of-0000N (N > 1: a merged Saver(sharded=True) bundle)
    shards = [np.memmap(prefix + ".data-%05d-of-%05d" % (k, num_shards), dtype=np.uint8, mode="r") for k in range(num_shards)]
    out = {}
    for key, val in rows:
        dtype, shape, offset, size, crc, shard = 0, [], 0, 0, None, 0
        for num, _, v in fields(val):
            if num == 1:
                dtype = v
            elif num == 3:
                shard = v
            elif num == 2:
                for n2, _, v2 in fields(v):
                    if n2 == 2:
                        d = 0
                        for n3, _, v3 in fields(v2):
                            if n3 == 1:
                                d = v3
                        shape.append(d)
            elif num == 4:
                offset = v
            elif num == 5:
                size = v
            elif num == 6:
                crc = v
        if dtype != 1:
            continue
        raw = bytes(shards[shard][offset:offset + size])
        if verify and crc is not None and _mask(_crc32c(raw)) != crc:
            raise ValueError(f"crc mismatch for {key!r}")
        out[key.decode()] = np.frombuffer(raw, "<f4").reshape(shape).copy()
    return out
