"""TEST INFRASTRUCTURE — TensorFlow's graph / checkpoint protobuf schemas, declared to Google's protobuf runtime.

The product's loader (csrc/tfmodel.cpp) and the oracle's reader (oracle/tf_reader.py) are hand-written wire-format
walkers, and unrealgo_b200/tfformat.py is a hand-written encoder: all three come from one reading of the format.  This
module gives the tests a FOURTH, foreign party: the message types below are declared from TensorFlow's published
.proto files (tensorflow/core/framework/{graph,node_def,attr_value,tensor,tensor_shape,types,versions}.proto,
tensorflow/core/protobuf/{meta_graph,saver,tensor_bundle}.proto — field names and numbers as published; only the
fields a graph/checkpoint export carries) and serialised by the real protobuf library, with its own choices of field
order, packed encodings and map-entry layout.  tests/foreign_graphs.py builds TensorFlow-1.x-idiom graphs with it.
"""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto
_T = {"int32": F.TYPE_INT32, "int64": F.TYPE_INT64, "float": F.TYPE_FLOAT, "double": F.TYPE_DOUBLE, "bool": F.TYPE_BOOL,
      "string": F.TYPE_STRING, "bytes": F.TYPE_BYTES, "fixed32": F.TYPE_FIXED32}


def _field(msg, name, number, typ, repeated=False, oneof=None):
    f = msg.field.add()
    f.name, f.number = name, number
    f.label = F.LABEL_REPEATED if repeated else F.LABEL_OPTIONAL
    if typ in _T:
        f.type = _T[typ]
    elif typ.startswith("enum:"):
        f.type, f.type_name = F.TYPE_ENUM, typ[5:]
    else:
        f.type, f.type_name = F.TYPE_MESSAGE, typ
    if oneof is not None:
        f.oneof_index = oneof
    return f


def _map_entry(parent, entry_name, key_type, value_type):
    e = parent.nested_type.add()
    e.name = entry_name
    e.options.map_entry = True
    _field(e, "key", 1, key_type)
    _field(e, "value", 2, value_type)


def _build():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "ugeval_test_tensorflow_subset.proto"
    fd.package = "tensorflow"
    fd.syntax = "proto3"

    dt = fd.enum_type.add()
    dt.name = "DataType"  # types.proto
    for name, num in (("DT_INVALID", 0), ("DT_FLOAT", 1), ("DT_DOUBLE", 2), ("DT_INT32", 3), ("DT_UINT8", 4),
                      ("DT_STRING", 7), ("DT_INT64", 9), ("DT_BOOL", 10), ("DT_RESOURCE", 20)):
        v = dt.value.add()
        v.name, v.number = name, num

    m = fd.message_type.add()
    m.name = "TensorShapeProto"  # tensor_shape.proto
    d = m.nested_type.add()
    d.name = "Dim"
    _field(d, "size", 1, "int64")
    _field(d, "name", 2, "string")
    _field(m, "dim", 2, ".tensorflow.TensorShapeProto.Dim", repeated=True)
    _field(m, "unknown_rank", 3, "bool")

    m = fd.message_type.add()
    m.name = "TensorProto"  # tensor.proto
    _field(m, "dtype", 1, "enum:.tensorflow.DataType")
    _field(m, "tensor_shape", 2, ".tensorflow.TensorShapeProto")
    _field(m, "version_number", 3, "int32")
    _field(m, "tensor_content", 4, "bytes")
    _field(m, "float_val", 5, "float", repeated=True)
    _field(m, "double_val", 6, "double", repeated=True)
    _field(m, "int_val", 7, "int32", repeated=True)
    _field(m, "string_val", 8, "bytes", repeated=True)
    _field(m, "int64_val", 10, "int64", repeated=True)
    _field(m, "bool_val", 11, "bool", repeated=True)

    m = fd.message_type.add()
    m.name = "AttrValue"  # attr_value.proto
    lv = m.nested_type.add()
    lv.name = "ListValue"
    _field(lv, "s", 2, "bytes", repeated=True)
    _field(lv, "i", 3, "int64", repeated=True)
    _field(lv, "f", 4, "float", repeated=True)
    _field(lv, "b", 5, "bool", repeated=True)
    _field(lv, "type", 6, "enum:.tensorflow.DataType", repeated=True)
    _field(lv, "shape", 7, ".tensorflow.TensorShapeProto", repeated=True)
    _field(lv, "tensor", 8, ".tensorflow.TensorProto", repeated=True)
    m.oneof_decl.add().name = "value"
    _field(m, "list", 1, ".tensorflow.AttrValue.ListValue", oneof=0)
    _field(m, "s", 2, "bytes", oneof=0)
    _field(m, "i", 3, "int64", oneof=0)
    _field(m, "f", 4, "float", oneof=0)
    _field(m, "b", 5, "bool", oneof=0)
    _field(m, "type", 6, "enum:.tensorflow.DataType", oneof=0)
    _field(m, "shape", 7, ".tensorflow.TensorShapeProto", oneof=0)
    _field(m, "tensor", 8, ".tensorflow.TensorProto", oneof=0)
    _field(m, "placeholder", 9, "string", oneof=0)

    m = fd.message_type.add()
    m.name = "NodeDef"  # node_def.proto
    _field(m, "name", 1, "string")
    _field(m, "op", 2, "string")
    _field(m, "input", 3, "string", repeated=True)
    _field(m, "device", 4, "string")
    _map_entry(m, "AttrEntry", "string", ".tensorflow.AttrValue")
    _field(m, "attr", 5, ".tensorflow.NodeDef.AttrEntry", repeated=True)

    m = fd.message_type.add()
    m.name = "VersionDef"  # versions.proto
    _field(m, "producer", 1, "int32")
    _field(m, "min_consumer", 2, "int32")
    _field(m, "bad_consumers", 3, "int32", repeated=True)

    m = fd.message_type.add()
    m.name = "GraphDef"  # graph.proto
    _field(m, "node", 1, ".tensorflow.NodeDef", repeated=True)
    _field(m, "version", 3, "int32")
    _field(m, "versions", 4, ".tensorflow.VersionDef")

    m = fd.message_type.add()
    m.name = "SaverDef"  # saver.proto
    _field(m, "filename_tensor_name", 1, "string")
    _field(m, "save_tensor_name", 2, "string")
    _field(m, "restore_op_name", 3, "string")
    _field(m, "max_to_keep", 4, "int32")
    _field(m, "sharded", 5, "bool")
    _field(m, "keep_checkpoint_every_n_hours", 6, "float")
    _field(m, "version", 7, "int32")  # CheckpointFormatVersion enum: V2 = 2

    m = fd.message_type.add()
    m.name = "CollectionDef"  # meta_graph.proto (bytes_list only: what trainable_variables / variables carry)
    b = m.nested_type.add()
    b.name = "BytesList"
    _field(b, "value", 1, "bytes", repeated=True)
    n = m.nested_type.add()
    n.name = "NodeList"
    _field(n, "value", 1, "string", repeated=True)
    m.oneof_decl.add().name = "kind"
    _field(m, "node_list", 1, ".tensorflow.CollectionDef.NodeList", oneof=0)
    _field(m, "bytes_list", 2, ".tensorflow.CollectionDef.BytesList", oneof=0)

    m = fd.message_type.add()
    m.name = "MetaGraphDef"  # meta_graph.proto
    mi = m.nested_type.add()
    mi.name = "MetaInfoDef"
    _field(mi, "meta_graph_version", 1, "string")
    _field(mi, "tags", 4, "string", repeated=True)
    _field(mi, "tensorflow_version", 5, "string")
    _field(mi, "tensorflow_git_version", 6, "string")
    _field(m, "meta_info_def", 1, ".tensorflow.MetaGraphDef.MetaInfoDef")
    _field(m, "graph_def", 2, ".tensorflow.GraphDef")
    _field(m, "saver_def", 3, ".tensorflow.SaverDef")
    _map_entry(m, "CollectionDefEntry", "string", ".tensorflow.CollectionDef")
    _field(m, "collection_def", 4, ".tensorflow.MetaGraphDef.CollectionDefEntry", repeated=True)

    m = fd.message_type.add()
    m.name = "BundleHeaderProto"  # tensor_bundle.proto
    _field(m, "num_shards", 1, "int32")
    _field(m, "endianness", 2, "int32")
    _field(m, "version", 3, ".tensorflow.VersionDef")

    m = fd.message_type.add()
    m.name = "BundleEntryProto"
    _field(m, "dtype", 1, "enum:.tensorflow.DataType")
    _field(m, "shape", 2, ".tensorflow.TensorShapeProto")
    _field(m, "shard_id", 3, "int32")
    _field(m, "offset", 4, "int64")
    _field(m, "size", 5, "int64")
    _field(m, "crc32c", 6, "fixed32")

    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda name: message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow." + name))
    return {n: get(n) for n in ("TensorShapeProto", "TensorProto", "AttrValue", "NodeDef", "GraphDef", "VersionDef",
                                "SaverDef", "MetaGraphDef", "CollectionDef", "BundleHeaderProto", "BundleEntryProto")}


_M = _build()
TensorShapeProto, TensorProto, AttrValue, NodeDef, GraphDef = (_M[k] for k in
                                                               ("TensorShapeProto", "TensorProto", "AttrValue", "NodeDef", "GraphDef"))
VersionDef, SaverDef, MetaGraphDef, CollectionDef = (_M[k] for k in ("VersionDef", "SaverDef", "MetaGraphDef", "CollectionDef"))
BundleHeaderProto, BundleEntryProto = _M["BundleHeaderProto"], _M["BundleEntryProto"]

DT_FLOAT, DT_DOUBLE, DT_INT32, DT_STRING, DT_INT64, DT_BOOL = 1, 2, 3, 7, 9, 10
