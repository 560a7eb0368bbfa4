"""Python mirror of the reference's evaluator interface, over the C ABI (include/ugeval.h).

Names, argument meaning and error behaviour follow
  tensorflow::DlTFNetworkEvaluator<T>  funcapproximator/DlTFNetworkEvaluator.h:33-83, .cc:27-352
  UctBoardEvaluator                    search/UctBoardEvaluator.cpp:6-33
  DlConfig                             funcapproximator/DlConfig.cc:13-41, :164-246
so that the parity tests read like the reference's own call sites.  All arithmetic happens in libugeval.so on
the GPU; these classes only marshal numpy buffers.
"""
import ctypes as C
import json

import numpy as np

from . import _lib
from ._lib import (MASK_WORDS, NUM_MOVES, NUM_PLANES, NUM_POINTS, PACKED_POS_BYTES, PREC_BF16, PREC_FP16, PREC_FP32,  # noqa: F401
                   TR_FLIP_PROB, TR_FLIP_SIGN, TR_IDENTITY)

DF_HWC, DF_CHW = 0, 1
MAX_BATCHES = 16  # funcapproximator/DlTFNetworkEvaluator.h:15 (the reference's cap; not a limit here)


class DlConfig:
    """dlcfg-<size> reader (funcapproximator/DlConfig.cc)."""

    def __init__(self, filename):
        self._lib = _lib.load()
        self._c = self._lib.ug_dlcfg_open(filename.encode())

    def __del__(self):
        if getattr(self, "_c", None):
            self._lib.ug_dlcfg_close(self._c)
            self._c = None

    def get(self, key, default=""):
        buf = C.create_string_buffer(4096)
        self._lib.ug_dlcfg_get(self._c, key.encode(), default.encode(), buf, len(buf))
        return buf.value.decode()

    def get_metagraph(self):
        return self.get("meta_graph", "bootstrap-ckpt.meta")

    def default_checkpoint(self):
        return self.get("checkpoint", "bootstrap-ckpt")

    def get_network_input(self):
        return self.get("nn_input", "input")

    def get_network_outputs(self):
        cap, mx = 512, 8
        buf = C.create_string_buffer(cap * mx)
        n = self._lib.ug_dlcfg_network_outputs(self._c, buf, cap, mx)
        return [buf.raw[i * cap:(i + 1) * cap].split(b"\0", 1)[0].decode() for i in range(max(n, 0))]

    def reuse_search_tree(self):
        return bool(self._lib.ug_dlcfg_reuse_search_tree(self._c))

    def get_value_transform(self):
        return self._lib.ug_dlcfg_value_transform(self._c)


def inspect_model(graph_path, ckpt_prefix=None, input_name=None, outputs=None):
    """host-only description of a graph (+checkpoint): dict, or raises RuntimeError with the loader's message"""
    lib = _lib.load()
    buf = C.create_string_buffer(1 << 22)
    rc = lib.ug_inspect_model(graph_path.encode(), ckpt_prefix.encode() if ckpt_prefix else None,
                              input_name.encode() if input_name else None,
                              outputs[0].encode() if outputs else None, outputs[1].encode() if outputs else None,
                              buf, len(buf))
    if rc != 0:
        raise RuntimeError(buf.value.decode())
    return json.loads(buf.value.decode())


def bundle_read_float(ckpt_prefix, key):
    """host-only: one DT_FLOAT variable of a TF V2 checkpoint bundle through the product's reader -> float32 array
    (shaped as stored); raises RuntimeError with the reader's message"""
    lib = _lib.load()
    dims = (C.c_longlong * 8)()
    ndim = C.c_int(0)
    err = C.create_string_buffer(1024)
    n = lib.ug_bundle_read_float(ckpt_prefix.encode(), key.encode(), None, 0, dims, C.byref(ndim), err, len(err))
    if n == -1:
        raise RuntimeError(err.value.decode())
    shape = tuple(int(dims[i]) for i in range(ndim.value))
    out = np.empty(int(np.prod(shape, dtype=np.int64)) if shape else 1, np.float32)
    n = lib.ug_bundle_read_float(ckpt_prefix.encode(), key.encode(), out.ctypes.data, out.size, dims, C.byref(ndim), err, len(err))
    if n < 0:
        raise RuntimeError(err.value.decode() or f"ug_bundle_read_float returned {n}")
    return out[:n].reshape(shape)


class DlTFNetworkEvaluator:
    """DlTFNetworkEvaluator<bool>: LoadGraph / UpdateCheckPoint / Evaluate, plus the native packed entry points."""

    def __init__(self, graphPath="", df=DF_HWC, feature_input=None, outputs=None, device=0, precision=PREC_FP16,
                 max_batch=256):
        self._lib = _lib.load()
        self._h = self._lib.ug_eval_create(device, precision, max_batch)
        if not self._h:
            raise RuntimeError("ug_eval_create failed: " + (self._lib.ug_last_error(None) or b"").decode())
        self.max_batch = max_batch
        self.m_dataFormat = df
        if feature_input is not None or outputs is not None:
            self._set_io(feature_input, outputs)
        self.LoadGraph(graphPath)  # empty path is legal: returns False, object usable later (.cc:77-83)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ug_eval_destroy(self._h)
            self._h = None

    __del__ = close

    def _err(self):
        return (self._lib.ug_last_error(self._h) or b"").decode()

    def _set_io(self, feature_input, outputs):
        arr = None
        n = 0
        if outputs is not None:
            n = len(outputs)
            arr = (C.c_char_p * n)(*[o.encode() for o in outputs])
        self._lib.ug_eval_set_io(self._h, feature_input.encode() if feature_input is not None else None, arr, n)

    def SetNetworkInput(self, name):
        self._set_io(name, None)

    def SetNetworkOutput(self, outputs):
        self._set_io(None, list(outputs))

    def LoadGraph(self, graphPath):
        rc = self._lib.ug_eval_load_graph(self._h, graphPath.encode())
        if rc < 0:
            raise RuntimeError(self._err())  # the reference throws std::runtime_error (.cc:91,:132,:158)
        return rc == 0

    def UpdateCheckPoint(self, checkpointPath=""):
        return self._lib.ug_eval_load_checkpoint(self._h, checkpointPath.encode()) == 0

    def MetaGraphLoaded(self):
        return bool(self._lib.ug_eval_ready(self._h))

    def set_value_transform(self, vt):
        self._lib.ug_eval_set_value_transform(self._h, int(vt))

    def set_conv_ctas(self, ctas):
        if self._lib.ug_eval_set_conv_ctas(self._h, ctas) != 0:
            raise ValueError("ctas must be 1 or 2")

    def register_host(self, *arrays):
        """page-lock caller arrays that Evaluate() will be called with (see ug_eval_register_host); returns the
        number accepted.  The arrays must outlive the evaluator or be passed to unregister_host() first."""
        ok = 0
        for a in arrays:
            if self._lib.ug_eval_register_host(self._h, a.ctypes.data, a.nbytes) == 0:
                ok += 1
        return ok

    def unregister_host(self, *arrays):
        for a in arrays:
            self._lib.ug_eval_unregister_host(self._h, a.ctypes.data)

    def set_conv_impl(self, impl):
        if self._lib.ug_eval_set_conv_impl(self._h, impl) != 0:
            raise ValueError("impl must be 1 or 2")

    def set_fp16_range_policy(self, policy):
        """a checkpoint outside fp16's range: 0/'auto' = rescale by powers of two, else bf16 (default); 1/'refuse';
        2/'ignore' = load anyway; 3/'bf16' = bf16 operands without rescaling"""
        if self._lib.ug_eval_set_fp16_range_policy(self._h, {"auto": 0, "refuse": 1, "ignore": 2, "bf16": 3}.get(policy, policy)) != 0:
            raise ValueError("policy: 0/'auto', 1/'refuse', 2/'ignore', 3/'bf16'")

    def set_small_batch(self, max_batch, nt=0):
        """batches <= max_batch run the small-batch tower kernel (0 disables); nt = output channels per tile (0 = auto),
        + 256 to force single-CTA tiles, + 512 to force CTA-pair tiles (nt >= 32)"""
        if self._lib.ug_eval_set_small_batch(self._h, max_batch, nt) != 0:
            raise ValueError("max_batch >= 0, nt in (0, 16, 32, 64, 128) [+ 256 | + 512]")

    def set_residual_split(self, on):
        self._lib.ug_eval_set_residual_split(self._h, 1 if on else 0)

    def info(self):
        ni = _lib.NetInfo()
        if self._lib.ug_eval_info(self._h, C.byref(ni)) != 0:
            raise RuntimeError("no network loaded")
        return {f: getattr(ni, f) for f, _ in _lib.NetInfo._fields_}

    @staticmethod
    def _is_hwc(feature):
        return feature.ndim == 4 and feature.shape[-1] == NUM_PLANES and feature.shape[1] == 19

    def CreateTensor(self, batches):
        """.cc:271-284: a tensor shaped {batches, 19, 19, 17} whatever m_dataFormat says (the device workspace itself
        is sized once per architecture); kept per batch size like m_input_tensors[batches - 1]"""
        if not hasattr(self, "m_input_tensors"):
            self.m_input_tensors = {}
        if batches not in self.m_input_tensors:
            self.m_input_tensors[batches] = np.zeros((batches, 19, 19, NUM_PLANES), np.bool_)
        return self.m_input_tensors[batches]

    def TransformFeature(self, feature, numBatches=None):
        """.cc:366-412, both overloads: fills and returns the bool tensor Session::Run would be fed.  DF_HWC: the NHWC
        re-layout of the planes (CHW planes go through the GPU kernel Evaluate() runs, ug_transform_planes); DF_CHW: the
        planes' CHW order written under the same {n, 19, 19, 17} dims, as the reference does."""
        feature = np.ascontiguousarray(feature, np.int8)
        n = feature.shape[0] if numBatches is None else numBatches
        t = self.CreateTensor(n)
        flat = t.reshape(-1)
        if self._is_hwc(feature):
            src = feature[:n] if self.m_dataFormat == DF_HWC else feature[:n].transpose(0, 3, 1, 2)
            flat[:] = np.ascontiguousarray(src).reshape(-1) != 0
        elif self.m_dataFormat == DF_HWC:
            out = np.empty((n, 19, 19, NUM_PLANES), np.int8)
            if self._lib.ug_transform_planes(self._h, feature.ctypes.data, n, out.ctypes.data) != 0:
                raise RuntimeError(self._err())
            flat[:] = out.reshape(-1) != 0
        else:
            flat[:] = feature[:n].reshape(-1) != 0
        return t

    @staticmethod
    def printFeature(data):
        """.cc:353-363"""
        d = np.asarray(data).reshape(-1)[:19 * 19 * NUM_PLANES].reshape(19 * 19, NUM_PLANES)
        print("\n".join(" ".join(str(int(v)) for v in row) + " " for row in d) + "\n")

    def Evaluate(self, feature, policies_, value_, numBatches=None):
        """feature: int8 [n][17][19][19] (or [n][19][19][17]); policies_: float64 [n][362]; value_: float64 [n].
        As in the reference a failed run logs and leaves the outputs untouched (.cc:315-317); returns bool.
        With m_dataFormat == DF_CHW the graph is handed the CHW-ordered buffer under NHWC dims, as the reference's
        TransformFeature + CreateTensor do (no reference caller passes DF_CHW, search/UctBoardEvaluator.cpp:6)."""
        feature = np.ascontiguousarray(feature, np.int8)
        n = feature.shape[0] if numBatches is None else numBatches
        hwc = self._is_hwc(feature)
        assert policies_.dtype == np.float64 and value_.dtype == np.float64
        assert policies_.flags.c_contiguous and value_.flags.c_contiguous
        if self.m_dataFormat != DF_HWC:
            if hwc:
                feature = np.ascontiguousarray(feature[:n].transpose(0, 3, 1, 2))
            hwc = True   # the buffer is read as [n][19][19][17] whatever order it was written in
        fn = self._lib.ug_eval_run_planes_hwc if hwc else self._lib.ug_eval_run_planes
        rc = fn(self._h, feature.ctypes.data, n, policies_.ctypes.data, value_.ctypes.data)
        if rc != 0:
            import sys
            print("Running model failed:", self._err(), file=sys.stderr)
        return rc == 0

    def evaluate_packed(self, packed, legal=None):
        """packed: uint32 [n][196] (ug_packed_pos); legal: optional uint32 [n][12].  -> (float32 [n][362], [n])"""
        packed = np.ascontiguousarray(packed, np.uint32).reshape(-1, PACKED_POS_BYTES // 4)
        n = packed.shape[0]
        pol = np.empty((n, NUM_MOVES), np.float32)
        val = np.empty(n, np.float32)
        lp = None
        if legal is not None:
            legal = np.ascontiguousarray(legal, np.uint32).reshape(n, MASK_WORDS)
            lp = legal.ctypes.data
        rc = self._lib.ug_eval_run_packed(self._h, packed.ctypes.data, n, lp, pol.ctypes.data, val.ctypes.data)
        if rc != 0:
            raise RuntimeError(self._err())
        return pol, val

    def run_packed_device(self, d_pos, n, d_policy, d_value, d_legal=None, stream=None):
        rc = self._lib.ug_eval_run_packed_device(self._h, d_pos, n, d_legal, d_policy, d_value, stream)
        if rc != 0:
            raise RuntimeError(self._err())

    def sync(self):
        if self._lib.ug_eval_sync(self._h) != 0:
            raise RuntimeError(self._err())

    # ---- encoder seam (CollectFeatures / TransformFeature)
    def encode_planes(self, packed):
        packed = np.ascontiguousarray(packed, np.uint32).reshape(-1, PACKED_POS_BYTES // 4)
        out = np.empty((packed.shape[0], NUM_PLANES, 19, 19), np.int8)
        if self._lib.ug_encode_planes(self._h, packed.ctypes.data, packed.shape[0], out.ctypes.data) != 0:
            raise RuntimeError(self._err())
        return out

    def encode_nhwc(self, packed):
        packed = np.ascontiguousarray(packed, np.uint32).reshape(-1, PACKED_POS_BYTES // 4)
        out = np.empty((packed.shape[0], 19, 19, NUM_PLANES), np.int8)
        if self._lib.ug_encode_nhwc(self._h, packed.ctypes.data, packed.shape[0], out.ctypes.data) != 0:
            raise RuntimeError(self._err())
        return out

    def set_debug_stop(self, layer):
        self._lib.ug_eval_set_debug_stop(self._h, layer)

    def debug_layer(self, layer, n):
        c = self.info()["filters"]
        out = np.empty((n, NUM_POINTS, c), np.float32)
        rc = self._lib.ug_eval_debug_layer(self._h, layer, n, out.ctypes.data)
        if rc < 0:
            raise RuntimeError(self._err())
        return out

    def bench(self, packed, iters):
        packed = np.ascontiguousarray(packed, np.uint32).reshape(-1, PACKED_POS_BYTES // 4)
        res = _lib.BenchResult()
        if self._lib.ug_eval_bench(self._h, packed.ctypes.data, packed.shape[0], iters, C.byref(res)) != 0:
            raise RuntimeError(self._err())
        return {f: getattr(res, f) for f, _ in _lib.BenchResult._fields_}


class DlTFRecordWriter:
    """tensorflow::io::DlTFRecordWriter (funcapproximator/DlTFRecordWriter.cc): self-play examples as a TFRecord file
    of tf.train.Example messages, written by libugeval without TensorFlow."""

    def __init__(self, filename):
        self._lib = _lib.load()
        self._w = self._lib.ug_tfrecord_open(filename.encode())
        if not self._w:
            raise OSError(f"cannot create {filename}")

    def WriteExample(self, feature17, policy, reward, names=None):
        """feature17: int8/uint8 array (6137 plane bytes in the reference); policy: float32 array; reward: float"""
        f = np.ascontiguousarray(feature17).view(np.uint8).reshape(-1)
        p = np.ascontiguousarray(policy, np.float32).reshape(-1)
        if names is None:
            rc = self._lib.ug_tfrecord_write_example(self._w, f.ctypes.data, f.size, p.ctypes.data, p.size, float(reward))
        else:
            fn, pn, rn = (n.encode() for n in names)
            rc = self._lib.ug_tfrecord_write_example_named(self._w, fn, f.ctypes.data, f.size, pn, p.ctypes.data, p.size,
                                                           rn, float(reward))
        if rc != 0:
            raise OSError("TFRecord write failed")
        return 0

    def Flush(self):
        self._lib.ug_tfrecord_flush(self._w)

    def close(self):
        """returns the number of records written"""
        n = -1
        if getattr(self, "_w", None):
            n = self._lib.ug_tfrecord_close(self._w)
            self._w = None
        return n

    __del__ = close


class UctBoardEvaluator:
    """search/UctBoardEvaluator.cpp:6-33 — config -> evaluator wiring + forwarding."""

    def __init__(self, graphPath=None, checkpoint=None, config=None, **kw):
        if graphPath is None:
            # default ctor: DlTFNetworkEvaluator<bool>("", DF_HWC) + names from DlConfig (.cpp:6-11)
            self.m_evaluator = DlTFNetworkEvaluator("", DF_HWC, **kw)
            if config is not None:
                self.m_evaluator.SetNetworkInput(config.get_network_input())
                outs = config.get_network_outputs()
                self.m_evaluator.SetNetworkOutput(outs)
                self.m_evaluator.set_value_transform(config.get_value_transform())
        else:
            self.m_evaluator = DlTFNetworkEvaluator(graphPath, **kw)
            if checkpoint:
                self.m_evaluator.UpdateCheckPoint(checkpoint)

    def LoadGraph(self, graphPath):
        self.m_evaluator.LoadGraph(graphPath)

    def UpdateCheckPoint(self, checkpoint):
        self.m_evaluator.UpdateCheckPoint(checkpoint)

    def EvaluateState(self, feature, actions_, value_, numBatches=None):
        self.m_evaluator.Evaluate(feature, actions_, value_, numBatches)

    def GraphLoaded(self):
        return self.m_evaluator.MetaGraphLoaded()


class LeafQueue:
    """ug_queue_*: multi-producer leaf batching in front of one evaluator."""

    def __init__(self, evaluator, max_batch, timeout_us=200, capacity=0):
        self._lib = _lib.load()
        self._ev = evaluator
        self._q = self._lib.ug_queue_create(evaluator._h, max_batch, timeout_us, capacity)
        if not self._q:
            raise RuntimeError("ug_queue_create failed: " + evaluator._err())

    def close(self):
        if getattr(self, "_q", None):
            self._lib.ug_queue_destroy(self._q)
            self._q = None

    __del__ = close

    def submit(self, packed_row, legal=None):
        row = np.ascontiguousarray(packed_row, np.uint32)
        lp = None
        if legal is not None:
            legal = np.ascontiguousarray(legal, np.uint32)
            lp = legal.ctypes.data
        t = self._lib.ug_queue_submit(self._q, row.ctypes.data, lp)
        if t < 0:
            raise RuntimeError("queue closed")
        return t

    def wait(self, ticket):
        pol = np.empty(NUM_MOVES, np.float32)
        val = C.c_float()
        if self._lib.ug_queue_wait(self._q, ticket, pol.ctypes.data, C.byref(val)) != 0:
            raise RuntimeError("queue evaluation failed: " + self._ev._err())
        return pol, val.value

    def stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._lib.ug_queue_stats(self._q, C.byref(a), C.byref(b), C.byref(c))
        return {"batches": a.value, "leaves": b.value, "flush_timeouts": c.value}

    def swap_checkpoint(self, prefix):
        return self._lib.ug_queue_swap_checkpoint(self._q, prefix.encode()) == 0

    def set_compact_priors(self, on=True):
        """leaves submitted with a legal mask get back only their #legal priors, packed in ascending move order"""
        self._lib.ug_queue_set_compact_priors(self._q, 1 if on else 0)

    def stats_ex(self):
        out = (C.c_uint64 * 9)()
        self._lib.ug_queue_stats_ex(self._q, out, 9)
        names = ("batches", "leaves", "flush_timeouts", "full_batches", "wave_cut_batches", "failed_batches",
                 "max_queue_depth", "gather_sleeps", "gpu_busy_us")
        return dict(zip(names, (int(v) for v in out)))


class ProductBoard:
    """test hook over the self-play driver's incremental board (csrc/goboard.h)"""

    def __init__(self):
        self._lib = _lib.load()
        self._b = self._lib.ug_board_new()

    def __del__(self):
        if getattr(self, "_b", None):
            self._lib.ug_board_free(self._b)
            self._b = None

    @property
    def to_play(self):
        return self._lib.ug_board_to_play(self._b)

    def legal(self, idx):
        return bool(self._lib.ug_board_legal(self._b, idx))

    def play(self, idx):
        self._lib.ug_board_play(self._b, idx)

    def pack(self):
        out = np.zeros(PACKED_POS_BYTES // 4, np.uint32)
        self._lib.ug_board_pack(self._b, out.ctypes.data)
        return out

    def score(self, komi=7.5):
        return float(self._lib.ug_board_score(self._b, komi))

    def planes(self):
        out = np.zeros((NUM_PLANES, 19, 19), np.int8)
        self._lib.ug_board_planes(self._b, out.ctypes.data)
        return out

    def generate(self):
        moves = np.zeros(NUM_MOVES, np.uint16)
        legal = np.zeros(MASK_WORDS, np.uint32)
        n = self._lib.ug_board_generate(self._b, moves.ctypes.data, legal.ctypes.data)
        return moves[:n].copy(), legal


def pack_planes(planes):
    """ug_pack_planes: CollectFeatures planes int8 [n][17][19][19] (or one [17][19][19]) -> packed uint32 [n][196]"""
    lib = _lib.load()
    planes = np.ascontiguousarray(planes, np.int8).reshape(-1, NUM_PLANES * 361)
    out = np.zeros((planes.shape[0], PACKED_POS_BYTES // 4), np.uint32)
    for i in range(planes.shape[0]):
        lib.ug_pack_planes(planes[i].ctypes.data, out[i].ctypes.data)
    return out


def selfplay(evaluator, games, playouts_per_move, threads=4, max_moves=0, max_batch=256, timeout_us=200,
             reuse_tree=True, random_opening_moves=0, puct=0.0, seed=1, record_moves=False, record_dir=None,
             komi=0.0, game_index_base=0, leaves_per_game=1, prelude_moves=0):
    """ug_selfplay_run: returns (stats dict, list of move lists or None)"""
    lib = _lib.load()
    cfg = _lib.SelfplayConfig(games, threads, playouts_per_move, max_moves, max_batch, timeout_us,
                              1 if reuse_tree else 0, random_opening_moves, puct, seed,
                              record_dir.encode() if record_dir else None, komi, game_index_base, leaves_per_game,
                              prelude_moves)
    st = _lib.SelfplayStats()
    cap = (max_moves if max_moves > 0 else 722) + prelude_moves
    moves = n_moves = None
    if record_moves:
        moves = np.zeros((games, cap), np.int16)
        n_moves = np.zeros(games, np.int32)
    rc = lib.ug_selfplay_run(evaluator._h, C.byref(cfg), C.byref(st), moves.ctypes.data if record_moves else None,
                             n_moves.ctypes.data if record_moves else None)
    if rc != 0:
        raise RuntimeError(f"ug_selfplay_run failed ({rc}): " + evaluator._err())
    stats = {f: getattr(st, f) for f, _ in _lib.SelfplayStats._fields_}
    return stats, ([moves[i, :n_moves[i]].tolist() for i in range(games)] if record_moves else None)
