"""ctypes binding of libugeval.so (C ABI declared in include/ugeval.h).

The library is the product; there is no Python or CPU fallback.  Importing this module without the built
library raises, and creating an evaluator without a B200 raises (ug_eval_create returns NULL).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libugeval.so")

NUM_POINTS = 361
NUM_MOVES = 362
NUM_PLANES = 17
HISTORY = 8
MASK_WORDS = 12
PACKED_POS_BYTES = 784

TR_IDENTITY, TR_FLIP_SIGN, TR_FLIP_PROB = 0, 1, 2
PREC_BF16, PREC_FP32, PREC_FP16 = 0, 1, 2


class PackedPos(C.Structure):
    _fields_ = [
        ("black", (C.c_uint32 * MASK_WORDS) * HISTORY),
        ("white", (C.c_uint32 * MASK_WORDS) * HISTORY),
        ("to_play", C.c_uint32),
        ("reserved", C.c_uint32 * 3),
    ]


class NetInfo(C.Structure):
    _fields_ = [
        ("blocks", C.c_int), ("filters", C.c_int), ("policy_channels", C.c_int), ("value_channels", C.c_int),
        ("value_hidden", C.c_int), ("input_planes", C.c_int), ("flops_per_position", C.c_double),
        ("conv_ctas", C.c_int), ("sm_count", C.c_int), ("conv_impl", C.c_int), ("small_batch_max", C.c_int),
        ("precision", C.c_int), ("calib_max_activation", C.c_float), ("calib_max_weight", C.c_float),
        ("fp16_headroom", C.c_float), ("fp16_stream_shift", C.c_int), ("fp16_max_block_shift", C.c_int),
        ("saturation_events", C.c_uint64),
    ]


class SelfplayConfig(C.Structure):
    _fields_ = [("games", C.c_int), ("threads", C.c_int), ("playouts_per_move", C.c_int), ("max_moves", C.c_int),
                ("max_batch", C.c_int), ("timeout_us", C.c_int), ("reuse_tree", C.c_int),
                ("random_opening_moves", C.c_int), ("puct", C.c_float), ("seed", C.c_uint64),
                ("record_dir", C.c_char_p), ("komi", C.c_float), ("game_index_base", C.c_int),
                ("leaves_per_game", C.c_int), ("prelude_moves", C.c_int)]


class SelfplayStats(C.Structure):
    _fields_ = [("playouts", C.c_uint64), ("evaluations", C.c_uint64), ("batches", C.c_uint64),
                ("flush_timeouts", C.c_uint64), ("moves", C.c_uint64), ("games_finished", C.c_uint64),
                ("terminal_leaves", C.c_uint64), ("seconds", C.c_double), ("playouts_per_s", C.c_double),
                ("mean_batch", C.c_double), ("mean_tree_nodes", C.c_double),
                ("search_busy_fraction", C.c_double), ("host_us_per_playout", C.c_double),
                ("full_batches", C.c_uint64), ("wave_cut_batches", C.c_uint64), ("max_queue_depth", C.c_uint64),
                ("gather_sleeps", C.c_uint64), ("gpu_busy_fraction", C.c_double),
                ("idle_launches", C.c_uint64), ("idle_gap_seconds", C.c_double)]


class BenchResult(C.Structure):
    _fields_ = [("ms_total", C.c_float), ("ms_encoder", C.c_float), ("ms_tower", C.c_float),
                ("ms_heads", C.c_float), ("launches_per_pass", C.c_int)]


# name -> (restype, argtypes); every symbol include/ugeval.h declares
SIGNATURES = {
    "ug_eval_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int]),
    "ug_eval_destroy": (None, [C.c_void_p]),
    "ug_last_error": (C.c_char_p, [C.c_void_p]),
    "ug_eval_set_io": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int]),
    "ug_eval_load_graph": (C.c_int, [C.c_void_p, C.c_char_p]),
    "ug_eval_load_checkpoint": (C.c_int, [C.c_void_p, C.c_char_p]),
    "ug_eval_ready": (C.c_int, [C.c_void_p]),
    "ug_eval_has_weights": (C.c_int, [C.c_void_p]),
    "ug_eval_set_value_transform": (C.c_int, [C.c_void_p, C.c_int]),
    "ug_eval_info": (C.c_int, [C.c_void_p, C.POINTER(NetInfo)]),
    "ug_eval_set_conv_ctas": (C.c_int, [C.c_void_p, C.c_int]),
    "ug_eval_efficient_batch": (C.c_int, [C.c_void_p, C.c_int]),
    "ug_efficient_batch_for_sms": (C.c_int, [C.c_int, C.c_int]),
    "ug_eval_set_conv_impl": (C.c_int, [C.c_void_p, C.c_int]),
    "ug_eval_set_residual_split": (C.c_int, [C.c_void_p, C.c_int]),
    "ug_eval_run_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ug_eval_run_planes_hwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ug_eval_register_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ug_eval_unregister_host": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ug_eval_run_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ug_eval_run_packed_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "ug_eval_sync": (C.c_int, [C.c_void_p]),
    "ug_encode_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ug_encode_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ug_transform_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ug_eval_debug_layer": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ug_eval_set_debug_stop": (C.c_int, [C.c_void_p, C.c_int]),
    "ug_eval_bench": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(BenchResult)]),
    "ug_k_conv3x3_bf16": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ug_k_conv3x3_v2": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_float), C.c_void_p]),
    "ug_k_conv3x3_sb": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "ug_eval_set_small_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ug_eval_set_fp16_range_policy": (C.c_int, [C.c_void_p, C.c_int]),
    "ug_debug_max_active_clusters": (C.c_int, [C.c_int, C.c_int]),
    "ug_queue_create": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "ug_queue_destroy": (None, [C.c_void_p]),
    "ug_queue_submit": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ug_queue_wait": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ug_queue_poll": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ug_selfplay_run": (C.c_int, [C.c_void_p, C.POINTER(SelfplayConfig), C.POINTER(SelfplayStats), C.c_void_p, C.c_void_p]),
    "ug_board_new": (C.c_void_p, []),
    "ug_board_free": (None, [C.c_void_p]),
    "ug_board_legal": (C.c_int, [C.c_void_p, C.c_int]),
    "ug_board_play": (None, [C.c_void_p, C.c_int]),
    "ug_board_pack": (None, [C.c_void_p, C.c_void_p]),
    "ug_board_generate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ug_board_to_play": (C.c_int, [C.c_void_p]),
    "ug_board_score": (C.c_float, [C.c_void_p, C.c_float]),
    "ug_board_planes": (None, [C.c_void_p, C.c_void_p]),
    "ug_pack_planes": (None, [C.c_void_p, C.c_void_p]),
    "ug_queue_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "ug_queue_swap_checkpoint": (C.c_int, [C.c_void_p, C.c_char_p]),
    "ug_queue_set_compact_priors": (C.c_int, [C.c_void_p, C.c_int]),
    "ug_queue_stats_ex": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]),
    "ug_tfrecord_open": (C.c_void_p, [C.c_char_p]),
    "ug_tfrecord_write_example": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_float]),
    "ug_tfrecord_write_example_named": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_char_p, C.c_void_p,
                                                  C.c_size_t, C.c_char_p, C.c_float]),
    "ug_tfrecord_flush": (C.c_int, [C.c_void_p]),
    "ug_tfrecord_close": (C.c_int64, [C.c_void_p]),
    "ug_ckptinfo_read": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    "ug_ckptinfo_write": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    "ug_checkpoint_files_exist": (C.c_int, [C.c_char_p]),
    "ug_dlcfg_open": (C.c_void_p, [C.c_char_p]),
    "ug_dlcfg_close": (None, [C.c_void_p]),
    "ug_dlcfg_get": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    "ug_dlcfg_value_transform": (C.c_int, [C.c_void_p]),
    "ug_dlcfg_reuse_search_tree": (C.c_int, [C.c_void_p]),
    "ug_dlcfg_network_outputs": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "ug_inspect_model": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    "ug_bundle_read_float": (C.c_longlong, [C.c_char_p, C.c_char_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                            C.c_char_p, C.c_int]),
    "ug_version": (C.c_char_p, []),
    "ug_crc32c": (C.c_uint32, [C.c_void_p, C.c_size_t]),
}

_lib = None


def load():
    """Load libugeval.so and bind every declared symbol; raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
