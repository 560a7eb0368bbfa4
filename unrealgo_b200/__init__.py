"""unrealgo_b200 — B200-native (sm_100a) drop-in for UnrealGo's batched network-evaluation path.

The product is the C-ABI CUDA library unrealgo_b200/lib/libugeval.so (include/ugeval.h); this package holds its
sources (csrc/), the C++ host shim mirroring the reference classes (host/), and thin Python bindings used by
the tests and bench.py.  No CPU fallback exists anywhere in this package.
"""
from ._lib import PREC_BF16, PREC_FP16, PREC_FP32, TR_FLIP_PROB, TR_FLIP_SIGN, TR_IDENTITY  # noqa: F401
from .evaluator import (DF_CHW, DF_HWC, DlConfig, DlTFNetworkEvaluator, DlTFRecordWriter, LeafQueue, ProductBoard,  # noqa: F401
                        UctBoardEvaluator, bundle_read_float, inspect_model, pack_planes, selfplay)
