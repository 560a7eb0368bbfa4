"""TEST INFRASTRUCTURE — CPU oracle for the UnrealGo network-evaluation path.

ctypes front-end of oracle/ugo_oracle.c (integer part: board, CollectFeatures, TransformFeature, UpdatePrior,
value_transform) plus the seeded position generator of SURVEY.md 8(d).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package; the product (unrealgo_b200/)
never does.  Parity status: see the header of ugo_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

NPTS, PASS, NUM_MAPS = 361, 361, 17
BLACK, WHITE, EMPTY = 0, 1, 2


def build():
    """compile liboracle.so (and oracle/_ref when the reference tree is present)"""
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.og_board_new.restype = C.c_void_p
        for name, args, res in [
            ("og_board_free", [C.c_void_p], None),
            ("og_board_to_play", [C.c_void_p], C.c_int),
            ("og_board_set_to_play", [C.c_void_p, C.c_int], None),
            ("og_board_move_number", [C.c_void_p], C.c_int),
            ("og_board_color", [C.c_void_p, C.c_int], C.c_int),
            ("og_board_colors", [C.c_void_p, C.c_void_p], None),
            ("og_board_is_legal", [C.c_void_p, C.c_int, C.c_int], C.c_int),
            ("og_board_play", [C.c_void_p, C.c_int, C.c_int], None),
            ("og_board_play_to_play", [C.c_void_p, C.c_int], None),
            ("og_board_undo", [C.c_void_p], None),
            ("og_board_last_move", [C.c_void_p], C.c_int),
            ("og_collect_features", [C.c_void_p, C.c_void_p], None),
            ("og_pack_position", [C.c_void_p, C.c_void_p], None),
            ("og_transform_feature_chw", [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p], None),
            ("og_transform_feature_hwc", [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p], None),
            ("og_point2index", [C.c_int], C.c_int),
            ("og_value_transform", [C.c_void_p, C.c_int, C.c_int], None),
            ("og_update_prior", [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p], None),
            ("og_get_value", [C.c_void_p, C.c_void_p, C.c_int], None),
            ("og_crc32c", [C.c_char_p, C.c_size_t], C.c_uint32),
        ]:
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _LIB = L
    return _LIB


def ref_lib():
    """oracle/_ref/libugref.so: the reference's own DlConfig/StringUtil/ArrayUtil compiled here; None if absent"""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libugref.so")
        if not os.path.exists(path):
            try:
                build()
            except Exception:
                return None
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_dlcfg_open.restype = C.c_void_p
        R.ref_dlcfg_open.argtypes = [C.c_char_p]
        R.ref_dlcfg_close.argtypes = [C.c_void_p]
        R.ref_dlcfg_get.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        for n in ("ref_dlcfg_network_input", "ref_dlcfg_metagraph", "ref_dlcfg_checkpoint"):
            getattr(R, n).argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        R.ref_dlcfg_network_outputs.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        R.ref_dlcfg_value_transform.argtypes = [C.c_void_p]
        R.ref_dlcfg_reuse_search_tree.argtypes = [C.c_void_p]
        R.ref_get_offset.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        if hasattr(R, "ref_board_new"):
            R.ref_board_new.restype = C.c_void_p
            R.ref_board_free.argtypes = [C.c_void_p]
            for n in ("ref_board_to_play", "ref_board_move_number", "ref_board_undo", "ref_board_last_move",
                      "ref_board_ko_point"):
                getattr(R, n).argtypes = [C.c_void_p]
            R.ref_board_undo.restype = None
            R.ref_board_set_to_play.argtypes = [C.c_void_p, C.c_int]
            R.ref_board_is_legal.argtypes = [C.c_void_p, C.c_int, C.c_int]
            R.ref_board_play.argtypes = [C.c_void_p, C.c_int, C.c_int]
            R.ref_board_colors.argtypes = [C.c_void_p, C.c_void_p]
            R.ref_collect_features.argtypes = [C.c_void_p, C.c_void_p]
            R.ref_board_tromp_taylor.argtypes = [C.c_void_p, C.c_float]
            R.ref_board_tromp_taylor.restype = C.c_float
            R.ref_board_is_simple_eye.argtypes = [C.c_void_p, C.c_int, C.c_int]
            R.ref_board_two_passes.argtypes = [C.c_void_p]
            R.ref_board_generate.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _REF = R
    return _REF


class Board:
    """restated GoBoard subset (go/GoBoard.cpp:648-702): play/undo with captures and simple ko"""

    def __init__(self):
        self._b = lib().og_board_new()

    def __del__(self):
        if getattr(self, "_b", None):
            lib().og_board_free(self._b)
            self._b = None

    @property
    def to_play(self):
        return lib().og_board_to_play(self._b)

    def set_to_play(self, c):
        lib().og_board_set_to_play(self._b, c)

    @property
    def move_number(self):
        return lib().og_board_move_number(self._b)

    def colors(self):
        out = np.empty(NPTS, np.int8)
        lib().og_board_colors(self._b, out.ctypes.data)
        return out

    def is_legal(self, p, color=None):
        return bool(lib().og_board_is_legal(self._b, p, self.to_play if color is None else color))

    def play(self, p, color=None):
        if color is None:
            lib().og_board_play_to_play(self._b, p)
        else:
            lib().og_board_play(self._b, p, color)

    def undo(self):
        lib().og_board_undo(self._b)

    def last_move(self):
        return lib().og_board_last_move(self._b)

    def collect_features(self):
        """CollectFeatures restatement -> int8 [17][19][19]"""
        out = np.empty((NUM_MAPS, 19, 19), np.int8)
        lib().og_collect_features(self._b, out.ctypes.data)
        return out

    def pack(self):
        """-> uint32[196] in ug_packed_pos layout"""
        out = np.zeros(196, np.uint32)
        lib().og_pack_position(self._b, out.ctypes.data)
        return out


class RefBoard:
    """The reference's OWN GoBoard (go/GoBoard.cpp compiled into oracle/_ref/libugref.so by oracle/Makefile), ko rule
    SIMPLEKO as GoUctState::Execute sets it (gouct/GoUctSearch.cpp:90-100).  Same interface as Board; exists only
    where the reference tree was present at build time — used to pin Board and to generate tests/golden/ref_board.npz."""

    def __init__(self):
        self._r = ref_lib()
        if self._r is None or not hasattr(self._r, "ref_board_new"):
            raise RuntimeError("oracle/_ref/libugref.so (with the board) is not built")
        self._b = self._r.ref_board_new()

    def __del__(self):
        if getattr(self, "_b", None):
            self._r.ref_board_free(self._b)
            self._b = None

    @property
    def to_play(self):
        return self._r.ref_board_to_play(self._b)

    def set_to_play(self, c):
        self._r.ref_board_set_to_play(self._b, c)

    @property
    def move_number(self):
        return self._r.ref_board_move_number(self._b)

    def colors(self):
        out = np.empty(NPTS, np.int8)
        self._r.ref_board_colors(self._b, out.ctypes.data)
        return out

    def is_legal(self, p, color=None):
        return bool(self._r.ref_board_is_legal(self._b, p, self.to_play if color is None else color))

    def play(self, p, color=None):
        """-> True when the reference board flagged the move illegal"""
        return bool(self._r.ref_board_play(self._b, p, self.to_play if color is None else color))

    def undo(self):
        self._r.ref_board_undo(self._b)

    def last_move(self):
        return self._r.ref_board_last_move(self._b)

    def ko_point(self):
        return self._r.ref_board_ko_point(self._b)

    def is_simple_eye(self, p, color=None):
        """GoEyeUtil::IsSimpleEye of the reference (go/GoEyeUtil.h:133-176) on an empty point"""
        return bool(self._r.ref_board_is_simple_eye(self._b, p, self.to_play if color is None else color))

    def generate(self, initial_move_number=0):
        """the move list of GenerateLegalMoves (gouct/GoUctGlobalSearch.h:308-330), as policy indices"""
        out = np.zeros(362, np.int16)
        n = self._r.ref_board_generate(self._b, initial_move_number, out.ctypes.data)
        return out[:n].copy()

    def score(self, komi=7.5):
        """GoBoardUtil::TrompTaylorScore of the reference (go/GoBoardUtil.h:626-680), positive = Black ahead"""
        return float(self._r.ref_board_tromp_taylor(self._b, komi))

    def collect_features(self):
        out = np.empty((NUM_MAPS, 19, 19), np.int8)
        self._r.ref_collect_features(self._b, out.ctypes.data)
        return out


def transform_feature(planes, data_format=0, as_bool=True, hwc_input=False):
    """TransformFeature restatement.  planes int8 [n][17][19][19] (or [n][19][19][17] with hwc_input)."""
    planes = np.ascontiguousarray(planes, np.int8)
    n = planes.shape[0]
    out = np.empty(planes.size, np.int8)
    fn = lib().og_transform_feature_hwc if hwc_input else lib().og_transform_feature_chw
    fn(planes.ctypes.data, n, data_format, 1 if as_bool else 0, out.ctypes.data)
    return out.reshape((n, 19, 19, NUM_MAPS) if data_format == 0 else (n, NUM_MAPS, 19, 19))


def value_transform(values, vt):
    v = np.ascontiguousarray(values, np.float64).copy()
    lib().og_value_transform(v.ctypes.data, v.size, vt)
    return v


def update_prior(policy, children):
    policy = np.ascontiguousarray(policy, np.float64)
    ch = np.ascontiguousarray(children, np.int32)
    out = np.empty(ch.size, np.float64)
    lib().og_update_prior(policy.ctypes.data, ch.ctypes.data, ch.size, out.ctypes.data)
    return out


def random_game(rng, length, pass_prob=0.02):
    """seeded random-legal-move playout (no eye filling heuristics needed: it only has to be legal)"""
    b = Board()
    for _ in range(length):
        if rng.random() < pass_prob:
            b.play(PASS)
            continue
        for _try in range(40):
            p = int(rng.integers(0, NPTS))
            if b.is_legal(p):
                b.play(p)
                break
        else:
            b.play(PASS)
    return b


def synthetic_positions(n, seed=20260119):
    """SURVEY.md 8(d) position set: game lengths uniform in [0, 300], both colours to move, >= 10 % with fewer
    than 8 moves of history, >= 5 % with a pass in the last 8 moves.
    Returns (packed uint32 [n][196], planes int8 [n][17][19][19])."""
    rng = np.random.default_rng(seed)
    packed = np.zeros((n, 196), np.uint32)
    planes = np.zeros((n, NUM_MAPS, 19, 19), np.int8)
    for i in range(n):
        u = i / max(n, 1)
        if u < 0.12:
            length = int(rng.integers(0, 8))
            b = random_game(rng, length)
        elif u < 0.19:
            length = int(rng.integers(8, 301))
            b = random_game(rng, length)
            b.play(PASS)  # a pass inside the last 8 moves
            for _ in range(int(rng.integers(0, 6))):
                for _try in range(40):
                    p = int(rng.integers(0, NPTS))
                    if b.is_legal(p):
                        b.play(p)
                        break
        else:
            length = int(rng.integers(0, 301))
            b = random_game(rng, length)
        packed[i] = b.pack()
        planes[i] = b.collect_features()
    return packed, planes
