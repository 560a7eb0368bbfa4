/*
 * ugo_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the integer part of UnrealGo's network-evaluation path, used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg to check the CUDA path.  Nothing under
 * unrealgo_b200/ may link or call this file.
 *
 * Pinning: the reference holds no golden vectors or tests for this path (CollectFeatures is
 * an empty stub in search/test/UctSearchTest.cpp:98-102).  The functions below are pinned by
 *  (1) outputs of the reference itself run in the build container: its own GoBoard, GoBoardUtil and GoEyeUtil
 *      (go/GoBoard.cpp ... compiled where they lie into oracle/_ref/libugref.so, oracle/Makefile) played through
 *      seeded games, with the planes produced by driving that board the way CollectFeatures does — committed as
 *      tests/golden/ref_board.npz (generator tests/golden/make_ref_board_golden.py) and replayed bit for bit by
 *      tests/test_ref_board_golden.py: colours, side to move, GetLastMove, legality of all 362 moves for both
 *      colours (captures, simple ko, suicide), undo, and the 17 planes;
 *  (2) the reference's own DlConfig/ArrayUtil sources compiled into the same library;
 *  (3) self-made known-answer vectors in tests/golden/encoder_kat.npz.
 *      The CollectFeatures template itself is executed as well (oracle/_ref/ref_search `features`: the reference's
 *      whole search stack compiles): its planes after every move of 21 games are tests/golden/ref_collect_features.npz
 *      and og_collect_features reproduces all 2966.
 * What stays a restatement: TransformFeature's loop (TensorFlow-dependent file; its index formula GetOffset IS
 * pinned), value_transform and UpdatePrior as stand-alone functions (UpdatePrior runs inside the engine-level search
 * parity, tests/test_reference_search_parity.py).
 *
 * Each function cites the reference lines it follows.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N 19
#define NPTS 361
#define PASS 361        /* policy index of GO_PASS (GoPointUtil::Point2Index, board/GoPoint.h:139-146) */
#define NULLMOVE (-1)   /* GO_NULLMOVE, config/BoardStaticConfig.h:20 */
#define BLACK 0         /* SG_BLACK, board/GoBlackWhite.h:8 */
#define WHITE 1         /* SG_WHITE, board/GoBlackWhite.h:9 */
#define EMPTY 2         /* SG_EMPTY, board/GoBoardColor.h:9 */
#define NUM_MAPS 17
#define MAX_STACK 4096

typedef struct {
  int16_t point;        /* 0..360 or PASS */
  int8_t color;
  int8_t to_play_before; /* StackEntry::m_toPlay saved by SaveState (go/GoBoard.cpp:731-744) */
  int16_t ko_before;
  int16_t n_captured;
  int16_t captured[NPTS];
} entry_t;

typedef struct {
  int8_t color[NPTS];
  int to_play;
  int ko_point;         /* point forbidden by simple ko, or -1 */
  int n_moves;
  entry_t* stack;
} board_t;

/* ------------------------------------------------------------------ board (input generator) */
board_t* og_board_new(void) {
  board_t* b = (board_t*)calloc(1, sizeof(board_t));
  b->stack = (entry_t*)calloc(MAX_STACK, sizeof(entry_t));
  memset(b->color, EMPTY, NPTS);
  b->to_play = BLACK;
  b->ko_point = -1;
  return b;
}
void og_board_free(board_t* b) { if (b) { free(b->stack); free(b); } }
int og_board_to_play(const board_t* b) { return b->to_play; }
void og_board_set_to_play(board_t* b, int c) { b->to_play = c; } /* GoBoard::SetToPlay */
int og_board_move_number(const board_t* b) { return b->n_moves; }
int og_board_color(const board_t* b, int p) { return b->color[p]; }
void og_board_colors(const board_t* b, int8_t* out) { memcpy(out, b->color, NPTS); }

static int neighbors(int p, int* out) {
  int r = p / N, c = p % N, n = 0;
  if (r > 0) out[n++] = p - N;
  if (r < N - 1) out[n++] = p + N;
  if (c > 0) out[n++] = p - 1;
  if (c < N - 1) out[n++] = p + 1;
  return n;
}

/* flood-fill the block at p; returns its liberty count (0 = dead); members listed in `members` */
static int block_liberties(const board_t* b, int p, int16_t* members, int* n_members) {
  static int8_t mark[NPTS];
  static int16_t queue[NPTS];
  memset(mark, 0, NPTS);
  int head = 0, tail = 0, libs = 0, col = b->color[p];
  queue[tail++] = (int16_t)p;
  mark[p] = 1;
  while (head < tail) {
    int q = queue[head++];
    int nb[4];
    int k = neighbors(q, nb);
    for (int i = 0; i < k; ++i) {
      int t = nb[i];
      if (mark[t]) continue;
      if (b->color[t] == EMPTY) { mark[t] = 2; ++libs; }
      else if (b->color[t] == col) { mark[t] = 1; queue[tail++] = (int16_t)t; }
    }
  }
  if (members) { memcpy(members, queue, (size_t)tail * sizeof(int16_t)); *n_members = tail; }
  return libs;
}

/* 1 if playing p for `color` is legal: empty, not the simple-ko point (for the side to move), not suicide
 * (GoBoard::IsLegal with SIMPLEKO, as used in-tree: gouct/GoUctSearch.cpp:93-96) */
int og_board_is_legal(board_t* b, int p, int color) {
  if (p == PASS) return 1;
  if (p < 0 || p >= NPTS || b->color[p] != EMPTY) return 0;
  if (p == b->ko_point && color == b->to_play) return 0; /* go/GoBoard.h:796: the ko binds the side to move only */
  int nb[4];
  int k = neighbors(p, nb);
  for (int i = 0; i < k; ++i) if (b->color[nb[i]] == EMPTY) return 1;
  b->color[p] = (int8_t)color;
  int ok = 0;
  for (int i = 0; i < k && !ok; ++i) {
    int t = nb[i];
    if (b->color[t] == color) { if (block_liberties(b, t, NULL, NULL) > 0) ok = 1; }
    else { if (block_liberties(b, t, NULL, NULL) == 0) ok = 1; } /* captures something */
  }
  b->color[p] = EMPTY;
  return ok;
}

/* GoBoard::Play(p, player), go/GoBoard.cpp:648-694: push a stack entry (point, colour, saved to-play and ko),
 * place the stone, remove opponent blocks without liberties, set the ko point, to_play = opp(player).
 * A pass only flips to_play (:663-666). */
void og_board_play(board_t* b, int p, int color) {
  entry_t* e = &b->stack[b->n_moves++];
  e->point = (int16_t)p;
  e->color = (int8_t)color;
  e->to_play_before = (int8_t)b->to_play;
  e->ko_before = (int16_t)b->ko_point;
  e->n_captured = 0;
  b->ko_point = -1;
  if (p != PASS) {
    b->color[p] = (int8_t)color;
    int nb[4];
    int k = neighbors(p, nb);
    int16_t members[NPTS];
    for (int i = 0; i < k; ++i) {
      int t = nb[i];
      if (b->color[t] != 1 - color) continue;
      int nm = 0;
      if (block_liberties(b, t, members, &nm) == 0) {
        for (int j = 0; j < nm; ++j) {
          b->color[members[j]] = EMPTY;
          e->captured[e->n_captured++] = members[j];
        }
      }
    }
    if (e->n_captured == 1) {
      /* ko iff the capturing stone is a single stone with exactly one liberty (:680-682) */
      int nm = 0;
      int libs = block_liberties(b, p, members, &nm);
      if (nm == 1 && libs == 1) b->ko_point = e->captured[0];
    }
  }
  b->to_play = 1 - color;
}
void og_board_play_to_play(board_t* b, int p) { og_board_play(b, p, b->to_play); } /* GoBoard::Play(p), go/GoBoard.h:958-960 */

/* GoBoard::Undo, go/GoBoard.cpp:696-702 (+RestoreState :719-730) */
void og_board_undo(board_t* b) {
  entry_t* e = &b->stack[--b->n_moves];
  if (e->point != PASS) {
    b->color[e->point] = EMPTY;
    for (int j = 0; j < e->n_captured; ++j) b->color[e->captured[j]] = (int8_t)(1 - e->color);
  }
  b->to_play = e->to_play_before;
  b->ko_point = e->ko_before;
}

/* GoBoard::GetLastMove, go/GoBoard.h:659-667 */
int og_board_last_move(const board_t* b) {
  if (b->n_moves == 0) return NULLMOVE;
  const entry_t* e = &b->stack[b->n_moves - 1];
  if (e->color != 1 - b->to_play) return NULLMOVE;
  return e->point;
}

/* ------------------------------------------------------------------ a1: CollectFeatures */
/* GoUctGlobalSearchState::CollectFeatures(features, 17), gouct/GoUctGlobalSearch.h:160-206.
 * features is char[17][19][19] with index [plane][row-1][col-1]; point index = (row-1)*19 + (col-1).
 * The loop undoes a move after EVERY plane pair (also after the 8th, :184-188) and replays them with
 * Execute(move) = Play(move, ToPlay()) (:201-205, gouct/GoUctSearch.cpp:90-100). */
void og_collect_features(board_t* b, int8_t* features) {
  int sequence[8];
  int depth = 0;
  const int cur = b->to_play;
  const int opp = 1 - cur;
  const int map_count = (NUM_MAPS - 1) / 2;
  for (int i = 0; i < map_count; ++i) {
    for (int p = 0; p < NPTS; ++p) {
      const int c = b->color[p];
      features[(2 * i) * NPTS + p] = (c == cur) ? 1 : 0;
      features[(2 * i + 1) * NPTS + p] = (c == opp) ? 1 : 0;
    }
    const int last = og_board_last_move(b);
    if (last != NULLMOVE) {
      sequence[depth++] = last;
      og_board_undo(b);
    }
  }
  memset(features + (NUM_MAPS - 1) * NPTS, 1 - cur, NPTS);
  while (depth > 0) og_board_play_to_play(b, sequence[--depth]);
}

/* Test-input generator for the packed entry point: the same walk, recording black/white stone masks per
 * history step instead of cur/opp planes.  Layout = ug_packed_pos (include/ugeval.h): black[8][12],
 * white[8][12], to_play, reserved[3]; bit i of word i>>5 = point i. */
void og_pack_position(board_t* b, uint32_t* out /* 196 words */) {
  int sequence[8];
  int depth = 0;
  const int cur = b->to_play;
  memset(out, 0, 196 * sizeof(uint32_t));
  for (int i = 0; i < 8; ++i) {
    for (int p = 0; p < NPTS; ++p) {
      const int c = b->color[p];
      if (c == BLACK) out[i * 12 + (p >> 5)] |= 1u << (p & 31);
      else if (c == WHITE) out[96 + i * 12 + (p >> 5)] |= 1u << (p & 31);
    }
    const int last = og_board_last_move(b);
    if (last != NULLMOVE) {
      sequence[depth++] = last;
      og_board_undo(b);
    }
  }
  out[192] = (uint32_t)cur;
  while (depth > 0) og_board_play_to_play(b, sequence[--depth]);
}

/* ------------------------------------------------------------------ a7: TransformFeature */
/* UnrealGo::ArrayUtil::GetOffset, lib/ArrayUtil.h:14-25 */
static int get_offset(const int* dims, int ndims, const int* idx) {
  int index = 0, sub = 1;
  for (int i = 0; i < ndims; ++i) sub *= dims[i];
  for (int i = 0; i < ndims; ++i) { sub /= dims[i]; index += idx[i] * sub; }
  return index;
}
/* DlTFNetworkEvaluator<T>::TransformFeature (CHW input), funcapproximator/DlTFNetworkEvaluator.cc:366-389.
 * data_format 0 = DF_HWC (the engine's setting, search/UctBoardEvaluator.cpp:6), 1 = DF_CHW.
 * as_bool: T = bool ((bool)char, any non-zero -> 1) vs T = float-like (value kept). */
void og_transform_feature_chw(const int8_t* feature, int n, int data_format, int as_bool, int8_t* data) {
  for (int b = 0; b < n; ++b)
    for (int d = 0; d < NUM_MAPS; ++d)
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
          int index;
          if (data_format == 0) {
            const int dims[4] = {n, N, N, NUM_MAPS}, idx[4] = {b, i, j, d};
            index = get_offset(dims, 4, idx);
          } else {
            const int dims[4] = {n, NUM_MAPS, N, N}, idx[4] = {b, d, i, j};
            index = get_offset(dims, 4, idx);
          }
          const int8_t v = feature[((b * NUM_MAPS + d) * N + i) * N + j];
          data[index] = as_bool ? (int8_t)(v != 0) : v;
        }
}
/* HWC-input overload, .cc:392-412 */
void og_transform_feature_hwc(const int8_t* feature, int n, int data_format, int as_bool, int8_t* data) {
  for (int b = 0; b < n; ++b)
    for (int d = 0; d < NUM_MAPS; ++d)
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
          int index;
          if (data_format == 0) {
            const int dims[4] = {n, N, N, NUM_MAPS}, idx[4] = {b, i, j, d};
            index = get_offset(dims, 4, idx);
          } else {
            const int dims[4] = {n, NUM_MAPS, N, N}, idx[4] = {b, d, i, j};
            index = get_offset(dims, 4, idx);
          }
          const int8_t v = feature[((b * N + i) * N + j) * NUM_MAPS + d];
          data[index] = as_bool ? (int8_t)(v != 0) : v;
        }
}

/* ------------------------------------------------------------------ a12 / a11 */
/* GoPointUtil::Pt + Point2Index, board/GoPoint.h:139-153: point = 20*row + col (1-based), index = (row-1)*19+(col-1) */
int og_point2index(int go_point) {
  const int go_pass = (N * N + 3 * (N + 1)) + 1; /* GO_MAXPOINT + 1, config/BoardStaticConfig.h:26,33 */
  if (go_point == go_pass) return NPTS;
  const int row = go_point / (N + 1), col = go_point % (N + 1);
  return (row - 1) * N + (col - 1);
}
/* value transform switch, funcapproximator/DlTFNetworkEvaluator.cc:300-310 (double arithmetic) */
void og_value_transform(double* value, int n, int vt) {
  for (int i = 0; i < n; ++i) {
    if (vt == 1) value[i] = -value[i];
    else if (vt == 2) value[i] = 1 - value[i];
  }
}
/* UctSearch::UpdatePrior, search/UctSearch.cpp:1257-1270: children = legal move indices */
void og_update_prior(const double* policies, const int* child_index, int n_children, double* prior_out) {
  double legal_sum = 0;
  for (int i = 0; i < n_children; ++i) legal_sum += policies[child_index[i]];
  for (int i = 0; i < n_children; ++i) prior_out[i] = policies[child_index[i]] / legal_sum;
}
/* DlTensorUtil<float>::GetValue, funcapproximator/DlTensorUtil.cc:247-250 */
void og_get_value(const float* tensor, double* out, int len) {
  for (int i = 0; i < len; ++i) out[i] = tensor[i];
}

/* CRC-32C for the independent bundle reader (oracle/tf_reader.py) */
uint32_t og_crc32c(const uint8_t* p, size_t n) {
  static uint32_t table[256];
  static int init = 0;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
      table[i] = c;
    }
    init = 1;
  }
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
