// Compact 19x19 Go board for the self-play driver (host side of the hot path).
//
// The reference keeps a full GoBoard per search thread and, at every leaf, rebuilds the network input by
// undoing and replaying up to 8 moves (GoUctGlobalSearchState::CollectFeatures, gouct/GoUctGlobalSearch.h:
// 160-206).  This board instead carries a ring of the last 8 positions as packed stone masks, updated
// incrementally on every move, so producing a ug_packed_pos for the GPU encoder is a 768-byte copy
// (SURVEY.md 8(f).2).  Rules as used in-tree by the reference: SIMPLEKO, no suicide
// (gouct/GoUctSearch.cpp:93-96), move generation = empty, legal, not a simple eye of the side to move, plus
// PASS (GoUctGlobalSearchState::GenerateLegalMoves, gouct/GoUctGlobalSearch.h:307-330; GoEyeUtil::IsSimpleEye,
// go/GoEyeUtil.h:133-176).
#pragma once
#include <cstdint>
#include <cstring>

#include "../../include/ugeval.h"

namespace ugo {

constexpr int N = 19;
constexpr int W = N + 2;          // bordered row stride
constexpr int NB = W * W;         // 441
constexpr int PASS = 361;         // policy index of the pass move
enum : uint8_t { BLACK = 0, WHITE = 1, EMPTY = 2, BORDER = 3 };

inline int to_b(int idx) { return (idx / N + 1) * W + (idx % N + 1); }       // policy index -> bordered index
inline int to_idx(int b) { return (b / W - 1) * N + (b % W - 1); }

struct Board {
  uint8_t color[NB];
  uint16_t id[NB];        // block id (bordered index of one of its stones)
  uint16_t next[NB];      // circular list of the block's stones
  int16_t plibs[NB];      // pseudo-liberties, valid at block ids (0 <=> no liberty)
  uint16_t bsize[NB];     // stones in block, valid at block ids
  uint32_t masks[2][UG_MASK_WORDS];            // current stones by colour, bit = policy index
  uint32_t ring[UG_HISTORY][2][UG_MASK_WORDS]; // ring[head] = current position, head-1 = one move ago ...
  int head = 0;
  int hist = 1;           // valid ring entries (1..8)
  int to_play = BLACK;
  int ko = -1;            // bordered index forbidden by simple ko
  int move_number = 0;
  int passes = 0;         // consecutive passes

  Board() { clear(); }
  void clear() {
    for (int b = 0; b < NB; ++b) {
      const int y = b / W, x = b % W;
      color[b] = (y == 0 || y == W - 1 || x == 0 || x == W - 1) ? BORDER : EMPTY;
      id[b] = (uint16_t)b;
      next[b] = (uint16_t)b;
      plibs[b] = 0;
      bsize[b] = 0;
    }
    memset(masks, 0, sizeof(masks));
    memset(ring, 0, sizeof(ring));
    head = 0; hist = 1; to_play = BLACK; ko = -1; move_number = 0; passes = 0;
  }

  static inline void nbrs(int b, int out[4]) { out[0] = b - W; out[1] = b - 1; out[2] = b + 1; out[3] = b + W; }

  // legality of a stone at bordered index b for colour c: empty, not the ko point (which binds the side to move
  // only, go/GoBoard.h:796), not suicide
  bool legal(int b, int c) const {
    if (color[b] != EMPTY || (b == ko && c == to_play)) return false;
    int nb[4];
    nbrs(b, nb);
    // count adjacencies of b to each neighbouring block (pseudo-liberties count an empty point once per adjacency)
    int ids[4], adj[4], k = 0;
    for (int i = 0; i < 4; ++i) {
      const uint8_t col = color[nb[i]];
      if (col == EMPTY) return true;
      if (col == BORDER) continue;
      const int g = id[nb[i]];
      int j = 0;
      for (; j < k; ++j) if (ids[j] == g) { ++adj[j]; break; }
      if (j == k) { ids[k] = g; adj[k] = 1; ++k; }
    }
    for (int j = 0; j < k; ++j) {
      const bool own = color[ids[j]] == c;
      const int left = plibs[ids[j]] - adj[j];
      if (own ? left > 0 : left == 0) return true;  // joins a block with another liberty, or captures
    }
    return false;
  }

  // GoEyeUtil::IsSimpleEye(bd, p, c), go/GoEyeUtil.h:133-176
  bool simple_eye(int b, int c) const {
    int nb[4];
    nbrs(b, nb);
    int anchors[2], na = 0;
    for (int i = 0; i < 4; ++i) {
      const uint8_t col = color[nb[i]];
      if (col == BORDER) continue;
      if (col != c) return false;  // empty or opponent neighbour
      const int g = id[nb[i]];
      if ((na > 0 && anchors[0] == g) || (na > 1 && anchors[1] == g)) continue;
      if (na > 1) return false;
      anchors[na++] = g;
    }
    if (na == 1) return true;
    if (na == 0) return false;
    // two blocks: a simple eye only if they share a second point surrounded by exactly these two blocks
    int s = anchors[0];
    do {
      int sn[4];
      nbrs(s, sn);
      for (int i = 0; i < 4; ++i) {
        const int lib = sn[i];
        if (color[lib] != EMPTY || lib == b) continue;
        bool shared = true;
        int found[2], nf = 0;
        int ln[4];
        nbrs(lib, ln);
        for (int j = 0; j < 4 && shared; ++j) {
          const uint8_t col = color[ln[j]];
          if (col == BORDER) continue;
          if (col != c) { shared = false; break; }
          const int g = id[ln[j]];
          if (g != anchors[0] && g != anchors[1]) { shared = false; break; }
          if (!((nf > 0 && found[0] == g) || (nf > 1 && found[1] == g))) found[nf++] = g;
        }
        if (shared && nf == 2) return true;
      }
      s = next[s];
    } while (s != anchors[0]);
    return false;
  }

  void set_bit(int c, int idx) { masks[c][idx >> 5] |= 1u << (idx & 31); }
  void clr_bit(int c, int idx) { masks[c][idx >> 5] &= ~(1u << (idx & 31)); }

  void remove_block(int g, int* captured_point, int* n_captured) {
    const int c = color[g];
    int s = g;
    do {
      const int nx = next[s];
      color[s] = EMPTY;
      clr_bit(c, to_idx(s));
      int nb[4];
      nbrs(s, nb);
      for (int i = 0; i < 4; ++i)
        if (color[nb[i]] == BLACK || color[nb[i]] == WHITE) ++plibs[id[nb[i]]];
      *captured_point = s;
      ++*n_captured;
      id[s] = (uint16_t)s;
      next[s] = (uint16_t)s;
      s = nx;
    } while (s != g);
  }

  // GoBoard::Play(p) for the side to move (go/GoBoard.cpp:648-694); idx = policy index, 361 = pass.
  // The caller guarantees legality.
  void play(int idx) {
    const int c = to_play;
    ko = -1;
    if (idx == PASS) {
      ++passes;
    } else {
      passes = 0;
      const int b = to_b(idx);
      color[b] = (uint8_t)c;
      set_bit(c, idx);
      id[b] = (uint16_t)b;
      next[b] = (uint16_t)b;
      bsize[b] = 1;
      int nb[4];
      nbrs(b, nb);
      int lib = 0;
      for (int i = 0; i < 4; ++i) {
        if (color[nb[i]] == EMPTY) ++lib;
        else if (color[nb[i]] != BORDER) --plibs[id[nb[i]]];
      }
      plibs[b] = (int16_t)lib;
      int captured_point = -1, n_captured = 0;
      for (int i = 0; i < 4; ++i)
        if (color[nb[i]] == 1 - c && plibs[id[nb[i]]] == 0) remove_block(id[nb[i]], &captured_point, &n_captured);
      for (int i = 0; i < 4; ++i) {
        if (color[nb[i]] != c) continue;
        int g1 = id[b], g2 = id[nb[i]];
        if (g1 == g2) continue;
        if (bsize[g1] < bsize[g2]) { const int t = g1; g1 = g2; g2 = t; }  // relabel the smaller block
        int s = g2;
        do { id[s] = (uint16_t)g1; s = next[s]; } while (s != g2);
        const uint16_t t = next[g1];
        next[g1] = next[g2];
        next[g2] = t;
        plibs[g1] = (int16_t)(plibs[g1] + plibs[g2]);
        bsize[g1] = (uint16_t)(bsize[g1] + bsize[g2]);
      }
      if (n_captured == 1 && bsize[id[b]] == 1 && plibs[id[b]] == 1) ko = captured_point;
    }
    to_play = 1 - c;
    ++move_number;
    head = (head + 1) & (UG_HISTORY - 1);
    memcpy(ring[head], masks, sizeof(masks));
    if (hist < UG_HISTORY) ++hist;
  }

  // ug_packed_pos of the current position: history step k = ring[head-k]; beyond the recorded history the
  // oldest position repeats (the reference stops undoing when GetLastMove() is GO_NULLMOVE).
  void pack(ug_packed_pos* out) const {
    for (int k = 0; k < UG_HISTORY; ++k) {
      const int kk = k < hist ? k : hist - 1;
      const int slot = (head - kk) & (UG_HISTORY - 1);
      memcpy(out->black[k], ring[slot][BLACK], sizeof(out->black[k]));
      memcpy(out->white[k], ring[slot][WHITE], sizeof(out->white[k]));
    }
    out->to_play = (uint32_t)to_play;
    out->reserved[0] = out->reserved[1] = out->reserved[2] = 0;
  }

  // GoBoardUtil::TrompTaylorScore(bd, komi), go/GoBoardUtil.h:626-680: stones plus empty regions that touch one
  // colour only, minus komi; positive = Black ahead.
  float tromp_taylor(float komi) const {
    float score = -komi;
    bool mark[NB] = {false};
    int stack[NB];
    for (int b = 0; b < NB; ++b) {
      if (mark[b] || color[b] == BORDER) continue;
      if (color[b] == BLACK) { score += 1.f; continue; }
      if (color[b] == WHITE) { score -= 1.f; continue; }
      int sp = 0, size = 0;
      bool adj[2] = {false, false};
      stack[sp++] = b;
      mark[b] = true;
      while (sp > 0) {
        const int p = stack[--sp];
        ++size;
        int nb[4];
        nbrs(p, nb);
        for (int i = 0; i < 4; ++i) {
          const uint8_t c = color[nb[i]];
          if (c == BLACK || c == WHITE) adj[c] = true;
          else if (c == EMPTY && !mark[nb[i]]) { mark[nb[i]] = true; stack[sp++] = nb[i]; }
        }
      }
      if (adj[BLACK] && !adj[WHITE]) score += (float)size;
      else if (!adj[BLACK] && adj[WHITE]) score -= (float)size;
    }
    return score;
  }

  // GenerateLegalMoves: returns the count; moves[] are policy indices (PASS last); legal_mask gets one bit per move.
  // No moves at all after two consecutive passes (gouct/GoUctGlobalSearch.h:311-314).
  int generate_moves(uint16_t* moves, uint32_t* legal_mask) const {
    memset(legal_mask, 0, UG_MASK_WORDS * sizeof(uint32_t));
    if (passes >= 2) return 0;
    int n = 0;
    for (int idx = 0; idx < 361; ++idx) {
      if (((masks[0][idx >> 5] | masks[1][idx >> 5]) >> (idx & 31)) & 1u) continue;
      const int b = to_b(idx);
      // fast path: an empty neighbour makes the move legal (it keeps a liberty) and rules out an eye
      const bool open = color[b - W] == EMPTY || color[b - 1] == EMPTY || color[b + 1] == EMPTY || color[b + W] == EMPTY;
      if (open ? b != ko : (!simple_eye(b, to_play) && legal(b, to_play))) {
        moves[n++] = (uint16_t)idx;
        legal_mask[idx >> 5] |= 1u << (idx & 31);
      }
    }
    moves[n++] = PASS;
    legal_mask[PASS >> 5] |= 1u << (PASS & 31);
    return n;
  }
};

}  // namespace ugo
