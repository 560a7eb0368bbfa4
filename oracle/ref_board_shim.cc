// extern "C" wrappers over the reference's OWN GoBoard, compiled from /root/reference where it lies
// (go/GoBoard.cpp, go/GoBoardUtil.cpp, go/GoRules.cpp, go/GoKomi.cpp, board/*.cpp it links, search/SgProp.cpp,
// lib/SgRandom.cpp, platform/Sg{System,Debug,Exception}.cpp — the list is REF_SRC in oracle/Makefile; Boost stand-ins
// under ref_shim/).  Output: oracle/_ref/libugref.so.  Test infrastructure only — used to pin the oracle's board
// restatement (captures, simple ko, suicide, undo, the GetLastMove null rule), its feature planes, and the product
// host board's move generator (simple eyes, two-pass rule) and Tromp-Taylor score against the real board.
//
// ref_collect_features drives the real board exactly the way GoUctGlobalSearchState::CollectFeatures does
// (gouct/GoUctGlobalSearch.h:160-206): the template itself sits on top of the whole search stack
// (UctSearch, threads, players) and does not compile here, but every board operation it performs —
// GetColor, GetLastMove, Undo (GoUctState::TakeBackInTree, gouct/GoUctSearch.cpp:154-157) and Play under the
// SIMPLEKO rule (GoUctState::Execute, :90-100) — is executed by the reference's own code.
#include <cstdint>
#include <cstring>
#include <vector>

#include "platform/SgSystem.h"

#include "board/GoPoint.h"
#include "go/GoBoard.h"
#include "go/GoBoardUtil.h"
#include "go/GoEyeUtil.h"

// go/GoInit.cpp registers SGF properties and needs the whole go/ library; the board only asks whether it ran.
void GoInitCheck() {}

namespace {
inline GoPoint to_point(int idx) {
  if (idx == GO_MAX_ONBOARD) return GO_PASS;
  return GoPointUtil::Pt(idx % GO_MAX_SIZE + 1, idx / GO_MAX_SIZE + 1);
}
inline int to_index(GoMove m) { return m == GO_NULLMOVE ? -1 : GoPointUtil::Point2Index(m); }
struct RefBoard {
  GoBoard bd;
  RefBoard() : bd(GO_MAX_SIZE) { bd.Rules().SetKoRule(GoRules::SIMPLEKO); }
};
}  // namespace

extern "C" {

void* ref_board_new() { return new RefBoard(); }
void ref_board_free(void* b) { delete static_cast<RefBoard*>(b); }
int ref_board_to_play(void* b) { return static_cast<RefBoard*>(b)->bd.ToPlay(); }
void ref_board_set_to_play(void* b, int c) { static_cast<RefBoard*>(b)->bd.SetToPlay(c); }
int ref_board_move_number(void* b) { return static_cast<RefBoard*>(b)->bd.MoveNumber(); }
int ref_board_is_legal(void* b, int idx, int color) {
  return static_cast<RefBoard*>(b)->bd.IsLegal(to_point(idx), color) ? 1 : 0;
}
// returns 1 when the board flagged the move illegal (it is still on the stack, as in the reference)
int ref_board_play(void* b, int idx, int color) {
  GoBoard& bd = static_cast<RefBoard*>(b)->bd;
  bd.Play(to_point(idx), color);
  return bd.LastMoveInfo(GO_MOVEFLAG_ILLEGAL) ? 1 : 0;
}
void ref_board_undo(void* b) { static_cast<RefBoard*>(b)->bd.Undo(); }
int ref_board_last_move(void* b) { return to_index(static_cast<RefBoard*>(b)->bd.GetLastMove()); }
void ref_board_colors(void* b, int8_t* out) {
  const GoBoard& bd = static_cast<RefBoard*>(b)->bd;
  for (GoBoard::Iterator it(bd); it; ++it) out[GoPointUtil::Point2Index(*it)] = (int8_t)bd.GetColor(*it);
}
int ref_board_ko_point(void* b) {
  const GoPoint k = static_cast<RefBoard*>(b)->bd.KoPoint();
  return k == GO_NULLPOINT ? -1 : GoPointUtil::Point2Index(k);
}

// GoBoardUtil::TrompTaylorScore (go/GoBoardUtil.h:626-680, a header template): positive = Black ahead
float ref_board_tromp_taylor(void* b, float komi) {
  return GoBoardUtil::TrompTaylorScore(static_cast<RefBoard*>(b)->bd, komi);
}

int ref_board_is_simple_eye(void* b, int idx, int color) {
  const GoBoard& bd = static_cast<RefBoard*>(b)->bd;
  const GoPoint p = to_point(idx);
  return bd.IsEmpty(p) && GoEyeUtil::IsSimpleEye(bd, p, color) ? 1 : 0;
}
int ref_board_two_passes(void* b) { return GoBoardUtil::TwoPasses(static_cast<RefBoard*>(b)->bd) ? 1 : 0; }

// The move list GoUctGlobalSearchState::GenerateLegalMoves builds (gouct/GoUctGlobalSearch.h:308-330) with no
// safety solver (AllSafe false everywhere, the default): nothing after two passes, else every empty point that is
// neither a simple eye of the side to move nor illegal, then the pass.  IsSimpleEye, IsLegal and TwoPasses are the
// reference's own; the template around them belongs to the search stack and does not compile here.
int ref_board_generate(void* b, int initial_move_number, int16_t* out /* [362] */) {
  const GoBoard& bd = static_cast<RefBoard*>(b)->bd;
  if (GoBoardUtil::TwoPasses(bd) && (bd.Rules().CaptureDead() || bd.MoveNumber() - initial_move_number >= 2)) return 0;
  const SgBlackWhite mover = bd.ToPlay();
  int n = 0;
  for (GoBoard::Iterator it(bd); it; ++it)
    if (bd.IsEmpty(*it) && !GoEyeUtil::IsSimpleEye(bd, *it, mover) && bd.IsLegal(*it, mover))
      out[n++] = (int16_t)GoPointUtil::Point2Index(*it);
  out[n++] = (int16_t)GO_MAX_ONBOARD;
  return n;
}

void ref_collect_features(void* b, int8_t* features /* [17][19][19] */) {
  GoBoard& bd = static_cast<RefBoard*>(b)->bd;
  const int plane = GO_MAX_SIZE * GO_MAX_SIZE;
  const SgBlackWhite mover = bd.ToPlay();
  std::vector<GoMove> undone;
  for (int k = 0; k < 8; ++k) {
    for (GoBoard::Iterator it(bd); it; ++it) {
      const int at = (GoPointUtil::Row(*it) - 1) * GO_MAX_SIZE + GoPointUtil::Col(*it) - 1;
      const int c = bd.GetColor(*it);
      features[(2 * k) * plane + at] = c == mover;
      features[(2 * k + 1) * plane + at] = c == SgOppBW(mover);
    }
    const GoMove last = bd.GetLastMove();
    if (last != GO_NULLMOVE) {
      undone.push_back(last);
      bd.Undo();
    }
  }
  memset(features + 16 * plane, 1 - mover, (size_t)plane);
  while (!undone.empty()) {
    bd.Play(undone.back());
    undone.pop_back();
  }
}

}  // extern "C"
