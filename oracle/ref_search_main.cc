// The reference's OWN search on the product library, test infrastructure only.  Everything below main() is the
// reference's unmodified code compiled where it lies (search/UctSearch.cpp with its NetworkEvalThread and
// ExpandAndBackup, search/UctBoardEvaluator.cpp, gouct/, go/, board/, lib/, platform/ — oracle/Makefile `refsearch`),
// with unrealgo_b200/host/overlay ahead of the reference root on the include path and libugeval.so +
// unrealgo_b200/host/DlTFRecordWriter.cc on the link line where TensorFlow used to be.  main() sets the search up the
// way UctDeepPlayer does (search/UctDeepPlayer.cpp:18-51, :288-320) and runs it from the empty board, one search per
// move, carrying the subtree over when ./dlcfg-19 says reuse_searchtree=true.
//   usage: ref_search <playouts per move> <moves> [threads]        (./dlcfg-19 names the graph and the checkpoint)
// Output: one line per move — move index, games played, seconds — then the root's children (move:visits) of the last
// search and the totals.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <unistd.h>

#include "platform/SgSystem.h"

#include "GoBoard.h"
#include "GoInit.h"
#include "GoUctDefaultMoveFilter.h"
#include "GoUctGlobalSearch.h"
#include "GoUctPlayoutPolicy.h"
#include "SgInit.h"
#include "UctTreeUtil.h"
#include "funcapproximator/DlConfig.h"
#include "platform/SgTimer.h"

typedef GoUctGlobalSearch<GoUctPlayoutPolicy<GoUctBoard>, GoUctPlayoutPolicyFactory<GoUctBoard> > SearchType;

// `ref_search features <games.txt> <out.bin>`: no search, no evaluator — for every game (one line of policy indices,
// 361 = pass) the moves are played on the engine's board and after each one the search thread's state is brought up
// to date (GoUctState::StartSearch -> GoBoardSynchronizer) and asked for its planes with the reference's own
// GoUctGlobalSearchState::CollectFeatures (gouct/GoUctGlobalSearch.h:160-206), the very call ExpandAndBackup makes
// (search/UctSearch.cpp:512).  out.bin receives char[17][19][19] per position, the initial position first.
static int dump_features(const char* games_path, const char* out_path) {
  SgInit();
  GoInit();
  GoBoard bd(GO_MAX_SIZE);
  GoUctPlayoutPolicyParam policyParam;
  GoUctDefaultMoveFilterParam filterParam;
  SearchType search(bd, new GoUctPlayoutPolicyFactory<GoUctBoard>(policyParam), policyParam, filterParam);
  search.CreateThreads();
  typedef GoUctGlobalSearchState<GoUctPlayoutPolicy<GoUctBoard> > StateType;
  StateType& state = dynamic_cast<StateType&>(search.ThreadState(0));
  FILE* in = fopen(games_path, "r");
  FILE* out = fopen(out_path, "wb");
  if (!in || !out) return 2;
  static char line[1 << 16];
  static char planes[NUM_MAPS][GO_MAX_SIZE][GO_MAX_SIZE];
  long positions = 0;
  while (fgets(line, sizeof(line), in)) {
    bd.Init(GO_MAX_SIZE);
    char* p = line;
    for (;;) {
      state.StartSearch();
      state.CollectFeatures(planes, NUM_MAPS);
      fwrite(planes, 1, sizeof(planes), out);
      ++positions;
      char* end;
      const long idx = strtol(p, &end, 10);
      if (end == p) break;
      p = end;
      bd.Play(idx == GO_MAX_ONBOARD ? GO_PASS : GoPointUtil::Pt((int)(idx % GO_MAX_SIZE) + 1, (int)(idx / GO_MAX_SIZE) + 1));
    }
  }
  fclose(out);
  printf("positions %ld\n", positions);
  fflush(stdout);
  _exit(0);
}

int main(int argc, char** argv) {
  if (argc == 4 && std::string(argv[1]) == "features") return dump_features(argv[2], argv[3]);
  if (argc < 3) { fprintf(stderr, "usage: %s playouts moves [threads] | features games.txt out.bin\n", argv[0]); return 2; }
  const double playouts = atof(argv[1]);
  const int moves = atoi(argv[2]);
  SgInit();
  GoInit();
  GoBoard bd(GO_MAX_SIZE);
  GoUctPlayoutPolicyParam policyParam;
  GoUctDefaultMoveFilterParam filterParam;
  SearchType search(bd, new GoUctPlayoutPolicyFactory<GoUctBoard>(policyParam), policyParam, filterParam);
  search.SetPruneMinCount(0);
  if (argc > 3) search.SetNumberThreads((size_t)atoi(argv[3]));
  search.UpdateCheckPoint(DlConfig::GetInstance().default_checkpoint());
  double total_games = 0, total_time = 0;
  for (int m = 0; m < moves; ++m) {
    std::vector<GoMove> sequence;
    std::vector<UctNode*> best;
    std::vector<GoMove> rootFilter;
    search.SetToPlay(bd.ToPlay());
    SgTimer timer;
    // tree reuse as UctDeepPlayer::DoSearch / FindInitTree do it (search/UctDeepPlayer.cpp:223-231, :290-297): the
    // subtree below the moves played since the last search becomes the initial tree of this one
    UctSearchTree* initTree = 0;
    if (DlConfig::GetInstance().reuse_search_tree()) {
      initTree = &search.GetTempTree();
      std::vector<GoPoint> played;
      if (static_cast<GoUctSearch&>(search).BoardHistory().SequenceToCurrent(bd, played))
        UctTreeUtil::ExtractSubtree(search.Tree(), *initTree, played, true, 1e9, search.PruneMinCount());
    }
    search.StartDeepUCTSearchThread(playouts, 1e9, sequence, &best, 0, 1.0, rootFilter, initTree, 0, true);
    const double t = timer.GetTime();
    const GoMove mv = sequence.empty() ? GO_PASS : sequence[0];
    if (mv == GO_PASS && bd.GetLastMove() == GO_PASS) {  // the game is over, the second pass is not played
      printf("game over after %d moves\n", m);         // (SelfPlayOneGame, search/UctDeepPlayer.cpp:82-88)
      break;
    }
    printf("move %d idx %d games %.0f seconds %.4f\n", m, GoPointUtil::Point2Index(mv), (double)search.GamesPlayed(), t);
    total_games += search.GamesPlayed();
    total_time += t;
    if (m + 1 == moves) {
      printf("root");
      for (UctChildNodeIterator it(search.Tree(), search.Tree().Root()); it; ++it)
        if ((*it).VisitCount() > 0) printf(" %d:%.0f", GoPointUtil::Point2Index((*it).Move()), (double)(*it).VisitCount());
      printf("\n");
    }
    bd.Play(mv);
  }
  printf("total games %.0f seconds %.4f playouts_per_s %.1f\n", total_games, total_time, total_games / total_time);
  // Leave without running destructors, as the engine itself does: ~NetworkEvalThread (search/UctSearch.cpp:153-158) sets
  // to_quit and joins, but a paused evaluation thread sits in wait(lock, !paused) (:170-172) and never looks at to_quit,
  // so destroying a search that has been stopped blocks forever.
  fflush(stdout);
  _exit(0);
}
