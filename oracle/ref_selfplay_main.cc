// The reference's OWN self-play player on the product library, test infrastructure only.  Built by oracle/Makefile
// (`refselfplay`) from the same unmodified sources as ref_search plus search/UctDeepPlayer.cpp; ZeroMQ/libcurl headers
// and the download/upload control plane are stood in for (ref_shim/zmq.hpp, ref_shim/curl/curl.h,
// ref_controlplane_stubs.cc).
//
// UctDeepPlayer::SelfPlayOneGame (search/UctDeepPlayer.cpp:61-121) only returns after a whole game of 4000-playout
// searches, far too long for a test, and the state it builds is private.  main() therefore opens the class up
// (the two #defines below: a test-only liberty, the class itself is compiled unmodified in its own translation
// unit) and runs that function's loop body for a given number of moves — every call in it is the reference's own:
// DoSearch with FindInitTree (tree reuse per ./dlcfg-19), the node/policy allocators, GoGame::AddMove — then calls the
// reference's WriteTFRecord, whose records leave through unrealgo_b200/host/DlTFRecordWriter.cc.
//   usage: ref_selfplay <moves> <out.tfrecords>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include <unistd.h>

#include "platform/SgSystem.h"

#define private public
#define protected public
#include "UctDeepPlayer.h"
#undef private
#undef protected

#include "GoGame.h"
#include "GoInit.h"
#include "SgInit.h"
#include "funcapproximator/DlConfig.h"

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s moves out.tfrecords\n", argv[0]); return 2; }
  const int moves = atoi(argv[1]);
  SgInit();
  GoInit();
  GoGame game(GO_MAX_SIZE);
  GoRules rules;
  UctDeepPlayer player(game, rules);
  player.m_search.UpdateCheckPoint(DlConfig::GetInstance().default_checkpoint());

  UctNode root(GO_NULLMOVE);  // from here: the body of SelfPlayOneGame, :62-107
  std::vector<UctNode*> selectedChild;
  SgBlackWhite toPlay = SG_BLACK;
  int length = 0;
  const double tau = 1.0;
  player.m_policyAllocator->Clear();
  player.m_nodeAllocator->Clear();
  player.m_nodeSequence.clear();
  player.m_nodeSequence.push_back(&root);
  player.ClearBoard();
  while (length < moves) {
    selectedChild.clear();
    float* policy = player.m_policyAllocator->CreateOneBlock();
    const GoMove move = player.DoSearch(toPlay, DBL_MAX, tau, &selectedChild, policy);
    if (move == UCT_RESIGN || (move == GO_PASS && player.m_nodeSequence.back()->Move() == GO_PASS)) break;
    UctNode* node = player.m_nodeAllocator->CreateOne(move, player.m_nodeSequence.back());
    player.m_nodeSequence.back()->SetFirstChild(node);
    player.m_nodeSequence.back()->SetNumChildren(1);
    node->SetPolicy(policy);
    node->CopyNonPointerData(*selectedChild[0]);
    player.m_nodeSequence.push_back(node);
    player.m_game.AddMove(move, toPlay);
    player.UpdateSubscriber();
    printf("move %d idx %d\n", length, GoPointUtil::Point2Index(move));
    toPlay = SgOppBW(toPlay);
    ++length;
  }
  // A game ends inside the loop with a search whose move (the second pass) is not added, so WriteTFRecord expects
  // the search state to hold every added move (:365-366); cut short after `moves`, one more search — its result
  // dropped the same way — brings the state there.
  if (length == moves) {
    selectedChild.clear();
    player.DoSearch(toPlay, DBL_MAX, tau, &selectedChild, player.m_policyAllocator->CreateOneBlock());
  }
  player.WriteTFRecord(&root, argv[2]);
  printf("records written for %d moves\n", length);
  fflush(stdout);
  _exit(0);  // see ref_search_main.cc: destroying a stopped search never returns
}
