// Drop-in proof, test infrastructure only.  This program is built from the REFERENCE's own, unmodified
// search/UctBoardEvaluator.{h,cpp}, funcapproximator/DlConfig.{h,cc} and lib/StringUtil.cc (compiled where they lie
// under /root/reference) with ONE change to the build: `-I unrealgo_b200/host/overlay` ahead of the reference root, so
// that their `#include "funcapproximator/DlTFNetworkEvaluator.h"` finds the ugeval header instead of the
// TensorFlow one, and libugeval.so on the link line instead of TensorFlow.  main() below does what the reference's
// NetworkEvalThread start-up and loop do with that class (search/UctSearch.cpp:164-205): default-construct (config
// from ./dlcfg-19 through the reference's DlConfig singleton), LoadGraph, UpdateCheckPoint, EvaluateState on an
// EvalBuffer.  Output: oracle/_ref/ref_dropin_eval; tests/test_gpu_reference_dropin.py runs it on the GPU.
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "platform/SgSystem.h"

#include "search/UctBoardEvaluator.h"

struct Buffer {  // the shape of EvalBuffer, search/UctSearch.h:126-130
  char feature_buf[MAX_BATCHES][NUM_MAPS][GO_MAX_SIZE][GO_MAX_SIZE];
  UctValueType policy_out[MAX_BATCHES][GO_MAX_MOVES];
  UctValueType values_out[MAX_BATCHES];
};

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s planes.bin n(<=16)\n", argv[0]); return 2; }
  const int n = atoi(argv[2]);
  static Buffer eb;
  FILE* f = fopen(argv[1], "rb");
  if (n < 1 || n > MAX_BATCHES || !f || fread(eb.feature_buf, NUM_MAPS * GO_MAX_ONBOARD, (size_t)n, f) != (size_t)n) {
    fprintf(stderr, "bad planes file\n");
    return 2;
  }
  fclose(f);
  UctBoardEvaluator evaluator;
  printf("loaded_before %d\n", evaluator.GraphLoaded() ? 1 : 0);
  bool threw = false;
  try { evaluator.LoadGraph("/nonexistent/graph.meta"); } catch (const std::runtime_error&) { threw = true; }
  printf("bad_path_throws %d\n", threw ? 1 : 0);
  evaluator.LoadGraph(DlConfig::GetInstance().get_metagraph());
  printf("loaded_after %d\n", evaluator.GraphLoaded() ? 1 : 0);
  evaluator.UpdateCheckPoint(DlConfig::GetInstance().default_checkpoint());
  evaluator.EvaluateState(eb.feature_buf, eb.policy_out, eb.values_out, n);
  for (int i = 0; i < n; ++i) {
    printf("value %d %.17g\n", i, eb.values_out[i]);
    printf("policy %d", i);
    for (int j = 0; j < GO_MAX_MOVES; ++j) printf(" %.17g", eb.policy_out[i][j]);
    printf("\n");
  }
  return 0;
}
