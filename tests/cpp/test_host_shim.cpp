// Exercises the C++ host shim exactly the way the reference's call sites do
// (search/UctSearch.cpp:178-186, funcapproximator/test/DlTFTestVariableBatchSize.cc:80-90): default-constructed
// UctBoardEvaluator configured from ./dlcfg-19, LoadGraph, UpdateCheckPoint, EvaluateState on an EvalBuffer.
// Prints policies/values as text; the pytest wrapper compares them with the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "../../unrealgo_b200/host/UctBoardEvaluator.h"

struct EvalBuffer {  // search/UctSearch.h:126-130
  char feature_buf[MAX_BATCHES][NUM_MAPS][GO_MAX_SIZE][GO_MAX_SIZE];
  UctValueType policy_out[MAX_BATCHES][GO_MAX_MOVES];
  UctValueType values_out[MAX_BATCHES];
};

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s planes.bin n\n", argv[0]); return 2; }
  const int n = atoi(argv[2]);
  static EvalBuffer eb;
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(eb.feature_buf, NUM_MAPS * 361, (size_t)n, f) != (size_t)n) { fprintf(stderr, "bad planes file\n"); return 2; }
  fclose(f);
  UctBoardEvaluator evaluator;  // reads ./dlcfg-19 through the DlConfig singleton
  printf("loaded_before %d\n", evaluator.GraphLoaded() ? 1 : 0);
  bool threw = false;
  try { evaluator.LoadGraph("/nonexistent/graph.meta"); } catch (const std::runtime_error& e) { threw = true; }
  printf("bad_path_throws %d\n", threw ? 1 : 0);
  for (int i = 0; i < n; ++i) { eb.values_out[i] = 77.0; eb.policy_out[i][0] = 77.0; }
  evaluator.EvaluateState(eb.feature_buf, eb.policy_out, eb.values_out, n);  // no network yet: outputs untouched
  printf("untouched %d\n", eb.values_out[0] == 77.0 && eb.policy_out[0][0] == 77.0 ? 1 : 0);
  evaluator.LoadGraph(DlConfig::GetInstance().get_metagraph());
  printf("loaded_after %d\n", evaluator.GraphLoaded() ? 1 : 0);
  evaluator.UpdateCheckPoint(DlConfig::GetInstance().default_checkpoint());
  evaluator.EvaluateState(eb.feature_buf, eb.policy_out, eb.values_out, n);
  // the same call again with the EvalBuffer page-locked (extension): results must be bit-identical
  static EvalBuffer eb2;
  memcpy(eb2.feature_buf, eb.feature_buf, sizeof(eb.feature_buf));
  const bool reg = evaluator.Evaluator().RegisterHostBuffer(&eb2, sizeof(eb2));
  evaluator.EvaluateState(eb2.feature_buf, eb2.policy_out, eb2.values_out, n);
  bool same = true;
  for (int i = 0; i < n; ++i) {
    same = same && eb2.values_out[i] == eb.values_out[i];
    for (int j = 0; j < GO_MAX_MOVES; ++j) same = same && eb2.policy_out[i][j] == eb.policy_out[i][j];
  }
  printf("registered %d identical %d\n", reg ? 1 : 0, same ? 1 : 0);
  for (int i = 0; i < n; ++i) {
    printf("value %d %.17g\n", i, eb.values_out[i]);
    printf("policy %d", i);
    for (int j = 0; j < GO_MAX_MOVES; ++j) printf(" %.9g", eb.policy_out[i][j]);
    printf("\n");
  }
  return 0;
}
