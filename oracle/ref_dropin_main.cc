// Drop-in proof, test infrastructure only.  This program is built from the REFERENCE's own, unmodified
// search/UctBoardEvaluator.{h,cpp}, funcapproximator/DlConfig.{h,cc} and lib/StringUtil.cc (compiled where they lie
// under /root/reference) with ONE change to the build: `-I unrealgo_b200/host/overlay` ahead of the reference root, so
// that their `#include "funcapproximator/DlTFNetworkEvaluator.h"` finds the ugeval header instead of the
// TensorFlow one, and libugeval.so on the link line instead of TensorFlow.  main() below does what the reference's
// NetworkEvalThread start-up and loop do with that class (search/UctSearch.cpp:164-205): default-construct (config
// from ./dlcfg-19 through the reference's DlConfig singleton), LoadGraph, UpdateCheckPoint, EvaluateState on an
// EvalBuffer.  Then the class itself the way the reference's own evaluator program drives it
// (funcapproximator/test/DlTFTestVariableBatchSize.cc:80-90: DlTFNetworkEvaluator<float> with explicit node names,
// UpdateCheckPoint, Evaluate at several batch sizes) plus the members only that class exposes: CreateTensor,
// TransformFeature (both overloads, both DataFormats), the HWC Evaluate overload.
// Output: oracle/_ref/ref_dropin_eval; tests/test_gpu_reference_dropin.py runs it on the GPU.
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

  // ---- tensorflow::DlTFNetworkEvaluator<float>, as in DlTFTestVariableBatchSize.cc:80-90
  std::vector<std::string> outputs;
  DlConfig::GetInstance().get_network_outputs(outputs);
  tensorflow::DlTFNetworkEvaluator<float> direct(DlConfig::GetInstance().get_metagraph(),
                                                 DlConfig::GetInstance().get_network_input(), outputs);
  printf("direct_loaded %d\n", direct.MetaGraphLoaded() ? 1 : 0);
  printf("direct_checkpoint %d\n", direct.UpdateCheckPoint(DlConfig::GetInstance().default_checkpoint()) ? 1 : 0);
  static char hwc[MAX_BATCHES][BD_SIZE][BD_SIZE][NUM_MAPS];
  for (int b = 0; b < n; ++b)
    for (int d = 0; d < NUM_MAPS; ++d)
      for (int i = 0; i < BD_SIZE; ++i)
        for (int j = 0; j < BD_SIZE; ++j) hwc[b][i][j][d] = eb.feature_buf[b][d][i][j];
  // TransformFeature: the CHW planes re-laid out (on the GPU) must be the HWC planes, element for element
  direct.CreateTensor(n);
  int tensor_zeroed = direct.InputTensor(n) != nullptr;
  for (int i = 0; tensor_zeroed && i < n * FEATURE_SIZE; ++i) tensor_zeroed = direct.InputTensor(n)[i] == 0.f;
  printf("create_tensor %d\n", tensor_zeroed);
  direct.TransformFeature(eb.feature_buf, n);
  int same = direct.InputTensor(n) != nullptr;
  for (int i = 0; same && i < n * FEATURE_SIZE; ++i) same = direct.InputTensor(n)[i] == (float)(&hwc[0][0][0][0])[i];
  printf("transform_chw_is_hwc %d\n", same);
  direct.TransformFeature(hwc, n);
  same = 1;
  for (int i = 0; same && i < n * FEATURE_SIZE; ++i) same = direct.InputTensor(n)[i] == (float)(&hwc[0][0][0][0])[i];
  printf("transform_hwc_is_copy %d\n", same);
  {  // DF_CHW keeps the planes' own order (.cc:378-380, :404-406)
    tensorflow::DlTFNetworkEvaluator<bool> chw_fmt("", DF_CHW);
    chw_fmt.TransformFeature(hwc, n);
    same = chw_fmt.InputTensor(n) != nullptr;
    for (int i = 0; same && i < n * FEATURE_SIZE; ++i) same = chw_fmt.InputTensor(n)[i] == (bool)(&eb.feature_buf[0][0][0][0])[i];
    chw_fmt.TransformFeature(eb.feature_buf, n);
    for (int i = 0; same && i < n * FEATURE_SIZE; ++i) same = chw_fmt.InputTensor(n)[i] == (bool)(&eb.feature_buf[0][0][0][0])[i];
    printf("transform_df_chw %d\n", same);
  }
  // Evaluate, both overloads, at the batch sizes of DlTFTestVariableBatchSize.cc:86-89: every position's answer must
  // be the one UctBoardEvaluator gave above, bit for bit, whatever batch it rides in and whichever layout it came in
  static double pol[MAX_BATCHES][GO_MAX_MOVES], val[MAX_BATCHES];
  int identical = 1;
  const int sizes[4] = {1, 4, 8, 16};
  for (int s = 0; s < 4; ++s) {
    const int m = sizes[s] < n ? sizes[s] : n;
    for (int overload = 0; overload < 2; ++overload) {
      for (int b = 0; b < m; ++b) { val[b] = -7; pol[b][0] = -7; }
      if (overload == 0) direct.Evaluate(eb.feature_buf, pol, val, m);
      else direct.Evaluate(hwc, pol, val, m);
      for (int b = 0; b < m; ++b) {
        identical = identical && val[b] == eb.values_out[b];
        for (int j = 0; j < GO_MAX_MOVES; ++j) identical = identical && pol[b][j] == eb.policy_out[b][j];
      }
    }
  }
  printf("direct_identical %d\n", identical);
  return 0;
}
