// Drop-in replacement for funcapproximator/DlTFNetworkEvaluator.h (reference lines :33-83) on top of the C ABI
// in include/ugeval.h.  Same class name, namespace, template parameter, method names, argument types, return
// values and throw behaviour, so search/UctBoardEvaluator.{h,cpp} and search/UctSearch.cpp compile unchanged
// against it; no TensorFlow header is needed any more.
//
//   reference member (.cc line)                         C export
//   DlTFNetworkEvaluator(graphPath, df)       :27-53    ug_eval_create + ug_eval_load_graph
//   SetNetworkInput / SetNetworkOutput        :56-64    ug_eval_set_io
//   LoadGraph                                 :77-83    ug_eval_load_graph   (<0 -> throw std::runtime_error)
//   UpdateCheckPoint                          :230-236  ug_eval_load_checkpoint
//   MetaGraphLoaded                           :239-241  ug_eval_ready
//   CreateTensor                              :271-284  (device workspace sized once per architecture; host tensor here)
//   TransformFeature (CHW / HWC overloads)    :366-412  ug_transform_planes (the GPU kernel Evaluate runs) / host cast
//   printFeature                              :353-363  same text
//   Evaluate (CHW / HWC overloads)            :287-352  ug_eval_run_planes / ug_eval_run_planes_hwc
//   ~DlTFNetworkEvaluator                     :67-69    ug_eval_destroy
// Not provided: WritePbText / GetMetaGraph (:244-252) hand out TensorFlow's own protobuf objects; their only callers
// are the graph-printing programs under funcapproximator/test/, which are not on the evaluation path.
#ifndef UGEVAL_HOST_DLTFNETWORKEVALUATOR_H
#define UGEVAL_HOST_DLTFNETWORKEVALUATOR_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ugeval.h"
// Built inside the reference tree (host/overlay/funcapproximator/DlTFNetworkEvaluator.h), the reference's own
// BoardStaticConfig.h, DataFormat.h and DlConfig.h are already in; standalone, the host-side stand-ins are used.
#ifndef UGEVAL_WITH_REFERENCE_TREE
#include "DlConfig.h"
static const int GO_MAX_SIZE = UG_BOARD_SIZE;
static const int GO_MAX_MOVES = UG_NUM_MOVES;
#else
static_assert(GO_MAX_SIZE == UG_BOARD_SIZE && GO_MAX_MOVES == UG_NUM_MOVES, "ugeval is built for 19x19");
#endif
const int BD_SIZE = UG_BOARD_SIZE;
const int NUM_MAPS = UG_NUM_PLANES;
const int FEATURE_SIZE = BD_SIZE * BD_SIZE * NUM_MAPS;
// The reference caps batches at 16 (DlTFNetworkEvaluator.h:15) because EvalBuffer is statically sized; the
// evaluator below accepts up to UGEVAL_MAX_BATCH per call.
const int MAX_BATCHES = 16;
#ifndef UGEVAL_MAX_BATCH
#define UGEVAL_MAX_BATCH 512
#endif

#ifndef UGEVAL_WITH_REFERENCE_TREE
enum DataFormat { DF_HWC = UG_DF_HWC, DF_CHW = UG_DF_CHW };  // funcapproximator/DataFormat.h:22-25
#else
static_assert(DF_HWC == UG_DF_HWC && DF_CHW == UG_DF_CHW, "DataFormat values differ from the C ABI's");
#endif

namespace tensorflow {  // kept so that `tensorflow::DlTFNetworkEvaluator<bool>` in UctBoardEvaluator.h still names it

template <typename T>
class DlTFNetworkEvaluator {
 public:
  explicit DlTFNetworkEvaluator(const std::string& graphPath, DataFormat _df = DF_HWC)
      : m_dataFormat(_df), m_h(ug_eval_create(device_from_env(), precision_from_env(), UGEVAL_MAX_BATCH)) {
    if (!m_h) throw std::runtime_error(std::string("Could not create evaluator: ") + ug_last_error(nullptr));
    LoadGraph(graphPath);
  }
  DlTFNetworkEvaluator(const std::string& graphPath, const std::string& feature_input,
                       const std::vector<std::string>& outputs, DataFormat _df = DF_HWC)
      : m_dataFormat(_df), m_h(ug_eval_create(device_from_env(), precision_from_env(), UGEVAL_MAX_BATCH)) {
    if (!m_h) throw std::runtime_error(std::string("Could not create evaluator: ") + ug_last_error(nullptr));
    SetNetworkInput(feature_input);
    std::vector<std::string> o(outputs);
    SetNetworkOutput(o);
    LoadGraph(graphPath);
  }
  ~DlTFNetworkEvaluator() { ug_eval_destroy(m_h); }
  DlTFNetworkEvaluator(const DlTFNetworkEvaluator&) = delete;
  DlTFNetworkEvaluator& operator=(const DlTFNetworkEvaluator&) = delete;

  void SetNetworkInput(std::string input) { ug_eval_set_io(m_h, input.c_str(), nullptr, 0); }
  void SetNetworkOutput(std::vector<std::string>& output) {
    std::vector<const char*> p;
    for (const std::string& s : output) p.push_back(s.c_str());
    // an empty list clears the names, like m_outputs.clear() in the reference
    static const char* none = nullptr;
    ug_eval_set_io(m_h, nullptr, p.empty() ? &none : p.data(), (int)p.size());
  }

  bool LoadGraph(const std::string& graphPath) {
    const int rc = ug_eval_load_graph(m_h, graphPath.c_str());
    if (rc < 0) throw std::runtime_error(ug_last_error(m_h));
    return rc == 0;
  }
  bool UpdateCheckPoint(const std::string& checkpointPath) {
    return ug_eval_load_checkpoint(m_h, checkpointPath.c_str()) == 0;
  }
  bool MetaGraphLoaded() { return ug_eval_ready(m_h) != 0; }

  // .cc:271-284.  The reference allocates one TensorFlow tensor per batch size, always shaped {batches, 19, 19, 17}
  // whatever m_dataFormat says, and keeps it for good.  The device workspace here is sized once per architecture
  // (ug_eval_load_checkpoint); what CreateTensor keeps is the host-side tensor TransformFeature() fills.
  void CreateTensor(int batches) {
    if (batches < 1 || batches > UGEVAL_MAX_BATCH) return;
    std::unique_ptr<T[]>& t = m_input_tensors[batches];
    if (!t) t.reset(new T[(size_t)batches * FEATURE_SIZE]());  // value-initialised, like a fresh tensor
  }
  // .cc:366-389: data[GetOffset(...)] = (T)feature[b][depth][i][j].  DF_HWC is the NHWC re-layout — done by the same
  // GPU kernel Evaluate() runs (planes_to_nhwc_kernel through ug_transform_planes); DF_CHW keeps the planes' own order.
  void TransformFeature(char feature[][NUM_MAPS][BD_SIZE][BD_SIZE], int numBatches = MAX_BATCHES) {
    CreateTensor(numBatches);
    if (numBatches < 1 || numBatches > UGEVAL_MAX_BATCH) return;
    T* data = m_input_tensors[numBatches].get();
    const size_t total = (size_t)numBatches * FEATURE_SIZE;
    const char* flat = &feature[0][0][0][0];
    if (m_dataFormat == DF_HWC) {
      std::vector<int8_t> hwc(total);
      if (ug_transform_planes(m_h, reinterpret_cast<const int8_t*>(flat), numBatches, hwc.data()) != 0) {
        fprintf(stderr, "TransformFeature failed: %s\n", ug_last_error(m_h));
        return;
      }
      for (size_t i = 0; i < total; ++i) data[i] = (T)hwc[i];
    } else {
      for (size_t i = 0; i < total; ++i) data[i] = (T)flat[i];
    }
  }
  // .cc:391-412: the same for planes that already are [b][i][j][depth]
  void TransformFeature(char feature[][BD_SIZE][BD_SIZE][NUM_MAPS], int numBatches = MAX_BATCHES) {
    CreateTensor(numBatches);
    if (numBatches < 1 || numBatches > UGEVAL_MAX_BATCH) return;
    T* data = m_input_tensors[numBatches].get();
    const char* flat = &feature[0][0][0][0];
    if (m_dataFormat == DF_HWC) {
      for (size_t i = 0; i < (size_t)numBatches * FEATURE_SIZE; ++i) data[i] = (T)flat[i];
    } else {
      for (int b = 0; b < numBatches; ++b)
        for (int p = 0; p < BD_SIZE * BD_SIZE; ++p)
          for (int d = 0; d < NUM_MAPS; ++d)
            data[((size_t)b * NUM_MAPS + d) * BD_SIZE * BD_SIZE + p] = (T)flat[((size_t)b * BD_SIZE * BD_SIZE + p) * NUM_MAPS + d];
    }
  }
  // extension: the tensor the last CreateTensor/TransformFeature(.., batches) left (m_input_tensors[batches - 1] there)
  const T* InputTensor(int batches) const {
    auto it = m_input_tensors.find(batches);
    return it == m_input_tensors.end() ? nullptr : it->second.get();
  }
  void printFeature(T data[]) {  // .cc:353-363
    for (int i = 0; i < BD_SIZE * BD_SIZE; ++i) {
      for (int j = 0; j < NUM_MAPS; ++j) printf("%d ", (int)data[i * NUM_MAPS + j]);
      printf("\n");
    }
    printf("\n");
  }

  // value_transform comes from the DlConfig singleton, as at DlTFNetworkEvaluator.cc:300.  The graph always reads its
  // placeholder as [n][19][19][17] (CreateTensor's dims): with DF_HWC that is the NHWC re-layout of the planes; with
  // DF_CHW the reference hands TensorFlow the CHW-ordered buffer under those same dims, and so does this — no caller in
  // the reference passes DF_CHW (search/UctBoardEvaluator.cpp:6), it is kept for the interface's sake.
  void Evaluate(char feature[][NUM_MAPS][BD_SIZE][BD_SIZE], double actions_[][GO_MAX_MOVES], double value_[],
                int numBatches = MAX_BATCHES) {
    ug_eval_set_value_transform(m_h, (int)DlConfig::GetInstance().get_value_transform());
    const int8_t* flat = reinterpret_cast<const int8_t*>(feature);
    const int rc = m_dataFormat == DF_HWC ? ug_eval_run_planes(m_h, flat, numBatches, &actions_[0][0], value_)
                                          : ug_eval_run_planes_hwc(m_h, flat, numBatches, &actions_[0][0], value_);
    if (rc != 0) fprintf(stderr, "Running model failed: %s\n", ug_last_error(m_h));  // outputs untouched (.cc:315-317)
  }
  void Evaluate(char feature[][BD_SIZE][BD_SIZE][NUM_MAPS], double actions_[][GO_MAX_MOVES], double value_[],
                int batch_size = MAX_BATCHES) {
    ug_eval_set_value_transform(m_h, (int)DlConfig::GetInstance().get_value_transform());
    const int8_t* flat = reinterpret_cast<const int8_t*>(feature);
    std::vector<int8_t> chw;
    if (m_dataFormat != DF_HWC && batch_size >= 1 && batch_size <= UGEVAL_MAX_BATCH) {  // see TransformFeature
      chw.resize((size_t)batch_size * FEATURE_SIZE);
      for (int b = 0; b < batch_size; ++b)
        for (int p = 0; p < BD_SIZE * BD_SIZE; ++p)
          for (int d = 0; d < NUM_MAPS; ++d)
            chw[((size_t)b * NUM_MAPS + d) * BD_SIZE * BD_SIZE + p] = flat[((size_t)b * BD_SIZE * BD_SIZE + p) * NUM_MAPS + d];
      flat = chw.data();
    }
    const int rc = ug_eval_run_planes_hwc(m_h, flat, batch_size, &actions_[0][0], value_);
    if (rc != 0) fprintf(stderr, "Running model failed: %s\n", ug_last_error(m_h));
  }

  ug_eval* handle() { return m_h; }  // for the leaf queue (ug_queue_create)
  // extension (not in the reference): page-lock a caller array that Evaluate() is always called with — EvalBuffer's
  // feature_buf / policy_out / values_out — so it is copied from/to directly (ug_eval_register_host)
  bool RegisterHostBuffer(void* p, size_t bytes) { return ug_eval_register_host(m_h, p, bytes) == 0; }

 protected:
  // default arithmetic: fp16 operands / fp32 accumulation (meets the stated tolerance); UGEVAL_PRECISION=bf16|fp32
  static int precision_from_env() {
    const char* p = getenv("UGEVAL_PRECISION");
    if (p && !strcmp(p, "bf16")) return UG_PREC_BF16;
    if (p && !strcmp(p, "fp32")) return UG_PREC_FP32;
    return UG_PREC_FP16;
  }
  static int device_from_env() {
    const char* d = getenv("UGEVAL_DEVICE");
    return d ? atoi(d) : 0;
  }
  DataFormat m_dataFormat;
  ug_eval* m_h;
  std::map<int, std::unique_ptr<T[]>> m_input_tensors;  // by batch size (the reference: Tensor m_input_tensors[MAX_BATCHES])
};

}  // namespace tensorflow
#endif
