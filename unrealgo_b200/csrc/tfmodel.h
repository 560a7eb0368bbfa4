// TensorFlow-free readers for the two artefacts the reference hands to TensorFlow:
//   * MetaGraphDef / GraphDef (ReadBinaryProto at funcapproximator/DlTFNetworkEvaluator.cc:89,130)
//   * TF V2 checkpoint bundle restored by the Saver op (DlTFNetworkEvaluator.cc:170-200)
// The graph is walked BACKWARDS from the two nn_outputs nodes to the nn_input placeholder; anything the walk
// does not reach (gradients/, save/, queues ...) is ignored, anything it reaches but does not recognise is an
// error naming the node.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace ug {

// ---- protobuf wire format (just enough)
struct PbField {
  uint32_t number = 0;
  uint32_t wire = 0;         // 0 varint, 1 fixed64, 2 bytes, 5 fixed32
  uint64_t value = 0;        // varint / fixed
  const uint8_t* data = nullptr;  // bytes
  size_t size = 0;
};
class PbReader {
 public:
  PbReader(const uint8_t* p, size_t n) : p_(p), end_(p + n) {}
  bool next(PbField* f);     // false at end or on malformed input (check ok())
  bool ok() const { return ok_; }
 private:
  const uint8_t* p_;
  const uint8_t* end_;
  bool ok_ = true;
};

// ---- graph
struct TfAttr {
  std::string s;
  int64_t i = 0;
  float f = 0.f;
  bool b = false;
  int type = 0;
  std::vector<int64_t> list_i;
  bool has_s = false, has_i = false, has_f = false, has_b = false, has_type = false;
  // `tensor` (Const nodes): DT_FLOAT values only, expanded to the tensor's element count (a constant filled with one
  // value is stored as a single float_val)
  bool has_tensor = false;
  int tensor_dtype = 0;
  std::vector<float> tensor_f;
  std::vector<int64_t> tensor_shape;   // dims as stored (empty: a scalar, or no shape given)
};
struct TfNode {
  std::string name, op;
  std::vector<std::string> inputs;
  std::map<std::string, TfAttr> attrs;
};
struct TfGraph {
  std::vector<TfNode> nodes;
  std::map<std::string, int> index;
  std::string saver_filename_tensor, saver_restore_op;
  const TfNode* find(const std::string& name) const;
};
// is_meta: MetaGraphDef (graph_def = field 2) vs bare GraphDef
bool parse_tf_graph(const std::vector<uint8_t>& bytes, bool is_meta, TfGraph* g, std::string* err);

// ---- network plan extracted from the graph (variable names, not values)
struct ConvPlan {
  std::string kernel;            // variable holding HWIO weights
  std::string bias;              // optional BiasAdd variable
  bool has_bn = false;
  std::string gamma, beta, mean, var;  // FusedBatchNorm inputs: checkpoint keys.  scale=False / center=False make gamma /
  std::vector<float> gamma_const, beta_const;  // beta a Const node instead: key empty, values here ([cout])
  float eps = 1e-3f;
  int ksize = 3;
};
struct DensePlan {
  std::string kernel, bias;
};
struct NetPlan {
  std::string input;             // placeholder name
  int input_dtype = 0;           // TF DataType enum of the placeholder (1 float, 10 bool)
  ConvPlan stem;
  std::vector<ConvPlan> tower;   // 2 per residual block, in execution order
  ConvPlan policy_conv, value_conv;
  DensePlan policy_fc, value_fc1, value_fc2;
  bool policy_softmax = false, value_tanh = false;
  // Weights that are Const nodes of the graph instead of variables (a graph run through freeze_graph): their "checkpoint
  // key" is kConstKeyPrefix + node name.  frozen: every weight is one, so there is nothing to restore from a checkpoint
  // (the reference's Session evaluates such a .pb straight after LoadPbGraph, DlTFNetworkEvaluator.cc:85-111).
  int const_weights = 0;
  bool frozen = false;
};
extern const char* const kConstKeyPrefix;   // "@const:"
// every checkpoint key the plan needs, in a fixed order (stem, tower, head convs, FCs)
void plan_variable_keys(const NetPlan& plan, std::vector<std::string>* keys);
bool build_net_plan(const TfGraph& g, const std::string& input, const std::string& policy_out,
                    const std::string& value_out, NetPlan* plan, std::string* err);

// ---- checkpoint bundle
struct BundleEntry {
  int dtype = 0;
  std::vector<int64_t> shape;
  int shard = 0;
  int64_t offset = 0, size = 0;
  uint32_t crc32c = 0;
  bool has_crc = false;
};
class TfBundle {
 public:
  bool open(const std::string& prefix, std::string* err);
  const BundleEntry* find(const std::string& key) const;
  // reads a DT_FLOAT tensor, verifying its masked crc32c when present
  bool read_float(const std::string& key, std::vector<float>* out, std::vector<int64_t>* shape,
                  std::string* err) const;
  const std::map<std::string, BundleEntry>& entries() const { return entries_; }
  int num_shards() const { return (int)shards_.size(); }
  // a tensor held in memory under `key` (weights that are Const nodes of the graph); read_float serves it like a variable
  void add_memory(const std::string& key, const std::vector<int64_t>& shape, const std::vector<float>& values);
  size_t memory_entries() const { return mem_.size(); }
 private:
  struct MemTensor { std::vector<int64_t> shape; std::vector<float> values; };
  std::map<std::string, MemTensor> mem_;
  std::map<std::string, BundleEntry> entries_;
  std::vector<std::vector<uint8_t>> shards_;   // <prefix>.data-0000k-of-0000N, N = BundleHeaderProto.num_shards
};

// registers every kConstKeyPrefix key of the plan with the bundle, from the graph's Const nodes
bool add_graph_constants(const TfGraph& g, const NetPlan& plan, TfBundle* ck, std::string* err);

uint32_t crc32c(const uint8_t* p, size_t n);
inline uint32_t crc32c_mask(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }
bool read_file(const std::string& path, std::vector<uint8_t>* out);

}  // namespace ug
