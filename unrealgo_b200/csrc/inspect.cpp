// Host-only inspection of the model artefacts (no CUDA call): parses the graph, builds the network plan and,
// when a checkpoint prefix is given, checks every variable the plan needs (presence, dtype, shape, crc32c).
// Lets tooling and the CPU test-suite exercise the loader of LoadMetaGraph / UpdateCheckPointForMetaGraph
// (funcapproximator/DlTFNetworkEvaluator.cc:115-200) without a GPU.
#include <cstring>
#include <sstream>

#include "../../include/ugeval.h"
#include "tfmodel.h"

using namespace ug;

namespace {
void json_conv(std::ostringstream& o, const ConvPlan& c) {
  o << "{\"kernel\":\"" << c.kernel << "\",\"bias\":\"" << c.bias << "\",\"bn\":" << (c.has_bn ? "true" : "false")
    << ",\"gamma\":\"" << c.gamma << "\",\"beta\":\"" << c.beta << "\",\"mean\":\"" << c.mean << "\",\"var\":\""
    << c.var << "\",\"gamma_const\":" << c.gamma_const.size() << ",\"beta_const\":" << c.beta_const.size()
    << ",\"eps\":" << c.eps << "}";
}
int emit(const std::string& s, char* out, int cap) {
  if (!out || cap <= 0) return -1;
  const int n = (int)s.size() < cap - 1 ? (int)s.size() : cap - 1;
  memcpy(out, s.data(), (size_t)n);
  out[n] = 0;
  return n;
}
}  // namespace

extern "C" int ug_inspect_model(const char* graph_path, const char* ckpt_prefix, const char* input,
                                const char* policy_out, const char* value_out, char* out, int cap) {
  const std::string gp = graph_path ? graph_path : "";
  const bool meta = gp.size() >= 5 && gp.compare(gp.size() - 5, 5, ".meta") == 0;
  const bool pb = gp.size() >= 3 && gp.compare(gp.size() - 3, 3, ".pb") == 0;
  if (!meta && !pb) { emit("graph path must end in .meta or .pb", out, cap); return 1; }
  std::vector<uint8_t> bytes;
  std::string err;
  if (!read_file(gp, &bytes)) { emit("Error reading graph definition from " + gp, out, cap); return -1; }
  TfGraph g;
  if (!parse_tf_graph(bytes, meta, &g, &err)) { emit(err, out, cap); return -1; }
  NetPlan plan;
  if (!build_net_plan(g, input ? input : "feature_input",
                      policy_out ? policy_out : "resnet/tower_0/policy_head/policy_predict",
                      value_out ? value_out : "resnet/tower_0/value_head/reward_predict", &plan, &err)) {
    emit("unsupported graph: " + err, out, cap);
    return -1;
  }
  std::ostringstream o;
  o << "{\"nodes\":" << g.nodes.size() << ",\"input\":\"" << plan.input << "\",\"input_dtype\":" << plan.input_dtype
    << ",\"blocks\":" << plan.tower.size() / 2 << ",\"policy_softmax\":" << (plan.policy_softmax ? "true" : "false")
    << ",\"value_tanh\":" << (plan.value_tanh ? "true" : "false") << ",\"saver_filename_tensor\":\""
    << g.saver_filename_tensor << "\",\"saver_restore_op\":\"" << g.saver_restore_op << "\",\"stem\":";
  json_conv(o, plan.stem);
  o << ",\"tower\":[";
  for (size_t i = 0; i < plan.tower.size(); ++i) { if (i) o << ","; json_conv(o, plan.tower[i]); }
  o << "],\"policy_conv\":";
  json_conv(o, plan.policy_conv);
  o << ",\"value_conv\":";
  json_conv(o, plan.value_conv);
  o << ",\"policy_fc\":{\"kernel\":\"" << plan.policy_fc.kernel << "\",\"bias\":\"" << plan.policy_fc.bias << "\"}"
    << ",\"value_fc1\":{\"kernel\":\"" << plan.value_fc1.kernel << "\",\"bias\":\"" << plan.value_fc1.bias << "\"}"
    << ",\"value_fc2\":{\"kernel\":\"" << plan.value_fc2.kernel << "\",\"bias\":\"" << plan.value_fc2.bias << "\"}";
  o << ",\"const_weights\":" << plan.const_weights << ",\"frozen\":" << (plan.frozen ? "true" : "false");
  const bool have_prefix = ckpt_prefix && ckpt_prefix[0];
  if (have_prefix || plan.frozen) {
    // a frozen graph carries every weight as a Const node: it is checked like a checkpoint, without one
    TfBundle ck;
    if (have_prefix && !ck.open(ckpt_prefix, &err)) {
      if (!plan.frozen) { emit(err, out, cap); return 2; }
      err.clear();   // nothing of a frozen graph lives in the checkpoint files
    }
    if (!add_graph_constants(g, plan, &ck, &err)) { emit(err, out, cap); return 2; }
    std::vector<std::string> need;
    plan_variable_keys(plan, &need);
    size_t params = 0;
    std::vector<float> tmp;
    std::vector<int64_t> shape, stem_shape;
    for (const std::string& k : need) {
      if (!ck.read_float(k, &tmp, &shape, &err)) { emit(err, out, cap); return 2; }
      params += tmp.size();
      if (k == plan.stem.kernel) stem_shape = shape;
    }
    o << ",\"checkpoint\":{\"entries\":" << ck.entries().size() + ck.memory_entries() << ",\"variables_used\":" << need.size()
      << ",\"parameters\":" << params << ",\"filters\":" << (stem_shape.size() == 4 ? stem_shape[3] : 0)
      << "}";
  }
  o << "}";
  emit(o.str(), out, cap);
  return 0;
}

// One DT_FLOAT tensor of a checkpoint bundle through the product's reader (SSTable index walk, shard files, masked
// crc32c): the CPU test-suite compares it value for value with what the writers stored.  Returns the element count
// (and up to 8 dims), -1 with the reason in err when the bundle or the variable cannot be read, -2 when cap is too small.
extern "C" long long ug_bundle_read_float(const char* ckpt_prefix, const char* key, float* out, long long cap,
                                          long long* dims, int* ndim, char* err_out, int err_cap) {
  std::string err;
  TfBundle ck;
  std::vector<float> v;
  std::vector<int64_t> shape;
  if (!ckpt_prefix || !key || !ck.open(ckpt_prefix, &err) || !ck.read_float(key, &v, &shape, &err)) {
    emit(err.empty() ? "bad arguments" : err, err_out, err_cap);
    return -1;
  }
  if (ndim) *ndim = (int)shape.size();
  if (dims) for (size_t i = 0; i < shape.size() && i < 8; ++i) dims[i] = (long long)shape[i];
  if ((long long)v.size() > cap || (!out && !v.empty())) return -2;
  if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(float));
  return (long long)v.size();
}
