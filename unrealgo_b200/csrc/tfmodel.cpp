#include "tfmodel.h"

#include <cstdio>
#include <cstring>
#include <functional>
#include <set>

namespace ug {

// ------------------------------------------------------------------------------------------ files / crc
bool read_file(const std::string& path, std::vector<uint8_t>* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (n < 0) { fclose(f); return false; }
  out->resize((size_t)n);
  size_t got = n ? fread(out->data(), 1, (size_t)n, f) : 0;
  fclose(f);
  return got == (size_t)n;
}

uint32_t crc32c(const uint8_t* p, size_t n) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
      table[i] = c;
    }
    init = true;
  }
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

// ------------------------------------------------------------------------------------------ protobuf wire
static bool read_varint(const uint8_t*& p, const uint8_t* end, uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; shift < 64 && p < end; shift += 7) {
    const uint8_t b = *p++;
    r |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) { *v = r; return true; }
  }
  return false;
}

bool PbReader::next(PbField* f) {
  if (p_ >= end_ || !ok_) return false;
  uint64_t key;
  if (!read_varint(p_, end_, &key)) { ok_ = false; return false; }
  f->number = (uint32_t)(key >> 3);
  f->wire = (uint32_t)(key & 7);
  f->data = nullptr;
  f->size = 0;
  f->value = 0;
  switch (f->wire) {
    case 0:
      if (!read_varint(p_, end_, &f->value)) { ok_ = false; return false; }
      return true;
    case 1:
      if (end_ - p_ < 8) { ok_ = false; return false; }
      memcpy(&f->value, p_, 8);
      p_ += 8;
      return true;
    case 2: {
      uint64_t n;
      if (!read_varint(p_, end_, &n) || (uint64_t)(end_ - p_) < n) { ok_ = false; return false; }
      f->data = p_;
      f->size = (size_t)n;
      p_ += n;
      return true;
    }
    case 5: {
      if (end_ - p_ < 4) { ok_ = false; return false; }
      uint32_t v;
      memcpy(&v, p_, 4);
      f->value = v;
      p_ += 4;
      return true;
    }
    default:
      ok_ = false;
      return false;
  }
}

// ------------------------------------------------------------------------------------------ GraphDef
static void parse_list_value(const uint8_t* p, size_t n, TfAttr* a) {
  PbReader r(p, n);
  PbField f;
  while (r.next(&f)) {
    if (f.number == 3) {  // repeated int64 i (packed or not)
      if (f.wire == 0) a->list_i.push_back((int64_t)f.value);
      else if (f.wire == 2) {
        const uint8_t* q = f.data;
        const uint8_t* e = f.data + f.size;
        uint64_t v;
        while (q < e && read_varint(q, e, &v)) a->list_i.push_back((int64_t)v);
      }
    }
  }
}

// TensorProto of a Const node (tensor.proto): dtype = 1, tensor_shape = 2, tensor_content = 4, float_val = 5
static void parse_tensor_proto(const uint8_t* p, size_t n, TfAttr* a) {
  PbReader r(p, n);
  PbField f;
  int64_t elems = 1;
  bool have_shape = false;
  std::vector<float> vals;
  while (r.next(&f)) {
    if (f.number == 1 && f.wire == 0) a->tensor_dtype = (int)f.value;
    else if (f.number == 2 && f.wire == 2) {
      have_shape = true;
      PbReader s(f.data, f.size);
      PbField sf;
      while (s.next(&sf)) {
        if (sf.number != 2 || sf.wire != 2) continue;
        PbReader d(sf.data, sf.size);
        PbField df;
        int64_t size = 0;
        while (d.next(&df)) if (df.number == 1) size = (int64_t)df.value;
        elems *= size;
        a->tensor_shape.push_back(size);
      }
    } else if (f.number == 4 && f.wire == 2) {
      vals.resize(f.size / 4);
      memcpy(vals.data(), f.data, vals.size() * 4);
    } else if (f.number == 5) {
      if (f.wire == 5) { uint32_t u = (uint32_t)f.value; float v; memcpy(&v, &u, 4); vals.push_back(v); }
      else if (f.wire == 2) {
        const size_t at = vals.size();
        vals.resize(at + f.size / 4);
        memcpy(vals.data() + at, f.data, (f.size / 4) * 4);
      }
    }
  }
  a->has_tensor = r.ok();
  if (a->tensor_dtype != 1 || !a->has_tensor) return;  // only float constants are ever needed
  if (!have_shape) elems = (int64_t)vals.size();
  if (elems < 0 || elems > (1 << 24)) { a->has_tensor = false; return; }
  if (vals.size() == 1 && elems > 1) vals.assign((size_t)elems, vals[0]);  // splat
  if ((int64_t)vals.size() != elems) { a->has_tensor = false; return; }
  a->tensor_f = std::move(vals);
}

static void parse_attr_value(const uint8_t* p, size_t n, TfAttr* a) {
  PbReader r(p, n);
  PbField f;
  while (r.next(&f)) {
    switch (f.number) {
      case 1: if (f.wire == 2) parse_list_value(f.data, f.size, a); break;
      case 2: if (f.wire == 2) { a->s.assign((const char*)f.data, f.size); a->has_s = true; } break;
      case 3: a->i = (int64_t)f.value; a->has_i = true; break;
      case 4: { uint32_t u = (uint32_t)f.value; memcpy(&a->f, &u, 4); a->has_f = true; } break;
      case 5: a->b = f.value != 0; a->has_b = true; break;
      case 6: a->type = (int)f.value; a->has_type = true; break;
      case 8: if (f.wire == 2) parse_tensor_proto(f.data, f.size, a); break;
      default: break;  // shape / func: not needed on the inference path
    }
  }
}

static bool parse_node(const uint8_t* p, size_t n, TfNode* node) {
  PbReader r(p, n);
  PbField f;
  while (r.next(&f)) {
    if (f.wire != 2) continue;
    switch (f.number) {
      case 1: node->name.assign((const char*)f.data, f.size); break;
      case 2: node->op.assign((const char*)f.data, f.size); break;
      case 3: node->inputs.emplace_back((const char*)f.data, f.size); break;
      case 5: {  // map<string, AttrValue> entry
        PbReader e(f.data, f.size);
        PbField ef;
        std::string key;
        TfAttr val;
        while (e.next(&ef)) {
          if (ef.number == 1 && ef.wire == 2) key.assign((const char*)ef.data, ef.size);
          else if (ef.number == 2 && ef.wire == 2) parse_attr_value(ef.data, ef.size, &val);
        }
        if (!e.ok()) return false;
        node->attrs[key] = val;
        break;
      }
      default: break;
    }
  }
  return r.ok();
}

static bool parse_graph_def(const uint8_t* p, size_t n, TfGraph* g) {
  PbReader r(p, n);
  PbField f;
  while (r.next(&f)) {
    if (f.number == 1 && f.wire == 2) {
      TfNode node;
      if (!parse_node(f.data, f.size, &node)) return false;
      g->index[node.name] = (int)g->nodes.size();
      g->nodes.push_back(std::move(node));
    }
  }
  return r.ok();
}

bool parse_tf_graph(const std::vector<uint8_t>& bytes, bool is_meta, TfGraph* g, std::string* err) {
  g->nodes.clear();
  g->index.clear();
  if (!is_meta) {
    if (!parse_graph_def(bytes.data(), bytes.size(), g) || g->nodes.empty()) {
      *err = "not a GraphDef protobuf";
      return false;
    }
    g->saver_filename_tensor = "save/Const:0";  // LoadPbGraph defaults, DlTFNetworkEvaluator.cc:214-215
    g->saver_restore_op = "save/restore_all";
    return true;
  }
  PbReader r(bytes.data(), bytes.size());
  PbField f;
  bool have_graph = false;
  while (r.next(&f)) {
    if (f.wire != 2) continue;
    if (f.number == 2) {
      if (!parse_graph_def(f.data, f.size, g)) { *err = "malformed graph_def in MetaGraphDef"; return false; }
      have_graph = true;
    } else if (f.number == 3) {  // SaverDef
      PbReader s(f.data, f.size);
      PbField sf;
      while (s.next(&sf)) {
        if (sf.wire != 2) continue;
        if (sf.number == 1) g->saver_filename_tensor.assign((const char*)sf.data, sf.size);
        if (sf.number == 3) g->saver_restore_op.assign((const char*)sf.data, sf.size);
      }
    }
  }
  if (!r.ok() || !have_graph || g->nodes.empty()) {
    *err = "not a MetaGraphDef protobuf (no graph_def)";
    return false;
  }
  return true;
}

const TfNode* TfGraph::find(const std::string& name) const {
  std::string n = name;
  if (!n.empty() && n[0] == '^') n = n.substr(1);
  const size_t c = n.rfind(':');
  if (c != std::string::npos && c + 1 < n.size() &&
      n.find_first_not_of("0123456789", c + 1) == std::string::npos)
    n = n.substr(0, c);
  auto it = index.find(n);
  return it == index.end() ? nullptr : &nodes[it->second];
}

// ------------------------------------------------------------------------------------------ plan
namespace {

struct Walker {
  const TfGraph& g;
  std::string err;
  explicit Walker(const TfGraph& gr) : g(gr) {}

  bool fail(const std::string& m) { if (err.empty()) err = m; return false; }

  const TfNode* in(const TfNode* n, size_t k) {
    size_t data_idx = 0;
    for (const std::string& s : n->inputs) {
      if (!s.empty() && s[0] == '^') continue;
      if (data_idx == k) {
        const TfNode* p = g.find(s);
        if (!p) fail("node '" + n->name + "' input '" + s + "' not found in graph");
        return p;
      }
      ++data_idx;
    }
    fail("node '" + n->name + "' (" + n->op + ") has no data input " + std::to_string(k));
    return nullptr;
  }

  static bool is_bn(const std::string& op) {
    return op == "FusedBatchNorm" || op == "FusedBatchNormV2" || op == "FusedBatchNormV3";
  }

  // Follow pass-through ops.  tf.cond(training, ...) leaves Switch/Merge pairs: take the Merge input whose
  // producer is an inference-mode (is_training=false) batch-norm, otherwise the first input.
  const TfNode* resolve(const TfNode* n) {
    for (int guard = 0; n && guard < 64; ++guard) {
      if (n->op == "Identity" || n->op == "StopGradient" || n->op == "PlaceholderWithDefault" ||
          n->op == "Switch" || n->op == "Snapshot") {
        n = in(n, 0);
      } else if (n->op == "Merge") {
        const TfNode* pick = nullptr;
        size_t k = 0;
        for (const std::string& s : n->inputs) {
          if (!s.empty() && s[0] == '^') continue;
          const TfNode* c = in(n, k++);
          while (c && (c->op == "Identity" || c->op == "Switch")) c = in(c, 0);
          if (c && is_bn(c->op)) {
            auto it = c->attrs.find("is_training");
            const bool training = it == c->attrs.end() ? true : it->second.b;
            if (!training) pick = c;
          }
        }
        if (!pick) pick = in(n, 0);
        n = pick;
      } else {
        return n;
      }
    }
    return n;
  }
  const TfNode* rin(const TfNode* n, size_t k) {
    const TfNode* p = in(n, k);
    return p ? resolve(p) : nullptr;
  }

  // variable reference -> checkpoint key
  bool var_of(const TfNode* n, std::string* key) {
    const TfNode* v = n;
    for (int guard = 0; v && guard < 16; ++guard) {
      if (v->op == "VariableV2" || v->op == "Variable" || v->op == "VarHandleOp") { *key = v->name; return true; }
      if (v->op == "Identity" || v->op == "ReadVariableOp" || v->op == "Switch" || v->op == "Enter") { v = in(v, 0); continue; }
      if (v->op == "Const") {  // a frozen graph: the variable has been folded into a float constant
        auto it = v->attrs.find("value");
        if (it == v->attrs.end() || !it->second.has_tensor || it->second.tensor_dtype != 1 || it->second.tensor_f.empty())
          return fail("constant '" + v->name + "' is not a float tensor the loader can read");
        *key = std::string(kConstKeyPrefix) + v->name;
        return true;
      }
      break;
    }
    return fail("expected a variable, found '" + (v ? v->name + "' (" + v->op + ")" : std::string("?'")));
  }

  // batch-norm scale / offset: a variable, or (scale=False / center=False) a float Const
  bool var_or_const(const TfNode* n, std::string* key, std::vector<float>* values) {
    const TfNode* v = n;
    for (int guard = 0; v && guard < 16 && (v->op == "Identity" || v->op == "Switch"); ++guard) v = in(v, 0);
    if (v && v->op == "Const") {
      auto it = v->attrs.find("value");
      if (it == v->attrs.end() || !it->second.has_tensor || it->second.tensor_dtype != 1 || it->second.tensor_f.empty())
        return fail("constant '" + v->name + "' is not a float tensor the loader can read");
      key->clear();
      *values = it->second.tensor_f;
      return true;
    }
    values->clear();
    return var_of(n, key);
  }

  bool conv_bn(const TfNode* n, ConvPlan* out, const TfNode** src) {
    if (!n) return false;
    if (is_bn(n->op)) {
      auto tr = n->attrs.find("is_training");
      if (tr != n->attrs.end() && tr->second.b)
        return fail("batch-norm '" + n->name + "' is in training mode; an inference graph is required");
      auto df = n->attrs.find("data_format");
      if (df != n->attrs.end() && df->second.has_s && df->second.s != "NHWC")
        return fail("batch-norm '" + n->name + "' data_format " + df->second.s + " unsupported (NHWC only)");
      auto ep = n->attrs.find("epsilon");
      out->eps = (ep != n->attrs.end() && ep->second.has_f) ? ep->second.f : 1e-4f;
      out->has_bn = true;
      const TfNode* p;
      if (!(p = in(n, 1)) || !var_or_const(p, &out->gamma, &out->gamma_const)) return false;
      if (!(p = in(n, 2)) || !var_or_const(p, &out->beta, &out->beta_const)) return false;
      if (!(p = in(n, 3)) || !var_of(p, &out->mean)) return false;
      if (!(p = in(n, 4)) || !var_of(p, &out->var)) return false;
      n = rin(n, 0);
      if (!n) return false;
    }
    if (n->op == "BiasAdd" || n->op == "Add" || n->op == "AddV2") {
      const TfNode* p = in(n, 1);
      if (!p || !var_of(p, &out->bias)) return false;
      n = rin(n, 0);
      if (!n) return false;
    }
    if (n->op != "Conv2D") return fail("expected Conv2D at '" + n->name + "', found " + n->op);
    auto df = n->attrs.find("data_format");
    if (df != n->attrs.end() && df->second.has_s && df->second.s != "NHWC")
      return fail("Conv2D '" + n->name + "' data_format " + df->second.s + " unsupported (NHWC only)");
    auto pd = n->attrs.find("padding");
    if (pd == n->attrs.end() || pd->second.s != "SAME")
      return fail("Conv2D '" + n->name + "' padding must be SAME");
    for (const char* key : {"strides", "dilations"}) {
      auto st = n->attrs.find(key);
      if (st != n->attrs.end())
        for (int64_t v : st->second.list_i)
          if (v != 1) return fail("Conv2D '" + n->name + "' " + key + " must be 1");
    }
    const TfNode* k = in(n, 1);
    if (!k || !var_of(k, &out->kernel)) return false;
    *src = rin(n, 0);
    return *src != nullptr;
  }

  bool dense(const TfNode* n, DensePlan* out, const TfNode** src) {
    if (!n) return false;
    if (n->op == "BiasAdd" || n->op == "Add" || n->op == "AddV2") {
      const TfNode* p = in(n, 1);
      if (!p || !var_of(p, &out->bias)) return false;
      n = rin(n, 0);
      if (!n) return false;
    }
    if (n->op != "MatMul") return fail("expected MatMul at '" + n->name + "', found " + n->op);
    for (const char* key : {"transpose_a", "transpose_b"}) {
      auto t = n->attrs.find(key);
      if (t != n->attrs.end() && t->second.b) return fail("MatMul '" + n->name + "' with " + key + " unsupported");
    }
    const TfNode* k = in(n, 1);
    if (!k || !var_of(k, &out->kernel)) return false;
    *src = rin(n, 0);
    return *src != nullptr;
  }

  const TfNode* skip_shape_ops(const TfNode* n) {
    while (n && (n->op == "Reshape" || n->op == "Squeeze" || n->op == "ExpandDims")) n = rin(n, 0);
    return n;
  }
};

}  // namespace

bool build_net_plan(const TfGraph& g, const std::string& input, const std::string& policy_out,
                    const std::string& value_out, NetPlan* plan, std::string* err) {
  Walker w(g);
  *plan = NetPlan();
  const TfNode* pn = g.find(policy_out);
  const TfNode* vn = g.find(value_out);
  if (!pn) { *err = "policy output node '" + policy_out + "' not in graph"; return false; }
  if (!vn) { *err = "value output node '" + value_out + "' not in graph"; return false; }

  // ---- policy head
  const TfNode* n = w.skip_shape_ops(w.resolve(pn));
  if (n && n->op == "Softmax") { plan->policy_softmax = true; n = w.rin(n, 0); }
  const TfNode* src = nullptr;
  if (!w.dense(n, &plan->policy_fc, &src)) { *err = "policy head: " + w.err; return false; }
  src = w.skip_shape_ops(src);
  if (!src || src->op != "Relu") { *err = "policy head: expected Relu before the flatten"; return false; }
  const TfNode* tower_out_p = nullptr;
  if (!w.conv_bn(w.rin(src, 0), &plan->policy_conv, &tower_out_p)) { *err = "policy head: " + w.err; return false; }
  plan->policy_conv.ksize = 1;

  // ---- value head
  n = w.skip_shape_ops(w.resolve(vn));
  if (n && n->op == "Tanh") { plan->value_tanh = true; n = w.rin(n, 0); }
  n = w.skip_shape_ops(n);
  if (!w.dense(n, &plan->value_fc2, &src)) { *err = "value head: " + w.err; return false; }
  if (!src || src->op != "Relu") { *err = "value head: expected Relu between the two dense layers"; return false; }
  if (!w.dense(w.rin(src, 0), &plan->value_fc1, &src)) { *err = "value head: " + w.err; return false; }
  src = w.skip_shape_ops(src);
  if (!src || src->op != "Relu") { *err = "value head: expected Relu before the flatten"; return false; }
  const TfNode* tower_out_v = nullptr;
  if (!w.conv_bn(w.rin(src, 0), &plan->value_conv, &tower_out_v)) { *err = "value head: " + w.err; return false; }
  plan->value_conv.ksize = 1;
  if (tower_out_p != tower_out_v) {
    *err = "policy and value heads do not share one tower output ('" + tower_out_p->name + "' vs '" +
           tower_out_v->name + "')";
    return false;
  }

  // ---- tower, walked from its output back to the placeholder
  std::vector<ConvPlan> rev;  // conv2, conv1, conv2, conv1 ... from the top
  const TfNode* cur = tower_out_p;
  for (int guard = 0; guard < 1024; ++guard) {
    if (!cur || cur->op != "Relu") {
      *err = "tower: expected Relu, found '" + (cur ? cur->name + "' (" + cur->op + ")" : std::string("?'"));
      return false;
    }
    const TfNode* x = w.rin(cur, 0);
    if (!x) { *err = "tower: " + w.err; return false; }
    if (x->op == "Add" || x->op == "AddV2") {
      const TfNode* a = w.rin(x, 0);
      const TfNode* b = w.rin(x, 1);
      if (!a || !b) { *err = "tower: " + w.err; return false; }
      const TfNode* branch = (a->op == "Relu") ? b : a;
      const TfNode* skip = (a->op == "Relu") ? a : b;
      if (skip->op != "Relu" || branch->op == "Relu") {
        *err = "tower: residual Add '" + x->name + "' does not join a conv branch and a Relu skip";
        return false;
      }
      ConvPlan c2, c1;
      const TfNode* mid = nullptr;
      const TfNode* blk_in = nullptr;
      if (!w.conv_bn(branch, &c2, &mid)) { *err = "tower: " + w.err; return false; }
      if (!mid || mid->op != "Relu") { *err = "tower: expected Relu between the convs of block at '" + x->name + "'"; return false; }
      if (!w.conv_bn(w.rin(mid, 0), &c1, &blk_in)) { *err = "tower: " + w.err; return false; }
      if (blk_in != skip) {
        *err = "tower: block at '" + x->name + "': skip input '" + skip->name + "' is not the branch input '" +
               (blk_in ? blk_in->name : "?") + "'";
        return false;
      }
      rev.push_back(c2);
      rev.push_back(c1);
      cur = skip;
    } else {
      const TfNode* inp = nullptr;
      if (!w.conv_bn(x, &plan->stem, &inp)) { *err = "stem: " + w.err; return false; }
      while (inp && inp->op == "Cast") inp = w.rin(inp, 0);
      if (!inp || inp->op != "Placeholder") {
        *err = "stem: expected the input Placeholder, found '" + (inp ? inp->name + "' (" + inp->op + ")" : std::string("?'"));
        return false;
      }
      if (inp->name != input) {
        *err = "graph input is '" + inp->name + "' but nn_input is '" + input + "'";
        return false;
      }
      plan->input = inp->name;
      auto dt = inp->attrs.find("dtype");
      plan->input_dtype = (dt != inp->attrs.end()) ? dt->second.type : 0;
      break;
    }
  }
  if (plan->input.empty()) { *err = "tower: did not reach the input placeholder"; return false; }
  plan->tower.assign(rev.rbegin(), rev.rend());
  std::vector<std::string> keys;
  plan_variable_keys(*plan, &keys);
  const size_t plen = strlen(kConstKeyPrefix);
  plan->const_weights = 0;
  for (const std::string& k : keys) plan->const_weights += k.compare(0, plen, kConstKeyPrefix) == 0 ? 1 : 0;
  plan->frozen = !keys.empty() && plan->const_weights == (int)keys.size();
  return true;
}

const char* const kConstKeyPrefix = "@const:";

void plan_variable_keys(const NetPlan& plan, std::vector<std::string>* keys) {
  keys->clear();
  auto add_conv = [&](const ConvPlan& c) {
    keys->push_back(c.kernel);
    if (!c.bias.empty()) keys->push_back(c.bias);
    if (c.has_bn) {
      if (!c.gamma.empty()) keys->push_back(c.gamma);  // empty: values carried in the plan (scale=False / frozen)
      if (!c.beta.empty()) keys->push_back(c.beta);
      keys->push_back(c.mean);
      keys->push_back(c.var);
    }
  };
  add_conv(plan.stem);
  for (const ConvPlan& c : plan.tower) add_conv(c);
  add_conv(plan.policy_conv);
  add_conv(plan.value_conv);
  for (const DensePlan* d : {&plan.policy_fc, &plan.value_fc1, &plan.value_fc2}) {
    keys->push_back(d->kernel);
    if (!d->bias.empty()) keys->push_back(d->bias);
  }
}

bool add_graph_constants(const TfGraph& g, const NetPlan& plan, TfBundle* ck, std::string* err) {
  std::vector<std::string> keys;
  plan_variable_keys(plan, &keys);
  const size_t plen = strlen(kConstKeyPrefix);
  for (const std::string& k : keys) {
    if (k.compare(0, plen, kConstKeyPrefix) != 0) continue;
    const TfNode* n = g.find(k.substr(plen));
    auto it = n ? n->attrs.find("value") : std::map<std::string, TfAttr>::const_iterator();
    if (!n || it == n->attrs.end() || !it->second.has_tensor || it->second.tensor_dtype != 1) {
      *err = "constant '" + k.substr(plen) + "' has no float tensor";
      return false;
    }
    ck->add_memory(k, it->second.tensor_shape, it->second.tensor_f);
  }
  return true;
}

// ------------------------------------------------------------------------------------------ bundle (SSTable)
namespace {

struct BlockHandle { uint64_t offset = 0, size = 0; };

bool parse_handle(const uint8_t*& p, const uint8_t* end, BlockHandle* h) {
  return read_varint(p, end, &h->offset) && read_varint(p, end, &h->size);
}

// iterate a LevelDB block: prefix-compressed entries followed by a restart array
bool block_entries(const std::vector<uint8_t>& file, BlockHandle h,
                   const std::function<bool(const std::string&, const uint8_t*, size_t)>& fn, std::string* err) {
  if (h.offset + h.size + 5 > file.size() || h.size < 4) { *err = "index block out of range"; return false; }
  const uint8_t* base = file.data() + h.offset;
  if (base[h.size] != 0) { *err = "compressed index blocks are not supported"; return false; }
  uint32_t stored;
  memcpy(&stored, base + h.size + 1, 4);
  const uint32_t actual = crc32c_mask(crc32c(base, (size_t)h.size + 1));
  if (stored != actual) { *err = "index block checksum mismatch"; return false; }
  uint32_t num_restarts;
  memcpy(&num_restarts, base + h.size - 4, 4);
  if ((uint64_t)num_restarts * 4 + 4 > h.size) { *err = "bad restart array"; return false; }
  const uint8_t* p = base;
  const uint8_t* end = base + h.size - 4 - (size_t)num_restarts * 4;
  std::string key;
  while (p < end) {
    uint64_t shared, non_shared, vlen;
    if (!read_varint(p, end, &shared) || !read_varint(p, end, &non_shared) || !read_varint(p, end, &vlen) ||
        shared > key.size() || (uint64_t)(end - p) < non_shared + vlen) {
      *err = "malformed index block entry";
      return false;
    }
    key.resize((size_t)shared);
    key.append((const char*)p, (size_t)non_shared);
    p += non_shared;
    if (!fn(key, p, (size_t)vlen)) return false;
    p += vlen;
  }
  return true;
}

bool parse_entry(const uint8_t* p, size_t n, BundleEntry* e) {
  PbReader r(p, n);
  PbField f;
  while (r.next(&f)) {
    switch (f.number) {
      case 1: e->dtype = (int)f.value; break;
      case 2: {
        if (f.wire != 2) break;
        PbReader s(f.data, f.size);
        PbField sf;
        while (s.next(&sf)) {
          if (sf.number == 2 && sf.wire == 2) {
            PbReader d(sf.data, sf.size);
            PbField df;
            int64_t size = 0;
            while (d.next(&df)) if (df.number == 1) size = (int64_t)df.value;
            e->shape.push_back(size);
          }
        }
        break;
      }
      case 3: e->shard = (int)f.value; break;
      case 4: e->offset = (int64_t)f.value; break;
      case 5: e->size = (int64_t)f.value; break;
      case 6: e->crc32c = (uint32_t)f.value; e->has_crc = true; break;
      default: break;
    }
  }
  return r.ok();
}

}  // namespace

bool TfBundle::open(const std::string& prefix, std::string* err) {
  entries_.clear();
  shards_.clear();
  int num_shards = 1;
  std::vector<uint8_t> idx;
  if (!read_file(prefix + ".index", &idx)) { *err = "cannot read " + prefix + ".index"; return false; }
  if (idx.size() < 48) { *err = "index file too small"; return false; }
  uint64_t magic;
  memcpy(&magic, idx.data() + idx.size() - 8, 8);
  if (magic != 0xdb4775248b80fb57ull) { *err = "bad SSTable magic in " + prefix + ".index"; return false; }
  const uint8_t* p = idx.data() + idx.size() - 48;
  const uint8_t* end = idx.data() + idx.size() - 8;
  BlockHandle meta, index;
  if (!parse_handle(p, end, &meta) || !parse_handle(p, end, &index)) { *err = "bad SSTable footer"; return false; }
  bool ok = block_entries(idx, index, [&](const std::string&, const uint8_t* v, size_t vn) {
    const uint8_t* q = v;
    BlockHandle dh;
    if (!parse_handle(q, v + vn, &dh)) { *err = "bad data block handle"; return false; }
    return block_entries(idx, dh, [&](const std::string& key, const uint8_t* ev, size_t en) {
      if (key.empty()) {  // BundleHeaderProto: num_shards = 1, endianness = 2 (0 = little), version = 3
        PbReader r(ev, en);
        PbField f;
        while (r.next(&f)) {
          if (f.number == 1 && f.wire == 0) num_shards = (int)f.value;
          if (f.number == 2 && f.wire == 0 && f.value != 0) { *err = "big-endian checkpoint bundles are not supported"; return false; }
        }
        return true;
      }
      BundleEntry e;
      if (!parse_entry(ev, en, &e)) { *err = "malformed BundleEntryProto for " + key; return false; }
      entries_[key] = e;
      return true;
    }, err);
  }, err);
  if (!ok) return false;
  if (num_shards < 1 || num_shards > 4096) { *err = "implausible shard count in " + prefix + ".index"; return false; }
  // a Saver(sharded=True) spreads the tensors over <prefix>.data-0000k-of-0000N; the usual case is N = 1
  shards_.resize((size_t)num_shards);
  for (int k = 0; k < num_shards; ++k) {
    char suffix[40];
    snprintf(suffix, sizeof suffix, ".data-%05d-of-%05d", k, num_shards);
    if (!read_file(prefix + suffix, &shards_[(size_t)k])) {
      *err = "cannot read " + prefix + suffix;
      shards_.clear();
      return false;
    }
  }
  return true;
}

void TfBundle::add_memory(const std::string& key, const std::vector<int64_t>& shape, const std::vector<float>& values) {
  MemTensor& t = mem_[key];
  t.shape = shape;
  t.values = values;
}

const BundleEntry* TfBundle::find(const std::string& key) const {
  auto it = entries_.find(key);
  return it == entries_.end() ? nullptr : &it->second;
}

bool TfBundle::read_float(const std::string& key, std::vector<float>* out, std::vector<int64_t>* shape,
                          std::string* err) const {
  auto mt = mem_.find(key);
  if (mt != mem_.end()) {  // a Const node of the graph standing in for a variable
    int64_t count = 1;
    for (int64_t d : mt->second.shape) count *= d;
    if (count != (int64_t)mt->second.values.size()) { *err = "constant '" + key + "' has an inconsistent shape"; return false; }
    *out = mt->second.values;
    if (shape) *shape = mt->second.shape;
    return true;
  }
  const BundleEntry* e = find(key);
  if (!e) { *err = "variable '" + key + "' not in checkpoint"; return false; }
  if (e->dtype != 1) { *err = "variable '" + key + "' is not DT_FLOAT"; return false; }
  if (e->shard < 0 || e->shard >= (int)shards_.size()) {
    *err = "variable '" + key + "' lives in shard " + std::to_string(e->shard) + " of " + std::to_string(shards_.size());
    return false;
  }
  const std::vector<uint8_t>& data = shards_[(size_t)e->shard];
  int64_t count = 1;
  for (int64_t d : e->shape) count *= d;
  if (e->size != count * 4 || e->offset < 0 || (uint64_t)(e->offset + e->size) > data.size()) {
    *err = "variable '" + key + "' has inconsistent size/offset (a partitioned variable is stored as slices: not supported)";
    return false;
  }
  const uint8_t* src = data.data() + e->offset;
  if (e->has_crc && crc32c_mask(crc32c(src, (size_t)e->size)) != e->crc32c) {
    *err = "variable '" + key + "' fails its crc32c check";
    return false;
  }
  out->resize((size_t)count);
  memcpy(out->data(), src, (size_t)e->size);
  if (shape) *shape = e->shape;
  return true;
}

}  // namespace ug
