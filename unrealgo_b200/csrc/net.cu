// Evaluator: load (graph plan + checkpoint -> folded, packed, device-resident weights), workspace, and the
// launch sequence encoder -> stem -> residual tower -> heads.  Mirrors DlTFNetworkEvaluator<T>
// (funcapproximator/DlTFNetworkEvaluator.cc) behind the C ABI in include/ugeval.h.
#include "net.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>

#include <cuda_fp16.h>

#include "conv3x3_tc.h"
#include "conv3x3_sb.h"
#include "conv3x3_v2.h"
#include "tmap.h"

namespace ug {

bool DevBuf::alloc(size_t n) {
  release();
  if (n == 0) n = 16;
  if (cudaMalloc(&p, n) != cudaSuccess) { p = nullptr; return false; }
  bytes = n;
  return true;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}

HeadWeights Weights::head() const {
  HeadWeights h{};
  h.conv_w = head_conv_w.as<float>();
  h.conv_b = head_conv_b.as<float>();
  h.pc = pc; h.vc = vc; h.C = C;
  h.pfc_w = pfc_w.as<float>(); h.pfc_b = pfc_b.as<float>();
  h.v1_w = v1_w.as<float>(); h.v1_b = v1_b.as<float>();
  h.v2_w = v2_w.as<float>(); h.v2_b = v2_b.as<float>();
  h.H = H;
  h.policy_softmax = policy_softmax ? 1 : 0;
  h.value_tanh = value_tanh ? 1 : 0;
  return h;
}

namespace {

thread_local std::string g_create_error;

bool upload(DevBuf* b, const void* src, size_t bytes) {
  if (!b->alloc(bytes)) return false;
  return cudaMemcpy(b->p, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
}

uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

uint16_t f32_to_f16_rn(float f) {
  // clamp like the device path (numfmt.cuh): weights beyond the fp16 range saturate instead of becoming inf
  if (f > 65504.f) f = 65504.f;
  if (f < -65504.f) f = -65504.f;
  return __half_as_ushort(__float2half_rn(f));
}

struct HostConv {
  std::vector<float> w;  // HWIO, folded
  std::vector<float> b;  // [cout]
  int k = 3, cin = 0, cout = 0;
};

// read conv (+bias) (+batch-norm) variables and fold:  w' = w * g/sqrt(v+eps),  b' = beta + (bias - mean) * g/sqrt(v+eps)
bool load_conv(const TfBundle& ck, const ConvPlan& p, int ksize, HostConv* out, std::string* err) {
  std::vector<int64_t> shape;
  if (!ck.read_float(p.kernel, &out->w, &shape, err)) return false;
  if (shape.size() != 4 || shape[0] != ksize || shape[1] != ksize) {
    *err = "conv kernel '" + p.kernel + "' is not " + std::to_string(ksize) + "x" + std::to_string(ksize) + " HWIO";
    return false;
  }
  out->k = ksize;
  out->cin = (int)shape[2];
  out->cout = (int)shape[3];
  const int cout = out->cout;
  std::vector<float> bias(cout, 0.f);
  if (!p.bias.empty()) {
    std::vector<int64_t> bs;
    if (!ck.read_float(p.bias, &bias, &bs, err)) return false;
    if ((int)bias.size() != cout) { *err = "bias '" + p.bias + "' has wrong size"; return false; }
  }
  out->b.assign(cout, 0.f);
  if (p.has_bn) {
    std::vector<float> g = p.gamma_const, be = p.beta_const, mu, var;  // Const scale / offset (scale=False / center=False)
    if ((!p.gamma.empty() && !ck.read_float(p.gamma, &g, nullptr, err)) ||
        (!p.beta.empty() && !ck.read_float(p.beta, &be, nullptr, err)) ||
        !ck.read_float(p.mean, &mu, nullptr, err) || !ck.read_float(p.var, &var, nullptr, err))
      return false;
    if ((int)g.size() != cout || (int)be.size() != cout || (int)mu.size() != cout || (int)var.size() != cout) {
      *err = "batch-norm variables of '" + p.kernel + "' have wrong size";
      return false;
    }
    std::vector<float> scale(cout);
    for (int c = 0; c < cout; ++c) {
      scale[c] = g[c] / std::sqrt(var[c] + p.eps);
      out->b[c] = be[c] + (bias[c] - mu[c]) * scale[c];
    }
    const size_t rows = out->w.size() / cout;
    for (size_t r = 0; r < rows; ++r)
      for (int c = 0; c < cout; ++c) out->w[r * cout + c] *= scale[c];
  } else {
    out->b = bias;
  }
  return true;
}

float max_abs(const std::vector<float>& v) {
  float m = 0.f;
  for (float x : v) { const float a = std::fabs(x); if (a > m || a != a) m = a; }
  return m;
}

// in_shift / out_shift (fp16 auto-ranging): the layer's input is stored times 2^-in_shift and its output times
// 2^-out_shift, so the packed weights carry 2^(in_shift - out_shift) and the tensor-core bias 2^-out_shift — exact
// scalings; the fp32 copies (validation tower, calibration) stay as loaded.
bool make_conv_layer(const HostConv& hc, int cin_pad, bool f16, int in_shift, int out_shift, ConvLayer* L,
                     std::string* err) {
  L->cin = hc.cin;
  L->cin_pad = cin_pad;
  L->cout = hc.cout;
  const int ktot = 9 * cin_pad;
  const float wscale = std::ldexp(1.f, in_shift - out_shift);
  std::vector<uint16_t> packed((size_t)hc.cout * ktot, 0);
  for (int tap = 0; tap < 9; ++tap)
    for (int ci = 0; ci < hc.cin; ++ci)
      for (int co = 0; co < hc.cout; ++co) {
        const float w = hc.w[((size_t)tap * hc.cin + ci) * hc.cout + co] * wscale;
        packed[(size_t)co * ktot + tap * cin_pad + ci] = f16 ? f32_to_f16_rn(w) : f32_to_bf16_rn(w);
      }
  std::vector<float> btc(hc.b);
  for (float& b : btc) b = std::ldexp(b, -out_shift);
  if (!upload(&L->w_bf16, packed.data(), packed.size() * 2) || !upload(&L->bias_tc, btc.data(), btc.size() * 4) ||
      !upload(&L->w_f32, hc.w.data(), hc.w.size() * 4) || !upload(&L->bias, hc.b.data(), hc.b.size() * 4)) {
    *err = "out of device memory uploading weights";
    return false;
  }
  const int ck = (cin_pad % 64 == 0) ? 64 : 32;
  for (int ctas = 1; ctas <= 2; ++ctas)
    if (!make_tmap_weights(&L->tm_w[ctas - 1], L->w_bf16.p, hc.cout, ktot, ck, hc.cout / ctas, err)) return false;
  if (!make_tmap_weights(&L->tm_w_sb, L->w_bf16.p, hc.cout, ktot, ck, 16, err)) return false;
  return true;
}

bool load_dense(const TfBundle& ck, const DensePlan& p, int in, std::vector<float>* w, std::vector<float>* b,
                int* out_dim, std::string* err) {
  std::vector<int64_t> shape;
  if (!ck.read_float(p.kernel, w, &shape, err)) return false;
  if (shape.size() != 2 || shape[0] != in) {
    *err = "dense kernel '" + p.kernel + "' expected [" + std::to_string(in) + ", *]";
    return false;
  }
  *out_dim = (int)shape[1];
  b->assign(*out_dim, 0.f);
  if (!p.bias.empty()) {
    if (!ck.read_float(p.bias, b, nullptr, err)) return false;
    if ((int)b->size() != *out_dim) { *err = "dense bias '" + p.bias + "' has wrong size"; return false; }
  }
  return true;
}

// shifts (optional): {stream_shift, block_shift[0], block_shift[1], ...} from the range calibration
Weights* build_weights(const NetPlan& plan, const TfBundle& ck, bool f16, const std::vector<int>* shifts,
                       std::string* err) {
  std::unique_ptr<Weights> W(new Weights());
  const int n_blocks = (int)plan.tower.size() / 2;
  W->block_shift.assign((size_t)n_blocks, 0);
  if (shifts && (int)shifts->size() == 1 + n_blocks) {
    W->stream_shift = (*shifts)[0];
    for (int b = 0; b < n_blocks; ++b) W->block_shift[(size_t)b] = (*shifts)[1 + (size_t)b];
  }
  HostConv stem;
  if (!load_conv(ck, plan.stem, 3, &stem, err)) return nullptr;
  if (stem.cin != UG_NUM_PLANES) { *err = "stem conv expects " + std::to_string(stem.cin) + " input planes, the engine provides 17"; return nullptr; }
  const int C = stem.cout;
  if (C % 64 != 0 || C > 256) {
    *err = "tower width " + std::to_string(C) + " unsupported (multiple of 64, <= 256)";
    return nullptr;
  }
  W->C = C;
  W->f16 = f16;
  W->blocks = (int)plan.tower.size() / 2;
  W->convs.resize(1 + plan.tower.size());
  if (!make_conv_layer(stem, kCinPad, f16, 0, W->stream_shift, &W->convs[0], err)) return nullptr;
  W->max_weight = std::ldexp(max_abs(stem.w), -W->stream_shift);  // of the weights as packed
  W->max_weight_layer = 0;
  for (size_t i = 0; i < plan.tower.size(); ++i) {
    HostConv hc;
    if (!load_conv(ck, plan.tower[i], 3, &hc, err)) return nullptr;
    if (hc.cin != C || hc.cout != C) { *err = "tower conv '" + plan.tower[i].kernel + "' is not C x C"; return nullptr; }
    // conv1 of block b: stream -> T_b; conv2: T_b -> stream
    const int bs = W->block_shift[i / 2];
    const int in_shift = (i % 2 == 0) ? W->stream_shift : bs, out_shift = (i % 2 == 0) ? bs : W->stream_shift;
    if (!make_conv_layer(hc, C, f16, in_shift, out_shift, &W->convs[1 + i], err)) return nullptr;
    const float mw = std::ldexp(max_abs(hc.w), in_shift - out_shift);
    if (mw > W->max_weight || mw != mw) { W->max_weight = mw; W->max_weight_layer = 1 + (int)i; }
  }
  HostConv pcv, vcv;
  if (!load_conv(ck, plan.policy_conv, 1, &pcv, err) || !load_conv(ck, plan.value_conv, 1, &vcv, err)) return nullptr;
  if (pcv.cin != C || vcv.cin != C) { *err = "head 1x1 convs do not take the tower width"; return nullptr; }
  W->pc = pcv.cout;
  W->vc = vcv.cout;
  if (W->pc + W->vc > 8) { *err = "more than 8 head conv channels unsupported"; return nullptr; }
  std::vector<float> hw((size_t)(W->pc + W->vc) * C), hb(W->pc + W->vc);
  for (int h = 0; h < W->pc; ++h) {
    for (int c = 0; c < C; ++c) hw[(size_t)h * C + c] = pcv.w[(size_t)c * W->pc + h];
    hb[h] = pcv.b[h];
  }
  for (int h = 0; h < W->vc; ++h) {
    for (int c = 0; c < C; ++c) hw[(size_t)(W->pc + h) * C + c] = vcv.w[(size_t)c * W->vc + h];
    hb[W->pc + h] = vcv.b[h];
  }
  std::vector<float> w, b;
  int od = 0;
  if (!load_dense(ck, plan.policy_fc, UG_NUM_POINTS * W->pc, &w, &b, &od, err)) return nullptr;
  if (od != UG_NUM_MOVES) { *err = "policy dense layer has " + std::to_string(od) + " outputs, expected 362"; return nullptr; }
  std::vector<float> hw_tc(hw);  // the heads read the stream: stored times 2^-stream_shift
  for (float& x : hw_tc) x = std::ldexp(x, W->stream_shift);
  bool ok = upload(&W->head_conv_w, hw.data(), hw.size() * 4) && upload(&W->head_conv_b, hb.data(), hb.size() * 4) &&
            upload(&W->head_conv_w_tc, hw_tc.data(), hw_tc.size() * 4) &&
            upload(&W->pfc_w, w.data(), w.size() * 4) && upload(&W->pfc_b, b.data(), b.size() * 4);
  if (!load_dense(ck, plan.value_fc1, UG_NUM_POINTS * W->vc, &w, &b, &od, err)) return nullptr;
  W->H = od;
  ok = ok && upload(&W->v1_w, w.data(), w.size() * 4) && upload(&W->v1_b, b.data(), b.size() * 4);
  if (!load_dense(ck, plan.value_fc2, W->H, &w, &b, &od, err)) return nullptr;
  if (od != 1) { *err = "value output layer has " + std::to_string(od) + " outputs, expected 1"; return nullptr; }
  ok = ok && upload(&W->v2_w, w.data(), w.size() * 4) && upload(&W->v2_b, b.data(), b.size() * 4);
  if (!ok) { *err = "out of device memory uploading head weights"; return nullptr; }
  W->policy_softmax = plan.policy_softmax;
  W->value_tanh = plan.value_tanh;
  return W.release();
}

// Range calibration.  fp16 operands saturate at 65504 and the tower clamps instead of producing inf, so a checkpoint
// whose activations (or folded weights w * gamma / sqrt(var + eps): a small variance makes them large) come near that
// limit would evaluate *wrongly but plausibly*.  A few built-in positions (empty board to a full one, both colours to
// move) run through the fp32 validation tower on the loaded weights; the largest |output| of every layer — exactly the
// quantity the 16-bit epilogues convert — is recorded.  ~50 ms for 20x256, once per checkpoint load.
constexpr int kCalibPositions = 4;

void calibration_positions(ug_packed_pos* pos) {
  memset(pos, 0, sizeof(ug_packed_pos) * kCalibPositions);
  uint64_t rng = 0x9E3779B97F4A7C15ull;
  auto next = [&rng] { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
  for (int p = 0; p < kCalibPositions; ++p) {
    const int stones = p * 110;  // 0, 110, 220, 330 of 361 points
    int order[UG_NUM_POINTS];
    for (int i = 0; i < UG_NUM_POINTS; ++i) order[i] = i;
    for (int i = UG_NUM_POINTS - 1; i > 0; --i) { const int j = (int)(next() % (uint64_t)(i + 1)); std::swap(order[i], order[j]); }
    for (int k = 0; k < UG_HISTORY; ++k) {        // history step k: the last k stones not yet played
      const int keep = stones - k > 0 ? stones - k : 0;
      for (int i = 0; i < keep; ++i) {
        uint32_t* m = (i & 1) ? pos[p].white[k] : pos[p].black[k];
        m[order[i] >> 5] |= 1u << (order[i] & 31);
      }
    }
    pos[p].to_play = (uint32_t)(stones & 1);
  }
}

bool calibrate_range(ug_eval* h, Weights* W, std::string* err) {
  const int n = kCalibPositions, C = W->C, layers = 1 + 2 * W->blocks;
  const size_t act = (size_t)n * UG_NUM_POINTS * C;
  DevBuf d_pos, d_in16, f_in, f_a, f_b, f_t, d_max;
  if (!d_pos.alloc(n * sizeof(ug_packed_pos)) || !d_in16.alloc((size_t)n * UG_NUM_POINTS * kCinPad * 2) ||
      !f_in.alloc((size_t)n * UG_NUM_POINTS * UG_NUM_PLANES * 4) || !f_a.alloc(act * 4) || !f_b.alloc(act * 4) ||
      !f_t.alloc(act * 4) || !d_max.alloc((size_t)layers * 4)) {
    *err = "out of device memory calibrating the activation range";
    return false;
  }
  ug_packed_pos pos[kCalibPositions];
  calibration_positions(pos);
  cudaStream_t s = h->stream;
  cudaMemcpyAsync(d_pos.p, pos, sizeof(pos), cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(d_max.p, 0, d_max.bytes, s);
  launch_encode_packed_nhwc(d_pos.as<ug_packed_pos>(), d_in16.as<__nv_bfloat16>(), n, 0, s);
  launch_nhwc_to_f32(d_in16.as<__nv_bfloat16>(), f_in.as<float>(), n, 0, s);
  float* S[2] = {f_a.as<float>(), f_b.as<float>()};
  float* mx = d_max.as<float>();
  const ConvLayer& st = W->convs[0];
  launch_conv3x3_f32(f_in.as<float>(), st.w_f32.as<float>(), st.bias.as<float>(), nullptr, S[0], n, UG_NUM_PLANES, C, 1, s);
  launch_max_abs_f32(S[0], act, mx + 0, s);
  for (int b = 0; b < W->blocks; ++b) {
    const int x = b & 1, y = x ^ 1;
    const ConvLayer& c1 = W->convs[1 + 2 * b];
    const ConvLayer& c2 = W->convs[2 + 2 * b];
    launch_conv3x3_f32(S[x], c1.w_f32.as<float>(), c1.bias.as<float>(), nullptr, f_t.as<float>(), n, C, C, 1, s);
    launch_max_abs_f32(f_t.as<float>(), act, mx + 1 + 2 * b, s);
    launch_conv3x3_f32(f_t.as<float>(), c2.w_f32.as<float>(), c2.bias.as<float>(), S[x], S[y], n, C, C, 1, s);
    launch_max_abs_f32(S[y], act, mx + 2 + 2 * b, s);
  }
  std::vector<float> host((size_t)layers);
  cudaMemcpyAsync(host.data(), d_max.p, (size_t)layers * 4, cudaMemcpyDeviceToHost, s);
  const cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { *err = std::string("range calibration failed: ") + cudaGetErrorString(e); return false; }
  W->max_activation = 0.f;
  W->layer_max = host;
  for (int l = 0; l < layers; ++l)
    if (host[(size_t)l] > W->max_activation || host[(size_t)l] != host[(size_t)l]) {
      W->max_activation = host[(size_t)l];
      W->max_activation_layer = l;
    }
  return true;
}

bool ends_with(const std::string& s, const char* suf) {
  const size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

}  // namespace
}  // namespace ug

using namespace ug;

// ------------------------------------------------------------------------------------------ ug_eval members
bool ug_eval::is_registered(const void* p, size_t bytes) const {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  for (const auto& r : host_regs)
    if (a >= r.first && a + bytes <= r.first + r.second) return true;
  return false;
}

ug_eval::~ug_eval() {
  if (stream) cudaStreamSynchronize(stream);
  for (const auto& r : host_regs) cudaHostUnregister(reinterpret_cast<void*>(r.first));
  drop_graphs();
  delete weights;
  if (h_pinned) cudaFreeHost(h_pinned);
  if (stream) cudaStreamDestroy(stream);
}

void ug_eval::drop_graphs() {
  for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
  graphs.clear();
}

bool ug_eval::ensure_workspace(int C) {
  if (ws_C == C) return true;
  drop_graphs();
  // The buffers are resized in place.  Should an allocation fail midway (a hot swap to a wider tower on a full GPU),
  // the old width's buffers and tensor maps are partly gone: ws_C = 0 marks the workspace invalid and the caller drops
  // the old weights, so later runs fail with "no network loaded" instead of writing past smaller buffers.
  ws_C = 0;
  const size_t nb = (size_t)max_batch;
  const size_t act_elems = nb * UG_NUM_POINTS * C;
  bool ok = d_pos.alloc(nb * sizeof(ug_packed_pos)) && d_legal.alloc(nb * UG_MASK_WORDS * 4) &&
            d_planes.alloc(nb * UG_NUM_PLANES * UG_NUM_POINTS) && d_bytes.alloc(nb * UG_NUM_PLANES * UG_NUM_POINTS) &&
            d_in.alloc(nb * UG_NUM_POINTS * kCinPad * 2) && d_feat.alloc(nb * UG_NUM_POINTS * 8 * 4) &&
            d_policy.alloc(nb * UG_NUM_MOVES * 4) && d_value.alloc(nb * 4) &&
            d_policy64.alloc(nb * UG_NUM_MOVES * 8) && d_value64.alloc(nb * 8) &&
            d_flags.alloc(2 * ((nb * UG_NUM_POINTS + 255) / 256) * sizeof(uint32_t)) && d_sat.alloc(16);
  for (int i = 0; i < 2 && ok; ++i) ok = d_s_hi[i].alloc(act_elems * 2) && d_s_lo[i].alloc(act_elems * 2);
  ok = ok && d_t.alloc(act_elems * 2);
  if (ok && precision == UG_PREC_FP32) {
    ok = f_in.alloc(nb * UG_NUM_POINTS * UG_NUM_PLANES * 4) && f_t.alloc(act_elems * 4);
    for (int i = 0; i < 2 && ok; ++i) ok = f_s[i].alloc(act_elems * 4);
  }
  if (!ok) { last_error = "out of device memory allocating the workspace"; return false; }
  cudaMemset(d_sat.p, 0, d_sat.bytes);
  // zero once so that rows past n (read by the tail tile) are finite
  cudaMemset(d_in.p, 0, d_in.bytes);
  for (int i = 0; i < 2; ++i) { cudaMemset(d_s_hi[i].p, 0, d_s_hi[i].bytes); cudaMemset(d_s_lo[i].p, 0, d_s_lo[i].bytes); }
  cudaMemset(d_t.p, 0, d_t.bytes);
  std::string err;
  if (!make_tmap_act_im2col(&tm_in, d_in.p, max_batch, kCinPad, 32, &err)) { last_error = err; return false; }
  for (int i = 0; i < 2; ++i)
    if (!make_tmap_act_im2col(&tm_s[i], d_s_hi[i].p, max_batch, C, 64, &err)) { last_error = err; return false; }
  if (!make_tmap_act_im2col(&tm_t, d_t.p, max_batch, C, 64, &err)) { last_error = err; return false; }
  {
    const long rows = (long)max_batch * UG_NUM_POINTS;
    bool ok2 = make_tmap_rows(&v2_a_in, d_in.p, rows, kCinPad, 32, 168, &err) &&
               make_tmap_rows(&v2_a_t, d_t.p, rows, C, 64, 168, &err) &&
               make_tmap_rows(&v2_out_t, d_t.p, rows, C, 32, 32, &err);
    for (int i = 0; i < 2 && ok2; ++i)
      ok2 = make_tmap_rows(&v2_a_s[i], d_s_hi[i].p, rows, C, 64, 168, &err) &&
            make_tmap_rows(&v2_res_hi[i], d_s_hi[i].p, rows, C, 32, 128, &err) &&
            make_tmap_rows(&v2_res_lo[i], d_s_lo[i].p, rows, C, 32, 128, &err) &&
            make_tmap_rows(&v2_out_hi[i], d_s_hi[i].p, rows, C, 32, 32, &err) &&
            make_tmap_rows(&v2_out_lo[i], d_s_lo[i].p, rows, C, 32, 32, &err);
    if (!ok2) { last_error = err; return false; }
    if (!conv3x3_v2_prepare(device) || !conv3x3_sb_prepare(device)) {
      last_error = "cannot upload the conv lane-mask table";
      return false;
    }
  }
  if (!h_pinned) {
    h_pinned_bytes = nb * (sizeof(ug_packed_pos) + UG_MASK_WORDS * 4 + UG_NUM_PLANES * UG_NUM_POINTS +
                           (UG_NUM_MOVES + 1) * 4) + 256;
    if (cudaMallocHost(&h_pinned, h_pinned_bytes) != cudaSuccess) {
      h_pinned = nullptr;
      last_error = "cannot allocate pinned staging memory";
      return false;
    }
    void* alias = nullptr;
    const char* zc = getenv("UGEVAL_ZERO_COPY_OUT");
    if (!(zc && zc[0] == '0') && cudaHostGetDevicePointer(&alias, h_pinned, 0) == cudaSuccess)
      d_pinned_alias = static_cast<uint8_t*>(alias);
    else
      cudaGetLastError();
  }
  ws_C = C;
  return true;
}

int ug_eval::enqueue_encoder(Src kind, const void* d_src, int n, cudaStream_t s) {
  __nv_bfloat16* in = d_in.as<__nv_bfloat16>();
  const int f16 = precision == UG_PREC_FP16;
  if (kind == SRC_PACKED) launch_encode_packed_nhwc(static_cast<const ug_packed_pos*>(d_src), in, n, f16, s);
  else launch_planes_to_nhwc(static_cast<const int8_t*>(d_src), kind == SRC_HWC ? 1 : 0, in, n, f16, s);
  return 0;
}

// The small-batch kernel wins while its grid at the widest tile (128 output channels per CTA) still fits one wave of
// the SMs: 26 positions on a 148-SM part with 256 filters (measured on the 20x256 net: 0.62 ms against 0.82 ms at n = 24,
// 1.10 ms against 0.84 ms at n = 32, where it needs two waves).
bool ug_eval::use_small_batch(int n) const {
  if (!weights || n > sb_max_batch) return false;
  const int m_tiles = (n * UG_NUM_POINTS + 127) / 128;
  return m_tiles * ((weights->C + 127) / 128) <= sms;
}

int ug_eval::enqueue_tower(int n, cudaStream_t s, const void** final_act, const void** final_lo) {
  const Weights& W = *weights;
  const int C = W.C;
  // block b: x = S[b&1] (stream), t = T, y = S[(b+1)&1]
  if (precision == UG_PREC_FP32) {
    launch_nhwc_to_f32(d_in.as<__nv_bfloat16>(), f_in.as<float>(), n, 0, s);
    const ConvLayer& st = W.convs[0];
    launch_conv3x3_f32(f_in.as<float>(), st.w_f32.as<float>(), st.bias.as<float>(), nullptr, f_s[0].as<float>(),
                       n, UG_NUM_PLANES, C, 1, s);
    *final_act = f_s[0].p;
    *final_lo = nullptr;
    for (int b = 0; b < W.blocks && debug_stop != 0; ++b) {
      const int x = b & 1, y = x ^ 1;
      const ConvLayer& c1 = W.convs[1 + 2 * b];
      const ConvLayer& c2 = W.convs[2 + 2 * b];
      launch_conv3x3_f32(f_s[x].as<float>(), c1.w_f32.as<float>(), c1.bias.as<float>(), nullptr, f_t.as<float>(),
                         n, C, C, 1, s);
      *final_act = f_t.p;
      if (debug_stop == 2 * b + 1) break;
      launch_conv3x3_f32(f_t.as<float>(), c2.w_f32.as<float>(), c2.bias.as<float>(), f_s[x].as<float>(),
                         f_s[y].as<float>(), n, C, C, 1, s);
      *final_act = f_s[y].p;
      if (debug_stop == 2 * b + 2) break;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  if (conv_impl == 2 && C % 64 == 0)
    return use_small_batch(n) ? enqueue_tower_sb(n, s, final_act, final_lo) : enqueue_tower_v2(n, s, final_act, final_lo);
  const int ctas = conv_ctas;
  ConvArgs a{};
  a.M = n * UG_NUM_POINTS;
  a.relu = 1;
  a.f16 = precision == UG_PREC_FP16;
  a.cout = C;
  // stem: 32 stored input channels (17 planes + zeros), SWIZZLE_64B tiles; its output opens the residual stream
  a.cin = kCinPad;
  a.bias = W.convs[0].bias_tc.as<float>();
  a.out = d_s_hi[0].as<__nv_bfloat16>();
  const bool split = resid_split;
  a.out_lo = split ? d_s_lo[0].as<__nv_bfloat16>() : nullptr;
  cudaError_t e = conv3x3_tc_launch(tm_in, W.convs[0].tm_w[ctas - 1], a, 32, ctas, sms, s);
  if (e != cudaSuccess) return fail(-1, std::string("stem conv launch: ") + cudaGetErrorString(e));
  a.cin = C;
  *final_act = d_s_hi[0].p;
  *final_lo = split ? d_s_lo[0].p : nullptr;
  for (int b = 0; b < W.blocks && debug_stop != 0; ++b) {
    const int x = b & 1, y = x ^ 1;
    const ConvLayer& c1 = W.convs[1 + 2 * b];
    const ConvLayer& c2 = W.convs[2 + 2 * b];
    a.bias = c1.bias_tc.as<float>();
    a.residual = nullptr;
    a.residual_lo = nullptr;
    a.out = d_t.as<__nv_bfloat16>();
    a.out_lo = nullptr;
    e = conv3x3_tc_launch(tm_s[x], c1.tm_w[ctas - 1], a, 64, ctas, sms, s);
    if (e != cudaSuccess) return fail(-1, std::string("tower conv launch: ") + cudaGetErrorString(e));
    *final_act = d_t.p;
    *final_lo = nullptr;
    if (debug_stop == 2 * b + 1) break;
    a.bias = c2.bias_tc.as<float>();
    a.residual = d_s_hi[x].as<__nv_bfloat16>();
    a.residual_lo = split ? d_s_lo[x].as<__nv_bfloat16>() : nullptr;
    a.out = d_s_hi[y].as<__nv_bfloat16>();
    a.out_lo = split ? d_s_lo[y].as<__nv_bfloat16>() : nullptr;
    e = conv3x3_tc_launch(tm_t, c2.tm_w[ctas - 1], a, 64, ctas, sms, s);
    if (e != cudaSuccess) return fail(-1, std::string("tower conv launch: ") + cudaGetErrorString(e));
    *final_act = d_s_hi[y].p;
    *final_lo = split ? d_s_lo[y].p : nullptr;
    if (debug_stop == 2 * b + 2) break;
  }
  return 0;
}

int ug_eval::enqueue_tower_v2(int n, cudaStream_t s, const void** final_act, const void** final_lo) {
  const Weights& W = *weights;
  const int C = W.C;
  const bool split = resid_split;
  ConvV2Args a{};
  a.M = n * UG_NUM_POINTS;
  a.relu = 1;
  a.f16 = precision == UG_PREC_FP16;
  a.cout = C;
  a.pdl = use_pdl ? 1 : 0;
  a.sat_count = d_sat.as<unsigned long long>();
  // Tile-granular dependencies need the overlap that programmatic launch provides; counters restart every pass.
  const bool deps = tile_deps && use_pdl;
  const size_t tiles_alloc = ((size_t)max_batch * UG_NUM_POINTS + 255) / 256;
  uint32_t* F[2] = {d_flags.as<uint32_t>(), d_flags.as<uint32_t>() + tiles_alloc};
  if (deps) cudaMemsetAsync(d_flags.p, 0, d_flags.bytes, s);
  int layer = 0;  // 0 = stem, then the tower convolutions in order
  auto set_deps = [&](bool has_res) {
    if (!deps) return;
    a.tile_rot = (layer * 9) % 74;  // 361 tiles on 74 pairs leave 9 pairs one tile short: shift that group each layer
    a.flags_out = F[layer & 1];
    a.flags_in = layer > 0 ? F[(layer - 1) & 1] : nullptr;
    a.target_in = layer > 0 ? 16u * (uint32_t)((layer - 1) / 2 + 1) : 0u;  // 16 epilogue warps per CTA pair
    a.flags_res = has_res ? F[layer & 1] : nullptr;   // written two layers back, same parity
    a.target_res = has_res ? 16u * (uint32_t)(layer / 2) : 0u;
  };
  // stem: writes S[0] (hi [+ lo])
  a.cin = kCinPad;
  a.bias = W.convs[0].bias_tc.as<float>();
  a.has_out_lo = split;
  ConvV2Maps m{&v2_a_in, &W.convs[0].tm_w[1], nullptr, nullptr, &v2_out_hi[0], split ? &v2_out_lo[0] : nullptr};
  set_deps(false);
  cudaError_t e = conv3x3_v2_launch(m, a, 32, device, sms, s);
  if (e != cudaSuccess) return fail(-1, std::string("stem conv launch: ") + cudaGetErrorString(e));
  a.cin = C;
  *final_act = d_s_hi[0].p;
  *final_lo = split ? d_s_lo[0].p : nullptr;
  for (int b = 0; b < W.blocks && debug_stop != 0; ++b) {
    const int x = b & 1, y = x ^ 1;
    const ConvLayer& c1 = W.convs[1 + 2 * b];
    const ConvLayer& c2 = W.convs[2 + 2 * b];
    a.bias = c1.bias_tc.as<float>();
    a.has_res = a.has_res_lo = a.has_out_lo = 0;
    m = ConvV2Maps{&v2_a_s[x], &c1.tm_w[1], nullptr, nullptr, &v2_out_t, nullptr};
    layer = 2 * b + 1;
    set_deps(false);
    e = conv3x3_v2_launch(m, a, 64, device, sms, s);
    if (e != cudaSuccess) return fail(-1, std::string("tower conv launch: ") + cudaGetErrorString(e));
    *final_act = d_t.p;
    *final_lo = nullptr;
    if (debug_stop == 2 * b + 1) break;
    a.bias = c2.bias_tc.as<float>();
    a.has_res = 1;
    a.has_res_lo = a.has_out_lo = split;
    // last layer: the head 1x1 convolutions ride in its epilogue and the activation itself is never stored
    const bool fuse = fuse_heads && debug_stop < 0 && b == W.blocks - 1 &&
                      conv3x3_v2_can_fuse_heads(64, 1, split, split, W.pc + W.vc);
    if (fuse) {
      a.head_w = W.head_conv_w_tc.as<float>();
      a.head_b = W.head_conv_b.as<float>();
      a.head_out = d_feat.as<float>();
      a.head_channels = W.pc + W.vc;
    }
    m = ConvV2Maps{&v2_a_t, &c2.tm_w[1], &v2_res_hi[x], split ? &v2_res_lo[x] : nullptr, &v2_out_hi[y],
                   split ? &v2_out_lo[y] : nullptr};
    layer = 2 * b + 2;
    set_deps(true);
    e = conv3x3_v2_launch(m, a, 64, device, sms, s);
    if (e != cudaSuccess) return fail(-1, std::string("tower conv launch: ") + cudaGetErrorString(e));
    a.head_out = nullptr;
    if (fuse) {  // d_feat already holds relu(head conv): enqueue_heads() skips its first stage
      *final_act = nullptr;
      *final_lo = nullptr;
      break;
    }
    *final_act = d_s_hi[y].p;
    *final_lo = split ? d_s_lo[y].p : nullptr;
    if (debug_stop == 2 * b + 2) break;
  }
  return 0;
}

// The tower on the small-batch kernel: same buffers, same layer sequence and arithmetic as enqueue_tower_v2.
int ug_eval::enqueue_tower_sb(int n, cudaStream_t s, const void** final_act, const void** final_lo) {
  const Weights& W = *weights;
  const int C = W.C;
  const bool split = resid_split;
  ConvSbArgs a{};
  a.M = n * UG_NUM_POINTS;
  a.relu = 1;
  a.f16 = precision == UG_PREC_FP16;
  a.cout = C;
  a.pdl = use_pdl ? 1 : 0;
  a.sat_count = d_sat.as<unsigned long long>();
  a.nt = sb_nt;
  a.pair = sb_pair;
  a.cin = kCinPad;
  a.bias = W.convs[0].bias_tc.as<float>();
  a.out_hi = d_s_hi[0].p;
  a.out_lo = split ? d_s_lo[0].p : nullptr;
  cudaError_t e = conv3x3_sb_launch(v2_a_in, W.convs[0].tm_w_sb, &v2_out_hi[0], split ? &v2_out_lo[0] : nullptr, a, 32,
                                    device, sms, s);
  if (e != cudaSuccess) return fail(-1, std::string("stem conv launch (small batch): ") + cudaGetErrorString(e));
  a.cin = C;
  *final_act = d_s_hi[0].p;
  *final_lo = split ? d_s_lo[0].p : nullptr;
  for (int b = 0; b < W.blocks && debug_stop != 0; ++b) {
    const int x = b & 1, y = x ^ 1;
    const ConvLayer& c1 = W.convs[1 + 2 * b];
    const ConvLayer& c2 = W.convs[2 + 2 * b];
    a.bias = c1.bias_tc.as<float>();
    a.res_hi = a.res_lo = nullptr;
    a.out_hi = d_t.p;
    a.out_lo = nullptr;
    e = conv3x3_sb_launch(v2_a_s[x], c1.tm_w_sb, &v2_out_t, nullptr, a, 64, device, sms, s);
    if (e != cudaSuccess) return fail(-1, std::string("tower conv launch (small batch): ") + cudaGetErrorString(e));
    *final_act = d_t.p;
    *final_lo = nullptr;
    if (debug_stop == 2 * b + 1) break;
    // Last layer: through conv3x3_v2 with the head 1x1 convolutions fused into its epilogue, exactly as at large
    // batches, so that a position's policy and value are bit-identical whatever batch it rides in (the fused heads read
    // the fp32 accumulators; an unfused head kernel would read the 16-bit rounded activation).  One CTA-pair tile time
    // for this one layer.
    if (fuse_heads && debug_stop < 0 && b == W.blocks - 1 && conv3x3_v2_can_fuse_heads(64, 1, split, split, W.pc + W.vc)) {
      ConvV2Args v{};
      v.M = a.M;
      v.relu = 1;
      v.f16 = a.f16;
      v.cin = v.cout = C;
      v.pdl = a.pdl;
      v.sat_count = a.sat_count;
      v.bias = c2.bias_tc.as<float>();
      v.has_res = 1;
      v.has_res_lo = v.has_out_lo = split;
      v.head_w = W.head_conv_w_tc.as<float>();
      v.head_b = W.head_conv_b.as<float>();
      v.head_out = d_feat.as<float>();
      v.head_channels = W.pc + W.vc;
      const ConvV2Maps m{&v2_a_t, &c2.tm_w[1], &v2_res_hi[x], split ? &v2_res_lo[x] : nullptr, &v2_out_hi[y],
                         split ? &v2_out_lo[y] : nullptr};
      e = conv3x3_v2_launch(m, v, 64, device, sms, s);
      if (e != cudaSuccess) return fail(-1, std::string("tower conv launch: ") + cudaGetErrorString(e));
      *final_act = nullptr;  // d_feat already holds relu(head conv)
      *final_lo = nullptr;
      break;
    }
    a.bias = c2.bias_tc.as<float>();
    a.res_hi = d_s_hi[x].p;
    a.res_lo = split ? d_s_lo[x].p : nullptr;
    a.out_hi = d_s_hi[y].p;
    a.out_lo = split ? d_s_lo[y].p : nullptr;
    e = conv3x3_sb_launch(v2_a_t, c2.tm_w_sb, &v2_out_hi[y], split ? &v2_out_lo[y] : nullptr, a, 64, device, sms, s);
    if (e != cudaSuccess) return fail(-1, std::string("tower conv launch (small batch): ") + cudaGetErrorString(e));
    *final_act = d_s_hi[y].p;
    *final_lo = split ? d_s_lo[y].p : nullptr;
    if (debug_stop == 2 * b + 2) break;
  }
  return 0;
}

int ug_eval::enqueue_heads(const void* act, const void* act_lo, int n, const uint32_t* d_leg, int vt, float* d_pol,
                           float* d_val, cudaStream_t s, bool compact) {
  const HeadWeights hw = precision == UG_PREC_FP32 ? weights->head() : weights->head_tc();
  if (act != nullptr)  // null: the tower's last layer already produced d_feat (fused head convolutions)
    launch_head_conv(act, act_lo, precision == UG_PREC_FP32 ? 1 : (precision == UG_PREC_FP16 ? 2 : 0), hw,
                     d_feat.as<float>(), n, s);
  launch_head_fc(d_feat.as<float>(), hw, d_leg, vt, d_pol, d_val, n, s, compact ? 1 : 0);
  return cudaGetLastError() == cudaSuccess ? 0 : fail(-1, "head kernel launch failed");
}

int ug_eval::enqueue_raw(Src kind, const void* d_src, int n, const uint32_t* d_leg, int vt, float* d_pol,
                         float* d_val, cudaStream_t s, bool compact) {
  int rc = enqueue_encoder(kind, d_src, n, s);
  if (rc) return rc;
  const void* act = nullptr;
  const void* act_lo = nullptr;
  rc = enqueue_tower(n, s, &act, &act_lo);
  if (rc) return rc;
  return enqueue_heads(act, act_lo, n, d_leg, vt, d_pol, d_val, s, compact);
}

int ug_eval::enqueue(Src kind, const void* d_src, int n, const uint32_t* d_leg, int vt, float* d_pol,
                     float* d_val, cudaStream_t s, bool compact) {
  if (!weights) return fail(-1, "no network loaded (LoadGraph + UpdateCheckPoint first)");
  if (n <= 0 || n > max_batch) return fail(-1, "batch size " + std::to_string(n) + " outside [1, max_batch]");
  if (!use_graphs) return enqueue_raw(kind, d_src, n, d_leg, vt, d_pol, d_val, s, compact);
  const GraphKey key((int)kind, n, d_src, (const void*)d_leg, (const void*)d_pol, (const void*)d_val, vt,
                     conv_ctas * 4 + precision + (resid_split ? 8192 : 0) + conv_impl * 16384 + (use_pdl ? 65536 : 0) + (fuse_heads ? 131072 : 0) + (tile_deps ? 262144 : 0) + (debug_stop + 1) * 16 +
                         (use_small_batch(n) ? (1 << 19) + (sb_nt << 20) + ((sb_pair + 1) << 29) : 0) + (compact ? (1 << 28) : 0));
  auto it = graphs.find(key);
  if (it == graphs.end()) {
    if (graphs.size() > 1024) drop_graphs();
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
      return enqueue_raw(kind, d_src, n, d_leg, vt, d_pol, d_val, s, compact);
    const int rc = enqueue_raw(kind, d_src, n, d_leg, vt, d_pol, d_val, s, compact);
    const cudaError_t ee = cudaStreamEndCapture(s, &g);
    if (rc != 0 || ee != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      return rc ? rc : fail(-1, std::string("graph capture failed: ") + cudaGetErrorString(ee));
    }
    cudaGraphExec_t ex = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (ie != cudaSuccess) return fail(-1, std::string("graph instantiate failed: ") + cudaGetErrorString(ie));
    it = graphs.emplace(key, ex).first;
  }
  const cudaError_t le = cudaGraphLaunch(it->second, s);
  return le == cudaSuccess ? 0 : fail(-1, std::string("graph launch failed: ") + cudaGetErrorString(le));
}

// ------------------------------------------------------------------------------------------ C ABI
extern "C" {

const char* ug_version(void) { return "ugeval sm_100a " __DATE__; }

uint32_t ug_crc32c(const void* p, size_t n) { return ug::crc32c(static_cast<const uint8_t*>(p), n); }

const char* ug_last_error(const ug_eval* h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

ug_eval* ug_eval_create(int device, int precision, int max_batch) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    g_create_error = "no CUDA device: libugeval has no CPU fallback";
    return nullptr;
  }
  if (device < 0 || device >= count) { g_create_error = "bad device ordinal"; return nullptr; }
  if (precision != UG_PREC_BF16 && precision != UG_PREC_FP32 && precision != UG_PREC_FP16) {
    g_create_error = "bad precision";
    return nullptr;
  }
  if (max_batch <= 0 || max_batch > 65536) { g_create_error = "max_batch outside [1, 65536]"; return nullptr; }
  cudaDeviceProp prop;
  cudaSetDevice(device);
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
    g_create_error = "device is not sm_100 (Blackwell B200): this library is built for sm_100a only";
    return nullptr;
  }
  ug_eval* h = new ug_eval();
  h->device = device;
  h->precision = h->precision_requested = precision;
  if (const char* fr = getenv("UGEVAL_FP16_RANGE"))
    h->fp16_policy = !strcmp(fr, "refuse") ? 1 : (!strcmp(fr, "ignore") ? 2 : (!strcmp(fr, "bf16") ? 3 : 0));
  h->max_batch = max_batch;
  h->sms = prop.multiProcessorCount;
  const char* ng = getenv("UGEVAL_NO_GRAPH");
  h->use_graphs = !(ng && ng[0] == '1');
  const char* cc = getenv("UGEVAL_CONV_CTAS");
  if (cc && (cc[0] == '1' || cc[0] == '2')) h->conv_ctas = cc[0] - '0';
  const char* fh = getenv("UGEVAL_FUSE_HEADS");
  if (fh && (fh[0] == '0' || fh[0] == '1')) h->fuse_heads = fh[0] == '1';
  const char* td = getenv("UGEVAL_TILE_DEPS");
  if (td && (td[0] == '0' || td[0] == '1')) h->tile_deps = td[0] == '1';
  const char* pd = getenv("UGEVAL_PDL");
  if (pd && (pd[0] == '0' || pd[0] == '1')) h->use_pdl = pd[0] == '1';
  const char* ci = getenv("UGEVAL_CONV_IMPL");
  if (ci && (ci[0] == '1' || ci[0] == '2')) h->conv_impl = ci[0] - '0';
  if (const char* sm = getenv("UGEVAL_SB_MAX")) h->sb_max_batch = atoi(sm) > 0 ? atoi(sm) : 0;
  if (const char* sn = getenv("UGEVAL_SB_NT")) {
    const int v = atoi(sn);
    if (v == 16 || v == 32 || v == 64 || v == 128) h->sb_nt = v;
  }
  const char* sp = getenv("UGEVAL_SB_PAIR");
  if (sp && (sp[0] == '0' || sp[0] == '1')) h->sb_pair = sp[0] - '0';
  const char* rs = getenv("UGEVAL_RESID_SPLIT");
  if (rs && (rs[0] == '0' || rs[0] == '1')) h->resid_split = rs[0] == '1';
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_error = "cannot create stream";
    delete h;
    return nullptr;
  }
  return h;
}

void ug_eval_destroy(ug_eval* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  delete h;
}

int ug_eval_set_io(ug_eval* h, const char* input, const char* const* outputs, int n_outputs) {
  if (!h) return -1;
  if (input) h->input_name = input;
  if (outputs) {
    h->outputs.clear();
    for (int i = 0; i < n_outputs; ++i) h->outputs.emplace_back(outputs[i] ? outputs[i] : "");
  }
  h->plan_ready = false;
  return 0;
}

static int rebuild_plan(ug_eval* h) {
  if (h->outputs.size() < 2) return h->fail(-1, "nn_outputs must name the policy and the value node");
  std::string err;
  if (!build_net_plan(h->graph, h->input_name, h->outputs[0], h->outputs[1], &h->plan, &err))
    return h->fail(-1, "unsupported graph: " + err);
  h->plan_ready = true;
  return 0;
}

static int install_weights(ug_eval* h, const TfBundle& ck, const std::string& label);

int ug_eval_load_graph(ug_eval* h, const char* path) {
  if (!h) return -1;
  const std::string p = path ? path : "";
  const bool meta = ends_with(p, ".meta");
  const bool pb = ends_with(p, ".pb");
  if (!meta && !pb) return 1;  // reference: LoadGraph returns false for any other suffix (.cc:77-83)
  std::vector<uint8_t> bytes;
  if (!read_file(p, &bytes)) return h->fail(-1, "Error reading graph definition from " + p);
  std::string err;
  TfGraph g;
  if (!parse_tf_graph(bytes, meta, &g, &err)) return h->fail(-1, "Error reading graph definition from " + p + ": " + err);
  h->graph = std::move(g);
  h->graph_loaded = true;
  h->plan_ready = false;
  const int rc = rebuild_plan(h);
  if (rc) { h->graph_loaded = false; return rc; }
  if (h->plan.frozen) {
    // every weight is a Const node (a graph run through freeze_graph): evaluate straight after LoadGraph, as the
    // reference's session does after LoadPbGraph (DlTFNetworkEvaluator.cc:85-111)
    cudaSetDevice(h->device);
    TfBundle consts;
    if (!add_graph_constants(h->graph, h->plan, &consts, &err)) { h->graph_loaded = false; return h->fail(-1, "frozen graph " + p + ": " + err); }
    if (install_weights(h, consts, "frozen graph " + p) != 0) { h->graph_loaded = false; return -1; }
  }
  return 0;
}

// Weights from `ck` (a checkpoint bundle, or the Const nodes of a frozen graph) -> folded, packed, device-resident;
// range calibration and the fp16 policy; swapped in between batches.  `label` names the source in messages.
static int install_weights(ug_eval* h, const TfBundle& ck, const std::string& label) {
  std::string err;
  int precision = h->precision_requested;
  Weights* W = build_weights(h->plan, ck, precision == UG_PREC_FP16, nullptr, &err);
  if (!W) return h->fail(1, label + ": " + err);
  if (!calibrate_range(h, W, &err)) { delete W; return h->fail(1, label + ": " + err); }
  {
    const float worst = W->max_activation > W->max_weight ? W->max_activation : W->max_weight;
    if (worst != worst) { delete W; return h->fail(1, label + ": weights or calibration activations are NaN"); }
    if (precision == UG_PREC_FP16 && worst > 65504.f / 8 && h->fp16_policy != 2) {
      char msg[512];
      snprintf(msg, sizeof(msg), "%s leaves fp16 less than 8x headroom (largest |activation| %.4g after layer %d, "
               "largest |folded weight| %.4g in layer %d, fp16 saturates at 65504)", label.c_str(),
               (double)W->max_activation, W->max_activation_layer, (double)W->max_weight, W->max_weight_layer);
      if (h->fp16_policy == 1) { delete W; return h->fail(1, std::string(msg) + ": refused (UGEVAL_FP16_RANGE=refuse)"); }
      // auto-ranging: store every hot tensor times 2^-shift so that its calibrated maximum sits near 2^11 (32x headroom)
      const int blocks = W->blocks;
      std::vector<int> shifts((size_t)(1 + blocks), 0);
      auto shift_for = [](float mx) { int k = 0; while (std::ldexp(mx, -k) > 2048.f && k < 60) ++k; return k; };
      float stream_max = W->layer_max.empty() ? 0.f : W->layer_max[0];
      for (int b = 0; b < blocks; ++b) stream_max = std::max(stream_max, W->layer_max[(size_t)(2 * b + 2)]);
      shifts[0] = shift_for(stream_max);
      for (int b = 0; b < blocks; ++b) shifts[(size_t)(1 + b)] = shift_for(W->layer_max[(size_t)(2 * b + 1)]);
      const std::vector<float> lm = W->layer_max;
      const float ma = W->max_activation, mw = W->max_weight;
      const int la = W->max_activation_layer, lw = W->max_weight_layer;
      delete W;
      W = nullptr;
      if (h->fp16_policy == 0) {
        W = build_weights(h->plan, ck, true, &shifts, &err);
        if (!W) return h->fail(1, label + ": " + err);
        if (W->max_weight > 65504.f / 2 || !(W->max_weight == W->max_weight)) {  // rescaling cannot help the weights
          delete W;
          W = nullptr;
        } else {
          int bmax = 0;
          for (int v : W->block_shift) bmax = std::max(bmax, v);
          fprintf(stderr, "ugeval: %s: activations rescaled by powers of two (stream 2^-%d, blocks up to 2^-%d)\n", msg,
                  W->stream_shift, bmax);
        }
      }
      if (!W) {
        fprintf(stderr, "ugeval: %s: evaluating with bf16 operands instead\n", msg);
        precision = UG_PREC_BF16;
        W = build_weights(h->plan, ck, false, nullptr, &err);
        if (!W) return h->fail(1, label + ": " + err);
      }
      if (W->stream_shift == 0) W->max_weight = mw;  // (rescaled weights keep the maximum of what was packed)
      W->layer_max = lm; W->max_activation = ma; W->max_activation_layer = la; W->max_weight_layer = lw;
    }
  }
  if (precision != h->precision) {  // 16-bit formats share every buffer; the format is a run-time flag of the kernels
    cudaStreamSynchronize(h->stream);
    h->drop_graphs();
    h->precision = precision;
  }
  if (!h->ensure_workspace(W->C)) {
    delete W;
    if (h->weights && h->weights->C != h->ws_C) {  // the old network's workspace did not survive the attempt
      cudaStreamSynchronize(h->stream);
      delete h->weights;
      h->weights = nullptr;
    }
    return 1;
  }
  cudaStreamSynchronize(h->stream);
  h->drop_graphs();  // graphs bake weight pointers
  delete h->weights;
  h->weights = W;
  return 0;
}

int ug_eval_load_checkpoint(ug_eval* h, const char* prefix) {
  if (!h) return 1;
  if (prefix && prefix[0]) h->ckpt_prefix = prefix;
  if (h->ckpt_prefix.empty()) return h->fail(1, "empty checkpoint path");
  if (!h->graph_loaded) return h->fail(1, "UpdateCheckPoint before LoadGraph");
  if (!h->plan_ready && rebuild_plan(h) != 0) return 1;
  // A frozen graph has no restore op to run: the reference's UpdateCheckPointForPbGraph fails on it and the session
  // goes on evaluating the constants (DlTFNetworkEvaluator.cc:203-226); so does this (the weights came with LoadGraph).
  if (h->plan.frozen) return h->fail(1, "the graph is frozen (every weight is a constant): there is nothing to restore");
  cudaSetDevice(h->device);
  std::string err;
  TfBundle ck;
  if (!ck.open(h->ckpt_prefix, &err)) return h->fail(1, err);
  if (h->plan.const_weights > 0 && !add_graph_constants(h->graph, h->plan, &ck, &err)) return h->fail(1, err);
  return install_weights(h, ck, "checkpoint " + h->ckpt_prefix);
}

int ug_eval_ready(const ug_eval* h) { return h && h->graph_loaded ? 1 : 0; }
int ug_eval_has_weights(const ug_eval* h) { return h && h->weights ? 1 : 0; }

int ug_eval_set_value_transform(ug_eval* h, int vt) {
  if (!h) return -1;
  h->value_transform = (vt == UG_TR_FLIP_SIGN || vt == UG_TR_FLIP_PROB) ? vt : UG_TR_IDENTITY;
  return 0;
}

int ug_eval_set_debug_stop(ug_eval* h, int layer) {
  if (!h) return -1;
  h->debug_stop = layer < 0 ? -1 : layer;
  return 0;
}

int ug_eval_set_conv_ctas(ug_eval* h, int ctas) {
  if (!h || (ctas != 1 && ctas != 2)) return -1;
  h->conv_ctas = ctas;
  return 0;
}

int ug_eval_efficient_batch(const ug_eval* h, int n) { return h ? ug_efficient_batch_for_sms(h->sms, n) : n; }

int ug_efficient_batch_for_sms(int sm_count, int n) {
  if (sm_count < 2 || n <= 0) return n;
  // one wave of the persistent tower kernel = (SMs / 2) CTA pairs x 256 pixel rows
  const long wave_rows = (long)(sm_count / 2) * 256;
  // the largest batch of w waves is floor(w * wave_rows / 361); take the largest w whose batch still fits in n
  const long w = ((long)(n + 1) * UG_NUM_POINTS + wave_rows - 1) / wave_rows - 1;
  if (w <= 0) return n;  // less than one wave costs one wave whatever n is
  const long m = w * wave_rows / UG_NUM_POINTS;
  return (int)(m < n ? m : n);
}

int ug_eval_set_conv_impl(ug_eval* h, int impl) {
  if (!h || (impl != 1 && impl != 2)) return -1;
  h->conv_impl = impl;
  return 0;
}

int ug_eval_set_residual_split(ug_eval* h, int on) {
  if (!h) return -1;
  h->resid_split = on != 0;
  return 0;
}

int ug_eval_set_fp16_range_policy(ug_eval* h, int policy) {
  if (!h || policy < 0 || policy > 3) return -1;
  h->fp16_policy = policy;
  return 0;
}

int ug_eval_set_small_batch(ug_eval* h, int max_batch, int nt) {
  const int width = nt & 0xff, mode = nt >> 8;  // mode: 0 = choose, 1 = single CTAs, 2 = CTA pairs
  if (!h || max_batch < 0 || nt < 0 || mode > 2 || (width != 0 && width != 16 && width != 32 && width != 64 && width != 128))
    return -1;
  if (mode == 2 && width == 16) return -1;  // a pair stages nt / 2 weight rows per CTA in 16-row TMA boxes
  h->sb_max_batch = max_batch;
  h->sb_nt = width;
  h->sb_pair = mode - 1;
  return 0;
}

int ug_eval_info(const ug_eval* h, ug_net_info* out) {
  if (!h || !out || !h->weights) return -1;
  const Weights& W = *h->weights;
  out->blocks = W.blocks;
  out->filters = W.C;
  out->policy_channels = W.pc;
  out->value_channels = W.vc;
  out->value_hidden = W.H;
  out->input_planes = UG_NUM_PLANES;
  const double C = W.C, P = UG_NUM_POINTS;
  out->flops_per_position = 2 * P * 9 * 17 * C + W.blocks * 2.0 * (2 * P * 9 * C * C) + 2 * P * C * W.pc +
                            2.0 * (P * W.pc) * UG_NUM_MOVES + 2 * P * C * W.vc + 2.0 * (P * W.vc) * W.H + 2.0 * W.H;
  out->conv_ctas = h->conv_ctas;
  out->sm_count = h->sms;
  out->conv_impl = (h->conv_impl == 2 && W.C % 64 == 0) ? 2 : 1;
  out->small_batch_max = out->conv_impl == 2 ? h->sb_max_batch : 0;
  out->precision = h->precision;
  out->calib_max_activation = W.max_activation;
  out->calib_max_weight = W.max_weight;
  {  // headroom of what is actually stored: calibrated maxima after the power-of-two shifts, and the packed weights
    float worst = W.max_weight;
    for (size_t l = 0; l < W.layer_max.size(); ++l) {
      const int sh = (l == 0 || l % 2 == 0) ? W.stream_shift : W.block_shift[(l - 1) / 2];
      worst = std::max(worst, std::ldexp(W.layer_max[l], -sh));
    }
    out->fp16_headroom = worst > 0 ? 65504.f / worst : 0.f;
  }
  out->fp16_stream_shift = W.stream_shift;
  out->fp16_max_block_shift = 0;
  for (int v : W.block_shift) out->fp16_max_block_shift = std::max(out->fp16_max_block_shift, v);
  out->saturation_events = 0;
  if (h->d_sat.p) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    unsigned long long v = 0;
    if (cudaMemcpy(&v, h->d_sat.p, sizeof(v), cudaMemcpyDeviceToHost) == cudaSuccess) out->saturation_events = v;
  }
  return 0;
}

int ug_eval_sync(ug_eval* h) {
  if (!h) return -1;
  const cudaError_t e = cudaStreamSynchronize(h->stream);
  return e == cudaSuccess ? 0 : h->fail(-1, std::string("CUDA error: ") + cudaGetErrorString(e));
}

// shared by the two reference-facing overloads
static int run_planes_common(ug_eval* h, const int8_t* planes, int hwc, int n, double* policy, double* value) {
  if (!h) return -1;
  if (!h->weights) return h->fail(-1, "Running model failed: no network loaded");
  if (n <= 0 || n > h->max_batch) return h->fail(-1, "Running model failed: bad batch size");
  cudaSetDevice(h->device);
  const size_t in_bytes = (size_t)n * UG_NUM_PLANES * UG_NUM_POINTS;
  const size_t pol_count = (size_t)n * UG_NUM_MOVES;
  uint8_t* stage = static_cast<uint8_t*>(h->h_pinned);
  float* out_stage = reinterpret_cast<float*>(stage + (((size_t)h->max_batch * UG_NUM_PLANES * UG_NUM_POINTS + 255) & ~size_t(255)));
  cudaStream_t s = h->stream;
  // caller buffers registered with ug_eval_register_host are DMA sources/targets themselves (the reference's
  // EvalBuffer is one fixed allocation, search/UctSearch.h:126-130); anything else goes through pinned staging
  if (n > 32 && h->is_registered(planes, in_bytes)) {  // (small batches: the staged copy is faster, see below)
    cudaMemcpyAsync(h->d_planes.p, planes, in_bytes, cudaMemcpyHostToDevice, s);
  } else {
    memcpy(stage, planes, in_bytes);
    cudaMemcpyAsync(h->d_planes.p, stage, in_bytes, cudaMemcpyHostToDevice, s);
  }
  // At the reference engine's own batch sizes the head kernel writes its 363 floats per position straight into the
  // pinned staging area (posted writes over PCIe) instead of two device-to-host copies being queued behind it: two
  // API calls and two copy-engine start-ups less per Evaluate() (measured at n = 1: see DESIGN.md 5a).
  const bool zero_copy_out = n <= 32 && h->d_pinned_alias != nullptr;
  float* d_pol = zero_copy_out ? reinterpret_cast<float*>(h->d_pinned_alias + (reinterpret_cast<uint8_t*>(out_stage) - stage))
                               : h->d_policy.as<float>();
  float* d_val = zero_copy_out ? d_pol + pol_count : h->d_value.as<float>();
  // value_transform is applied in double after widening, exactly as the reference does (.cc:298-310)
  int rc = h->enqueue(hwc ? ug_eval::SRC_HWC : ug_eval::SRC_CHW, h->d_planes.p, n, nullptr, UG_TR_IDENTITY, d_pol, d_val, s);
  if (rc) return rc;
  // Registered outputs are worth it for large batches only (the widening kernel and two DMA writes of doubles replace
  // a host loop over n * 362 floats); at the reference engine's own batch sizes the extra launch costs more than the
  // loop: measured 0.41 ms against 0.35 ms per call at n = 1.
  const bool direct_out = n > 32 && h->is_registered(policy, pol_count * 8) && h->is_registered(value, (size_t)n * 8);
  if (direct_out) {
    // Outputs must stay untouched on failure (.cc:315-317): a faulting kernel poisons the stream, so the copies
    // queued behind it never run.
    launch_widen_outputs(h->d_policy.as<float>(), h->d_value.as<float>(), h->d_policy64.as<double>(),
                         h->d_value64.as<double>(), n, h->value_transform, s);
    cudaMemcpyAsync(policy, h->d_policy64.p, pol_count * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(value, h->d_value64.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s);
    const cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return h->fail(-1, std::string("Running model failed: ") + cudaGetErrorString(e));
    return 0;
  }
  if (!zero_copy_out) {
    cudaMemcpyAsync(out_stage, h->d_policy.p, pol_count * 4, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(out_stage + pol_count, h->d_value.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s);
  }
  const cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return h->fail(-1, std::string("Running model failed: ") + cudaGetErrorString(e));
  for (size_t i = 0; i < pol_count; ++i) policy[i] = (double)out_stage[i];  // DlTensorUtil::GetValue
  const float* v = out_stage + pol_count;
  for (int i = 0; i < n; ++i) {
    double d = (double)v[i];
    if (h->value_transform == UG_TR_FLIP_SIGN) d = -d;
    else if (h->value_transform == UG_TR_FLIP_PROB) d = 1 - d;
    value[i] = d;
  }
  return 0;
}

int ug_eval_register_host(ug_eval* h, void* p, size_t bytes) {
  if (!h || !p || bytes == 0) return -1;
  if (h->is_registered(p, bytes)) return 0;
  if (h->host_regs.size() >= 64) return h->fail(1, "too many registered host buffers");
  cudaSetDevice(h->device);
  const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return h->fail(1, std::string("cudaHostRegister failed: ") + cudaGetErrorString(e));
  }
  h->host_regs.emplace_back(reinterpret_cast<uintptr_t>(p), bytes);
  return 0;
}

int ug_eval_unregister_host(ug_eval* h, void* p) {
  if (!h || !p) return -1;
  for (size_t i = 0; i < h->host_regs.size(); ++i)
    if (h->host_regs[i].first == reinterpret_cast<uintptr_t>(p)) {
      cudaSetDevice(h->device);
      cudaStreamSynchronize(h->stream);
      cudaHostUnregister(p);
      h->host_regs.erase(h->host_regs.begin() + (long)i);
      return 0;
    }
  return 1;
}

int ug_eval_run_planes(ug_eval* h, const int8_t* chw, int n, double* policy, double* value) {
  return run_planes_common(h, chw, 0, n, policy, value);
}
int ug_eval_run_planes_hwc(ug_eval* h, const int8_t* hwc, int n, double* policy, double* value) {
  return run_planes_common(h, hwc, 1, n, policy, value);
}

int ug_eval_run_packed(ug_eval* h, const ug_packed_pos* pos, int n, const uint32_t* legal, float* policy,
                       float* value) {
  if (!h) return -1;
  if (!h->weights) return h->fail(-1, "no network loaded");
  if (n <= 0 || n > h->max_batch) return h->fail(-1, "bad batch size");
  cudaSetDevice(h->device);
  uint8_t* stage = static_cast<uint8_t*>(h->h_pinned);
  const size_t pos_bytes = (size_t)n * sizeof(ug_packed_pos);
  const size_t legal_off = (size_t)h->max_batch * sizeof(ug_packed_pos);
  const size_t out_off = legal_off + (size_t)h->max_batch * UG_MASK_WORDS * 4;
  memcpy(stage, pos, pos_bytes);
  cudaStream_t s = h->stream;
  cudaMemcpyAsync(h->d_pos.p, stage, pos_bytes, cudaMemcpyHostToDevice, s);
  const uint32_t* d_legal = nullptr;
  if (legal) {
    memcpy(stage + legal_off, legal, (size_t)n * UG_MASK_WORDS * 4);
    cudaMemcpyAsync(h->d_legal.p, stage + legal_off, (size_t)n * UG_MASK_WORDS * 4, cudaMemcpyHostToDevice, s);
    d_legal = h->d_legal.as<uint32_t>();
  }
  float* out_stage = reinterpret_cast<float*>(stage + out_off);
  const bool zero_copy_out = n <= 32 && h->d_pinned_alias != nullptr;   // as in run_planes_common
  float* d_pol = zero_copy_out ? reinterpret_cast<float*>(h->d_pinned_alias + out_off) : h->d_policy.as<float>();
  float* d_val = zero_copy_out ? d_pol + (size_t)n * UG_NUM_MOVES : h->d_value.as<float>();
  int rc = h->enqueue(ug_eval::SRC_PACKED, h->d_pos.p, n, d_legal, h->value_transform, d_pol, d_val, s);
  if (rc) return rc;
  if (!zero_copy_out) {
    cudaMemcpyAsync(out_stage, h->d_policy.p, (size_t)n * UG_NUM_MOVES * 4, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(out_stage + (size_t)n * UG_NUM_MOVES, h->d_value.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s);
  }
  const cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return h->fail(-1, std::string("CUDA error: ") + cudaGetErrorString(e));
  memcpy(policy, out_stage, (size_t)n * UG_NUM_MOVES * 4);
  memcpy(value, out_stage + (size_t)n * UG_NUM_MOVES, (size_t)n * 4);
  return 0;
}

int ug_eval_run_packed_device(ug_eval* h, const ug_packed_pos* d_pos, int n, const uint32_t* d_legal,
                              float* d_policy, float* d_value, void* stream) {
  if (!h) return -1;
  cudaSetDevice(h->device);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  return h->enqueue(ug_eval::SRC_PACKED, d_pos, n, d_legal, h->value_transform, d_policy, d_value, s);
}

// ---- encoder hooks
static int encode_common(ug_eval* h, int what, const void* src, size_t src_bytes, int n, int8_t* out) {
  if (!h) return -1;
  if (n <= 0 || n > h->max_batch) return h->fail(-1, "bad batch size");
  cudaSetDevice(h->device);
  if (!h->ensure_workspace(h->ws_C ? h->ws_C : 64)) return -1;
  cudaStream_t s = h->stream;
  const size_t out_bytes = (size_t)n * UG_NUM_PLANES * UG_NUM_POINTS;
  void* d_src = (what == 2) ? h->d_planes.p : h->d_pos.p;
  cudaMemcpyAsync(d_src, src, src_bytes, cudaMemcpyHostToDevice, s);
  if (what == 0) {
    launch_encode_packed_chw(h->d_pos.as<ug_packed_pos>(), h->d_bytes.as<int8_t>(), n, s);
  } else {
    const int f16 = h->precision == UG_PREC_FP16;
    if (what == 1) launch_encode_packed_nhwc(h->d_pos.as<ug_packed_pos>(), h->d_in.as<__nv_bfloat16>(), n, f16, s);
    else launch_planes_to_nhwc(h->d_planes.as<int8_t>(), 0, h->d_in.as<__nv_bfloat16>(), n, f16, s);
    launch_nhwc_to_bytes(h->d_in.as<__nv_bfloat16>(), h->d_bytes.as<int8_t>(), n, f16, s);
  }
  cudaMemcpyAsync(out, h->d_bytes.p, out_bytes, cudaMemcpyDeviceToHost, s);
  const cudaError_t e = cudaStreamSynchronize(s);
  return e == cudaSuccess ? 0 : h->fail(-1, std::string("CUDA error: ") + cudaGetErrorString(e));
}
int ug_encode_planes(ug_eval* h, const ug_packed_pos* pos, int n, int8_t* chw_out) {
  return encode_common(h, 0, pos, (size_t)n * sizeof(ug_packed_pos), n, chw_out);
}
int ug_encode_nhwc(ug_eval* h, const ug_packed_pos* pos, int n, int8_t* hwc_out) {
  return encode_common(h, 1, pos, (size_t)n * sizeof(ug_packed_pos), n, hwc_out);
}
int ug_transform_planes(ug_eval* h, const int8_t* chw, int n, int8_t* hwc_out) {
  return encode_common(h, 2, chw, (size_t)n * UG_NUM_PLANES * UG_NUM_POINTS, n, hwc_out);
}

int ug_eval_debug_layer(ug_eval* h, int layer, int n, float* out) {
  if (!h || !h->weights) return -1;
  const Weights& W = *h->weights;
  if (layer < 0 || layer > 2 * W.blocks || n <= 0 || n > h->max_batch) return h->fail(-1, "bad layer or n");
  // layer 0 (stem) -> S[0]; conv1 of block b -> T; conv2 of block b -> S[(b+1)&1].  Buffers are reused by later
  // blocks: run with ug_eval_set_debug_stop(h, layer) first so that `layer` is the last writer of its buffer.
  const bool is_t = layer > 0 && (layer & 1);
  const int sidx = layer == 0 ? 0 : ((layer / 2) & 1);
  cudaSetDevice(h->device);
  const size_t count = (size_t)n * UG_NUM_POINTS * W.C;
  cudaStream_t s = h->stream;
  if (h->precision == UG_PREC_FP32) {
    cudaMemcpyAsync(out, is_t ? h->f_t.p : h->f_s[sidx].p, count * 4, cudaMemcpyDeviceToHost, s);
  } else {
    DevBuf tmp;
    if (!tmp.alloc(count * 4)) return h->fail(-1, "out of memory");
    launch_bf16_to_f32(is_t ? h->d_t.as<__nv_bfloat16>() : h->d_s_hi[sidx].as<__nv_bfloat16>(), tmp.as<float>(), count,
                       h->precision == UG_PREC_FP16, s);
    cudaMemcpyAsync(out, tmp.p, count * 4, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
  }
  const cudaError_t e = cudaStreamSynchronize(s);
  return e == cudaSuccess ? W.C : h->fail(-1, std::string("CUDA error: ") + cudaGetErrorString(e));
}

int ug_eval_bench(ug_eval* h, const ug_packed_pos* host_pos, int n, int iters, ug_bench_result* out) {
  if (!h || !out) return -1;
  if (!h->weights) return h->fail(-1, "no network loaded");
  if (n <= 0 || n > h->max_batch || iters <= 0) return h->fail(-1, "bad arguments");
  cudaSetDevice(h->device);
  cudaStream_t s = h->stream;
  cudaMemcpyAsync(h->d_pos.p, host_pos, (size_t)n * sizeof(ug_packed_pos), cudaMemcpyHostToDevice, s);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto timed = [&](const std::function<int()>& body, float* ms_avg) -> int {
    int rc = body();  // warm
    if (rc) return rc;
    cudaEventRecord(e0, s);
    for (int i = 0; i < iters && rc == 0; ++i) rc = body();
    cudaEventRecord(e1, s);
    if (cudaEventSynchronize(e1) != cudaSuccess) return h->fail(-1, "CUDA error in bench");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_avg = ms / iters;
    return rc;
  };
  float full = 0.f;
  int rc = timed([&] { return h->enqueue(ug_eval::SRC_PACKED, h->d_pos.p, n, nullptr, h->value_transform,
                                         h->d_policy.as<float>(), h->d_value.as<float>(), s); }, &full);
  const void* act = nullptr;
  const void* act_lo = nullptr;
  if (!rc) rc = timed([&] { return h->enqueue_encoder(ug_eval::SRC_PACKED, h->d_pos.p, n, s); }, &out->ms_encoder);
  if (!rc) rc = timed([&] { return h->enqueue_tower(n, s, &act, &act_lo); }, &out->ms_tower);
  if (!rc) rc = timed([&] { return h->enqueue_heads(act, act_lo, n, nullptr, h->value_transform, h->d_policy.as<float>(),
                                                    h->d_value.as<float>(), s); }, &out->ms_heads);
  out->ms_total = full * iters;
  out->launches_per_pass = 1 + (h->precision == UG_PREC_FP32 ? 1 : 0) + 1 + 2 * h->weights->blocks + (act == nullptr ? 1 : 2);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

int ug_k_conv3x3_bf16(int device, const void* d_in, const void* d_w, const float* d_bias, const void* d_res,
                      void* d_out, int n, int cin, int cout, int relu, int ck, int ctas) {
  cudaSetDevice(device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
  std::string err;
  CUtensorMap ta, tb;
  if (!make_tmap_act_im2col(&ta, d_in, n, cin, ck, &err)) { g_create_error = err; return -2; }
  if (!make_tmap_weights(&tb, d_w, cout, 9 * cin, ck, cout / ctas, &err)) { g_create_error = err; return -3; }
  ConvArgs a{};
  a.bias = d_bias;
  a.residual = static_cast<const __nv_bfloat16*>(d_res);
  a.out = static_cast<__nv_bfloat16*>(d_out);
  a.M = n * UG_NUM_POINTS;
  a.cin = cin;
  a.cout = cout;
  a.relu = relu;
  cudaError_t e = conv3x3_tc_launch(ta, tb, a, ck, ctas, prop.multiProcessorCount, 0);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return -4; }
  return 0;
}

int ug_k_conv3x3_sb(int device, const void* d_in, const void* d_w, const float* d_bias, const void* d_res_hi,
                    const void* d_res_lo, void* d_out_hi, void* d_out_lo, int n, int n_alloc, int cin, int cout,
                    int relu, int f16, int nt, int pdl, int iters, float* ms_out, long long* d_prof) {
  cudaSetDevice(device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
  std::string err;
  const int ck = (cin % 64 == 0) ? 64 : 32;
  CUtensorMap ta, tw;
  CUtensorMap toh, tol;
  if (!make_tmap_rows(&ta, d_in, (long)n_alloc * UG_NUM_POINTS, cin, ck, 168, &err) ||
      !make_tmap_weights(&tw, d_w, cout, 9 * cin, ck, 16, &err) ||
      !make_tmap_rows(&toh, d_out_hi, (long)n_alloc * UG_NUM_POINTS, cout, 32, 32, &err) ||
      (d_out_lo && !make_tmap_rows(&tol, d_out_lo, (long)n_alloc * UG_NUM_POINTS, cout, 32, 32, &err))) {
    g_create_error = err;
    return -2;
  }
  ConvSbArgs a{};
  a.bias = d_bias;
  a.res_hi = d_res_hi;
  a.res_lo = d_res_lo;
  a.out_hi = d_out_hi;
  a.out_lo = d_out_lo;
  a.M = n * UG_NUM_POINTS;
  a.cin = cin;
  a.cout = cout;
  a.relu = relu;
  a.f16 = f16;
  a.nt = nt & 0xff;
  a.pair = ((nt >> 8) & 3) - 1;  // as ug_eval_set_small_batch: +256 = single CTAs, +512 = CTA pairs, else the launcher chooses
  a.tma_out = (nt & 1024) ? -1 : 1;   // +1024: per-thread stores even for wide tiles; else TMA stores from 64 columns
  a.pdl = pdl;
  cudaStream_t s = nullptr;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaError_t e = conv3x3_sb_launch(ta, tw, &toh, d_out_lo ? &tol : nullptr, a, ck, device, prop.multiProcessorCount, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess && iters > 0) {
    cudaEventRecord(e0, s);
    for (int i = 0; i < iters && e == cudaSuccess; ++i) {
      // time stamps: launch i writes its own [ctas][16] block (the caller sizes d_prof for iters x 148 CTAs)
      if (d_prof) a.prof = d_prof + (size_t)i * 148 * 16;
      e = conv3x3_sb_launch(ta, tw, &toh, d_out_lo ? &tol : nullptr, a, ck, device, prop.multiProcessorCount, s);
    }
    cudaEventRecord(e1, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = ms / iters;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaStreamDestroy(s);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return -4; }
  return 0;
}

int ug_debug_max_active_clusters(int device, int cluster_size) {
  cudaSetDevice(device);
  return conv3x3_v2_max_active_clusters(cluster_size);
}

int ug_k_conv3x3_v2(int device, const void* d_in, const void* d_w, const float* d_bias, const void* d_res_hi,
                    const void* d_res_lo, void* d_out_hi, void* d_out_lo, int n, int n_alloc, int cin, int cout,
                    int relu, int f16, int desc_mode, int iters, float* ms_out, long long* d_prof) {
  cudaSetDevice(device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
  std::string err;
  const int ck = (cin % 64 == 0) ? 64 : 32;
  const long rows = (long)n_alloc * UG_NUM_POINTS;
  CUtensorMap ta, tw, trh, trl, toh, tol;
  if (!make_tmap_rows(&ta, d_in, rows, cin, ck, 168, &err) ||
      !make_tmap_weights(&tw, d_w, cout, 9 * cin, ck, cout / 2, &err) ||
      !make_tmap_rows(&toh, d_out_hi, rows, cout, 32, 32, &err) ||
      (d_out_lo && !make_tmap_rows(&tol, d_out_lo, rows, cout, 32, 32, &err)) ||
      (d_res_hi && !make_tmap_rows(&trh, d_res_hi, rows, cout, 32, 128, &err)) ||
      (d_res_lo && !make_tmap_rows(&trl, d_res_lo, rows, cout, 32, 128, &err))) {
    g_create_error = err;
    return -2;
  }
  ConvV2Maps m{&ta, &tw, d_res_hi ? &trh : nullptr, d_res_lo ? &trl : nullptr, &toh, d_out_lo ? &tol : nullptr};
  ConvV2Args a{};
  a.bias = d_bias;
  a.M = n * UG_NUM_POINTS;
  a.cin = cin;
  a.cout = cout;
  a.relu = relu;
  a.f16 = f16;
  a.has_res = d_res_hi != nullptr;
  a.has_res_lo = d_res_lo != nullptr;
  a.has_out_lo = d_out_lo != nullptr;
  a.desc_mode = desc_mode & (7 | 16);
  a.pdl = (desc_mode & 8) ? 1 : 0;  // timing probe: programmatic dependent launch between the repeated launches
  a.prof = d_prof;
  cudaStream_t s = nullptr;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaError_t e = conv3x3_v2_launch(m, a, ck, device, prop.multiProcessorCount, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess && iters > 0) {
    cudaEventRecord(e0, s);
    for (int i = 0; i < iters && e == cudaSuccess; ++i) e = conv3x3_v2_launch(m, a, ck, device, prop.multiProcessorCount, s);
    cudaEventRecord(e1, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = ms / iters;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaStreamDestroy(s);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return -4; }
  return 0;
}

}  // extern "C"
