// The evaluator object behind the C ABI: graph plan + resident weights + workspace + launch sequence.
// Host-side mirror of tensorflow::DlTFNetworkEvaluator<T> (funcapproximator/DlTFNetworkEvaluator.h:33-83).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/ugeval.h"
#include "kernels.h"
#include "tfmodel.h"

namespace ug {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  bool alloc(size_t n);
  void release();
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct ConvLayer {
  int cin = 0, cin_pad = 0, cout = 0;
  DevBuf w_bf16;   // [cout][9*cin_pad] K-major, batch-norm folded
  DevBuf w_f32;    // [3][3][cin][cout] folded (fp32 validation path)
  DevBuf bias;     // [cout] fp32 (fp32 validation tower, range calibration)
  DevBuf bias_tc;  // [cout] fp32 for the 16-bit tensor-core tower: bias * 2^-out_shift (fp16 auto-ranging), else == bias
  CUtensorMap tm_w[2];  // index = ctas-1 (box rows differ)
  CUtensorMap tm_w_sb;  // box {ck, 16}: the small-batch kernel loads its NT weight rows 16 at a time
};

struct Weights {
  int blocks = 0, C = 0, pc = 0, vc = 0, H = 0;
  bool f16 = false;              // operand format the conv weights were packed in
  // range calibration (ug_eval_load_checkpoint): the largest |folded weight| of any tower layer, and the largest
  // |layer output| seen when a few calibration positions run through the fp32 validation tower
  float max_weight = 0.f, max_activation = 0.f;
  int max_activation_layer = -1, max_weight_layer = -1;
  std::vector<float> layer_max;  // calibration: largest |output| per layer (0 = stem, 2b+1 / 2b+2 = convs of block b)
  // fp16 auto-ranging: activations are stored times 2^-shift (exact: powers of two), folded into the packed weights and
  // biases — stream_shift for the residual stream (one value: the skip connection adds without rescaling), block_shift[b]
  // for block b's internal tensor.  All zero unless the calibration found fp16 short of headroom.
  int stream_shift = 0;
  std::vector<int> block_shift;
  DevBuf head_conv_w_tc;         // head 1x1 weights * 2^stream_shift for the tensor-core tower (else a copy)
  HeadWeights head_tc() const { HeadWeights h = head(); h.conv_w = head_conv_w_tc.as<float>(); return h; }
  bool policy_softmax = true, value_tanh = true;
  std::vector<ConvLayer> convs;  // [0] stem, then 2 per block
  DevBuf head_conv_w, head_conv_b, pfc_w, pfc_b, v1_w, v1_b, v2_w, v2_b;
  HeadWeights head() const;
};

}  // namespace ug

struct ug_eval {
  int device = 0;
  int precision = UG_PREC_BF16;            // arithmetic in use
  int precision_requested = UG_PREC_BF16;  // what ug_eval_create asked for (a checkpoint outside fp16's range can
                                           // move an fp16 evaluator to bf16: fp16_policy)
  // fp16 operands saturate at 65504.  When a checkpoint's calibrated range leaves less than 8x headroom:
  // 0 = auto (default): rescale the layers by powers of two (exact; folded into weights and biases), and only if that
  //     cannot help (a folded weight out of range) evaluate with bf16 operands; 1 = refuse the checkpoint;
  // 2 = load anyway (ug_net_info.saturation_events then tells); 3 = bf16 operands without trying to rescale.
  // UGEVAL_FP16_RANGE=auto|refuse|ignore|bf16
  int fp16_policy = 0;
  int max_batch = 0;
  int sms = 148;
  int conv_ctas = 2;
  int conv_impl = 2;  // 1 = im2col kernel (conv3x3_tc.cu), 2 = resident-tile kernel (conv3x3_v2.cu)
  bool fuse_heads = true;   // head 1x1 convolutions in the last tower layer's epilogue (UGEVAL_FUSE_HEADS=0 disables)
  // Layers depend on each other tile by tile (per-tile completion counters) instead of grid by grid.  Measured
  // -2 % at batch 256 against the final kernel (the tail it can recover is the 2.4 % of 361 tiles on 74 CTA pairs):
  // an experiment kept behind UGEVAL_TILE_DEPS=1.
  bool tile_deps = false;
  bool use_pdl = true;      // programmatic dependent launch between tower layers (UGEVAL_PDL=0 disables)
  // Batches of up to sb_max_batch positions run the tower on conv3x3_sb.cu (128 x NT tiles over many SMs) instead of
  // conv3x3_v2.cu (256 x 256 CTA-pair tiles): the batch sizes the reference's own engine issues (1 .. 16).
  // UGEVAL_SB_MAX / ug_eval_set_small_batch_max; 0 disables.  sb_nt: tile width override (0 = from the grid size).
  // sb_pair: 1 = CTA pairs (cta_group::2: each SM stages half of the weights), 0 = single CTAs, -1 = chosen from the
  // batch (UGEVAL_SB_PAIR=0|1 forces).
  int sb_max_batch = 32;
  int sb_nt = 0;
  int sb_pair = -1;
  bool resid_split = true;  // carry the residual stream as hi + lo 16-bit pair (numerics)
  int value_transform = UG_TR_IDENTITY;
  bool use_graphs = true;
  int debug_stop = -1;  // >= 0: stop the tower after that layer (0 = stem, 2b+1 / 2b+2 = convs of block b)
  mutable std::string last_error;

  std::string input_name = "feature_input";
  std::vector<std::string> outputs = {"resnet/tower_0/policy_head/policy_predict",
                                      "resnet/tower_0/value_head/reward_predict"};
  std::string ckpt_prefix = "bootstrap-ckpt";
  bool graph_loaded = false;
  bool plan_ready = false;
  ug::TfGraph graph;
  ug::NetPlan plan;
  ug::Weights* weights = nullptr;  // swapped atomically by load_checkpoint

  cudaStream_t stream = nullptr;
  // workspace (allocated once the architecture is known)
  int ws_C = 0;
  ug::DevBuf d_pos, d_legal, d_planes, d_in, d_feat, d_policy, d_value, d_bytes;
  ug::DevBuf d_sat;    // one 64-bit counter: outputs clamped to fp16's range by the tower epilogues
  ug::DevBuf d_flags;  // [2][tiles] per-tile completion counters of the tower layers (tile-granular dependencies)
  ug::DevBuf d_policy64, d_value64;  // double images of the outputs for registered (pinned) caller buffers
  std::vector<std::pair<uintptr_t, size_t>> host_regs;  // caller buffers registered with ug_eval_register_host
  // residual stream ping-pong S[0]/S[1] as bf16 hi + bf16 lo (skip = hi + lo), block-internal tensor T (bf16)
  ug::DevBuf d_s_hi[2], d_s_lo[2], d_t;
  ug::DevBuf f_in, f_s[2], f_t;  // fp32 validation build
  CUtensorMap tm_in, tm_s[2], tm_t;
  // conv v2 maps over the same buffers: halo-tile loads, residual loads (hi/lo), staged stores (hi/lo)
  CUtensorMap v2_a_in, v2_a_s[2], v2_a_t, v2_res_hi[2], v2_res_lo[2], v2_out_hi[2], v2_out_lo[2], v2_out_t;
  void* h_pinned = nullptr;  // staging for host-buffer entry points
  // Device-side alias of h_pinned (equal to it under unified addressing), or null: small batches let the head kernel
  // write policy and value straight into the staging area instead of queueing two device-to-host copies behind it.
  uint8_t* d_pinned_alias = nullptr;
  size_t h_pinned_bytes = 0;

  // cuda graphs keyed by (kind, n, legal?, ctas, pointers)
  typedef std::tuple<int, int, const void*, const void*, const void*, const void*, int, int> GraphKey;
  std::map<GraphKey, cudaGraphExec_t> graphs;

  ~ug_eval();
  int fail(int code, const std::string& msg) const { last_error = msg; return code; }
  bool ensure_workspace(int C);
  bool is_registered(const void* p, size_t bytes) const;
  bool use_small_batch(int n) const;
  void drop_graphs();
  // source kinds for enqueue()
  enum Src { SRC_PACKED = 0, SRC_CHW = 1, SRC_HWC = 2 };
  // enqueue encoder + tower + heads on `s`; all pointers are device pointers
  // compact: with a legal mask, policy rows hold only the legal priors packed in ascending move order (leaf queue)
  int enqueue(Src kind, const void* d_src, int n, const uint32_t* d_legal, int vt, float* d_policy,
              float* d_value, cudaStream_t s, bool compact = false);
  int enqueue_raw(Src kind, const void* d_src, int n, const uint32_t* d_legal, int vt, float* d_policy,
                  float* d_value, cudaStream_t s, bool compact = false);
  int enqueue_encoder(Src kind, const void* d_src, int n, cudaStream_t s);
  int enqueue_tower(int n, cudaStream_t s, const void** final_act, const void** final_lo);
  int enqueue_tower_v2(int n, cudaStream_t s, const void** final_act, const void** final_lo);
  int enqueue_tower_sb(int n, cudaStream_t s, const void** final_act, const void** final_lo);
  int enqueue_heads(const void* act, const void* act_lo, int n, const uint32_t* d_legal, int vt, float* d_policy, float* d_value,
                    cudaStream_t s, bool compact = false);
};
