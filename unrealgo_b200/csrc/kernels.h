// Launchers of the non-GEMM kernels: encoder (K1/K2), heads (K5-K7), fp32 validation tower.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ugeval.h"

namespace ug {

constexpr int kCinPad = 32;  // stem input channels as stored (17 planes + zero padding)

// K1+K2: packed leaves -> NHWC bf16 [n][361][kCinPad]   (CollectFeatures + TransformFeature in one pass)
void launch_encode_packed_nhwc(const ug_packed_pos* pos, __nv_bfloat16* out, int n, int f16, cudaStream_t s);
// K1 test hook: packed leaves -> int8 [n][17][19][19]
void launch_encode_packed_chw(const ug_packed_pos* pos, int8_t* out, int n, cudaStream_t s);
// K2: reference planes int8 [n][17][19][19] (hwc=0) or [n][19][19][17] (hwc=1) -> NHWC bf16 [n][361][kCinPad]
void launch_planes_to_nhwc(const int8_t* planes, int hwc, __nv_bfloat16* out, int n, int f16, cudaStream_t s);
// readback helper: NHWC bf16 [n][361][kCinPad] -> int8 [n][361][17]
void launch_nhwc_to_bytes(const __nv_bfloat16* in, int8_t* out, int n, int f16, cudaStream_t s);
// fp32 validation input: NHWC bf16 [n][361][kCinPad] -> fp32 [n][361][17]
void launch_nhwc_to_f32(const __nv_bfloat16* in, float* out, int n, int f16, cudaStream_t s);

struct HeadWeights {
  // 1x1 convs with batch-norm folded: rows [0,pc) policy, [pc,pc+vc) value
  const float* conv_w;  // [pc+vc][C]
  const float* conv_b;  // [pc+vc]
  int pc, vc, C;
  const float* pfc_w;  // [361*pc][362]  (TF dense kernel, [in][out])
  const float* pfc_b;  // [362]
  const float* v1_w;   // [361*vc][H]
  const float* v1_b;   // [H]
  const float* v2_w;   // [H]
  const float* v2_b;   // [1]
  int H;
  int policy_softmax;  // graph output is Softmax(logits) (else raw logits)
  int value_tanh;      // graph output is Tanh(fc2)
};

// K5/K6 stage 1: features[n][361][pc+vc] = relu(conv1x1(act)); act is bf16 (+ optional bf16 low part) or fp32
// NHWC [n][361][C]
void launch_head_conv(const void* act, const void* act_lo, int act_fmt /*0 bf16, 1 fp32, 2 fp16*/, const HeadWeights& w, float* feat, int n, cudaStream_t s);
// K5-K7 stage 2: FC + softmax (optionally renormalised over a legal mask) and FC-ReLU-FC-tanh-value_transform
// compact (with a legal mask only): row p holds just its legal priors, packed in ascending move order
void launch_head_fc(const float* feat, const HeadWeights& w, const uint32_t* legal, int vt, float* policy,
                    float* value, int n, cudaStream_t s, int compact = 0);

// fp32 validation tower (SIMT, direct convolution): in/out fp32 NHWC, weights HWIO [3][3][cin][cout]
void launch_conv3x3_f32(const float* in, const float* w_hwio, const float* bias, const float* res, float* out,
                        int n, int cin, int cout, int relu, cudaStream_t s);
// *slot = max(*slot, max |p[i]|) (slot zero-initialised by the caller; fp32 values, non-negative result)
void launch_max_abs_f32(const float* p, size_t count, float* slot, cudaStream_t s);
// bf16 <-> fp32 images for debug readback
// out64[i] = (double)in[i] for the policy; value: widened, then value_transform applied in double exactly as the
// reference does on the host (DlTensorUtil::GetValue + DlTFNetworkEvaluator.cc:300-310)
void launch_widen_outputs(const float* policy, const float* value, double* policy64, double* value64, int n, int vt,
                          cudaStream_t s);
void launch_bf16_to_f32(const __nv_bfloat16* in, float* out, size_t count, int f16, cudaStream_t s);

}  // namespace ug
