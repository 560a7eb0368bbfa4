// K5-K7 — policy and value heads.
//
// In the reference these are the tail of the TensorFlow graph run by Session::Run
// (funcapproximator/DlTFNetworkEvaluator.cc:295) plus the value_transform switch (.cc:300-310) and,
// optionally, UctSearch::UpdatePrior's legal-move renormalisation (search/UctSearch.cpp:1257-1270).
//
// Stage 1 (head_conv): both 1x1 convolutions (+ folded batch-norm + ReLU) in one pass over the final tower
//   activation: one warp per board point, 16-byte loads, warp-shuffle reduction.  HBM-bound: reads
//   361*C*2 bytes per position, writes 361*(pc+vc) floats.
// Stage 2 (head_fc): per CTA kPosPerCta positions: policy FC -> softmax (optionally masked and
//   renormalised), value FC -> ReLU -> FC -> tanh -> value_transform.  Weights stream once per CTA from L2.
#include <math_constants.h>

#include "kernels.h"
#include "numfmt.cuh"

namespace ug {

namespace {

constexpr int kPoints = UG_NUM_POINTS;
constexpr int kMoves = UG_NUM_MOVES;
constexpr int kMaxHeadCh = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T>
struct Load8;
template <>
struct Load8<__nv_bfloat16> {  // 16-bit storage: bf16 or fp16 selected by f16
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8], int f16) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = unpack2(w[i], f16);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
};
template <>
struct Load8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8], int) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};

// feat[point][h] = relu(sum_c act[point][c] * w[h][c] + b[h]);  HC = pc+vc <= 8, C % 8 == 0.
// One warp handles kPtsPerIter board points per iteration (all their 16-byte loads are issued before the
// arithmetic, so 2*kPtsPerIter loads per lane are in flight); a point's C channels are spread over the lanes.
constexpr int kPtsPerIter = 4;
template <typename T, int HC>
__global__ void __launch_bounds__(256) head_conv_kernel(const T* __restrict__ act, const T* __restrict__ act_lo,
                                                        const float* __restrict__ w,
                                                        const float* __restrict__ b, float* __restrict__ feat,
                                                        long total_points, int C, int f16) {
  extern __shared__ float sw[];  // [HC][C]
  for (int i = threadIdx.x; i < HC * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long warp_global = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long warp_stride = (long)gridDim.x * (blockDim.x >> 5);
  const int chunks = C >> 3;
  for (long pt0 = warp_global * kPtsPerIter; pt0 < total_points; pt0 += warp_stride * kPtsPerIter) {
    float acc[kPtsPerIter][HC];
#pragma unroll
    for (int u = 0; u < kPtsPerIter; ++u)
#pragma unroll
      for (int h = 0; h < HC; ++h) acc[u][h] = 0.f;
    for (int ch = lane; ch < chunks; ch += 32) {
      float x[kPtsPerIter][8];
#pragma unroll
      for (int u = 0; u < kPtsPerIter; ++u) {
        const long pt = min(pt0 + u, total_points - 1);  // clamped: the duplicate is not stored
        Load8<T>::load(act + pt * C + ch * 8, x[u], f16);
      }
      if (act_lo != nullptr) {
#pragma unroll
        for (int u = 0; u < kPtsPerIter; ++u) {
          const long pt = min(pt0 + u, total_points - 1);
          float xl[8];
          Load8<T>::load(act_lo + pt * C + ch * 8, xl, f16);
#pragma unroll
          for (int i = 0; i < 8; ++i) x[u][i] += xl[i];
        }
      }
#pragma unroll
      for (int h = 0; h < HC; ++h) {
        const float4 w0 = *reinterpret_cast<const float4*>(sw + h * C + ch * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(sw + h * C + ch * 8 + 4);
#pragma unroll
        for (int u = 0; u < kPtsPerIter; ++u)
          acc[u][h] += x[u][0] * w0.x + x[u][1] * w0.y + x[u][2] * w0.z + x[u][3] * w0.w + x[u][4] * w1.x +
                       x[u][5] * w1.y + x[u][6] * w1.z + x[u][7] * w1.w;
      }
    }
#pragma unroll
    for (int u = 0; u < kPtsPerIter; ++u)
#pragma unroll
      for (int h = 0; h < HC; ++h) acc[u][h] = warp_sum(acc[u][h]);
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < kPtsPerIter; ++u) {
        if (pt0 + u < total_points) {
#pragma unroll
          for (int h = 0; h < HC; ++h) feat[(pt0 + u) * HC + h] = fmaxf(acc[u][h] + __ldg(b + h), 0.f);
        }
      }
    }
  }
}

constexpr int kPosPerCta = 4;
constexpr int kPolThreads = 384;                 // policy FC: one output column per thread (362 used) ...
constexpr int kPolSplit = 2;                     // ... and the K range split over 2 thread groups
constexpr int kValThreads = 256;                 // value FC1: one hidden unit per thread (strided when H > 256)
constexpr int kFcThreads = kPolSplit * kPolThreads + kValThreads;  // 1024
constexpr int kFcUnroll = 16;                    // independent weight loads in flight per thread

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();  // scratch reuse
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float r = (lane < nw) ? scratch[lane] : (is_max ? -CUDART_INF_F : 0.f);
  r = is_max ? warp_max(r) : warp_sum(r);
  return r;
}

// acc[p] += sum_i f[p][i] * w[i*stride]  for kPosPerCta positions; weights stream from L2 (coalesced across the
// threads of a warp: consecutive threads own consecutive columns), kFcUnroll loads issued before their FMAs.
// The features are read from shared memory four at a time (one 16-byte broadcast load per position and group):
// f and fstride must be multiples of 4 floats.
__device__ __forceinline__ void fc_column(const float* __restrict__ wcol, int stride, const float* f, int fstride,
                                          int len, float (&acc)[kPosPerCta]) {
  int i = 0;
  for (; i + kFcUnroll <= len; i += kFcUnroll) {
    float wv[kFcUnroll];
#pragma unroll
    for (int u = 0; u < kFcUnroll; ++u) wv[u] = __ldg(wcol + (size_t)(i + u) * stride);
#pragma unroll
    for (int g = 0; g < kFcUnroll / 4; ++g)
#pragma unroll
      for (int p = 0; p < kPosPerCta; ++p) {
        const float4 x = *reinterpret_cast<const float4*>(f + p * fstride + i + 4 * g);
        acc[p] = fmaf(x.x, wv[4 * g], acc[p]);
        acc[p] = fmaf(x.y, wv[4 * g + 1], acc[p]);
        acc[p] = fmaf(x.z, wv[4 * g + 2], acc[p]);
        acc[p] = fmaf(x.w, wv[4 * g + 3], acc[p]);
      }
  }
  for (; i < len; ++i) {
    const float wv = __ldg(wcol + (size_t)i * stride);
#pragma unroll
    for (int p = 0; p < kPosPerCta; ++p) acc[p] = fmaf(f[p * fstride + i], wv, acc[p]);
  }
}

// feat layout per position: [361][pc+vc] (policy channels first).  Policy FC input index = point*pc + c,
// value FC input index = point*vc + c (the NHWC flatten TensorFlow's Reshape produces).
// Threads [0,768) own the policy columns (two groups, each half of the K range), threads [768,1024) the value
// hidden units: both FCs run concurrently and every thread walks ~361 weight rows.
__global__ void __launch_bounds__(kFcThreads) head_fc_kernel(const float* __restrict__ feat, HeadWeights hw,
                                                             const uint32_t* __restrict__ legal, int vt,
                                                             float* __restrict__ policy,
                                                             float* __restrict__ value, int n) {
  extern __shared__ float sm[];
  const int pc = hw.pc, vc = hw.vc, hc = pc + vc, H = hw.H;
  const int pin = kPoints * pc, vin = kPoints * vc;
  const int pstr = (pin + 3) & ~3, vstr = (vin + 3) & ~3;  // row strides padded for 16-byte reads
  float* fp = sm;                          // [kPosPerCta][pstr]
  float* fv = fp + kPosPerCta * pstr;      // [kPosPerCta][vstr]
  float* hid = fv + kPosPerCta * vstr;     // [kPosPerCta][H]
  float* scratch = hid + kPosPerCta * H;   // [32]
  float* part = scratch + 32;              // [kPosPerCta][kPolThreads] partial logits of the second K half
  const int p0 = blockIdx.x * kPosPerCta;
  const int np = min(kPosPerCta, n - p0);

  for (int i = threadIdx.x; i < kPosPerCta * kPoints * hc; i += blockDim.x) {
    const int p = i / (kPoints * hc);
    const int r = i - p * (kPoints * hc);
    const int point = r / hc, c = r - point * hc;
    const float v = (p < np) ? __ldg(feat + (size_t)(p0 + p) * kPoints * hc + r) : 0.f;
    if (c < pc) fp[p * pstr + point * pc + c] = v;
    else fv[p * vstr + point * vc + (c - pc)] = v;
  }
  __syncthreads();

  const int tid = threadIdx.x;
  const int j = tid < kPolSplit * kPolThreads ? tid % kPolThreads : kMoves;  // policy column (>= 362: none)
  const int khalf = tid / kPolThreads;                                       // 0/1 for policy threads
  float logit[kPosPerCta];
#pragma unroll
  for (int p = 0; p < kPosPerCta; ++p) logit[p] = 0.f;
  if (tid < kPolSplit * kPolThreads) {
    // ---- policy: logits[j] = b[j] + sum_i f[i] * W[i][j]; group 0 sums i in [0, pin/2), group 1 the rest
    if (j < kMoves) {
      const int mid = (pin / 2) & ~3;  // multiple of 4: both halves start 16-byte aligned
      const int k0 = khalf == 0 ? 0 : mid;
      const int k1 = khalf == 0 ? mid : pin;
      fc_column(hw.pfc_w + (size_t)k0 * kMoves + j, kMoves, fp + k0, pstr, k1 - k0, logit);
      if (khalf == 1) {
#pragma unroll
        for (int p = 0; p < kPosPerCta; ++p) part[p * kPolThreads + j] = logit[p];
      }
    }
  } else {
    // ---- value: hidden = relu(W1^T f + b1)
    for (int h = tid - kPolSplit * kPolThreads; h < H; h += kValThreads) {
      float acc[kPosPerCta];
#pragma unroll
      for (int p = 0; p < kPosPerCta; ++p) acc[p] = 0.f;
      fc_column(hw.v1_w + h, H, fv, vstr, vin, acc);
      const float b1 = __ldg(hw.v1_b + h);
#pragma unroll
      for (int p = 0; p < kPosPerCta; ++p) hid[p * H + h] = fmaxf(acc[p] + b1, 0.f);
    }
  }
  __syncthreads();
  const bool pol = tid < kPolThreads && j < kMoves;  // the threads that own a finished logit
  if (pol) {
    const float bj = __ldg(hw.pfc_b + j);
#pragma unroll
    for (int p = 0; p < kPosPerCta; ++p) logit[p] += part[p * kPolThreads + j] + bj;
  }
#pragma unroll
  for (int p = 0; p < kPosPerCta; ++p) {
    if (p >= np) break;  // np is block-uniform
    const float mx = block_reduce(pol ? logit[p] : -CUDART_INF_F, true, scratch);
    float prob = logit[p];
    if (hw.policy_softmax) {
      const float e = pol ? expf(logit[p] - mx) : 0.f;
      const float sum = block_reduce(e, false, scratch);
      prob = e / sum;
    }
    if (legal != nullptr) {
      // UpdatePrior: prior = p[idx] / sum over legal children of p[idx]; illegal moves get 0
      const uint32_t* lm = legal + (size_t)(p0 + p) * UG_MASK_WORDS;
      const bool ok = pol && ((__ldg(lm + (j >> 5)) >> (j & 31)) & 1u);
      const float lsum = block_reduce(ok ? prob : 0.f, false, scratch);
      prob = ok ? prob / lsum : 0.f;
    }
    if (pol) policy[(size_t)(p0 + p) * kMoves + j] = prob;
  }

  // ---- v = tanh(w2 . hidden + b2); hid was completed before the first block_reduce's barrier
  __syncthreads();
  for (int p = 0; p < np; ++p) {
    float part = 0.f;
    for (int h = threadIdx.x; h < H; h += blockDim.x) part = fmaf(hid[p * H + h], __ldg(hw.v2_w + h), part);
    const float tot = block_reduce(part, false, scratch);
    if (threadIdx.x == 0) {
      float v = tot + __ldg(hw.v2_b);
      if (hw.value_tanh) v = tanhf(v);
      if (vt == UG_TR_FLIP_SIGN) v = -v;
      else if (vt == UG_TR_FLIP_PROB) v = 1.f - v;
      value[p0 + p] = v;
    }
  }
}

template <typename T>
void head_conv_dispatch(const T* act, const T* act_lo, int f16, const HeadWeights& w, float* feat, int n,
                        cudaStream_t s) {
  const long total = (long)n * kPoints;
  const int hc = w.pc + w.vc;
  int grid = (int)((total + 8 * kPtsPerIter - 1) / (8 * kPtsPerIter));
  if (grid > 148 * 8) grid = 148 * 8;
  const size_t smem = (size_t)hc * w.C * sizeof(float);
  switch (hc) {
#define UG_HC_CASE(N)                                                                                   \
  case N:                                                                                               \
    head_conv_kernel<T, N><<<grid, 256, smem, s>>>(act, act_lo, w.conv_w, w.conv_b, feat, total, w.C, f16);          \
    break;
    UG_HC_CASE(1) UG_HC_CASE(2) UG_HC_CASE(3) UG_HC_CASE(4) UG_HC_CASE(5) UG_HC_CASE(6) UG_HC_CASE(7)
    UG_HC_CASE(8)
#undef UG_HC_CASE
    default: break;
  }
}

}  // namespace

void launch_head_conv(const void* act, const void* act_lo, int act_fmt, const HeadWeights& w, float* feat, int n,
                      cudaStream_t s) {
  if (n <= 0) return;
  if (act_fmt == 1) head_conv_dispatch<float>(static_cast<const float*>(act), nullptr, 0, w, feat, n, s);
  else head_conv_dispatch<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(act),
                                         static_cast<const __nv_bfloat16*>(act_lo), act_fmt == 2, w, feat, n, s);
}

void launch_head_fc(const float* feat, const HeadWeights& w, const uint32_t* legal, int vt, float* policy,
                    float* value, int n, cudaStream_t s) {
  if (n <= 0) return;
  const size_t smem =
      sizeof(float) * ((size_t)kPosPerCta * (((kPoints * w.pc + 3) & ~3) + ((kPoints * w.vc + 3) & ~3) + w.H + kPolThreads) + 32);
  static size_t smem_set[64] = {};  // per device (function attributes are per device)
  int dev = 0;
  cudaGetDevice(&dev);
  size_t& set = smem_set[dev & 63];
  if (smem > 48 * 1024 && smem > set) {
    cudaFuncSetAttribute(head_fc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    set = smem;
  }
  const int grid = (n + kPosPerCta - 1) / kPosPerCta;
  head_fc_kernel<<<grid, kFcThreads, smem, s>>>(feat, w, legal, vt, policy, value, n);
}

}  // namespace ug
