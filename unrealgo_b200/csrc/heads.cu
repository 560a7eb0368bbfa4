// K5-K7 — policy and value heads.
//
// In the reference these are the tail of the TensorFlow graph run by Session::Run
// (funcapproximator/DlTFNetworkEvaluator.cc:295) plus the value_transform switch (.cc:300-310) and,
// optionally, UctSearch::UpdatePrior's legal-move renormalisation (search/UctSearch.cpp:1257-1270).
//
// Stage 1 (head_conv): both 1x1 convolutions (+ folded batch-norm + ReLU) in one pass over the final tower
//   activation: one warp per board point, 16-byte loads, warp-shuffle reduction.  HBM-bound: reads
//   361*C*2 bytes per position, writes 361*(pc+vc) floats.
// Stage 2 (head_fc): policy FC -> softmax (optionally masked and renormalised), value FC -> ReLU -> FC -> tanh ->
//   value_transform.  Default: head_fc_cluster_kernel (4-CTA cluster per 8 positions, reductions through distributed
//   shared memory); head_fc_kernel (one CTA per 4 positions) is the single-CTA variant.
#include <math_constants.h>

#include <cstdlib>

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
// compact priors: position of move j among the set bits of the legal mask (ascending move order)
__device__ __forceinline__ int legal_rank(const uint32_t* __restrict__ lm, int j) {
  int r = 0;
  const int w = j >> 5;
  for (int i = 0; i < w; ++i) r += __popc(__ldg(lm + i));
  return r + __popc(__ldg(lm + w) & ((1u << (j & 31)) - 1u));
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
                                                             float* __restrict__ value, int n, int compact) {
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
      if (compact) {  // only the legal priors, packed in ascending move order
        if (ok) policy[(size_t)(p0 + p) * kMoves + legal_rank(lm, j)] = prob;
        continue;
      }
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

// ------------------------------------------------------------------------------------------------
// Clustered variant of stage 2: a cluster of 4 CTAs shares 8 positions.  Every CTA owns a quarter of the policy
// columns and of the value hidden units (so the FC weights are read once per cluster, a quarter per CTA), the K
// range of every column is split over several threads, and the softmax / legal-mask / value reductions run across
// the cluster through distributed shared memory.  Same arithmetic as head_fc_kernel up to summation order.
constexpr int kClu = 4;
constexpr int kCluPos = 8;                     // positions per cluster
constexpr int kCluThreads = 512;
constexpr int kPolKSplit = 4, kValKSplit = 2;
constexpr int kColsPerCta = (kMoves + kClu - 1) / kClu;  // 91

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store v at the same shared-memory offset in CTA `cta` of this cluster
__device__ __forceinline__ void dsmem_store(float* local, uint32_t cta, float v) {
  uint32_t la = static_cast<uint32_t>(__cvta_generic_to_shared(local)), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(cta));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

// acc[p] += sum_{i<len} f[p*fstride + i] * w[i*stride], 8 positions.  Software-pipelined: the kFcDepth weight loads of
// the next step are in flight while the current ones are consumed; features are 16-byte broadcast reads.  A thread walks
// ~181 weight rows and every step is one L2 round trip, so the number of steps sets the kernel's latency: the remainder
// (len % kFcDepth rows) is one more batched step with predicated loads, not a row-at-a-time loop (round 1: 11 pipelined
// steps + 5 single rows = 16 round trips; now 7 + 1).  The remainder step reads features up to 3 floats past `len`
// (multiplied by zero weights): the caller keeps those finite (rows are padded with zeros).
constexpr int kFcDepth = 24;
__device__ __forceinline__ void fc_column8(const float* __restrict__ wcol, int stride, const float* f, int fstride,
                                           int len, float (&acc)[kCluPos]) {
  const int full = len - len % kFcDepth;
  float cur[kFcDepth], nxt[kFcDepth];
  if (full > 0) {
#pragma unroll
    for (int u = 0; u < kFcDepth; ++u) cur[u] = __ldg(wcol + (size_t)u * stride);
  }
  const int rem = len - full;
  for (int i = 0; i < full; i += kFcDepth) {
    if (i + kFcDepth < full) {
#pragma unroll
      for (int u = 0; u < kFcDepth; ++u) nxt[u] = __ldg(wcol + (size_t)(i + kFcDepth + u) * stride);
    } else {
#pragma unroll
      for (int u = 0; u < kFcDepth; ++u) nxt[u] = u < rem ? __ldg(wcol + (size_t)(full + u) * stride) : 0.f;
    }
#pragma unroll
    for (int g = 0; g < kFcDepth / 4; ++g)
#pragma unroll
      for (int p = 0; p < kCluPos; ++p) {
        const float4 x = *reinterpret_cast<const float4*>(f + p * fstride + i + 4 * g);
        acc[p] = fmaf(x.x, cur[4 * g], acc[p]);
        acc[p] = fmaf(x.y, cur[4 * g + 1], acc[p]);
        acc[p] = fmaf(x.z, cur[4 * g + 2], acc[p]);
        acc[p] = fmaf(x.w, cur[4 * g + 3], acc[p]);
      }
#pragma unroll
    for (int u = 0; u < kFcDepth; ++u) cur[u] = nxt[u];
  }
  if (full == 0) {
#pragma unroll
    for (int u = 0; u < kFcDepth; ++u) cur[u] = u < rem ? __ldg(wcol + (size_t)u * stride) : 0.f;
  }
  if (rem > 0) {
#pragma unroll
    for (int g = 0; g < kFcDepth / 4; ++g) {
      if (4 * g < rem) {
#pragma unroll
        for (int p = 0; p < kCluPos; ++p) {
          const float4 x = *reinterpret_cast<const float4*>(f + p * fstride + full + 4 * g);
          acc[p] = fmaf(x.x, cur[4 * g], acc[p]);
          acc[p] = fmaf(x.y, cur[4 * g + 1], acc[p]);
          acc[p] = fmaf(x.z, cur[4 * g + 2], acc[p]);
          acc[p] = fmaf(x.w, cur[4 * g + 3], acc[p]);
        }
      }
    }
  }
}

__global__ void __cluster_dims__(kClu, 1, 1) __launch_bounds__(kCluThreads)
head_fc_cluster_kernel(const float* __restrict__ feat, HeadWeights hw, const uint32_t* __restrict__ legal, int vt,
                       float* __restrict__ policy, float* __restrict__ value, int n, int compact) {
  extern __shared__ float sm[];
  const int pc = hw.pc, vc = hw.vc, hc = pc + vc, H = hw.H;
  const int pin = kPoints * pc, vin = kPoints * vc;
  const int pstr = (pin + 3) & ~3, vstr = (vin + 3) & ~3;
  const int hper = (H + kClu - 1) / kClu;       // hidden units of this CTA
  float* fp = sm;                               // [8][pstr]
  float* fv = fp + kCluPos * pstr;              // [8][vstr]
  float* part = fv + kCluPos * vstr;            // [kPolKSplit][8][96] partial logits
  float* vpart = part + kPolKSplit * kCluPos * 96;  // [kValKSplit][8][hper] partial hidden units
  float* logit = vpart + kValKSplit * kCluPos * hper;  // [8][96]
  float* hid = logit + kCluPos * 96;            // [8][hper]
  float* xch = hid + kCluPos * hper;            // [4 kinds][kClu][8]: max, sum, legal sum, value partial (DSMEM targets)
  const uint32_t rank = cluster_rank();
  const int p0 = (blockIdx.x / kClu) * kCluPos;
  const int np = min(kCluPos, n - p0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = (int)rank * kColsPerCta;
  const int ncol = max(0, min(kColsPerCta, kMoves - c0));
  const int h0 = (int)rank * hper;
  const int nh = max(0, min(hper, H - h0));

  // row padding [pin, pstr) / [vin, vstr): read (times zero weights) by fc_column8's remainder step, must be finite
  if (tid < kCluPos) {
    for (int i = pin; i < pstr; ++i) fp[tid * pstr + i] = 0.f;
    for (int i = vin; i < vstr; ++i) fv[tid * vstr + i] = 0.f;
    if (tid < 4) part[tid] = 0.f;  // the last value row's remainder step may look up to 2 floats past fv
  }
  {  // features of the cluster's 8 positions -> shared memory, 4 independent loads in flight per thread
    const int total = kCluPos * kPoints * hc;
    for (int base = 0; base < total; base += 4 * (int)blockDim.x) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * (int)blockDim.x + tid;
        const int p = i / (kPoints * hc);
        v[u] = (i < total && p < np) ? __ldg(feat + (size_t)p0 * kPoints * hc + i) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * (int)blockDim.x + tid;
        if (i < total) {
          const int p = i / (kPoints * hc);
          const int r = i - p * (kPoints * hc);
          const int point = r / hc, c = r - point * hc;
          if (c < pc) fp[p * pstr + point * pc + c] = v[u];
          else fv[p * vstr + point * vc + (c - pc)] = v[u];
        }
      }
    }
  }
  __syncthreads();

  // ---- partial dot products, both FCs at once: policy threads [0, 4*91) (K split in 4), value threads
  //      [384, 384 + 2*hper) (K split in 2); every thread walks ~181 weight rows for 8 positions
  if (tid < kPolKSplit * kColsPerCta) {
    const int col = tid % kColsPerCta, ks = tid / kColsPerCta;
    float acc[kCluPos];
#pragma unroll
    for (int p = 0; p < kCluPos; ++p) acc[p] = 0.f;
    const int klen = (((pin + kPolKSplit - 1) / kPolKSplit) + 3) & ~3;
    const int k0 = ks * klen, k1 = min(pin, k0 + klen);
    if (col < ncol && k0 < k1) fc_column8(hw.pfc_w + (size_t)k0 * kMoves + c0 + col, kMoves, fp + k0, pstr, k1 - k0, acc);
#pragma unroll
    for (int p = 0; p < kCluPos; ++p) part[(ks * kCluPos + p) * 96 + col] = acc[p];
  } else if (tid >= 384) {
    for (int t = tid - 384; t < kValKSplit * hper; t += kCluThreads - 384) {
      const int u = t % hper, ks = t / hper;
      float acc[kCluPos];
#pragma unroll
      for (int p = 0; p < kCluPos; ++p) acc[p] = 0.f;
      const int klen = (((vin + kValKSplit - 1) / kValKSplit) + 3) & ~3;
      const int k0 = ks * klen, k1 = min(vin, k0 + klen);
      if (u < nh && k0 < k1) fc_column8(hw.v1_w + (size_t)k0 * H + h0 + u, H, fv + k0, vstr, k1 - k0, acc);
#pragma unroll
      for (int p = 0; p < kCluPos; ++p) vpart[(ks * kCluPos + p) * hper + u] = acc[p];
    }
  }
  __syncthreads();
  for (int i = tid; i < kCluPos * kColsPerCta; i += blockDim.x) {
    const int p = i / kColsPerCta, col = i - p * kColsPerCta;
    float sacc = 0.f;
#pragma unroll
    for (int ks = 0; ks < kPolKSplit; ++ks) sacc += part[(ks * kCluPos + p) * 96 + col];
    logit[p * 96 + col] = col < ncol ? sacc + __ldg(hw.pfc_b + c0 + col) : -CUDART_INF_F;
  }
  for (int i = tid; i < kCluPos * hper; i += blockDim.x) {
    const int p = i / hper, u = i - p * hper;
    float sacc = 0.f;
#pragma unroll
    for (int ks = 0; ks < kValKSplit; ++ks) sacc += vpart[(ks * kCluPos + p) * hper + u];
    hid[p * hper + u] = u < nh ? fmaxf(sacc + __ldg(hw.v1_b + h0 + u), 0.f) : 0.f;
  }
  __syncthreads();

  // ---- per position (warp p): local max of the logits, local value partial; publish to every CTA of the cluster
  float* x_max = xch;                       // [kClu][8]
  float* x_sum = xch + kClu * kCluPos;
  float* x_leg = xch + 2 * kClu * kCluPos;
  float* x_val = xch + 3 * kClu * kCluPos;
  if (warp < kCluPos) {
    const int p = warp;
    float m = -CUDART_INF_F, vpart = 0.f;
    for (int c = lane; c < ncol; c += 32) m = fmaxf(m, logit[p * 96 + c]);
    for (int u = lane; u < nh; u += 32) vpart = fmaf(hid[p * hper + u], __ldg(hw.v2_w + h0 + u), vpart);
    m = warp_max(m);
    vpart = warp_sum(vpart);
    if (lane < kClu) {
      dsmem_store(x_max + rank * kCluPos + p, lane, m);
      dsmem_store(x_val + rank * kCluPos + p, lane, vpart);
    }
  }
  cluster_barrier();
  float e_keep[3] = {0.f, 0.f, 0.f};  // this lane's exponentials (ncol <= 96 = 3 per lane)
  if (warp < kCluPos) {
    const int p = warp;
    float gmax = -CUDART_INF_F;
#pragma unroll
    for (int r = 0; r < kClu; ++r) gmax = fmaxf(gmax, x_max[r * kCluPos + p]);
    float lsum = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int c = lane + 32 * k;
      if (c < ncol) {
        e_keep[k] = hw.policy_softmax ? expf(logit[p * 96 + c] - gmax) : logit[p * 96 + c];
        lsum += e_keep[k];
      }
    }
    lsum = warp_sum(lsum);
    if (lane < kClu) dsmem_store(x_sum + rank * kCluPos + p, lane, lsum);
    if (rank == 0 && lane == 0 && p < np) {  // value: tanh(w2 . hidden + b2), value_transform
      float v = __ldg(hw.v2_b);
#pragma unroll
      for (int r = 0; r < kClu; ++r) v += x_val[r * kCluPos + p];
      if (hw.value_tanh) v = tanhf(v);
      if (vt == UG_TR_FLIP_SIGN) v = -v;
      else if (vt == UG_TR_FLIP_PROB) v = 1.f - v;
      value[p0 + p] = v;
    }
  }
  cluster_barrier();
  if (warp < kCluPos) {
    const int p = warp;
    float tot = 0.f;
#pragma unroll
    for (int r = 0; r < kClu; ++r) tot += x_sum[r * kCluPos + p];
    const float inv = hw.policy_softmax ? 1.f / tot : 1.f;
    float lleg = 0.f;
    bool ok[3] = {true, true, true};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int c = lane + 32 * k;
      e_keep[k] *= inv;
      if (legal != nullptr && c < ncol && p < np) {
        const int j = c0 + c;
        ok[k] = (__ldg(legal + (size_t)(p0 + p) * UG_MASK_WORDS + (j >> 5)) >> (j & 31)) & 1u;
        if (ok[k]) lleg += e_keep[k];
      }
    }
    if (legal != nullptr) {  // UpdatePrior: prior = p[idx] / sum over legal children; illegal moves get 0
      lleg = warp_sum(lleg);
      if (lane < kClu) dsmem_store(x_leg + rank * kCluPos + p, lane, lleg);
    }
    if (legal == nullptr && p < np) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int c = lane + 32 * k;
        if (c < ncol) policy[(size_t)(p0 + p) * kMoves + c0 + c] = e_keep[k];
      }
    }
    if (legal != nullptr) {  // illegal moves get 0; the legal ones are renormalised after one more exchange
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (!ok[k]) e_keep[k] = 0.f;
    }
  }
  if (legal != nullptr) {
    cluster_barrier();
    if (warp < kCluPos && warp < np) {
      const int p = warp;
      float lt = 0.f;
#pragma unroll
      for (int r = 0; r < kClu; ++r) lt += x_leg[r * kCluPos + p];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int c = lane + 32 * k;
        if (c < ncol) {
          if (!compact) policy[(size_t)(p0 + p) * kMoves + c0 + c] = e_keep[k] / lt;
          else if (e_keep[k] != 0.f || ((__ldg(legal + (size_t)(p0 + p) * UG_MASK_WORDS + ((c0 + c) >> 5)) >> ((c0 + c) & 31)) & 1u))
            // legal moves only (an exponential can underflow to 0: ask the mask, not the value), packed in ascending order
            policy[(size_t)(p0 + p) * kMoves + legal_rank(legal + (size_t)(p0 + p) * UG_MASK_WORDS, c0 + c)] = e_keep[k] / lt;
        }
      }
    }
  }
  // no CTA may exit while a peer can still write into its shared memory
  cluster_barrier();
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

static bool use_cluster_fc() {
  static const bool on = [] {
    const char* e = getenv("UGEVAL_HEAD_FC_CLUSTER");
    return !(e && e[0] == '0');
  }();
  return on;
}

void launch_head_fc(const float* feat, const HeadWeights& w, const uint32_t* legal, int vt, float* policy,
                    float* value, int n, cudaStream_t s, int compact) {
  if (n <= 0) return;
  if (legal == nullptr) compact = 0;
  const int hper = (w.H + kClu - 1) / kClu;
  if (use_cluster_fc() && hper <= 96 && w.pc + w.vc <= kMaxHeadCh) {
    const int pstr = (kPoints * w.pc + 3) & ~3, vstr = (kPoints * w.vc + 3) & ~3;
    const size_t smem = sizeof(float) * ((size_t)kCluPos * (pstr + vstr) + (size_t)kPolKSplit * kCluPos * 96 +
                                         (size_t)kValKSplit * kCluPos * hper + (size_t)kCluPos * 96 +
                                         (size_t)kCluPos * hper + 4 * kClu * kCluPos);
    static size_t set_c[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && smem > set_c[dev & 63]) {
      cudaFuncSetAttribute(head_fc_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      set_c[dev & 63] = smem;
    }
    const int clusters = (n + kCluPos - 1) / kCluPos;
    head_fc_cluster_kernel<<<clusters * kClu, kCluThreads, smem, s>>>(feat, w, legal, vt, policy, value, n, compact);
    return;
  }
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
  head_fc_kernel<<<grid, kFcThreads, smem, s>>>(feat, w, legal, vt, policy, value, n, compact);
}

}  // namespace ug
