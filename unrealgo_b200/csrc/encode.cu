// K1/K2 — feature-plane encoder.
//
// Reference: GoUctGlobalSearchState::CollectFeatures (gouct/GoUctGlobalSearch.h:160-206) builds
// char[17][19][19] on the search thread by undoing/replaying up to 8 moves, then
// DlTFNetworkEvaluator::TransformFeature (funcapproximator/DlTFNetworkEvaluator.cc:366-389) re-lays it out
// element by element into bool[n,19,19,17].  Here the search thread ships 784-byte packed stone masks
// (ug_packed_pos) and one kernel writes the tower's NHWC bf16 input directly:
//     plane 2k   = stones of the side to move on the board k moves ago
//     plane 2k+1 = stones of the opponent           "
//     plane 16   = 1 - to_play  (1 if black to move)
// HBM-bound integer work: 784 B in, 361*32*2 B out per position; every thread produces one 16-byte,
// fully coalesced store (8 channels of one point).
#include "kernels.h"
#include "numfmt.cuh"

namespace ug {

namespace {

constexpr int kPoints = UG_NUM_POINTS;
constexpr int kChunksPerPoint = kCinPad / 8;                  // 16-byte chunks per point
constexpr int kChunksPerPos = kPoints * kChunksPerPoint;      // 1444

__global__ void __launch_bounds__(256) encode_packed_nhwc_kernel(const ug_packed_pos* __restrict__ pos,
                                                                 __nv_bfloat16* __restrict__ out, int n,
                                                                 uint32_t kOneBf16) {
  __shared__ uint32_t masks[2][UG_HISTORY][UG_MASK_WORDS];  // [0] = side to move, [1] = opponent
  for (int p = blockIdx.x; p < n; p += gridDim.x) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(pos + p);
    // every thread fetches the to-play word itself (one broadcast load) together with its mask word: one global-memory
    // round trip before the stores instead of two (the kernel is latency-bound: 784 bytes in per position)
    const uint32_t tp = __ldg(src + 2 * UG_HISTORY * UG_MASK_WORDS) & 1u;
    if (threadIdx.x < 2 * UG_HISTORY * UG_MASK_WORDS) {
      // source order: black[8][12] then white[8][12]; colour c goes to slot (c ^ to_play)
      const int colour = threadIdx.x / (UG_HISTORY * UG_MASK_WORDS);
      const int rest = threadIdx.x - colour * (UG_HISTORY * UG_MASK_WORDS);
      (&masks[colour ^ tp][0][0])[rest] = __ldg(src + threadIdx.x);
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)p * kPoints * kCinPad);
    for (int idx = threadIdx.x; idx < kChunksPerPos; idx += blockDim.x) {
      const int point = idx / kChunksPerPoint;
      const int q = idx - point * kChunksPerPoint;
      const int word = point >> 5, bit = point & 31;
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (q < 2) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = 4 * q + e;  // history step; channels 2k (to-move) and 2k+1 (opponent)
          const uint32_t lo = (masks[0][k][word] >> bit) & 1u;
          const uint32_t hi = (masks[1][k][word] >> bit) & 1u;
          w[e] = (lo ? kOneBf16 : 0u) | ((hi ? kOneBf16 : 0u) << 16);
        }
      } else if (q == 2) {
        w[0] = tp ? 0u : kOneBf16;  // channel 16 = 1 - to_play
      }
      dst[idx] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) encode_packed_chw_kernel(const ug_packed_pos* __restrict__ pos,
                                                                int8_t* __restrict__ out, int n) {
  __shared__ uint32_t masks[2][UG_HISTORY][UG_MASK_WORDS];
  for (int p = blockIdx.x; p < n; p += gridDim.x) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(pos + p);
    const uint32_t tp = __ldg(src + 2 * UG_HISTORY * UG_MASK_WORDS) & 1u;
    if (threadIdx.x < 2 * UG_HISTORY * UG_MASK_WORDS) {
      const int colour = threadIdx.x / (UG_HISTORY * UG_MASK_WORDS);
      const int rest = threadIdx.x - colour * (UG_HISTORY * UG_MASK_WORDS);
      (&masks[colour ^ tp][0][0])[rest] = __ldg(src + threadIdx.x);
    }
    __syncthreads();
    int8_t* dst = out + (size_t)p * UG_NUM_PLANES * kPoints;
    for (int idx = threadIdx.x; idx < UG_NUM_PLANES * kPoints; idx += blockDim.x) {
      const int plane = idx / kPoints;
      const int point = idx - plane * kPoints;
      int8_t v;
      if (plane == 16) v = (int8_t)(1 - (int)tp);
      else v = (int8_t)((masks[plane & 1][plane >> 1][point >> 5] >> (point & 31)) & 1u);
      dst[idx] = v;
    }
    __syncthreads();
  }
}

// reference planes -> NHWC bf16.  T=bool semantics of DlTFNetworkEvaluator<bool>: any non-zero byte is 1.
__global__ void __launch_bounds__(256) planes_to_nhwc_kernel(const int8_t* __restrict__ planes, int hwc,
                                                             __nv_bfloat16* __restrict__ out, int n,
                                                             uint32_t kOneBf16) {
  __shared__ int8_t sp[UG_NUM_PLANES * kPoints + 7];
  for (int p = blockIdx.x; p < n; p += gridDim.x) {
    const int8_t* src = planes + (size_t)p * UG_NUM_PLANES * kPoints;
    for (int i = threadIdx.x; i < UG_NUM_PLANES * kPoints; i += blockDim.x) sp[i] = src[i];
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)p * kPoints * kCinPad);
    for (int idx = threadIdx.x; idx < kChunksPerPos; idx += blockDim.x) {
      const int point = idx / kChunksPerPoint;
      const int q = idx - point * kChunksPerPoint;
      uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ch = 8 * q + e;
        if (ch < UG_NUM_PLANES) {
          const int8_t b = hwc ? sp[point * UG_NUM_PLANES + ch] : sp[ch * kPoints + point];
          if (b != 0) w[e >> 1] |= kOneBf16 << (16 * (e & 1));
        }
      }
      dst[idx] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
  }
}

__global__ void nhwc_to_bytes_kernel(const __nv_bfloat16* __restrict__ in, int8_t* __restrict__ out, size_t total,
                                     int f16) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t point = i / UG_NUM_PLANES;
  const int ch = (int)(i - point * UG_NUM_PLANES);
  out[i] = (int8_t)unpack1(reinterpret_cast<const uint16_t*>(in)[point * kCinPad + ch], f16);
}

__global__ void nhwc_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t total,
                                   int f16) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t point = i / UG_NUM_PLANES;
  const int ch = (int)(i - point * UG_NUM_PLANES);
  out[i] = unpack1(reinterpret_cast<const uint16_t*>(in)[point * kCinPad + ch], f16);
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t count,
                                   int f16) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = unpack1(reinterpret_cast<const uint16_t*>(in)[i], f16);
}

__global__ void widen_outputs_kernel(const float* __restrict__ policy, const float* __restrict__ value,
                                     double* __restrict__ policy64, double* __restrict__ value64, int n, int vt) {
  const size_t total = (size_t)n * UG_NUM_MOVES;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) policy64[i] = (double)policy[i];
  if (i < (size_t)n) {
    double d = (double)value[i];
    if (vt == UG_TR_FLIP_SIGN) d = -d;
    else if (vt == UG_TR_FLIP_PROB) d = 1 - d;
    value64[i] = d;
  }
}

inline int pos_grid(int n) { return n < 148 * 16 ? n : 148 * 16; }

}  // namespace

void launch_encode_packed_nhwc(const ug_packed_pos* pos, __nv_bfloat16* out, int n, int f16, cudaStream_t s) {
  if (n <= 0) return;
  encode_packed_nhwc_kernel<<<pos_grid(n), 256, 0, s>>>(pos, out, n, one16(f16));
}
void launch_encode_packed_chw(const ug_packed_pos* pos, int8_t* out, int n, cudaStream_t s) {
  if (n <= 0) return;
  encode_packed_chw_kernel<<<pos_grid(n), 256, 0, s>>>(pos, out, n);
}
void launch_planes_to_nhwc(const int8_t* planes, int hwc, __nv_bfloat16* out, int n, int f16, cudaStream_t s) {
  if (n <= 0) return;
  planes_to_nhwc_kernel<<<pos_grid(n), 256, 0, s>>>(planes, hwc, out, n, one16(f16));
}
void launch_nhwc_to_bytes(const __nv_bfloat16* in, int8_t* out, int n, int f16, cudaStream_t s) {
  if (n <= 0) return;
  const size_t total = (size_t)n * kPoints * UG_NUM_PLANES;
  nhwc_to_bytes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, out, total, f16);
}
void launch_nhwc_to_f32(const __nv_bfloat16* in, float* out, int n, int f16, cudaStream_t s) {
  if (n <= 0) return;
  const size_t total = (size_t)n * kPoints * UG_NUM_PLANES;
  nhwc_to_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, out, total, f16);
}
void launch_widen_outputs(const float* policy, const float* value, double* policy64, double* value64, int n, int vt,
                          cudaStream_t s) {
  if (n <= 0) return;
  const size_t total = (size_t)n * UG_NUM_MOVES;
  widen_outputs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(policy, value, policy64, value64, n, vt);
}
void launch_bf16_to_f32(const __nv_bfloat16* in, float* out, size_t count, int f16, cudaStream_t s) {
  if (count == 0) return;
  bf16_to_f32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(in, out, count, f16);
}

}  // namespace ug
