// K3/K4 — 3x3 SAME convolution of the residual tower as a tcgen05/TMEM implicit GEMM.
//
// Replaces the Conv2D + FusedBatchNorm + Relu (+ Add) nodes that the reference executes inside
// tensorflow::Session::Run (funcapproximator/DlTFNetworkEvaluator.cc:295).  Batch-norm is folded into the
// weights/bias at load time (net.cu), so one launch computes
//     out[m, co] = act( sum_{tap, ci} in[pixel(m) + tap, ci] * W[co, tap*Cin + ci] + bias[co] (+ res[m, co]) )
// with m = (n*19 + y)*19 + x over NHWC bf16 activations.
//
// GEMM view: M = N*361 output pixels, N = Cout (<= 256, one tile), K = 9*Cin.
//   A (activations) : TMA *im2col* loads — 128 consecutive output pixels x CK channels for one filter tap
//                     per k-step; the halo is zero-filled by the TMA unit, so there is no padded copy of the
//                     activations and no wasted MMA rows.
//   B (weights)     : [Cout][9*Cin] K-major bf16, plain 2-D TMA tile Cout x CK.
//   D (accumulator) : 128 lanes x Cout fp32 columns in TMEM, double-buffered (2 x 256 columns) so the
//                     epilogue of tile i overlaps the MMAs of tile i+1.
// Warp roles (256 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// lane), warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM lane quarter = warp & 3).
//
// Two variants share this file: conv3x3_tc_kernel<CK,1> (one CTA per 128-row tile) and
// conv3x3_tc_kernel<CK,2> (CTA pair, cta_group::2: 256-row tile, each CTA stages its own 128 A rows and
// half of the B tile, which halves per-SM weight traffic).
#include <cuda_bf16.h>

#include <atomic>

#include "conv3x3_tc.h"
#include "numfmt.cuh"
#include "ptx.cuh"

namespace ug {

constexpr int kBoard = 19;
constexpr int kPix = kBoard * kBoard;  // 361
constexpr int kTileM = 128;
constexpr int kMaxCout = 256;
constexpr int kAccCols = 256;  // TMEM columns per accumulator stage

template <int CK>
struct ConvSmem {
  static constexpr int kRowBytes = CK * 2;
  static constexpr int kABytes = kTileM * kRowBytes;
  static constexpr int kBBytesMax = kMaxCout * kRowBytes;
};

template <int CK, int CTAS>
__host__ __device__ constexpr int conv_stage_bytes() {
  return ConvSmem<CK>::kABytes + ConvSmem<CK>::kBBytesMax / CTAS;
}
template <int CK, int CTAS>
__host__ __device__ constexpr int conv_num_stages() {
  // 1-CTA, CK=64: 48 KB/stage -> 4 stages; 2-CTA: 32 KB/stage -> 6 stages.
  return CK == 64 ? (CTAS == 1 ? 4 : 6) : 6;
}
template <int CK, int CTAS>
__host__ __device__ constexpr int conv_smem_bytes() {
  return conv_num_stages<CK, CTAS>() * conv_stage_bytes<CK, CTAS>() + 1024 /*align slack*/ + 256 /*barriers*/;
}

template <int CK, int CTAS>
__global__ void __launch_bounds__(256, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const ConvArgs args) {
  constexpr int kStages = conv_num_stages<CK, CTAS>();
  constexpr int kRowBytes = CK * 2;
  constexpr int kABytes = ConvSmem<CK>::kABytes;
  constexpr int kStageBytes = conv_stage_bytes<CK, CTAS>();
  constexpr int kMmaPerStep = CK / 16;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = bars;                  // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;       // [kStages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * kStages;    // [2]        MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;         // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;

  const int cout = args.cout;
  const int b_rows = cout / CTAS;  // weight rows staged by this CTA
  const uint32_t b_bytes = (uint32_t)b_rows * kRowBytes;
  const int chunks = args.cin / CK;
  const int ksteps = 9 * chunks;
  // tiles are handed out to clusters; a 2-CTA tile is 256 rows (this CTA: rows [cta_rank*128, +128))
  const int cluster_id = blockIdx.x / CTAS;
  const int num_clusters = gridDim.x / CTAS;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);     // one expect_tx arrival (leader); TMA bytes of both CTAs land here
      mbar_init(&empty_bar[s], 1);    // one tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 128 * CTAS);  // every epilogue thread of the group
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CTAS>(tmem_slot, 512);
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < args.num_tiles; tile += num_clusters) {
        const int m0 = tile * (kTileM * CTAS) + (int)cta_rank * kTileM;
        const int n0 = m0 / kPix;
        const int rem = m0 - n0 * kPix;
        const int y0 = rem / kBoard;
        const int x0 = rem - y0 * kBoard;
        for (int tap = 0; tap < 9; ++tap) {
          const int r = tap / 3, s = tap - 3 * r;
          for (int j = 0; j < chunks; ++j) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* a_dst = smem + stage * kStageBytes;
            uint8_t* b_dst = a_dst + kABytes;
            if constexpr (CTAS == 1) {
              mbar_expect_tx(&full_bar[stage], kABytes + b_bytes);
              tma_load_im2col_4d(a_dst, &tmA, &full_bar[stage], j * CK, x0 - 1, y0 - 1, n0, (uint16_t)s,
                                 (uint16_t)r);
              tma_load_2d(b_dst, &tmB, &full_bar[stage], tap * args.cin + j * CK, 0);
            } else {
              // both CTAs credit the leader's barrier; the leader's MMA consumes both halves
              if (leader) mbar_expect_tx(&full_bar[stage], 2 * (kABytes + b_bytes));
              tma_load_im2col_4d_2sm(a_dst, &tmA, &full_bar[stage], j * CK, x0 - 1, y0 - 1, n0, (uint16_t)s,
                                     (uint16_t)r);
              tma_load_2d_2sm(b_dst, &tmB, &full_bar[stage], tap * args.cin + j * CK, (int)cta_rank * b_rows);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc_16(kTileM * CTAS, cout, args.f16);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < args.num_tiles; tile += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccCols;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = make_kmajor_desc<kRowBytes>(a_addr);
          const uint64_t bdesc = make_kmajor_desc<kRowBytes>(a_addr + kABytes);
#pragma unroll
          for (int k = 0; k < kMmaPerStep; ++k) {
            // +32 B along K inside the swizzle span (encoded >>4)
            umma_f16<CTAS>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (ks | k) != 0);
          }
          if constexpr (CTAS == 1) umma_commit(&empty_bar[stage]); else umma_commit_2sm(&empty_bar[stage], 3);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if constexpr (CTAS == 1) umma_commit(&acc_full[as]); else umma_commit_2sm(&acc_full[as], 3);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32)
    const __nv_bfloat16* __restrict__ res = args.residual;
    const __nv_bfloat16* __restrict__ res_lo = args.residual_lo;
    __nv_bfloat16* __restrict__ out = args.out;
    __nv_bfloat16* __restrict__ out_lo = args.out_lo;
    const float* __restrict__ bias = args.bias;
    const int f16 = args.f16;
    int it = 0;
    for (int tile = cluster_id; tile < args.num_tiles; tile += num_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const long row = (long)tile * (kTileM * CTAS) + (long)cta_rank * kTileM + quarter * 32 + lane;
      const bool row_ok = row < args.M;
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * kAccCols;
      for (int c0 = 0; c0 < cout; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c0, v);
        tmem_ld_wait();
        if (row_ok) {
          const size_t off = (size_t)row * cout + c0;
          uint4 rr[4], rl[4];
          if (res != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(res + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) rr[q] = __ldg(rp + q);
            if (res_lo != nullptr) {
              const uint4* lp = reinterpret_cast<const uint4*>(res_lo + off);
#pragma unroll
              for (int q = 0; q < 4; ++q) rl[q] = __ldg(lp + q);
            }
          }
          uint4 oo[4], ol[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t packed[4], packed_lo[4];
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(&rr[q]);
            const uint32_t* lw = reinterpret_cast<const uint32_t*>(&rl[q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = q * 8 + e * 2;
              float f0 = __uint_as_float(v[c]) + __ldg(bias + c0 + c);
              float f1 = __uint_as_float(v[c + 1]) + __ldg(bias + c0 + c + 1);
              if (res != nullptr) {
                const float2 r2 = unpack2(rw[e], f16);
                f0 += r2.x;
                f1 += r2.y;
                if (res_lo != nullptr) {
                  const float2 l2 = unpack2(lw[e], f16);
                  f0 += l2.x;
                  f1 += l2.y;
                }
              }
              if (args.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
              packed[e] = pack2(f0, f1, f16);
              if (out_lo != nullptr) {
                const float2 hi = unpack2(packed[e], f16);
                packed_lo[e] = pack2(f0 - hi.x, f1 - hi.y, f16);
              }
            }
            oo[q] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            if (out_lo != nullptr) ol[q] = make_uint4(packed_lo[0], packed_lo[1], packed_lo[2], packed_lo[3]);
          }
          uint4* op = reinterpret_cast<uint4*>(out + off);
#pragma unroll
          for (int q = 0; q < 4; ++q) op[q] = oo[q];
          if (out_lo != nullptr) {
            uint4* lp = reinterpret_cast<uint4*>(out_lo + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) lp[q] = ol[q];
          }
        }
      }
      tc_fence_before();
      if constexpr (CTAS == 1) mbar_arrive(&acc_empty[as]); else mbar_arrive_cluster(&acc_empty[as], 0);
    }
  }

  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CTAS>(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host
template <int CK, int CTAS>
static cudaError_t launch_variant(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvArgs& a, int sms,
                                  cudaStream_t stream) {
  auto kfn = conv3x3_tc_kernel<CK, CTAS>;
  constexpr int smem = conv_smem_bytes<CK, CTAS>();
  // function attributes are per device: remember which devices this instantiation has been prepared on
  static std::atomic<uint64_t> attr_done{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  int clusters = sms / CTAS;
  if (clusters > a.num_tiles) clusters = a.num_tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CTAS);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kfn, tmA, tmB, a);
}

cudaError_t conv3x3_tc_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, ConvArgs a, int ck, int ctas,
                              int sms, cudaStream_t stream) {
  if (a.cout > kMaxCout || a.cout % (16 * ctas) != 0 || a.cin % ck != 0) return cudaErrorInvalidValue;
  a.num_tiles = (a.M + kTileM * ctas - 1) / (kTileM * ctas);
  if (ck == 64 && ctas == 1) return launch_variant<64, 1>(tmA, tmB, a, sms, stream);
  if (ck == 64 && ctas == 2) return launch_variant<64, 2>(tmA, tmB, a, sms, stream);
  if (ck == 32 && ctas == 1) return launch_variant<32, 1>(tmA, tmB, a, sms, stream);
  if (ck == 32 && ctas == 2) return launch_variant<32, 2>(tmA, tmB, a, sms, stream);
  return cudaErrorInvalidValue;
}

}  // namespace ug
