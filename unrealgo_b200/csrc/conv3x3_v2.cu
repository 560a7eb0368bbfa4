// K3/K4 (second generation) — 3x3 SAME convolution as a tcgen05/TMEM implicit GEMM with the input tile
// resident in shared memory.
//
// Same contract as conv3x3_tc.cu (the Conv2D + FusedBatchNorm + Relu (+ Add) nodes the reference runs inside
// tensorflow::Session::Run, funcapproximator/DlTFNetworkEvaluator.cc:295), different operand staging.  The
// im2col kernel re-reads the activations once per filter tap (9x) and was measured L2-delivery-bound and
// epilogue-bound (profiles/).  Here, per 128-pixel tile and 64-channel chunk:
//
//   A  one TMA load of 168 consecutive pixel rows (the tile plus a 20-row halo either side: 19+1) into a
//      SWIZZLE_128B slot.  Filter tap (dy,dx) is the SAME slot read through a shared-memory descriptor whose
//      start address is shifted by (20 + 19*dy + dx) rows — the flat NHWC pixel index makes every tap a
//      constant row shift.  Pixels whose tap source lies off the board (or in the neighbouring image) must
//      contribute zero: those output rows are switched off for that tap's MMAs with tcgen05.mma's
//      disable-output-lane mask (a 256-bit mask per (tile, tap), periodic in the tile index with period 361,
//      precomputed on the host).  No padded activation layout, no wasted MMA rows, exact zero padding.
//   B  [Cout/2 x 64] weight tile per (chunk, tap) through a 6-stage TMA ring (each CTA of the pair stages half).
//   D  2 x 256 TMEM columns, double-buffered across tiles.
//   epilogue  tcgen05.ld -> +bias (+residual hi/lo) -> ReLU -> 16-bit hi (+lo) pack; the residual tile arrives
//      by TMA into shared memory (prefetched while the tile's MMAs run), results leave through per-warp
//      swizzled staging and TMA stores — no uncoalesced per-thread global traffic.
//
// Warp roles (384 threads, persistent, CTA pair = cta_group::2): 0 = A-chunk + mask producer, 1 = B producer,
// 2 = TMEM allocator + MMA issuer (leader CTA, one lane), 3 = residual producer, 4-11 = epilogue (two warps per
// TMEM lane quarter).
#include <cuda_bf16.h>

#include <atomic>

#include <mutex>
#include <vector>

#include "conv3x3_v2.h"
#include "numfmt.cuh"
#include "ptx.cuh"

namespace ug {

namespace {

constexpr int kBoard = 19;
constexpr int kPix = 361;
constexpr int kTileM = 128;                 // rows per CTA
constexpr int kHalo = kBoard + 1;           // 20
constexpr int kARows = kTileM + 2 * kHalo;  // 168
constexpr int kAccCols = 256;
constexpr int kNA = 3;   // A chunk slots
constexpr int kNR = 2;   // residual slots (32 columns each, hi + lo), only in kernels that read a residual
constexpr int kEpiWarps = 8;  // two per TMEM lane quarter
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kEpiCols = 32;
constexpr int kResHalf = kTileM * kEpiCols * 2;   // 8192: 128 rows x 64 B
constexpr int kResSlot = 2 * kResHalf;            // hi + lo
constexpr int kOutHalf = 32 * kEpiCols * 2;       // 2048: 32 rows x 64 B (one warp)
constexpr int kOutSlot = 2 * kOutHalf;            // hi + lo

// Shared-memory plan.  Kernels without a residual input spend the residual ring's 32 KB on two more weight stages.
template <int CK, bool HAS_RES>
struct Cfg {
  static constexpr int kNB = HAS_RES ? 6 : 8;       // B stages
  static constexpr int kNRes = HAS_RES ? kNR : 0;   // residual slots
  static constexpr int kRowBytes = CK * 2;
  static constexpr int kASlot = kARows * kRowBytes;   // 21504 (CK=64) / 10752 (CK=32)
  static constexpr int kBStage = kTileM * kRowBytes;  // up to 128 weight rows per CTA
  static constexpr int kOffB = kNA * kASlot;
  static constexpr int kOffRes = kOffB + kNB * kBStage;
  static constexpr int kOffOut = kOffRes + kNRes * kResSlot;
  static constexpr int kOffBias = kOffOut + kEpiWarps * kOutSlot;  // one staging slot (hi + lo) per epilogue warp
  static constexpr int kOffBars = kOffBias + 256 * 4;
  static constexpr int kNumBars = 2 * kNA + 2 * kNB + 2 * kNR + 2 + 2;  // (residual barriers exist in every variant)
  static constexpr int kBytes = kOffBars + kNumBars * 8 + 16;
  static constexpr int kSmem = kBytes + 1024;  // alignment slack
  static_assert(kSmem <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
};

// Disable-output-lane words: c_edge_mask[tap][q] bit i is set when pixel (q + i) mod 361 of an image has no
// in-board source for filter tap `tap` (tap = 3*(dy+1) + (dx+1)): that output row must not receive the tap's
// contribution.  13 KB of constant memory, read through the uniform datapath by the MMA-issuing warp.
__constant__ uint32_t c_edge_mask[9][kPix];

// 16-bit pack with the fp16 saturation folded into the conversion instruction (F2FP.SATFINITE.*.PACK_AB)
template <bool F16>
__device__ __forceinline__ uint32_t pack2_sat(float lo, float hi) {
  uint32_t r;
  if constexpr (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <bool F16>
__device__ __forceinline__ float2 unpack2_t(uint32_t w) {
  if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  else return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

// wait and (when profiling) add the stalled cycles to *acc
__device__ __forceinline__ void timed_wait(uint64_t* bar, uint32_t parity, bool prof, long long* acc) {
  if (!prof) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  *acc += clock64() - t0;
}

// RES: 0 = no residual, 1 = one 16-bit residual tensor, 2 = hi + lo pair.  OUT_LO: also write (result - hi).
// HEADS > 0: last layer of the tower — instead of storing the activation, every thread feeds its row's fp32
// results into the HEADS 1x1 head convolutions (policy channels first) and writes relu(head) to args.head_out.
template <int CK, bool F16, int RES, bool OUT_LO, int HEADS>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_v2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmRhi, const __grid_constant__ CUtensorMap tmRlo,
                  const __grid_constant__ CUtensorMap tmOhi, const __grid_constant__ CUtensorMap tmOlo,
                  const ConvV2Args args) {
  using C = Cfg<CK, (RES > 0)>;
  constexpr int kNB = C::kNB;
  constexpr int kRowBytes = C::kRowBytes;
  constexpr int kMmaPerStep = CK / 16;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + C::kOffB;
  uint8_t* smRes = smem + C::kOffRes;
  uint8_t* smOut = smem + C::kOffOut;
  float* smBias = reinterpret_cast<float*>(smem + C::kOffBias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBars);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kNA;
  uint64_t* b_full = a_empty + kNA;
  uint64_t* b_empty = b_full + kNB;
  uint64_t* res_full = b_empty + kNB;
  uint64_t* res_empty = res_full + kNR;
  uint64_t* acc_full = res_empty + kNR;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler too
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;

  const int cout = args.cout;
  const int b_rows = cout >> 1;
  const uint32_t b_bytes = (uint32_t)b_rows * kRowBytes;
  const int chunks = args.cin / CK;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  // Which cluster gets the short tile list (num_tiles is rarely a multiple of the cluster count) rotates from
  // layer to layer, so that with tile-granular dependencies every SM pair does the same work over a few layers.
  const int tile0 = (cluster_id + args.tile_rot) % num_clusters;
  constexpr bool has_res = RES > 0;
  constexpr bool has_res_lo = RES > 1;
  constexpr bool has_out_lo = OUT_LO;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOhi);
    if (has_res) tma_prefetch_desc(&tmRhi);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kNA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kNB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < kNR; ++s) { mbar_init(&res_full[s], 1); mbar_init(&res_empty[s], 4); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 2 * kEpiWarps);  // one elected lane per epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<2>(tmem_slot, 512);
  for (int i = threadIdx.x; i < cout; i += blockDim.x) smBias[i] = __ldg(args.bias + i);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  // Programmatic dependent launch: this grid may have been scheduled while the previous layer was still
  // draining.  Everything above (barriers, TMEM, bias) and the weight ring below do not depend on it; the two
  // warps that read the previous layer's output (A chunks, residual) wait for it, and every later step of the
  // tile pipeline is ordered behind their loads.
  pdl_launch_dependents();

  // The three producer warps and the MMA warp run their loops with all 32 lanes on warp-uniform values and
  // predicate only the TMA / MMA instruction itself on one elected lane.  (Issuing from inside an `if (lane == 0)`
  // region makes ptxas keep the operands in vector registers and wrap every UTCHMMA/UTMALDG in an
  // ELECT + R2UR.BROADCAST loop — measured ~210 cycles per MMA issue against a 128-cycle MMA.)
  const uint32_t elected = elect_one();
  if (warp == 0) {
    // ===================== A-chunk producer =====================
    // What this layer reads was written by the previous kernel in the stream.  Either wait for that whole grid
    // (griddepcontrol.wait), or — tile-granular mode — for the three 256-row tiles this tile's rows and halo come
    // from: the previous layer's epilogue counts its finished tiles in args.flags_in, so the tail of one layer and
    // the head of the next overlap instead of draining the GPU at every layer boundary.
    const bool tile_deps = args.flags_in != nullptr;
    if (!tile_deps) pdl_wait();
    int sa = 0;
    uint32_t pha = 0;
    for (int tile = tile0; tile < args.num_tiles; tile += num_clusters) {
      const int mc = tile * (2 * kTileM) + (int)cta_rank * kTileM;
      if (tile_deps) {
        if (elected) {
          const int t_lo = tile > 0 ? tile - 1 : 0;
          const int t_hi = tile + 1 < args.num_tiles ? tile + 1 : args.num_tiles - 1;
          for (int t = t_lo; t <= t_hi; ++t) wait_flag_ge(args.flags_in + t, args.target_in);
          fence_proxy_async_all();  // the TMA loads below must observe what the flags published
        }
        __syncwarp();
      }
      for (int j = 0; j < chunks; ++j) {
        mbar_wait(&a_empty[sa], pha ^ 1);
        if (elected) {
          if (leader) mbar_expect_tx(&a_full[sa], 2 * C::kASlot);
          tma_load_2d_2sm(smA + sa * C::kASlot, &tmA, &a_full[sa], j * CK, mc - kHalo);
        }
        __syncwarp();
        if (++sa == kNA) { sa = 0; pha ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== B (weights) producer =====================
    int sb = 0;
    uint32_t phb = 0;
    int nb_issued = 0;
    for (int tile = tile0; tile < args.num_tiles; tile += num_clusters) {
      for (int j = 0; j < chunks; ++j) {
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
          const int tap = (t == 0) ? 4 : (t <= 4 ? t - 1 : t);  // centre first, then 0,1,2,3,5,6,7,8
          mbar_wait(&b_empty[sb], phb ^ 1);
          if (elected) {
            if ((args.desc_mode & 16) && nb_issued >= kNB) {
              // timing probe: after the first lap leave the stale weights in place (no L2 traffic for B)
              if (leader) mbar_arrive(&b_full[sb]);
            } else {
              if (leader) mbar_expect_tx(&b_full[sb], 2 * b_bytes);
              tma_load_2d_2sm(smB + sb * C::kBStage, &tmB, &b_full[sb], tap * args.cin + j * CK,
                              (int)cta_rank * b_rows);
            }
          }
          ++nb_issued;
          __syncwarp();
          if (++sb == kNB) { sb = 0; phb ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer =====================
    // One warp issues every MMA of the tile pair: 128 tensor-pipe cycles per 256x256x16 MMA, so the issue loop has
    // about 500 cycles per filter tap (4 MMAs) before it, not the tensor pipe, paces the kernel.  The nine taps are
    // fully unrolled (row shifts and descriptor offsets are immediates) and the 72 lane-mask words of a tile are
    // fetched once per tile instead of eight per tap: the first version of this loop (per-tap mask lookup with a
    // dynamic constant index, ~130 SASS instructions per tap) measured 139 cycles per MMA.
    if (leader) {
      const uint32_t idesc = make_idesc_16(2 * kTileM, cout, F16 ? 1 : 0);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int it = 0;
      const bool prof = args.prof != nullptr;
      long long w_acc = 0, w_a = 0, w_b = 0;
      const long long t_begin = clock64();
      for (int tile = tile0; tile < args.num_tiles; tile += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        // pixel-in-image index of the tile's first row; the lane mask of 32 consecutive rows depends only on
        // (tap, that index mod 361)
        const int q0 = (int)(((long)tile * (2 * kTileM)) % kPix);
        uint32_t msk[9][8];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          int q = q0 + 32 * w;
          if (q >= kPix) q -= kPix;  // q0 < 361 and 32*w <= 224, so one subtraction suffices
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) msk[tap][w] = c_edge_mask[tap][q];
        }
        timed_wait(&acc_empty[as], aphase ^ 1, prof, &w_acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccCols;
#pragma unroll 1
        for (int j = 0; j < chunks; ++j) {
          timed_wait(&a_full[sa], pha, prof, &w_a);
          const uint64_t adesc0 = make_kmajor_desc<kRowBytes>(smem_u32(smA + sa * C::kASlot));
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            constexpr int kOrder[9] = {4, 0, 1, 2, 3, 5, 6, 7, 8};  // centre first: it initialises the accumulator
            const int tap = kOrder[t];
            const int shift = kHalo + (tap / 3 - 1) * kBoard + (tap % 3 - 1);  // rows: a compile-time constant
            timed_wait(&b_full[sb], phb, prof, &w_b);
            tc_fence_after();
            const uint64_t adesc = adesc0 + (uint64_t)(shift * (kRowBytes / 16));
            const uint64_t bdesc = make_kmajor_desc<kRowBytes>(smem_u32(smB + sb * C::kBStage));
            if (elected) {
#pragma unroll
              for (int k = 0; k < kMmaPerStep; ++k) {
                if (tap == 4)  // the centre tap never needs a mask
                  umma_f16<2>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((j | k) != 0));
                else
                  umma_f16_2sm_masked(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u, msk[tap]);
              }
              umma_commit_2sm(&b_empty[sb], 3);
            }
            __syncwarp();
            if (++sb == kNB) { sb = 0; phb ^= 1; }
          }
          if (elected) umma_commit_2sm(&a_empty[sa], 3);
          __syncwarp();
          if (++sa == kNA) { sa = 0; pha ^= 1; }
        }
        if (elected) umma_commit_2sm(&acc_full[as], 3);
        __syncwarp();
      }
      if (prof && lane == 0) {
        long long* o = args.prof + (size_t)cluster_id * 64;
        o[0] = clock64() - t_begin; o[1] = w_acc; o[2] = 0; o[3] = w_a; o[4] = w_b; o[5] = it;
      }
    }
  } else if (warp == 3) {
    // ===================== residual producer =====================
    if constexpr (has_res) {
      const bool tile_deps = args.flags_res != nullptr;
      if (!tile_deps) pdl_wait();
      int rs = 0;
      uint32_t phr = 0;
      constexpr uint32_t bytes = has_res_lo ? kResSlot : kResHalf;
      for (int tile = tile0; tile < args.num_tiles; tile += num_clusters) {
        const int mc = tile * (2 * kTileM) + (int)cta_rank * kTileM;
        if (tile_deps) {  // the residual tile was written two layers back: its own counter
          if (elected) {
            wait_flag_ge(args.flags_res + tile, args.target_res);
            fence_proxy_async_all();
          }
          __syncwarp();
        }
        for (int c0 = 0; c0 < cout; c0 += kEpiCols) {
          mbar_wait(&res_empty[rs], phr ^ 1);
          if (elected) {
            mbar_expect_tx(&res_full[rs], bytes);
            tma_load_2d(smRes + rs * kResSlot, &tmRhi, &res_full[rs], c0, mc);
            if constexpr (has_res_lo) tma_load_2d(smRes + rs * kResSlot + kResHalf, &tmRlo, &res_full[rs], c0, mc);
          }
          __syncwarp();
          if (++rs == kNR) { rs = 0; phr ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    // Eight warps: a warp may only read the TMEM lane quarter (warp % 4), so two warps share each quarter and
    // split the 32-column steps between them (group 0: steps 0, 2, 4, ...; group 1: steps 1, 3, 5, ...).  The
    // epilogue of a layer's last tile is not hidden behind any MMA, so its length is paid once per launch.
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, +32)
    const int grp = (warp - 4) >> 2;       // 0 or 1
    const float act_floor = args.relu ? 0.f : -3.0e38f;  // ReLU as a branch-free max
    uint8_t* my_out = smOut + (grp * 4 + quarter) * kOutSlot;  // one staging slot (hi + lo) per warp
    const int row_l = quarter * 32 + lane;   // row inside the CTA tile
    const int sw = (row_l >> 1) & 3;         // SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3
    const int rs = grp;                      // residual slot: step c uses slot c % kNR, and kNR == 2
    uint32_t phr = 0;
    int it = 0;
    const bool prof = args.prof != nullptr && leader && warp == 4;
    long long w_full = 0, w_res = 0, w_st = 0;
    const long long t_begin = clock64();
    for (int tile = tile0; tile < args.num_tiles; tile += num_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int mc = tile * (2 * kTileM) + (int)cta_rank * kTileM;
      timed_wait(&acc_full[as], aphase, prof, &w_full);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * kAccCols;
      float hacc[HEADS > 0 ? HEADS : 1];
#pragma unroll
      for (int h = 0; h < (HEADS > 0 ? HEADS : 1); ++h) hacc[h] = 0.f;
      float out_max = 0.f;  // fp16 range watch: the largest |output| this thread converts in this tile
      for (int c0 = grp * kEpiCols; c0 < cout; c0 += 2 * kEpiCols) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c0, v);
        uint4 rh[4], rl[4];
        if constexpr (has_res) {
          timed_wait(&res_full[rs], phr, prof, &w_res);
          const uint8_t* rrow = smRes + rs * kResSlot + row_l * 64;
#pragma unroll
          for (int q = 0; q < 4; ++q) rh[q] = *reinterpret_cast<const uint4*>(rrow + ((q ^ sw) << 4));
          if constexpr (has_res_lo) {
#pragma unroll
            for (int q = 0; q < 4; ++q) rl[q] = *reinterpret_cast<const uint4*>(rrow + kResHalf + ((q ^ sw) << 4));
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&res_empty[rs]);
          phr ^= 1;
        }
        tmem_ld_wait();
        if (c0 + 2 * kEpiCols >= cout) {
          // this warp has read its share of the accumulator: hand the TMEM stage back before the arithmetic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&acc_empty[as], 0);
        }
        uint4 oh[4], ol[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t packed[4], packed_lo[4];
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(&rh[q]);
          const uint32_t* lw = reinterpret_cast<const uint32_t*>(&rl[q]);
          const float4 b0 = *reinterpret_cast<const float4*>(smBias + c0 + q * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(smBias + c0 + q * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = q * 8 + e * 2;
            float f0 = __uint_as_float(v[c]) + bb[e * 2];
            float f1 = __uint_as_float(v[c + 1]) + bb[e * 2 + 1];
            if constexpr (has_res) {
              const float2 r2 = unpack2_t<F16>(rw[e]);
              f0 += r2.x;
              f1 += r2.y;
              if constexpr (has_res_lo) {
                const float2 l2 = unpack2_t<F16>(lw[e]);
                f0 += l2.x;
                f1 += l2.y;
              }
            }
            f0 = fmaxf(f0, act_floor);
            f1 = fmaxf(f1, act_floor);
            if constexpr (F16 && HEADS == 0) out_max = fmaxf(out_max, fmaxf(fabsf(f0), fabsf(f1)));
            if constexpr (HEADS > 0) {
              // head weights [HEADS][cout] fp32: the same address in every lane (one L1 wavefront per load)
#pragma unroll
              for (int h = 0; h < HEADS; ++h) {
                const float2 hw = __ldg(reinterpret_cast<const float2*>(args.head_w + (size_t)h * cout + c0 + c));
                hacc[h] = fmaf(f0, hw.x, fmaf(f1, hw.y, hacc[h]));
              }
              continue;
            }
            packed[e] = pack2_sat<F16>(f0, f1);
            if constexpr (has_out_lo) {
              const float2 hi = unpack2_t<F16>(packed[e]);
              packed_lo[e] = pack2_sat<F16>(f0 - hi.x, f1 - hi.y);
            }
          }
          oh[q] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          if constexpr (has_out_lo) ol[q] = make_uint4(packed_lo[0], packed_lo[1], packed_lo[2], packed_lo[3]);
        }
        if constexpr (HEADS == 0) {
          // the staging slot must have been read by this warp's previous TMA store (its arithmetic is already done)
          if (lane == 0) {
            const long long t0 = prof ? clock64() : 0;
            bulk_wait_group_read<0>();
            if (prof) w_st += clock64() - t0;
          }
          __syncwarp();
          uint8_t* orow = my_out + lane * 64;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            *reinterpret_cast<uint4*>(orow + ((q ^ sw) << 4)) = oh[q];
            if constexpr (has_out_lo) *reinterpret_cast<uint4*>(orow + kOutHalf + ((q ^ sw) << 4)) = ol[q];
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int row0 = mc + quarter * 32;
            if (row0 < args.M) {  // rows past M inside the box are clipped by the tensor map (its height is M_alloc)
              tma_store_2d(&tmOhi, my_out, c0, row0);
              if (has_out_lo) tma_store_2d(&tmOlo, my_out + kOutHalf, c0, row0);
            }
            bulk_commit_group();
          }
        }
      }
      if constexpr (F16 && HEADS == 0) {
        // cvt.satfinite clamped something: the evaluator is outside fp16's range (ug_net_info.saturation_events)
        if (out_max > 65504.f && args.sat_count != nullptr) atomicAdd(args.sat_count, 1ull);
      }
      if constexpr (HEADS == 0) {
        // tile-granular mode: publish "this warp's share of the tile is in memory" (16 counts complete a tile)
        if (args.flags_out != nullptr && lane == 0) {
          bulk_wait_group_all();     // the TMA stores of this tile have completed, not merely been read
          fence_proxy_async_all();   // async-proxy writes before the generic-proxy release below
          red_release_gpu_add(args.flags_out + tile, 1u);
        }
      }
      if constexpr (HEADS > 0) {
        // Each group holds the head dot products over its half of the channels.  Group 1 hands its partial sums to
        // group 0 through shared memory (the unused output staging, double-buffered by tile parity; one named
        // barrier per tile: group 1 cannot reach the buffer again before group 0 has passed the next barrier).
        float* xch = reinterpret_cast<float*>(smOut) + (it & 1) * (kTileM * 4) + row_l * 4;
        if (grp == 1) {
#pragma unroll
          for (int h = 0; h < HEADS; ++h) xch[h] = hacc[h];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (grp == 0) {
          // feat[row][h] = relu(head conv + folded batch-norm offset): 4*HEADS bytes per row, rows consecutive
          const long row = (long)mc + row_l;
          if (row < args.M) {
#pragma unroll
            for (int h = 0; h < HEADS; ++h)
              args.head_out[row * HEADS + h] = fmaxf(hacc[h] + xch[h] + __ldg(args.head_b + h), 0.f);
          }
        }
      }
    }
    if (lane == 0) bulk_wait_group_all();
    if (prof && lane == 0) {
      long long* o = args.prof + (size_t)cluster_id * 64;
      o[8] = clock64() - t_begin; o[9] = w_full; o[10] = w_res; o[11] = w_st;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc<2>(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host
std::mutex g_mask_mu;
bool g_mask_done[64] = {};  // per device

bool upload_masks(int device) {
  std::lock_guard<std::mutex> lk(g_mask_mu);
  if (device < 0 || device >= 64) return false;
  if (g_mask_done[device]) return true;
  std::vector<uint32_t> h((size_t)9 * kPix, 0u);
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    for (int q = 0; q < kPix; ++q)
      for (int i = 0; i < 32; ++i) {
        const int p = (q + i) % kPix;
        const int y = p / kBoard + dy, x = p % kBoard + dx;
        if (y < 0 || y >= kBoard || x < 0 || x >= kBoard) h[(size_t)tap * kPix + q] |= 1u << i;
      }
  }
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(device);
  const cudaError_t e = cudaMemcpyToSymbol(c_edge_mask, h.data(), h.size() * 4);
  cudaSetDevice(cur);
  if (e != cudaSuccess) return false;
  g_mask_done[device] = true;
  return true;
}

template <int CK, bool F16, int RES, bool OUT_LO, int HEADS = 0>
cudaError_t launch_v2(const ConvV2Maps& m, const ConvV2Args& a, int sms, cudaStream_t stream) {
  auto kfn = conv3x3_v2_kernel<CK, F16, RES, OUT_LO, HEADS>;
  constexpr int smem = Cfg<CK, (RES > 0)>::kSmem;
  // function attributes are per device: remember which devices this instantiation has been prepared on
  static std::atomic<uint64_t> attr_done{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  int clusters = sms / 2;
  if (clusters > a.num_tiles) clusters = a.num_tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * 2);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a.pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kfn, *m.a, *m.w, *m.res_hi, *m.res_lo, *m.out_hi, *m.out_lo, a);
}

template <int CK, bool F16>
cudaError_t dispatch_v2(const ConvV2Maps& m, const ConvV2Args& a, int sms, cudaStream_t stream) {
  const int res = a.has_res ? (a.has_res_lo ? 2 : 1) : 0;
  const bool lo = a.has_out_lo != 0;
  if (a.head_out != nullptr) {
    // fused heads: only the shapes the tower's last layer takes (64-channel chunks, 3 or 4 head channels: the exchange
    // buffer between the two epilogue warp groups holds 4 floats per row)
    if constexpr (CK == 64) {
      if (a.head_channels == 3 && res == 2 && lo) return launch_v2<CK, F16, 2, true, 3>(m, a, sms, stream);
      if (a.head_channels == 3 && res == 1 && !lo) return launch_v2<CK, F16, 1, false, 3>(m, a, sms, stream);
      // 2 policy + 2 value channels (a two-channel value head)
      if (a.head_channels == 4 && res == 2 && lo) return launch_v2<CK, F16, 2, true, 4>(m, a, sms, stream);
      if (a.head_channels == 4 && res == 1 && !lo) return launch_v2<CK, F16, 1, false, 4>(m, a, sms, stream);
    }
    return cudaErrorInvalidValue;
  }
  if (res == 0 && !lo) return launch_v2<CK, F16, 0, false>(m, a, sms, stream);
  if (res == 0 && lo) return launch_v2<CK, F16, 0, true>(m, a, sms, stream);
  if (res == 1 && !lo) return launch_v2<CK, F16, 1, false>(m, a, sms, stream);
  if (res == 2 && lo) return launch_v2<CK, F16, 2, true>(m, a, sms, stream);
  if (res == 2 && !lo) return launch_v2<CK, F16, 2, false>(m, a, sms, stream);
  return cudaErrorInvalidValue;  // (one residual tensor, split output) is not a combination the tower produces
}

}  // namespace

bool conv3x3_v2_prepare(int device) { return upload_masks(device); }

int conv3x3_v2_max_active_clusters(int cluster_size) {
  auto kfn = conv3x3_v2_kernel<64, true, 2, true, 0>;
  constexpr int smem = Cfg<64, true>::kSmem;
  if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
  if (cluster_size > 8 &&
      cudaFuncSetAttribute(kfn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return -1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148 / cluster_size * cluster_size);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kfn, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

bool conv3x3_v2_can_fuse_heads(int ck, int has_res, int has_res_lo, int has_out_lo, int head_channels) {
  if (ck != 64 || (head_channels != 3 && head_channels != 4) || !has_res) return false;
  return (has_res_lo && has_out_lo) || (!has_res_lo && !has_out_lo);
}

int conv3x3_v2_smem_bytes(int ck) { return ck == 64 ? Cfg<64, true>::kSmem : Cfg<32, true>::kSmem; }

cudaError_t conv3x3_v2_launch(const ConvV2Maps& maps, ConvV2Args a, int ck, int device, int sms,
                              cudaStream_t stream) {
  // cout: whole pairs of 32-column epilogue steps (the two epilogue warp groups alternate steps, and the residual
  // ring pairs a step's parity with a slot)
  if (a.cout > 256 || a.cout % 64 != 0 || a.cin % ck != 0 || (ck != 64 && ck != 32)) return cudaErrorInvalidValue;
  if (!maps.a || !maps.w || !maps.out_hi) return cudaErrorInvalidValue;
  if (!upload_masks(device)) return cudaErrorUnknown;  // no-op after conv3x3_v2_prepare()
  a.num_tiles = (a.M + 2 * kTileM - 1) / (2 * kTileM);
  ConvV2Maps m = maps;
  // unused maps still have to be valid kernel parameters
  if (!m.res_hi) m.res_hi = m.out_hi;
  if (!m.res_lo) m.res_lo = m.res_hi;
  if (!m.out_lo) m.out_lo = m.out_hi;
  if (ck == 64) return a.f16 ? dispatch_v2<64, true>(m, a, sms, stream) : dispatch_v2<64, false>(m, a, sms, stream);
  return a.f16 ? dispatch_v2<32, true>(m, a, sms, stream) : dispatch_v2<32, false>(m, a, sms, stream);
}

}  // namespace ug
