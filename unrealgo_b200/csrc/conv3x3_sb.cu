// K3/K4 (small batches) — the 3x3 SAME tower convolution for 1 .. a few dozen positions.
//
// Same contract as conv3x3_v2.cu (Conv2D + FusedBatchNorm + Relu (+ Add) of the graph the reference runs inside
// tensorflow::Session::Run, funcapproximator/DlTFNetworkEvaluator.cc:295).  The reference's own engine never
// evaluates more than MAX_BATCHES = 16 positions per call (funcapproximator/DlTFNetworkEvaluator.h:15) and exactly
// one in the shipped FORCE_SINGLE_THREAD build (CMakeLists.txt:15; search/UctSearch.cpp:183-186): 361 .. 5776 GEMM
// rows.  conv3x3_v2's persistent 256-row x 256-column CTA-pair tiles put such a layer on 2 .. 23 of the 74 SM pairs and
// a layer then costs one full tile time (~30 us) whatever the batch.  Here the layer is cut the other way:
//
//   grid   (ceil(M / 128), cout / NT): one 128-row x NT-column tile per CTA, cta_group::1, NT in {16, 32, 64, 128}
//          chosen so that the grid is as large as one wave of the SMs allows (n = 1 -> 3 x 8 CTAs at NT = 32,
//          n = 16 -> 46 x 2 at NT = 128), so the 1.2 MB of a layer's weights are pulled through many SMs at once;
//   A      the whole K extent of the tile resident: cin/64 TMA loads of 168 consecutive pixel rows (tile + 20-row halo
//          either side), taps = the same slot read through a descriptor shifted by (20 + 19 dy + dx) rows, off-board
//          rows switched off per tap with tcgen05.mma's disable-output-lane mask (exactly conv3x3_v2's scheme);
//   B      weights in stages of three filter taps ([3][NT x 64] per chunk and stage) through a ring of up to 12 stages
//          (96 KB), 16 rows per TMA box: one mbarrier wait feeds twelve MMAs (see the MMA issuer);
//   D      NT TMEM columns; one tile per CTA, so no double buffering;
//   epilogue  all 8 warps (two per TMEM lane quarter, alternating 16-column steps): residual rows prefetched from
//          global memory before the accumulator is ready, tcgen05.ld -> +bias (+residual hi/lo) -> ReLU -> 16-bit hi
//          (+lo) -> 32-byte global stores per thread.  Small tensors: the uncoalesced per-row accesses that bounded
//          the first-generation kernel at batch 256 cost a few hundred nanoseconds here.
//
// Warp roles (256 threads): 0 = A producer (waits for the previous layer: griddepcontrol.wait), 1 = MMA issuer,
// 2 = TMEM allocator, 3 = weight producer (starts before the previous layer has finished); then every warp joins the
// epilogue.
//
// PAIR variant (cta_group::2, clusters of two CTAs along M): what bounds the single-CTA kernel from a handful of
// positions up is what each SM has to pull in from L2 — its A tile plus ALL the weights of its NT columns (n = 16,
// NT = 128: 86 + 590 KB per CTA at ~50 B/clk).  A CTA pair computes a 256-row x NT tile with one MMA stream issued by
// the leader: each CTA stages its own 128 rows of A and only HALF of the NT weight rows (the tensor cores of both SMs
// read both halves), so the weight bytes per SM halve.  Barriers as in conv3x3_v2.cu: both CTAs' TMA loads complete
// on the leader's full barriers, the leader's tcgen05.commit multicasts to both CTAs' empty / accumulator barriers.
// Same MMA order per output element, so the results are the same bits as the single-CTA variant's.
#include <cuda_bf16.h>

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "conv3x3_sb.h"
#include "numfmt.cuh"
#include "ptx.cuh"

namespace ug {

namespace {

constexpr int kBoard = 19;
constexpr int kPix = 361;
constexpr int kTileM = 128;
constexpr int kHalo = kBoard + 1;           // 20
constexpr int kARows = kTileM + 2 * kHalo;  // 168
constexpr int kThreads = 256;
constexpr int kTapsPerStage = 3;            // filter taps per weight stage
constexpr int kMaxStages = 12;              // 3 per 64-channel chunk of a 256-channel layer
constexpr int kBRing = 96 * 1024;
constexpr int kPairNarrowTiles = 6;           // CTA pairs: 32-column tiles up to this many 128-row tiles (two positions)
constexpr int kStageHalf = 32 * 64;           // TMA-store staging: 32 rows x 64 B (one warp, 32 output channels)
constexpr int kStageSlot = 2 * kStageHalf;    // hi + lo
constexpr int kMaxEpiSteps = 4;             // 16-column steps per warp: NT = 128 -> 8 steps over two warp groups

template <int CK>
struct SbCfg {
  static constexpr int kRowBytes = CK * 2;
  static constexpr int kMaxChunks = CK == 64 ? 4 : 1;   // cin <= 256 (CK = 64) / the 32-channel stem (CK = 32)
  static constexpr int kASlot = kARows * kRowBytes;     // 21504 / 10752
  static constexpr int kOffB = (kMaxChunks * kASlot + 1023) & ~1023;
  static constexpr int kOffBars = kOffB + kBRing;
  static constexpr int kNumBars = kMaxChunks + 2 * kMaxStages + 1;
  static constexpr int kBytes = kOffBars + kNumBars * 8 + 16;
  static constexpr int kSmem = kBytes + 1024;  // alignment slack
  static_assert(kSmem <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
};

// same table as conv3x3_v2.cu (a __constant__ symbol is private to its translation unit): c_edge_mask_sb[tap][q] bit i
// is set when pixel (q + i) mod 361 of an image has no in-board source for filter tap `tap`
__constant__ uint32_t c_edge_mask_sb[9][kPix];

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

__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// bring-up instrumentation: slot k of this CTA's 16 time stamps (SM clock; slots 8/9 hold the global timer in ns)
#define UG_SB_STAMP(k)                                                                                   \
  do {                                                                                                   \
    if (args.prof != nullptr && lane == 0)                                                               \
      args.prof[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + (k)] = clock64();                   \
  } while (0)

template <int CK, bool F16, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_sb_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmOhi, const __grid_constant__ CUtensorMap tmOlo,
                  const ConvSbArgs args) {
  using C = SbCfg<CK>;
  constexpr int kRowBytes = C::kRowBytes;
  constexpr int kMmaPerStep = CK / 16;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + C::kOffB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBars);
  uint64_t* a_full = bars;
  uint64_t* b_full = a_full + C::kMaxChunks;
  uint64_t* b_empty = b_full + kMaxStages;
  uint64_t* acc_full = b_empty + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler too
  const int lane = threadIdx.x & 31;
  const int nt = args.nt;
  const int stages = args.stages;
  const int chunks = args.cin / CK;
  // PAIR: the cluster is two consecutive CTAs along x, i.e. two consecutive 128-row tiles; rank 0 leads
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int m0 = (int)blockIdx.x * kTileM;
  const int n0 = (int)blockIdx.y * nt;
  const int nb = PAIR ? nt >> 1 : nt;                    // weight rows this CTA stages per filter tap
  const int nb0 = n0 + (int)cta_rank * nb;               // ... starting at this output channel
  const uint32_t tap_bytes = (uint32_t)nb * kRowBytes;   // one filter tap's [nb x CK] weight tile
  constexpr uint32_t kTxScale = PAIR ? 2u : 1u;          // the leader's full barriers count both CTAs' bytes

  if (warp == 0) {
    UG_SB_STAMP(0);
    if (args.prof != nullptr && lane == 0)
      args.prof[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + 8] = globaltimer_ns();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int j = 0; j < C::kMaxChunks; ++j) mbar_init(&a_full[j], 1);
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<PAIR ? 2 : 1>(tmem_slot, nt < 32 ? 32u : (uint32_t)nt);
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if constexpr (PAIR) {
    // the pair's MMAs address the accumulator of both CTAs with ONE TMEM address: both allocations (same size, into
    // an SM's empty TMEM — a CTA of this kernel has the SM to itself) must have landed on the same columns
    if (threadIdx.x == 0) {
      uint32_t peer_slot, peer_base;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_slot) : "r"(smem_u32(tmem_slot)), "r"(cta_rank ^ 1u));
      asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(peer_base) : "r"(peer_slot) : "memory");
      if (peer_base != tmem_base) {
        printf("ugeval: conv3x3_sb pair: TMEM bases differ (%x vs %x)\n", tmem_base, peer_base);
        __trap();
      }
    }
  }
  // Programmatic dependent launch: the next layer's CTAs may be scheduled from here on (their prologue and weight
  // loads do not depend on this layer); they read this layer's output only behind their own griddepcontrol.wait.
  pdl_launch_dependents();

  // Producer and MMA warps run their loops on all 32 lanes with warp-uniform values and predicate only the TMA / MMA
  // instruction on one elected lane (see conv3x3_v2.cu: issuing from inside `if (lane == 0)` costs an
  // ELECT + R2UR.BROADCAST loop per instruction).
  const uint32_t elected = elect_one();
  if (warp == 0) {
    // ===================== A producer: the previous layer's output =====================
    UG_SB_STAMP(1);
    pdl_wait();
    UG_SB_STAMP(2);
    for (int j = 0; j < chunks; ++j) {
      if (elected) {
        if (leader) mbar_expect_tx(&a_full[j], kTxScale * C::kASlot);
        if constexpr (PAIR) tma_load_2d_2sm(smA + j * C::kASlot, &tmA, &a_full[j], j * CK, m0 - kHalo);
        else tma_load_2d(smA + j * C::kASlot, &tmA, &a_full[j], j * CK, m0 - kHalo);
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ===================== weight producer =====================
    int sb = 0;
    uint32_t phb = 0;
    for (int j = 0; j < chunks; ++j) {
#pragma unroll 1
      for (int g = 0; g < 9 / kTapsPerStage; ++g) {
        mbar_wait(&b_empty[sb], phb ^ 1);
        if (elected) {
          if (leader) mbar_expect_tx(&b_full[sb], kTxScale * kTapsPerStage * tap_bytes);
          uint8_t* dst = smB + (size_t)sb * (kTapsPerStage * tap_bytes);
          for (int u = 0; u < kTapsPerStage; ++u) {
            const int t = g * kTapsPerStage + u;
            const int tap = (t == 0) ? 4 : (t <= 4 ? t - 1 : t);  // centre first, then 0,1,2,3,5,6,7,8
            for (int i = 0; i < nb; i += 16) {
              if constexpr (PAIR)
                tma_load_2d_2sm(dst + u * tap_bytes + i * kRowBytes, &tmW, &b_full[sb], tap * args.cin + j * CK, nb0 + i);
              else
                tma_load_2d(dst + u * tap_bytes + i * kRowBytes, &tmW, &b_full[sb], tap * args.cin + j * CK, nb0 + i);
            }
          }
        }
        __syncwarp();
        if (++sb == stages) { sb = 0; phb ^= 1; }
      }
    }
    if constexpr (PAIR) {
      // producer tail: the leader's commits arrive on BOTH CTAs' empty barriers; this CTA must not leave (and hand its
      // shared memory to the next layer's CTA) with one of those arrivals still in flight
      for (int i = 0; i < stages; ++i) {
        mbar_wait(&b_empty[sb], phb ^ 1);
        if (++sb == stages) { sb = 0; phb ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (PAIR: the leader issues for both CTAs) =====================
    // One warp issues every MMA of the tile, and at these tile widths an MMA occupies the tensor pipe for only 8 .. 64
    // cycles: the issue loop itself is the critical path of the kernel.  (First version: one mbarrier wait, one mask
    // lookup and ~130 SASS instructions per filter tap — measured 0.2 us per tap, 7.2 of a layer's 10.4 us at n = 1.)
    // Hence: the nine taps fully unrolled with their row shifts as immediates, the lane masks of all taps fetched
    // once per CTA, and one barrier wait / one commit per stage of three taps (12 MMAs).
    constexpr int kMaskWords = PAIR ? 8 : 4;   // 32 rows each; PAIR: words 4-7 are the peer CTA's rows
    const uint32_t idesc = make_idesc_16(PAIR ? 2 * kTileM : kTileM, nt, F16 ? 1 : 0);
    // pixel-in-image index of the tile's first row; the lane mask of 32 consecutive rows depends only on
    // (tap, that index mod 361)
    const int q0 = m0 % kPix;
    uint32_t msk[9][8];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        int q = q0 + 32 * w;
        if (q >= kPix) q -= kPix;  // q0 < 361 and 32*w <= 224
        msk[tap][w] = w < kMaskWords ? c_edge_mask_sb[tap][q] : 0u;
      }
    const uint32_t tap_desc_step = tap_bytes >> 4;  // descriptor start-address units (16 B) between two taps of a stage
    int sb = 0;
    uint32_t phb = 0;
#pragma unroll 1
    for (int j = 0; j < chunks; ++j) {
      mbar_wait(&a_full[j], 0);
      if (j == 0) UG_SB_STAMP(3);
      if (j == chunks - 1) UG_SB_STAMP(4);
      const uint64_t adesc0 = make_kmajor_desc<kRowBytes>(smem_u32(smA + j * C::kASlot));
#pragma unroll
      for (int g = 0; g < 9 / kTapsPerStage; ++g) {
        mbar_wait(&b_full[sb], phb);
        tc_fence_after();
        const uint64_t bdesc0 = make_kmajor_desc<kRowBytes>(smem_u32(smB + (size_t)sb * (kTapsPerStage * tap_bytes)));
        if (elected) {
#pragma unroll
          for (int u = 0; u < kTapsPerStage; ++u) {
            constexpr int kOrder[9] = {4, 0, 1, 2, 3, 5, 6, 7, 8};
            const int tap = kOrder[g * kTapsPerStage + u];
            const int shift = kHalo + (tap / 3 - 1) * kBoard + (tap % 3 - 1);  // rows: a compile-time constant
            const uint64_t adesc = adesc0 + (uint64_t)(shift * (kRowBytes / 16));
            const uint64_t bdesc = bdesc0 + (uint64_t)(u * tap_desc_step);
#pragma unroll
            for (int k = 0; k < kMmaPerStep; ++k) {
              if (tap == 4)  // the centre tap never needs a mask; it goes first and initialises the accumulator
                umma_f16<PAIR ? 2 : 1>(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((j | k) != 0));
              else if constexpr (PAIR)
                umma_f16_2sm_masked(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, 1u, msk[tap]);
              else
                umma_f16_1sm_masked(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, 1u, msk[tap]);
            }
          }
          if constexpr (PAIR) umma_commit_2sm(&b_empty[sb], 3);  // frees the stage in both CTAs
          else umma_commit(&b_empty[sb]);
        }
        __syncwarp();
        if (++sb == stages) { sb = 0; phb ^= 1; }
      }
    }
    if (elected) {
      if constexpr (PAIR) umma_commit_2sm(acc_full, 3);
      else umma_commit(acc_full);
    }
    __syncwarp();
    UG_SB_STAMP(5);
  }

  // ===================== epilogue: every warp =====================
  {
    const int quarter = warp & 3;            // TMEM lanes [32*quarter, +32)
    const int grp = warp < 4 ? 1 : 0;        // the role-free warps 4-7 take the even 16-column steps
    const int total_steps = nt >> 4;
    // Which 16-column steps this warp takes.  Direct stores (narrow tiles): the two warps of a lane quarter alternate
    // steps.  TMA stores (tiles of 64 columns and more): they alternate PAIRS of steps, so that each warp owns whole
    // 32-column (64-byte) boxes of the output tensor map.
    const bool tma_out = args.tma_out != 0;
    auto step_of = [&](int s) { return tma_out ? 4 * (s >> 1) + 2 * grp + (s & 1) : 2 * s + grp; };
    const long row = (long)m0 + quarter * 32 + lane;
    const bool row_ok = row < args.M;
    const float act_floor = args.relu ? 0.f : -3.0e38f;
    const bool has_res = args.res_hi != nullptr, has_res_lo = args.res_lo != nullptr, has_out_lo = args.out_lo != nullptr;
    const size_t row_off = (size_t)(row_ok ? row : 0) * args.cout + n0;
    pdl_wait();  // the residual was written two layers back; every warp must be ordered behind it
    uint4 rh[kMaxEpiSteps][2], rl[kMaxEpiSteps][2];
    if (has_res) {
#pragma unroll
      for (int s = 0; s < kMaxEpiSteps; ++s) {
        const int step = step_of(s);
        if (step < total_steps && row_ok) {
          const uint4* p = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(args.res_hi) + row_off + step * 16);
          rh[s][0] = __ldg(p);
          rh[s][1] = __ldg(p + 1);
          if (has_res_lo) {
            const uint4* q = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(args.res_lo) + row_off + step * 16);
            rl[s][0] = __ldg(q);
            rl[s][1] = __ldg(q + 1);
          }
        }
      }
    }
    // bias of this warp's columns: fetched before the accumulator is ready, like the residual
    float2 bs[kMaxEpiSteps][8];
#pragma unroll
    for (int s = 0; s < kMaxEpiSteps; ++s) {
      const int step = step_of(s);
      if (step < total_steps) {
#pragma unroll
        for (int e = 0; e < 8; ++e) bs[s][e] = __ldg(reinterpret_cast<const float2*>(args.bias + n0 + step * 16) + e);
      }
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // PAIR: every MMA has completed, so the leader no longer reads the peer's shared memory and every multicast commit
    // has been triggered: from here on the two CTAs only have to keep each other alive until both have seen this
    // point.  Arrive on the cluster barrier now, wait for it at the very end — its latency hides behind the epilogue
    // (a full barrier at the end cost 0.7 us per layer on the critical path to the next layer's griddepcontrol.wait).
    if constexpr (PAIR) cluster_arrive_relaxed();  // liveness only: no data passes between the CTAs through memory
    if (warp == 4) UG_SB_STAMP(6);
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // all of this warp's accumulator columns in one go: up to four tcgen05.ld in flight, one wait
    uint32_t v[kMaxEpiSteps][16];
#pragma unroll
    for (int s = 0; s < kMaxEpiSteps; ++s) {
      const int step = step_of(s);
      if (step < total_steps) tmem_ld_32x16(t_row + step * 16, v[s]);   // warp-uniform condition
    }
    tmem_ld_wait();
    float out_max = 0.f;  // fp16 range watch
    // TMA-store staging: the weight ring is free once the accumulator is complete (every stage has been consumed);
    // per warp two slots of [hi | lo] x (32 rows x 64 B, SWIZZLE_64B as the output map expects)
    uint8_t* stage = smB + (size_t)warp * (2 * kStageSlot);
    const int sw = (lane >> 1) & 3;          // 16-byte chunk index ^= (row >> 1) & 3
    const int row0 = m0 + quarter * 32;      // first row of this warp's boxes
#pragma unroll
    for (int s = 0; s < kMaxEpiSteps; ++s) {
      const int step = step_of(s);
      if (step < total_steps && (row_ok || tma_out)) {
        const int c0 = step * 16;
        uint32_t ph[8], pl[8];
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(&rh[s][0]);
        const uint32_t* lw = reinterpret_cast<const uint32_t*>(&rl[s][0]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float f0 = __uint_as_float(v[s][2 * e]) + bs[s][e].x;
          float f1 = __uint_as_float(v[s][2 * e + 1]) + bs[s][e].y;
          if (has_res) {
            const float2 r2 = unpack2_t<F16>(rw[e]);
            f0 += r2.x;
            f1 += r2.y;
            if (has_res_lo) {
              const float2 l2 = unpack2_t<F16>(lw[e]);
              f0 += l2.x;
              f1 += l2.y;
            }
          }
          f0 = fmaxf(f0, act_floor);
          f1 = fmaxf(f1, act_floor);
          if (!row_ok) f0 = f1 = 0.f;   // (TMA stores) rows past M inside a box: defined zeros, never counted below
          if constexpr (F16) out_max = fmaxf(out_max, fmaxf(fabsf(f0), fabsf(f1)));
          ph[e] = pack2_sat<F16>(f0, f1);
          if (has_out_lo) {
            const float2 hi = unpack2_t<F16>(ph[e]);
            pl[e] = pack2_sat<F16>(f0 - hi.x, f1 - hi.y);
          }
        }
        if (tma_out) {
          uint8_t* slot = stage + (s >> 1) * kStageSlot;
          uint8_t* orow = slot + lane * 64;
          const int q0 = (s & 1) * 2;        // this step's two 16-byte chunks of the 64-byte box row
          *reinterpret_cast<uint4*>(orow + ((q0 ^ sw) << 4)) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          *reinterpret_cast<uint4*>(orow + (((q0 + 1) ^ sw) << 4)) = make_uint4(ph[4], ph[5], ph[6], ph[7]);
          if (has_out_lo) {
            *reinterpret_cast<uint4*>(orow + kStageHalf + ((q0 ^ sw) << 4)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            *reinterpret_cast<uint4*>(orow + kStageHalf + (((q0 + 1) ^ sw) << 4)) = make_uint4(pl[4], pl[5], pl[6], pl[7]);
          }
          if (s & 1) {                       // both halves of the box are staged
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (row0 < args.M) {           // rows past M inside the box hold zeros (the map's height is the allocation)
                tma_store_2d(&tmOhi, slot, n0 + c0 - 16, row0);
                if (has_out_lo) tma_store_2d(&tmOlo, slot + kStageHalf, n0 + c0 - 16, row0);
              }
              bulk_commit_group();
            }
          }
        } else {
          uint4* o = reinterpret_cast<uint4*>(static_cast<uint16_t*>(args.out_hi) + row_off + c0);
          o[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          o[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
          if (has_out_lo) {
            uint4* ol = reinterpret_cast<uint4*>(static_cast<uint16_t*>(args.out_lo) + row_off + c0);
            ol[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            ol[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
          }
        }
      }
    }
    if (tma_out && lane == 0) bulk_wait_group_all();  // the stores have completed (not merely read their staging)
    if constexpr (F16) {
      if (out_max > 65504.f && args.sat_count != nullptr) atomicAdd(args.sat_count, 1ull);
    }
  }
  if (warp == 4) UG_SB_STAMP(7);
  tc_fence_before();
  // PAIR: neither CTA may leave (free its shared memory, barriers, TMEM) while the other can still touch them: the
  // barrier every thread arrived on after the accumulator was complete
  __syncthreads();
  if constexpr (PAIR) cluster_wait_acquire();
  if (warp == 2) tmem_dealloc<PAIR ? 2 : 1>(tmem_base, nt < 32 ? 32u : (uint32_t)nt);
  if (warp == 0) {
    UG_SB_STAMP(10);
    if (args.prof != nullptr && lane == 0)
      args.prof[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + 9] = globaltimer_ns();
  }
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
  const cudaError_t e = cudaMemcpyToSymbol(c_edge_mask_sb, h.data(), h.size() * 4);
  cudaSetDevice(cur);
  if (e != cudaSuccess) return false;
  g_mask_done[device] = true;
  return true;
}

template <int CK, bool F16, bool PAIR>
cudaError_t launch_sb(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmOhi, const CUtensorMap& tmOlo,
                      const ConvSbArgs& a, cudaStream_t stream) {
  auto kfn = conv3x3_sb_kernel<CK, F16, PAIR>;
  // PAIR: more than half of an SM's shared memory even for the small stem configuration, so that two CTAs never share
  // an SM — the pair's MMAs address both CTAs' accumulators with one TMEM address, which needs both allocations to
  // start at the same column (checked in the kernel)
  constexpr int smem = (PAIR && SbCfg<CK>::kSmem < 117 * 1024) ? 117 * 1024 : SbCfg<CK>::kSmem;
  static std::atomic<uint64_t> attr_done{0};  // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  unsigned m_tiles = (unsigned)((a.M + kTileM - 1) / kTileM);
  if (PAIR) m_tiles = (m_tiles + 1u) & ~1u;  // whole pairs: the odd tile's partner computes rows past M and stores nothing
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(m_tiles, (unsigned)(a.cout / a.nt));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (a.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kfn, tmA, tmW, tmOhi, tmOlo, a);
}

template <int CK, bool F16>
cudaError_t dispatch_sb(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmOhi, const CUtensorMap& tmOlo,
                        const ConvSbArgs& a, cudaStream_t stream) {
  return a.pair ? launch_sb<CK, F16, true>(tmA, tmW, tmOhi, tmOlo, a, stream)
                : launch_sb<CK, F16, false>(tmA, tmW, tmOhi, tmOlo, a, stream);
}

// CTAs of the grid for tile width nt
int grid_ctas(int m_tiles, int cout, int nt, bool pair) {
  return (pair ? (m_tiles + 1) / 2 * 2 : m_tiles) * (cout / nt);
}

}  // namespace

bool conv3x3_sb_prepare(int device) { return upload_masks(device); }

int conv3x3_sb_pick_nt(int M, int cout, int sms, int pair) {
  const int m_tiles = (M + kTileM - 1) / kTileM;
  // The smallest tile width (largest grid) that still fits one wave.  16-column tiles only pay their 144 tiny MMAs
  // back when nothing else fills the machine, so 32 is the floor unless asked for explicitly (and a CTA pair stages
  // nt / 2 weight rows per CTA in 16-row TMA boxes: 32 is its hard floor).  CTA pairs: 32 columns only for one or two
  // positions; from there 64-column tiles on half as many CTAs are as fast or faster (measured at n = 3, 4, 6: -1, -1,
  // -4 %), every MMA re-reads its 128 x 16 A operand from shared memory, which narrow tiles amortise badly.
  const int floor_nt = pair && m_tiles > kPairNarrowTiles ? 64 : 32;
  for (int nt = floor_nt; nt <= 128; nt <<= 1)
    if (cout % nt == 0 && grid_ctas(m_tiles, cout, nt, pair != 0) <= sms) return nt;
  for (int nt = 128; nt >= (pair ? 32 : 16); nt >>= 1)
    if (cout % nt == 0) return nt;
  return 0;
}

int conv3x3_sb_pick_pair(int M, int cout, int sms) {
  // A CTA pair halves the weight bytes each SM pulls from L2 and the shared-memory operand reads per MMA row.  Measured
  // on the 20x256 net (tests/gpu_small_batch_probe.py --pairs) against single CTAs at their best width:
  // -11 % at one or two positions, -13 % from the batch size at which single-CTA tiles have to be 128 columns wide
  // (n >= 14 for 256 filters), within +-1 % in between (-4 % at n = 6): pairs whenever the channel count allows.
  (void)M;
  (void)sms;
  static const int force = [] {
    const char* e = getenv("UGEVAL_SB_PAIR_MIN_TILES");   // experiments: pairs from this many tiles up only
    return e ? atoi(e) : 0;
  }();
  if (cout % 32 != 0) return 0;
  if (force > 0) return (M + kTileM - 1) / kTileM >= force ? 1 : 0;
  return 1;
}

cudaError_t conv3x3_sb_launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap* tmOhi,
                              const CUtensorMap* tmOlo, ConvSbArgs a, int ck, int device, int sms, cudaStream_t stream) {
  if (ck != 64 && ck != 32) return cudaErrorInvalidValue;
  if (a.cin % ck != 0 || a.cin / ck > (ck == 64 ? 4 : 1)) return cudaErrorInvalidValue;
  if (a.pair < 0) a.pair = conv3x3_sb_pick_pair(a.M, a.cout, sms);
  // a forced width that does not divide this layer's channel count (a narrow net) falls back to the automatic choice
  if (a.nt == 0 || (a.nt > 0 && a.cout % a.nt != 0)) a.nt = conv3x3_sb_pick_nt(a.M, a.cout, sms, a.pair);
  if (a.pair && a.nt == 16) a.pair = 0;  // 8 weight rows per CTA: below the TMA box
  if (a.nt != 16 && a.nt != 32 && a.nt != 64 && a.nt != 128) return cudaErrorInvalidValue;
  if (a.cout % a.nt != 0 || a.M <= 0 || !a.out_hi || !a.bias) return cudaErrorInvalidValue;
  if (a.res_lo && !a.res_hi) return cudaErrorInvalidValue;
  if (!upload_masks(device)) return cudaErrorUnknown;  // no-op after conv3x3_sb_prepare()
  const int nb = a.pair ? a.nt / 2 : a.nt;                       // weight rows a CTA stages per tap
  const int need = (9 / kTapsPerStage) * (a.cin / ck);          // stages a tile consumes
  const int fit = kBRing / (kTapsPerStage * nb * ck * 2);        // stages the ring holds (>= 2: 128 rows -> 48 KB each)
  a.stages = fit < need ? fit : need;
  if (a.stages > kMaxStages) a.stages = kMaxStages;
  // 128-column tiles leave through shared-memory staging and TMA stores when the caller has output maps (box {32
  // channels, 32 rows}); tma_out > 0 asks for them from 64 columns up (the kernel supports that; it is not faster).
  static const bool tma_allowed = [] { const char* e = getenv("UGEVAL_SB_TMA_OUT"); return !(e && e[0] == '0'); }();
  a.tma_out = (tma_allowed && a.tma_out >= 0 && tmOhi != nullptr && (a.out_lo == nullptr || tmOlo != nullptr) &&
               a.nt >= (a.tma_out > 0 ? 64 : 128)) ? 1 : 0;   // measured: +1.5 % at 128 columns, -1.5 % at 64
  const CUtensorMap& oh = tmOhi ? *tmOhi : tmA;   // unused maps still have to be valid kernel parameters
  const CUtensorMap& ol = tmOlo ? *tmOlo : oh;
  if (ck == 64)
    return a.f16 ? dispatch_sb<64, true>(tmA, tmW, oh, ol, a, stream) : dispatch_sb<64, false>(tmA, tmW, oh, ol, a, stream);
  return a.f16 ? dispatch_sb<32, true>(tmA, tmW, oh, ol, a, stream) : dispatch_sb<32, false>(tmA, tmW, oh, ol, a, stream);
}

}  // namespace ug
