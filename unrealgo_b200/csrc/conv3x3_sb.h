// Launch interface of the small-batch tcgen05 3x3 convolution (conv3x3_sb.cu): the tower kernel for the batch sizes
// the reference's own engine can issue (1 .. MAX_BATCHES = 16, funcapproximator/DlTFNetworkEvaluator.h:15) and for
// single-game search, where conv3x3_v2's 256-row CTA-pair tiles leave most of the 148 SMs idle.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ug {

struct ConvSbArgs {
  const float* bias;      // [cout] folded batch-norm offset (fp32)
  const void* res_hi;     // [rows][cout] 16-bit skip connection (or null)
  const void* res_lo;     // optional low part of the skip connection
  void* out_hi;           // [rows][cout] 16-bit result
  void* out_lo;           // optional: 16-bit (result - hi)
  int M;                  // n * 361 output pixels
  int cin;                // input channels as stored (multiple of ck)
  int cout;               // multiple of nt
  int relu;
  int f16;                // operands/outputs are fp16 (else bf16)
  int nt;                 // output channels per CTA: 16, 32, 64 or 128 (0 = choose from the grid size)
  int pair;               // 1 = CTA pairs (cta_group::2, 256-row x nt tiles, nt >= 32), 0 = single CTAs, < 0 = choose
  int tma_out;            // < 0: per-thread stores only; 0: staged TMA stores for 128-column tiles; > 0: from 64 columns
  int pdl;                // launch with programmatic stream serialization
  int stages;             // filled by the launcher: depth of the weight ring
  unsigned long long* sat_count;  // fp16 mode: += 1 per thread that had to clamp an output to +-65504 (or null)
  long long* prof;        // optional [ctas][16] time stamps of this launch (bring-up instrumentation), else null
};

// tmA: halo-tile map over the input [rows][cin] (box {ck, 168}, the same map conv3x3_v2 uses);
// tmW: packed weights [cout][9*cin] with box {ck, 16} (16 output channels per TMA load, nt/16 loads per stage);
// tmOhi / tmOlo (optional): maps over out_hi / out_lo [rows][cout] with box {32, 32} — the ones conv3x3_v2 stores
// through; with them, 128-column tiles are written by TMA from shared-memory staging.
cudaError_t conv3x3_sb_launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap* tmOhi,
                              const CUtensorMap* tmOlo, ConvSbArgs a, int ck, int device, int sms, cudaStream_t stream);
// output channels per tile the launcher would pick for M rows (largest grid that still fits one wave of `sms` CTAs)
int conv3x3_sb_pick_nt(int M, int cout, int sms, int pair);
// 1 when the launcher's automatic choice for M rows is the CTA-pair variant
int conv3x3_sb_pick_pair(int M, int cout, int sms);
// uploads the lane-mask table for `device`: call once outside any stream capture
bool conv3x3_sb_prepare(int device);

}  // namespace ug
