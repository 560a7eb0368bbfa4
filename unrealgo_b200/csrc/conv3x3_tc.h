// Launch interface of the tcgen05 implicit-GEMM 3x3 convolution (conv3x3_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace ug {

struct ConvArgs {
  const float* bias;              // [cout] folded batch-norm offset (fp32)
  const __nv_bfloat16* residual;     // [M][cout] skip connection (high part) or nullptr
  const __nv_bfloat16* residual_lo;  // optional low part: skip = float(hi) + float(lo)
  __nv_bfloat16* out;                // [M][cout] bf16(result): the next conv's operand
  __nv_bfloat16* out_lo;             // optional: bf16(result - float(out)), keeps the residual stream from
                                     // being re-rounded to 8 mantissa bits every block
  int M;                          // n * 361 output pixels
  int cin;                        // input channels as stored (multiple of ck)
  int cout;                       // <= 256, multiple of 16 (32 for the 2-CTA variant)
  int relu;
  int f16;                        // operands/outputs are fp16 (else bf16); see numfmt.cuh
  int num_tiles;                  // filled by the launcher
};

// tmA: im2col map over the input activations, tmB: tiled map over packed weights (see tmap.h).
// ck = channels per k-step (64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B); ctas = 1 or 2 (cta_group).
cudaError_t conv3x3_tc_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, ConvArgs a, int ck, int ctas,
                              int sms, cudaStream_t stream);

}  // namespace ug
