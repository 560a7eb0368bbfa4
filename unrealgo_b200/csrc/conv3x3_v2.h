// Launch interface of the second-generation tcgen05 3x3 convolution (conv3x3_v2.cu): input tile resident in
// shared memory, filter taps as shifted descriptors + lane masks, TMA-staged epilogue.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ug {

struct ConvV2Args {
  const float* bias;      // [cout] folded batch-norm offset (fp32)
  // fused head 1x1 convolutions (last tower layer only; see conv3x3_v2_can_fuse_heads): weights [head_channels][cout]
  // and offsets [head_channels] in fp32, output [M][head_channels] fp32 = relu(conv + offset).  head_out == null: off.
  const float* head_w;
  const float* head_b;
  float* head_out;
  int head_channels;
  int M;                  // n * 361 output pixels
  int cin;                // input channels as stored (multiple of ck)
  int cout;               // <= 256, multiple of 64
  int relu;
  int f16;                // operands/outputs are fp16 (else bf16)
  int has_res, has_res_lo, has_out_lo;
  int desc_mode;          // bring-up knob: bit 4 = no weight reloads after the first lap of the ring (wrong results: a
                          // timing probe for the weight traffic).  Bits 0-2 (descriptor base offset / unmasked MMAs /
                          // no tap shift) were probes of the first bring-up and are ignored since the issue loop was
                          // unrolled.
  // Tile-granular dependencies between consecutive layers (all null/0: whole-grid dependency through
  // griddepcontrol.wait).  Counters are per 256-row tile, +16 when a layer has stored the tile (one count per epilogue
  // warp of the CTA pair), zeroed by the caller at the start of every pass through the tower.
  const uint32_t* flags_in;   // previous layer's counters: tile t may start once tiles t-1..t+1 reached target_in
  const uint32_t* flags_res;  // counters of the layer that wrote the residual tile (two layers back)
  uint32_t* flags_out;        // this layer's counters
  uint32_t target_in, target_res;
  int tile_rot;               // rotation of the cluster -> first-tile assignment (0 .. clusters-1)
  int pdl;                // launch with programmatic stream serialization (overlap the prologue and the weight
                          // prefetch with the previous kernel's tail)
  int num_tiles;          // filled by the launcher
  unsigned long long* sat_count;  // fp16 mode: += 1 per (thread, tile) that had to clamp an output to +-65504 (or null)
  long long* prof;        // optional [clusters][64] cycle counters, zeroed by the caller (bring-up instrumentation), else null
};

// Tensor maps (see tmap.h): a = halo tile over the input [rows][cin] (box {ck, 168}); w = packed weights
// (box {ck, cout/2}); res_* = residual [rows][cout] (box {32, 128}, SWIZZLE_64B); out_* = outputs (box {32, 32}).
// res_* / out_lo may be null when the matching has_* flag is 0.
struct ConvV2Maps {
  const CUtensorMap* a;
  const CUtensorMap* w;
  const CUtensorMap* res_hi;
  const CUtensorMap* res_lo;
  const CUtensorMap* out_hi;
  const CUtensorMap* out_lo;
};

cudaError_t conv3x3_v2_launch(const ConvV2Maps& maps, ConvV2Args a, int ck, int device, int sms,
                              cudaStream_t stream);
int conv3x3_v2_smem_bytes(int ck);
// whether a launch with these flags and `head_channels` fused head outputs has a kernel instantiation
bool conv3x3_v2_can_fuse_heads(int ck, int has_res, int has_res_lo, int has_out_lo, int head_channels);
// uploads the lane-mask table for `device` (cudaMalloc + copy): call once outside any stream capture
bool conv3x3_v2_prepare(int device);
// how many clusters of `cluster_size` CTAs of the tower kernel the current device can hold at once (planning probe)
int conv3x3_v2_max_active_clusters(int cluster_size);

}  // namespace ug
