// Host-side CUtensorMap builders (driver entry points resolved at run time: no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <string>

namespace ug {

// NHWC bf16 activations [n][19][19][c] described as {C, W, H, N}; im2col box = 128 pixels x ck channels,
// 3x3 SAME window (corners -1/-1), halo zero-filled.  Returns false and sets err on failure.
bool make_tmap_act_im2col(CUtensorMap* out, const void* base, int n, int c, int ck, std::string* err);

// Packed weights [cout][ktot] bf16 (K-major); box = box_rows x ck.
bool make_tmap_weights(CUtensorMap* out, const void* base, int cout, int ktot, int ck, int box_rows,
                       std::string* err);

}  // namespace ug
