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

// Generic 2-D tiled map over a row-major 16-bit matrix [rows][cols]; box = box_rows x box_cols; the swizzle
// follows the box's inner extent (64 B -> SWIZZLE_64B, 128 B -> SWIZZLE_128B).  Out-of-range rows/columns
// (including negative start coordinates) are zero-filled on load and clipped on store.
bool make_tmap_rows(CUtensorMap* out, const void* base, long rows, int cols, int box_cols, int box_rows,
                    std::string* err);

}  // namespace ug
