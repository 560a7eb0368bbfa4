#include "tmap.h"

#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <mutex>

namespace ug {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_tiled = nullptr;
static EncodeIm2colFn g_im2col = nullptr;
static int g_driver_version = 0;

static bool resolve(std::string* err) {
  static std::once_flag once;
  static bool ok = false;
  std::call_once(once, [&] {
    void* f1 = nullptr;
    void* f2 = nullptr;
    cudaDriverEntryPointQueryResult q1, q2;
    cudaError_t e1 = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f1, cudaEnableDefault, &q1);
    cudaError_t e2 = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &q2);
    cudaDriverGetVersion(&g_driver_version);
    if (e1 == cudaSuccess && e2 == cudaSuccess && f1 && f2) {
      g_tiled = reinterpret_cast<EncodeTiledFn>(f1);
      g_im2col = reinterpret_cast<EncodeIm2colFn>(f2);
      ok = true;
    }
  });
  if (!ok && err) *err = "cuTensorMapEncode{Tiled,Im2col} driver entry points unavailable";
  return ok;
}

bool make_tmap_act_im2col(CUtensorMap* out, const void* base, int n, int c, int ck, std::string* err) {
  if (!resolve(err)) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)c, 19, 19, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)c * 2 * 19, (cuuint64_t)c * 2 * 361};
  const int lower[2] = {-1, -1};  // {W, H}: base pixel starts one left/above (pad 1)
  const int upper[2] = {-1, -1};  // pad - (filter-1)*dilation = 1 - 2
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = (ck == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = g_im2col(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower,
                        upper, (cuuint32_t)ck, 128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeIm2col failed: " + std::to_string((int)r);
    return false;
  }
  // Drivers up to 13.1 mis-encode im2col maps of tensors smaller than 128 KiB (same guard CUTLASS applies in
  // make_im2col_tma_copy_desc): clear bit 21 of the second descriptor word.
  if (g_driver_version <= 13010 && (size_t)n * 361 * c * 2 < 131072) {
    uint64_t w[16];
    memcpy(w, out, sizeof(w));
    w[1] &= ~(1ull << 21);
    memcpy(out, w, sizeof(w));
  }
  return true;
}

bool make_tmap_weights(CUtensorMap* out, const void* base, int cout, int ktot, int ck, int box_rows,
                       std::string* err) {
  if (!resolve(err)) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)cout};
  const cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  const cuuint32_t box[2] = {(cuuint32_t)ck, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = (ck == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = g_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r);
    return false;
  }
  return true;
}

bool make_tmap_rows(CUtensorMap* out, const void* base, long rows, int cols, int box_cols, int box_rows,
                    std::string* err) {
  if (!resolve(err)) return false;
  if (box_cols != 32 && box_cols != 64) {
    if (err) *err = "make_tmap_rows: box_cols must be 32 or 64 (64 B / 128 B swizzle span)";
    return false;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = (box_cols == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = g_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r);
    return false;
  }
  return true;
}

}  // namespace ug
