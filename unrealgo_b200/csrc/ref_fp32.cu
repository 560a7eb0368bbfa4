// fp32 validation tower: direct 3x3 SAME convolution on CUDA cores, NHWC fp32, HWIO weights exactly as
// stored in the checkpoint (batch-norm folded in fp32).  Slow by design — it exists so the bf16 tcgen05 path
// can be checked layer by layer on the GPU and so parity against the oracle can be stated at fp32 tolerance.
#include "kernels.h"

namespace ug {

namespace {

// one thread per (pixel, cout); cout is the fastest index so weight loads and stores are coalesced
__global__ void __launch_bounds__(256) conv3x3_f32_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                          const float* __restrict__ bias,
                                                          const float* __restrict__ res, float* __restrict__ out,
                                                          long total, int cin, int cout, int relu) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = (int)(idx % cout);
  const long m = idx / cout;
  const int n = (int)(m / UG_NUM_POINTS);
  const int rem = (int)(m - (long)n * UG_NUM_POINTS);
  const int y = rem / UG_BOARD_SIZE, x = rem - y * UG_BOARD_SIZE;
  float acc = 0.f;
  for (int r = 0; r < 3; ++r) {
    const int yy = y + r - 1;
    if (yy < 0 || yy >= UG_BOARD_SIZE) continue;
    for (int s = 0; s < 3; ++s) {
      const int xx = x + s - 1;
      if (xx < 0 || xx >= UG_BOARD_SIZE) continue;
      const float* ip = in + ((size_t)n * UG_NUM_POINTS + yy * UG_BOARD_SIZE + xx) * cin;
      const float* wp = w + (size_t)(r * 3 + s) * cin * cout + co;
      for (int ci = 0; ci < cin; ++ci) acc = fmaf(__ldg(ip + ci), __ldg(wp + (size_t)ci * cout), acc);
    }
  }
  acc += __ldg(bias + co);
  if (res != nullptr) acc += __ldg(res + idx);
  if (relu) acc = fmaxf(acc, 0.f);
  out[idx] = acc;
}

// range calibration: the largest magnitude in a tensor (non-negative floats order like their bit patterns)
__global__ void __launch_bounds__(256) max_abs_f32_kernel(const float* __restrict__ p, size_t count, float* slot) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const float v = fabsf(p[i]);
    m = (v > m || v != v) ? v : m;   // a NaN wins, so that it is seen
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, m, o);
    m = (t > m || t != t) ? t : m;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(slot), __float_as_uint(m));
}

}  // namespace

void launch_max_abs_f32(const float* p, size_t count, float* slot, cudaStream_t s) {
  if (count == 0) return;
  unsigned grid = (unsigned)((count + 255) / 256);
  if (grid > 1184) grid = 1184;
  max_abs_f32_kernel<<<grid, 256, 0, s>>>(p, count, slot);
}

void launch_conv3x3_f32(const float* in, const float* w_hwio, const float* bias, const float* res, float* out,
                        int n, int cin, int cout, int relu, cudaStream_t s) {
  if (n <= 0) return;
  const long total = (long)n * UG_NUM_POINTS * cout;
  conv3x3_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, w_hwio, bias, res, out, total, cin, cout,
                                                                     relu);
}

}  // namespace ug
