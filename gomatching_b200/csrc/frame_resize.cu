// Test-time frame resize on the GPU, bit-identical to Pillow's 8-bit bilinear resample -- the
// `ResizeShortestEdge(...).get_transform(x).apply_image(x)` step of the reference's predictors
// (gomatching/text_track_visualizer.py:283-284, :318-319; detectron2 ResizeTransform -> PIL.Image.resize(BILINEAR)).
//
// Pillow resamples uint8 images in two separable passes with 22-bit fixed-point coefficients
// (libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc):
//     out = clip8((2^21 + sum_i in[first + i] * k[i]) >> 22)      horizontal pass first, its uint8 result feeds the vertical one
// The coefficient tables (bounds, k) are built on the host exactly as Pillow builds them (gomatching_b200/video/resize.py);
// this file is the two passes.  HBM-bound byte work: one thread per output pixel (all channels), rows of the source are
// read contiguously in the horizontal pass and coalesced across x in the vertical one.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msda_b200.h"

namespace msda {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= kPrecisionBits;
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// axis 1 (horizontal): in (N, H, W, C) -> out (N, H, out_size, C);  axis 0 (vertical): -> out (N, out_size, W, C)
template <int C>
__global__ void __launch_bounds__(256) resample_u8_kernel(const unsigned char* __restrict__ in, int N, int H, int W, int axis,
                                                          const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                          int out_size, unsigned char* __restrict__ out) {
  const int OH = axis == 0 ? out_size : H, OW = axis == 1 ? out_size : W;
  const long long total = (long long)N * OH * OW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % OW);
    const long long r = i / OW;
    const int y = (int)(r % OH), n = (int)(r / OH);
    const int o = axis == 1 ? x : y;
    const int first = bounds[2 * o], cnt = bounds[2 * o + 1];
    const int* k = kk + (long long)o * ksize;
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (kPrecisionBits - 1);
    const unsigned char* src = in + ((long long)n * H * W) * C;
    const long long step = axis == 1 ? C : (long long)W * C;
    const unsigned char* p = axis == 1 ? src + ((long long)y * W + first) * C : src + ((long long)first * W + x) * C;
    for (int j = 0; j < cnt; ++j, p += step) {
      const int w = k[j];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] += (int)p[c] * w;
    }
    unsigned char* q = out + (((long long)n * OH + y) * OW + x) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) q[c] = clip8(acc[c]);
  }
}

}  // namespace
}  // namespace msda

extern "C" int msda_b200_resample_u8_hwc(const unsigned char* in, int N, int H, int W, int C, int axis, const int* bounds,
                                         const int* coeffs, int ksize, int out_size, unsigned char* out, void* stream) {
  using namespace msda;
  if (!in || !bounds || !coeffs || !out) return MSDA_E_NULLPTR;
  if (N <= 0 || H <= 0 || W <= 0 || ksize <= 0 || out_size <= 0 || (axis != 0 && axis != 1)) return MSDA_E_DIMS;
  if (C != 3 && C != 1 && C != 4) return MSDA_E_UNSUPPORTED;
  const long long total = (long long)N * (axis == 0 ? out_size : H) * (axis == 1 ? out_size : W);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  cudaStream_t s = (cudaStream_t)stream;
  if (C == 3) resample_u8_kernel<3><<<(int)blocks, 256, 0, s>>>(in, N, H, W, axis, bounds, coeffs, ksize, out_size, out);
  else if (C == 1) resample_u8_kernel<1><<<(int)blocks, 256, 0, s>>>(in, N, H, W, axis, bounds, coeffs, ksize, out_size, out);
  else resample_u8_kernel<4><<<(int)blocks, 256, 0, s>>>(in, N, H, W, axis, bounds, coeffs, ksize, out_size, out);
  return (int)cudaGetLastError();
}
