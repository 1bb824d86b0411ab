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

// ---- vectorised passes (same arithmetic, same bits) ---------------------------------------------------------------------
// Vertical: a row is W*C bytes and the pass never mixes bytes of a row, so it is channel-agnostic: one thread owns four
// consecutive bytes of an output row (32-bit loads / stores, fully coalesced).  Needs W*C % 4 == 0.
__global__ void __launch_bounds__(256) resample_v_u8x4_kernel(const unsigned char* __restrict__ in, int N, int H, int row_words,
                                                              const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                              int OH, unsigned char* __restrict__ out) {
  const long long total = (long long)N * OH * row_words;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xw = (int)(i % row_words);
    const long long r = i / row_words;
    const int y = (int)(r % OH), n = (int)(r / OH);
    const int first = bounds[2 * y], cnt = bounds[2 * y + 1];
    const int* k = kk + (long long)y * ksize;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0, a3 = a0;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(in) + ((long long)n * H + first) * row_words + xw;
    for (int j = 0; j < cnt; ++j, p += row_words) {
      const uint32_t v = __ldg(p);
      const int w = k[j];
      a0 += (int)(v & 0xffu) * w; a1 += (int)((v >> 8) & 0xffu) * w; a2 += (int)((v >> 16) & 0xffu) * w; a3 += (int)(v >> 24) * w;
    }
    reinterpret_cast<uint32_t*>(out)[((long long)n * OH + y) * row_words + xw] =
        (uint32_t)clip8(a0) | ((uint32_t)clip8(a1) << 8) | ((uint32_t)clip8(a2) << 16) | ((uint32_t)clip8(a3) << 24);
  }
}

// Horizontal: one CTA per image row.  The source row is staged in shared memory with 16-byte loads, every thread
// computes output pixels from it into a shared output row placed at the destination's byte alignment, and the row
// leaves as aligned 32-bit words (edge bytes singly).  Needs W*C % 16 == 0 (the staged loads) -- 1280 and 1920 qualify.
template <int C>
__global__ void __launch_bounds__(256) resample_h_row_u8_kernel(const unsigned char* __restrict__ in, long long rows, int W,
                                                                const int* __restrict__ bounds, const int* __restrict__ kk,
                                                                int ksize, int OW, unsigned char* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_row[];
  const int in_bytes = W * C, out_bytes = OW * C;
  unsigned char* sIn = smem_row;
  unsigned char* sOut = smem_row + ((in_bytes + 15) & ~15);
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(in + row * in_bytes);
    for (int i = threadIdx.x; i < in_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sIn)[i] = __ldg(src + i);
    __syncthreads();
    unsigned char* dst = out + row * out_bytes;
    const int shift = (int)(reinterpret_cast<uintptr_t>(dst) & 3u);           // sOut[shift + i] <-> dst[i]
    for (int xx = threadIdx.x; xx < OW; xx += blockDim.x) {
      const int first = bounds[2 * xx], cnt = bounds[2 * xx + 1];
      const int* k = kk + (long long)xx * ksize;
      int acc[C];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = 1 << (kPrecisionBits - 1);
      const unsigned char* p = sIn + first * C;
      for (int j = 0; j < cnt; ++j, p += C) {
        const int w = k[j];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += (int)p[c] * w;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) sOut[shift + xx * C + c] = clip8(acc[c]);
    }
    __syncthreads();
    // bytes [0, head) and [tail, out_bytes) singly, the aligned words in between from shared memory
    const int head = (4 - shift) & 3, body_end = shift + out_bytes - ((shift + out_bytes) & 3);
    if (threadIdx.x < head && threadIdx.x < out_bytes) dst[threadIdx.x] = sOut[shift + threadIdx.x];
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst - shift);
    for (int wd = (shift + head) / 4 + threadIdx.x; wd * 4 < body_end; wd += blockDim.x)
      dw[wd] = reinterpret_cast<const uint32_t*>(sOut)[wd];
    for (int b = max(body_end, shift + head) + (int)threadIdx.x; b < shift + out_bytes; b += blockDim.x) dst[b - shift] = sOut[b];
    __syncthreads();
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
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  if (axis == 0 && aligned && ((long long)W * C) % 4 == 0) {                  // vertical: 4 bytes per thread
    const int row_words = W * C / 4;
    long long b4 = ((long long)N * out_size * row_words + 255) / 256;
    if (b4 > 148 * 32) b4 = 148 * 32;
    resample_v_u8x4_kernel<<<(int)b4, 256, 0, s>>>(in, N, H, row_words, bounds, coeffs, ksize, out_size, out);
    return (int)cudaGetLastError();
  }
  if (axis == 1 && aligned && C == 3 && ((long long)W * C) % 16 == 0) {       // horizontal: one staged row per CTA
    const size_t smem = (size_t)((W * C + 15) & ~15) + (size_t)out_size * C + 16;
    if (smem <= 48 * 1024) {
      const long long rows = (long long)N * H;
      resample_h_row_u8_kernel<3><<<(int)(rows < 148 * 16 ? rows : 148 * 16), 256, smem, s>>>(in, rows, W, bounds, coeffs, ksize,
                                                                                             out_size, out);
      return (int)cudaGetLastError();
    }
  }
  if (C == 3) resample_u8_kernel<3><<<(int)blocks, 256, 0, s>>>(in, N, H, W, axis, bounds, coeffs, ksize, out_size, out);
  else if (C == 1) resample_u8_kernel<1><<<(int)blocks, 256, 0, s>>>(in, N, H, W, axis, bounds, coeffs, ksize, out_size, out);
  else resample_u8_kernel<4><<<(int)blocks, 256, 0, s>>>(in, N, H, W, axis, bounds, coeffs, ksize, out_size, out);
  return (int)cudaGetLastError();
}
