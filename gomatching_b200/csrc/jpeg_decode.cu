// JPEG frame decode on the device (the data format BEFORE the path).  Replaces the frame read of the
// reference's video loop -- `read_image(path, format="BGR")`, eval.py:324-327 = detectron2 -> PIL.Image.open ->
// convert("RGB") -> numpy -> BGR, i.e. Pillow's libjpeg-turbo with default settings -- bit for bit:
//   host   jpeg_entropy.h: markers + Huffman decoding -> quantised coefficient blocks (sequential by nature)
//   device jpeg_idct_kernel:   dequantise + jidctint.c's 13-bit integer inverse DCT (jpeg_idct_islow), range limit with
//                              the +128 level shift, 8 threads per 8x8 block, one 8-byte store per output row
//          jpeg_colour_kernel: jdsample.c's fancy (triangle) chroma upsampling evaluated per output pixel from the
//                              component planes + jdcolor.c's fixed-point YCbCr -> RGB, written as BGR or RGB HWC
// so a frame crosses PCIe as 2 bytes per DCT coefficient and never exists as pixels on the host.  The arithmetic is the
// published IJG release-6b algorithm that libjpeg-turbo reproduces exactly; parity is pinned on Pillow itself
// (tests/test_jpeg_decode.py: every pixel equal to PIL.Image.open for 4:4:4 / 4:2:2 / 4:2:0 / grey, qualities 2..100, odd
// sizes, restart markers, optimised Huffman tables) and on the CPU restatement oracle/jpeg_oracle.cpp.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msda_b200.h"
#include "jpeg_entropy.h"
#include "msda_launch.h"

namespace msda {
namespace {

struct JpegPlanes {
  int ncomp;
  int block_begin[4];          // first 8x8 block of component c in the coefficient array; [ncomp] = total
  int blocks_w[3];
  int quant[3];                // quantisation table of component c
  int pitch[3];                // bytes per row of the component plane (= blocks_w * 8)
  int cw[3], ch[3];            // real samples (downsampled_width / height)
  int hs[3], vs[3];            // upsampling factors to full resolution
  long long plane_offset[3];   // byte offset of the component plane in the plane buffer
};

constexpr int kCB = 13, kP1 = 2;
constexpr int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270, F_0_899976223 = 7373,
              F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137, F_1_961570560 = 16069,
              F_2_053119869 = 16819, F_2_562915447 = 20995, F_3_072711026 = 25172;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// jidctint.c: one 8-point pass of the LL&M inverse DCT on in[0..7] (already scaled as the pass expects); out[k] before the
// final descale
__device__ __forceinline__ void idct8(const int* in, int* o) {
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * F_0_541196100;
  int tmp2 = z1 + z3 * (-F_1_847759065);
  int tmp3 = z1 + z2 * F_0_765366865;
  int tmp0 = (in[0] + in[4]) * (1 << kCB), tmp1 = (in[0] - in[4]) * (1 << kCB);
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * F_1_175875602;
  tmp0 *= F_0_298631336; tmp1 *= F_2_053119869; tmp2 *= F_3_072711026; tmp3 *= F_1_501321110;
  z1 *= -F_0_899976223; z2 *= -F_2_562915447; z3 *= -F_1_961570560; z4 *= -F_0_390180644;
  z3 += z5; z4 += z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  o[0] = tmp10 + tmp3; o[7] = tmp10 - tmp3;
  o[1] = tmp11 + tmp2; o[6] = tmp11 - tmp2;
  o[2] = tmp12 + tmp1; o[5] = tmp12 - tmp1;
  o[3] = tmp13 + tmp0; o[4] = tmp13 - tmp0;
}

// the post-IDCT range-limit table addressed with (x & 1023): 128..255, 255 x 384, 0 x 384, 0..127 (jdmaster.c)
__device__ __forceinline__ unsigned idct_limit(int x) {
  x &= 1023;
  return x < 128 ? (unsigned)(x + 128) : (x < 512 ? 255u : (x < 896 ? 0u : (unsigned)(x - 896)));
}

constexpr int kBlocksPerCta = 32, kWs = 72;      // 32 8x8 blocks x 8 threads; workspace rows padded against bank conflicts

__global__ void __launch_bounds__(256) jpeg_idct_kernel(const int16_t* __restrict__ coef, const uint16_t* __restrict__ quant,
                                                        JpegPlanes g, unsigned char* __restrict__ planes) {
  __shared__ int ws[kBlocksPerCta * kWs];
  const int jb = threadIdx.x >> 3, i = threadIdx.x & 7;
  const int b = blockIdx.x * kBlocksPerCta + jb;
  const bool live = b < g.block_begin[g.ncomp];
  int c = 0;
  if (live) {
    while (c + 1 < g.ncomp && b >= g.block_begin[c + 1]) ++c;
    const int16_t* cf = coef + (size_t)b * 64;
    const uint16_t* q = quant + g.quant[c] * 64;
    int in[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) in[r] = (int)cf[8 * r + i] * (int)q[8 * r + i];       // column i
    idct8(in, o);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[jb * kWs + 8 * r + i] = descale(o[r], kCB - kP1);
  }
  __syncthreads();
  if (live) {
    int in[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) in[k] = ws[jb * kWs + 8 * i + k];                      // row i
    idct8(in, o);
    unsigned lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo |= idct_limit(descale(o[k], kCB + kP1 + 3)) << (8 * k);
      hi |= idct_limit(descale(o[k + 4], kCB + kP1 + 3)) << (8 * k);
    }
    const int lb = b - g.block_begin[c];
    const int by = lb / g.blocks_w[c], bx = lb % g.blocks_w[c];
    unsigned char* dst = planes + g.plane_offset[c] + (size_t)(by * 8 + i) * g.pitch[c] + bx * 8;
    *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
  }
}

// component c at full resolution for output pixel (y, x): jdsample.c fullsize / h2v1_fancy / h2v2_fancy / plain replication
__device__ __forceinline__ int sample_at(const unsigned char* __restrict__ planes, const JpegPlanes& g, int c, int y, int x) {
  const unsigned char* p = planes + g.plane_offset[c];
  const int pitch = g.pitch[c], hs = g.hs[c], vs = g.vs[c], cw = g.cw[c], ch = g.ch[c];
  if (hs == 1 && vs == 1) return p[(size_t)y * pitch + x];
  const int r = vs == 2 ? y >> 1 : y, i = hs == 2 ? x >> 1 : x;
  if (cw <= 2) return p[(size_t)r * pitch + i];                       // jinit_upsampler: fancy only if downsampled_width > 2
  const unsigned char* row0 = p + (size_t)r * pitch;
  if (vs == 2) {
    int rn = (y & 1) ? r + 1 : r - 1;
    rn = rn < 0 ? 0 : (rn > ch - 1 ? ch - 1 : rn);
    const unsigned char* row1 = p + (size_t)rn * pitch;
    const int cur = 3 * row0[i] + row1[i];
    if (!(x & 1)) return i == 0 ? (cur * 4 + 8) >> 4 : (cur * 3 + 3 * row0[i - 1] + row1[i - 1] + 8) >> 4;
    return i == cw - 1 ? (cur * 4 + 7) >> 4 : (cur * 3 + 3 * row0[i + 1] + row1[i + 1] + 7) >> 4;
  }
  const int cur = row0[i];
  if (!(x & 1)) return i == 0 ? cur : (cur * 3 + row0[i - 1] + 1) >> 2;
  return i == cw - 1 ? cur : (cur * 3 + row0[i + 1] + 2) >> 2;
}

__device__ __forceinline__ unsigned char clamp8(int v) { return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

__global__ void __launch_bounds__(256) jpeg_colour_kernel(const unsigned char* __restrict__ planes, JpegPlanes g, int width, int height,
                                                          int ycc, int bgr, unsigned char* __restrict__ out) {
  const long long n = (long long)width * height;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(idx / width), x = (int)(idx % width);
    int r, gg, b;
    const int c0 = sample_at(planes, g, 0, y, x);
    if (g.ncomp == 1) {
      r = gg = b = c0;
    } else {
      const int c1 = sample_at(planes, g, 1, y, x), c2 = sample_at(planes, g, 2, y, x);
      if (ycc) {                                                       // jdcolor.c build_ycc_rgb_table / ycc_rgb_convert
        const int cb = c1 - 128, cr = c2 - 128;
        r = clamp8(c0 + ((91881 * cr + 32768) >> 16));
        gg = clamp8(c0 + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
        b = clamp8(c0 + ((116130 * cb + 32768) >> 16));
      } else {
        r = c0; gg = c1; b = c2;
      }
    }
    unsigned char* o = out + idx * 3;
    o[0] = (unsigned char)(bgr ? b : r);
    o[1] = (unsigned char)gg;
    o[2] = (unsigned char)(bgr ? r : b);
  }
}

int map_error(int e) {
  switch (e) {
    case msda_jpeg::kOk: return 0;
    case msda_jpeg::kUnsupported: return MSDA_E_UNSUPPORTED;
    default: return MSDA_E_DIMS;                                       // truncated / corrupt stream
  }
}

}  // namespace
}  // namespace msda

extern "C" {

int msda_b200_jpeg_info(const unsigned char* data, size_t len, int* width, int* height, int* components) {
  if (!data || !width || !height) return MSDA_E_NULLPTR;
  // SOFn scan only: no entropy decoding
  if (len < 4 || data[0] != 0xff || data[1] != 0xd8) return MSDA_E_DIMS;
  size_t pos = 2;
  while (pos + 4 <= len) {
    if (data[pos] != 0xff) { ++pos; continue; }
    const int m = data[pos + 1];
    if (m == 0xff) { ++pos; continue; }
    if (m == 0xd8 || m == 0x01 || (m >= 0xd0 && m <= 0xd7)) { pos += 2; continue; }
    const int seglen = (data[pos + 2] << 8) | data[pos + 3];
    if (m >= 0xc0 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc) {
      if (pos + 10 > len) return MSDA_E_DIMS;
      *height = (data[pos + 5] << 8) | data[pos + 6];
      *width = (data[pos + 7] << 8) | data[pos + 8];
      if (components) *components = data[pos + 9];
      if (m != 0xc0 && m != 0xc1) return MSDA_E_UNSUPPORTED;           // progressive / lossless / arithmetic
      return (*width > 0 && *height > 0) ? 0 : MSDA_E_DIMS;
    }
    if (m == 0xda || m == 0xd9) break;
    pos += 2 + seglen;
  }
  return MSDA_E_DIMS;
}

int msda_b200_jpeg_decode_u8(const unsigned char* data, size_t len, int bgr, unsigned char* out, int width, int height,
                             void* stream) {
  using namespace msda;
  if (!data || !out) return MSDA_E_NULLPTR;
  msda_jpeg::Decoded d;
  const int rc = msda_jpeg::entropy_decode(data, len, d);
  if (rc != 0) return map_error(rc);
  if (d.width != width || d.height != height) return MSDA_E_DIMS;
  cudaStream_t st = (cudaStream_t)stream;

  JpegPlanes g;
  memset(&g, 0, sizeof(g));
  g.ncomp = d.ncomp;
  uint16_t quant[3 * 64];
  long long plane_bytes = 0;
  int blocks = 0;
  for (int c = 0; c < d.ncomp; ++c) {
    const msda_jpeg::Component& cp = d.comp[c];
    g.block_begin[c] = blocks;
    blocks += cp.blocks_w * cp.blocks_h;
    g.blocks_w[c] = cp.blocks_w;
    g.quant[c] = c;
    memcpy(quant + 64 * c, d.quant[cp.tq], 64 * sizeof(uint16_t));
    g.pitch[c] = cp.blocks_w * 8;
    g.cw[c] = cp.width;
    g.ch[c] = cp.height;
    g.hs[c] = d.max_h / cp.h;
    g.vs[c] = d.max_v / cp.v;
    g.plane_offset[c] = plane_bytes;
    plane_bytes += (long long)cp.blocks_w * 8 * cp.blocks_h * 8;
  }
  g.block_begin[d.ncomp] = blocks;

  // one stream-ordered scratch buffer: coefficients | quantisation tables | component planes
  const size_t coef_bytes = d.coef.size() * sizeof(int16_t);
  const size_t quant_off = (coef_bytes + 255) & ~(size_t)255;
  const size_t plane_off = (quant_off + sizeof(quant) + 255) & ~(size_t)255;
  unsigned char* scratch = nullptr;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&scratch), plane_off + (size_t)plane_bytes, st);
  if (e != cudaSuccess) return (int)e;
  // pageable sources: cudaMemcpyAsync returns once they have been staged, so `d` and `quant` may go out of scope
  e = cudaMemcpyAsync(scratch, d.coef.data(), coef_bytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(scratch + quant_off, quant, sizeof(quant), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    jpeg_idct_kernel<<<(blocks + kBlocksPerCta - 1) / kBlocksPerCta, 256, 0, st>>>(
        reinterpret_cast<const int16_t*>(scratch), reinterpret_cast<const uint16_t*>(scratch + quant_off), g, scratch + plane_off);
    const long long n = (long long)width * height;
    const int grid = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
    jpeg_colour_kernel<<<grid, 256, 0, st>>>(scratch + plane_off, g, width, height, d.ycc ? 1 : 0, bgr, out);
    e = cudaGetLastError();
  }
  const cudaError_t fe = cudaFreeAsync(scratch, st);
  return e != cudaSuccess ? (int)e : (int)fe;
}

}  // extern "C"
