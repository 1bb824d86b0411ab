// Frame batcher: the per-frame input side of the video loop, one pass over HBM.
//
// Reference (per frame, three eager steps and an 4x larger host->device copy):
//   GoMBatchPredictor.__call__   gomatching/text_track_visualizer.py:313-321  optional channel flip x[:, :, ::-1],
//                                 x.astype("float32").transpose(2, 0, 1) on the HOST (so 12 B/pixel cross PCIe)
//   GoMatching.preprocess_image  gomatching/modeling/meta_arch/gom_lstmatcher.py:159-170
//                                 (x - pixel_mean) / pixel_std on the device, then ImageList.from_tensors (zero padding
//                                 up to the backbone's size divisibility, applied AFTER the normalisation)
// Here the uint8 HWC frames are copied as they are (3 B/pixel) and one kernel writes the normalised, padded
// (N, 3, Hp, Wp) float32 batch:  out = fdiv_rn(fsub_rn((float)u8, mean_c), std_c)  -- the two IEEE operations the eager
// ops perform, so the result is bit-identical.  HBM-bound: 3 B read + 12 B written per pixel (+ padding).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msda_b200.h"

namespace msda {
namespace {

struct BatcherParams {
  const unsigned char* in;   // (N, H, W, 3)
  float* out;                // (N, 3, Hp, Wp)
  int N, H, W, Hp, Wp;
  int flip;                  // output channel c reads input channel 2 - c
  float mean[3], stdv[3];
};

// VEC = 4: one thread = 4 horizontally adjacent pixels (12 input bytes as three 32-bit loads where the row is 4-byte
// aligned, byte loads otherwise; one float4 store per channel plane); needs Wp % 4 == 0 and a 16-byte aligned output.
// VEC = 1: any geometry.
template <int VEC>
__global__ void __launch_bounds__(256) frames_u8_to_chw_f32_kernel(const BatcherParams p) {
  const int groups_per_row = p.Wp / VEC;
  const long long total = (long long)p.N * p.Hp * groups_per_row;
  const size_t plane = (size_t)p.Hp * p.Wp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int gx = (int)(i % groups_per_row);
    const long long r = i / groups_per_row;
    const int y = (int)(r % p.Hp), n = (int)(r / p.Hp);
    const int x = gx * VEC;
    float* o = p.out + (size_t)n * 3 * plane + (size_t)y * p.Wp + x;
    if constexpr (VEC == 4) {
      float v[3][4];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) v[c][k] = 0.0f;
      if (y < p.H && x < p.W) {
        const unsigned char* src = p.in + (((size_t)n * p.H + y) * p.W + x) * 3;
        unsigned char b[12];
        const int npx = min(4, p.W - x);                       // pixels of this group inside the frame
        if (npx == 4 && (reinterpret_cast<uintptr_t>(src) & 3u) == 0) {
          const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
          const uint32_t w0 = __ldg(s4), w1 = __ldg(s4 + 1), w2 = __ldg(s4 + 2);
#pragma unroll
          for (int k = 0; k < 4; ++k) { b[k] = (w0 >> (8 * k)) & 0xff; b[4 + k] = (w1 >> (8 * k)) & 0xff; b[8 + k] = (w2 >> (8 * k)) & 0xff; }
        } else {                                               // odd widths: rows are not 4-byte aligned
#pragma unroll
          for (int k = 0; k < 12; ++k) b[k] = (k < 3 * npx) ? __ldg(src + k) : 0;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int ci = p.flip ? 2 - c : c;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < npx) v[c][k] = __fdiv_rn(__fsub_rn((float)b[3 * k + ci], p.mean[c]), p.stdv[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) __stcs(reinterpret_cast<float4*>(o + c * plane), make_float4(v[c][0], v[c][1], v[c][2], v[c][3]));
    } else {
      const bool inside = (y < p.H) && (x < p.W);
      const unsigned char* src = p.in + (((size_t)n * p.H + y) * p.W + x) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int ci = p.flip ? 2 - c : c;
        o[c * plane] = inside ? __fdiv_rn(__fsub_rn((float)__ldg(src + ci), p.mean[c]), p.stdv[c]) : 0.0f;
      }
    }
  }
}

}  // namespace
}  // namespace msda

extern "C" int msda_b200_frames_u8_to_chw_f32(const unsigned char* frames, int N, int H, int W, int flip_channels,
                                              const float* mean3, const float* std3, int Hp, int Wp, float* out,
                                              void* stream) {
  using namespace msda;
  if (!frames || !mean3 || !std3 || !out) return MSDA_E_NULLPTR;
  if (N <= 0 || H <= 0 || W <= 0 || Hp < H || Wp < W) return MSDA_E_DIMS;
  BatcherParams p;
  p.in = frames; p.out = out; p.N = N; p.H = H; p.W = W; p.Hp = Hp; p.Wp = Wp; p.flip = flip_channels ? 1 : 0;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean3[c]; p.stdv[c] = std3[c]; }
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return MSDA_E_NOCUDA;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  const bool vec = (Wp % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
  const long long total = (long long)N * Hp * (vec ? Wp / 4 : Wp);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sms * 16;      // persistent grid-stride loop, a multiple of the SM count
  if (blocks > cap) blocks = cap;
  if (vec) frames_u8_to_chw_f32_kernel<4><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  else frames_u8_to_chw_f32_kernel<1><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}
