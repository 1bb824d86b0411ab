// The elementwise glue around every layer of DeepSolo's point-query decoder, one kernel each instead
// of ~28 eager launches on 2 500 points per layer (profiles/r02_launches_clip_graph_heads.csv):
//   point_pos_embed   gen_point_pos_embed(reference_points_input[:, :, :, 0, :], d_model, temp)
//                     third_party/adet/modeling/model/utils.py:24-37, called at deformable_transformer.py:477
//   refine_points     (tmp + inverse_sigmoid(reference_points)).sigmoid()
//                     deformable_transformer.py:483-486, adet/utils/misc.py:115-119
// Same fp32 operations in the same order as the eager code (IEEE multiply / divide / add, sinf / cosf / logf / expf of
// libdevice -- what ATen's elementwise kernels call), so the results are bit-identical (tests/test_encoder_layer.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msda_b200.h"
#include "msda_launch.h"

namespace msda {
namespace {

// out[p][c]: c < half_dim -> x coordinate, else y; even index -> sin, odd -> cos; argument = (ref * ratio) * 2pi / dim_t[i]
__global__ void __launch_bounds__(256) point_pos_embed_kernel(const float* __restrict__ ref, const float* __restrict__ ratio,
                                                              int ratio_stride, const float* __restrict__ dim_t,
                                                              long long points, int per_batch, int half_dim,
                                                              float* __restrict__ out) {
  const long long total = points * 2 * half_dim;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long p = idx / (2 * half_dim);
    const int c = (int)(idx % (2 * half_dim));
    const int xy = c / half_dim, i = c % half_dim;
    float v = ref[p * 2 + xy];
    if (ratio) v = __fmul_rn(v, ratio[(p / per_batch) * ratio_stride + xy]);     // reference_points * valid_ratios (level 0)
    const float e = __fmul_rn(v, 6.283185307179586f);                             // * (2 * math.pi), rounded to fp32 by torch
    const float a = __fdiv_rn(e, dim_t[i]);
    out[idx] = (i & 1) ? cosf(a) : sinf(a);
  }
}

__global__ void __launch_bounds__(256) refine_points_kernel(const float* __restrict__ tmp, const float* __restrict__ ref, long long n,
                                                            float eps, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float x = ref[i];
    x = fminf(fmaxf(x, 0.0f), 1.0f);                       // clamp(min=0, max=1); NaN propagates like torch's clamp
    if (ref[i] != ref[i]) x = ref[i];
    float x1 = x < eps ? eps : x;                          // clamp(min=eps)
    float y = __fsub_rn(1.0f, x);
    float x2 = y < eps ? eps : y;
    const float l = logf(__fdiv_rn(x1, x2));
    const float s = __fadd_rn(tmp[i], l);
    out[i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-s)));   // ATen sigmoid: 1 / (1 + exp(-x))
  }
}

}  // namespace
}  // namespace msda

extern "C" {

int msda_b200_point_pos_embed_f32(const float* ref, const float* ratio, int ratio_stride, const float* dim_t, long long points,
                                  int points_per_batch, int half_dim, float* out, void* stream) {
  using namespace msda;
  if (!ref || !dim_t || !out) return MSDA_E_NULLPTR;
  if (points <= 0 || half_dim <= 0 || points_per_batch <= 0 || (ratio && ratio_stride < 2)) return MSDA_E_DIMS;
  const long long total = points * 2 * half_dim;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  point_pos_embed_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ref, ratio, ratio_stride, dim_t, points, points_per_batch, half_dim, out);
  return (int)cudaGetLastError();
}

int msda_b200_refine_points_f32(const float* tmp, const float* ref, long long n, float eps, float* out, void* stream) {
  using namespace msda;
  if (!tmp || !ref || !out) return MSDA_E_NULLPTR;
  if (n <= 0) return MSDA_E_DIMS;
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  refine_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tmp, ref, n, eps, out);
  return (int)cudaGetLastError();
}

}  // extern "C"
