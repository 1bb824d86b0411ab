// Multi-head self-attention core for the SHORT sequences of DeepSolo's point-query decoder:
//   attn_intra  25 points of one proposal          (third_party/adet/layers/deformable_transformer.py:386-394)
//   attn_inter  100 proposals at one point index   (:396-404)
// head dimension 32, 8 heads.  The reference runs nn.MultiheadAttention's functional path (need_weights=True): packed
// input projection, q * d^-1/2, bmm, softmax, bmm, head-averaged attention weights it throws away, output projection --
// a dozen launches on 2 500 tokens per call, twelve calls per frame.  Here the projections are two GEMMs of
// proj_gemm.cu and everything between them is this one kernel:
//     out[b, i, h, :] = softmax_j( (q[b,i,h,:] * d^-1/2) . k[b,j,h,:] ) . v[b,j,h,:]
// q, k, v are COLUMN SLICES of the packed projection output (row pitch ld floats), rows addressed as
// row(b, i) = b * batch_stride + i * seq_stride -- so the "inter" attention reads the (proposal, point) row order it is
// given and needs no transposed copy.  One CTA per (batch, head, chunk of query rows): K and V of the sequence live in
// shared memory (rows padded to 33 floats; the 100-key sequences are split into row chunks so that ~6 CTAs per SM are in
// flight instead of 200 long serial ones), a warp owns a query row at a time: lanes = keys for the scores and the softmax (shuffle
// reductions, exp via expf like torch), lanes = channels for the weighted sum.  fp32 throughout; same operation order as
// the reference up to the summation order inside the two dot products.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msda_b200.h"
#include "msda_launch.h"

namespace msda {
namespace {

constexpr int kHd = 32;          // head dimension
constexpr int kMaxL = 128;       // longest sequence (keys per lane: 4)
constexpr int kPad = kHd + 1;

__global__ void __launch_bounds__(128) small_mha_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                        const float* __restrict__ v, int ld, float* __restrict__ out, int ldo,
                                                        int L, int H, long long batch_stride, long long seq_stride, float scale) {
  __shared__ float sK[kMaxL * kPad], sV[kMaxL * kPad];
  __shared__ float sP[4][kMaxL];
  __shared__ float sQ[4][kHd];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)b * batch_stride;
  for (int i = tid; i < L * kHd; i += blockDim.x) {
    const int j = i / kHd, c = i % kHd;
    const long long r = (row0 + (long long)j * seq_stride) * ld + h * kHd + c;
    sK[j * kPad + c] = k[r];
    sV[j * kPad + c] = v[r];
  }
  __syncthreads();
  const int rows_per_cta = (L + (int)gridDim.y - 1) / (int)gridDim.y;
  const int i_begin = blockIdx.y * rows_per_cta, i_end = min(L, i_begin + rows_per_cta);
  for (int i = i_begin + warp; i < i_end; i += 4) {
    const long long r = row0 + (long long)i * seq_stride;
    sQ[warp][lane] = __fmul_rn(q[r * ld + h * kHd + lane], scale);          // q * d^-1/2 first, like the reference
    __syncwarp();
    float s[4], mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = lane + 32 * t;
      float acc = 0.0f;
      if (j < L) {
#pragma unroll
        for (int c = 0; c < kHd; ++c) acc = __fmaf_rn(sQ[warp][c], sK[j * kPad + c], acc);
        mx = fmaxf(mx, acc);
      }
      s[t] = acc;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = lane + 32 * t;
      s[t] = j < L ? expf(__fsub_rn(s[t], mx)) : 0.0f;
      sum += s[t];
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = lane + 32 * t;
      if (j < L) sP[warp][j] = __fdiv_rn(s[t], sum);
    }
    __syncwarp();
    float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f, o3 = 0.0f;       // four independent chains: the sum is latency-bound otherwise
    int j = 0;
    for (; j + 3 < L; j += 4) {
      o0 = __fmaf_rn(sP[warp][j], sV[j * kPad + lane], o0);
      o1 = __fmaf_rn(sP[warp][j + 1], sV[(j + 1) * kPad + lane], o1);
      o2 = __fmaf_rn(sP[warp][j + 2], sV[(j + 2) * kPad + lane], o2);
      o3 = __fmaf_rn(sP[warp][j + 3], sV[(j + 3) * kPad + lane], o3);
    }
    for (; j < L; ++j) o0 = __fmaf_rn(sP[warp][j], sV[j * kPad + lane], o0);
    out[r * ldo + h * kHd + lane] = __fadd_rn(__fadd_rn(o0, o1), __fadd_rn(o2, o3));
    __syncwarp();
  }
}

}  // namespace
}  // namespace msda

extern "C" int msda_b200_small_mha_f32(const float* q, const float* k, const float* v, int ld, float* out, int ldo, int B, int L,
                                       int H, int head_dim, long long batch_stride, long long seq_stride, void* stream) {
  using namespace msda;
  if (!q || !k || !v || !out) return MSDA_E_NULLPTR;
  if (B <= 0 || L <= 0 || H <= 0 || ld < H * head_dim || ldo < H * head_dim) return MSDA_E_DIMS;
  if (head_dim != kHd || L > kMaxL) return MSDA_E_UNSUPPORTED;
  const float scale = 0.17677669529663687f;          // float(32 ** -0.5), the value torch multiplies q by
  const int sms = msda_b200_sm_count();
  if (sms < 0) return sms;
  int chunks = (6 * sms) / (B * H);                  // ~6 resident CTAs per SM
  if (chunks > (L + 3) / 4) chunks = (L + 3) / 4;    // at least one row per warp
  if (chunks < 1) chunks = 1;
  small_mha_kernel<<<dim3(B * H, chunks), 128, 0, (cudaStream_t)stream>>>(q, k, v, ld, out, ldo, L, H, batch_stride, seq_stride, scale);
  return (int)cudaGetLastError();
}
