// Backward of the core operator (sm_100a).
//
// Replaces ms_deformable_col2im_cuda and its seven kernel variants
// (third_party/adet/layers/csrc/DeformAttn/ms_deform_im2col_cuda.cuh:301-920, dispatch :956-1327) and
// ms_deform_attn_col2im_bilinear (:87-159).  Gradients:
//   grad_value[corner]   += w_corner * attn * grad_out            (atomics, like the reference)
//   grad_attn[sample]     = sum_c grad_out_c * bilinear(value)_c
//   grad_loc[sample].x    = sum_c  W * (-hh*v1 + hh*v2 - lh*v3 + lh*v4)_c * attn * grad_out_c     (cuh:126-151)
//   grad_loc[sample].y    = sum_c  H * (-hw*v1 - lw*v2 + hw*v3 + lw*v4)_c * attn * grad_out_c
// Samples outside the map get zero gradients (cuh:352-355).
//
// Same lane layout as the forward: LPR lanes own one (b,q,m) unit and 4 channels each, so
//   * the four corner rows are read with one 16-byte load per lane,
//   * grad_value is accumulated with ONE vector reduction (red.global.add.v4.f32) per corner per lane
//     instead of four scalar atomics,
//   * the channel sums for grad_loc / grad_attn are xor-shuffle reductions over the unit's lanes; no
//     shared-memory reduction tree, no per-D kernel zoo.
// grad_value accumulation order is non-deterministic (atomics) exactly as in the reference.
#include "msda_device.cuh"
#include "msda_launch.h"
#include "../../include/msda_b200.h"

namespace msda {

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int D, int NW>
__global__ void __launch_bounds__(NW * 32) msda_bwd_tiled_kernel(const float* __restrict__ value,
                                                                const int64_t* __restrict__ shapes,
                                                                const int64_t* __restrict__ lsi,
                                                                const float* __restrict__ loc,
                                                                const float* __restrict__ attn,
                                                                const float* __restrict__ grad_out, int N, int S, int M,
                                                                int L, int Lq, int P, float* __restrict__ grad_value,
                                                                float* __restrict__ grad_loc,
                                                                float* __restrict__ grad_attn) {
  constexpr int VEC = 4, LPR = D / VEC, UPW = 32 / LPR;
  __shared__ int sH[kMaxLevels], sW[kMaxLevels], sStart[kMaxLevels];
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = lane / LPR, k = lane % LPR;
  if (tid < L) {
    sH[tid] = (int)shapes[2 * tid];
    sW[tid] = (int)shapes[2 * tid + 1];
    sStart[tid] = (int)lsi[tid];
  }
  __syncthreads();
  const long long units = (long long)N * Lq * M;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + tid) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int cstride = M * D;   // elements between horizontally adjacent pixels

  for (long long u0 = warp_global * UPW; u0 < units; u0 += nwarps * UPW) {
    const long long unit_raw = u0 + g;
    const bool valid = unit_raw < units;
    const size_t unit = valid ? (size_t)unit_raw : 0;
    const int m = (int)(unit % M);
    const size_t b = unit / ((size_t)M * Lq);
    const float4 go = valid ? *reinterpret_cast<const float4*>(grad_out + unit * D + k * VEC) : make_float4(0, 0, 0, 0);
    const float gof[4] = {go.x, go.y, go.z, go.w};
    const size_t vb = (b * S * M + m) * (size_t)D + (size_t)k * VEC;
    const float* locp = loc + unit * L * P * 2;
    const float* attp = attn + unit * L * P;
    for (int l = 0; l < L; ++l) {
      const int H = sH[l], W = sW[l];
      const size_t vl = vb + (size_t)sStart[l] * M * D;
      const int rstride = W * cstride;
      for (int pt = 0; pt < P; ++pt) {
        const int s = l * P + pt;
        // every lane of the unit evaluates the same sample (broadcast loads)
        const float2 xy = valid ? *reinterpret_cast<const float2*>(locp + 2 * s) : make_float2(-9.f, -9.f);
        const float a = valid ? attp[s] : 0.0f;
        const SampleGeom sg = sample_setup(xy.x, xy.y, H, W);
        float g_a = 0.0f, g_w = 0.0f, g_h = 0.0f;
        if (sg.in_range) {
          const ptrdiff_t o1 = (ptrdiff_t)vl + ((ptrdiff_t)sg.h_low * W + sg.w_low) * cstride;
          float4 z = make_float4(0, 0, 0, 0), q1 = z, q2 = z, q3 = z, q4 = z;
          if (sg.mask & 1) q1 = *reinterpret_cast<const float4*>(value + o1);
          if (sg.mask & 2) q2 = *reinterpret_cast<const float4*>(value + o1 + cstride);
          if (sg.mask & 4) q3 = *reinterpret_cast<const float4*>(value + o1 + rstride);
          if (sg.mask & 8) q4 = *reinterpret_cast<const float4*>(value + o1 + rstride + cstride);
          const float v1[4] = {q1.x, q1.y, q1.z, q1.w}, v2[4] = {q2.x, q2.y, q2.z, q2.w};
          const float v3[4] = {q3.x, q3.y, q3.z, q3.w}, v4[4] = {q4.x, q4.y, q4.z, q4.w};
          const float lh = sg.lh, lw = sg.lw, hh = 1.0f - lh, hw = 1.0f - lw;
          const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
          float t[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            t[c] = gof[c] * a;                                                     // top_grad_value (cuh:105)
            const float val = w1 * v1[c] + w2 * v2[c] + w3 * v3[c] + w4 * v4[c];
            g_a += gof[c] * val;                                                   // cuh:149
            const float ghw = -hw * v1[c] - lw * v2[c] + hw * v3[c] + lw * v4[c];  // grad_h_weight
            const float gww = -hh * v1[c] + hh * v2[c] - lh * v3[c] + lh * v4[c];  // grad_w_weight
            g_w += (float)W * gww * t[c];                                          // cuh:150
            g_h += (float)H * ghw * t[c];                                          // cuh:151
          }
          if (sg.mask & 1) red_add_v4(grad_value + o1, w1 * t[0], w1 * t[1], w1 * t[2], w1 * t[3]);
          if (sg.mask & 2) red_add_v4(grad_value + o1 + cstride, w2 * t[0], w2 * t[1], w2 * t[2], w2 * t[3]);
          if (sg.mask & 4) red_add_v4(grad_value + o1 + rstride, w3 * t[0], w3 * t[1], w3 * t[2], w3 * t[3]);
          if (sg.mask & 8) red_add_v4(grad_value + o1 + rstride + cstride, w4 * t[0], w4 * t[1], w4 * t[2], w4 * t[3]);
        }
        // channel sums over the unit's LPR lanes (lanes of other units never mix: groups are aligned)
#pragma unroll
        for (int off = LPR / 2; off >= 1; off >>= 1) {
          g_a += __shfl_xor_sync(0xffffffffu, g_a, off);
          g_w += __shfl_xor_sync(0xffffffffu, g_w, off);
          g_h += __shfl_xor_sync(0xffffffffu, g_h, off);
        }
        if (valid && k == 0) {
          *reinterpret_cast<float2*>(grad_loc + (unit * L * P + s) * 2) = make_float2(g_w, g_h);
          grad_attn[unit * L * P + s] = g_a;
        }
      }
    }
  }
}

// any D: one thread per (unit, channel); all three gradients through global atomics, the scheme of the
// reference's fallback ms_deformable_col2im_gpu_kernel_gm (cuh:845-920).  grad_loc / grad_attn are zeroed first.
__global__ void __launch_bounds__(256) msda_bwd_generic_kernel(const float* __restrict__ value,
                                                               const int64_t* __restrict__ shapes,
                                                               const int64_t* __restrict__ lsi,
                                                               const float* __restrict__ loc,
                                                               const float* __restrict__ attn,
                                                               const float* __restrict__ grad_out, int N, int S, int M,
                                                               int D, int L, int Lq, int P,
                                                               float* __restrict__ grad_value,
                                                               float* __restrict__ grad_loc,
                                                               float* __restrict__ grad_attn) {
  const long long total = (long long)N * Lq * M * D;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    const long long unit = idx / D;
    const int m = (int)(unit % M);
    const long long b = unit / ((long long)M * Lq);
    const float go = grad_out[idx];
    const size_t vb = ((size_t)b * S * M + m) * D + c;
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const size_t vl = vb + (size_t)((int)lsi[l]) * M * D;
      const ptrdiff_t cs = (ptrdiff_t)M * D, rs = (ptrdiff_t)W * M * D;
      for (int pt = 0; pt < P; ++pt) {
        const size_t s = (size_t)unit * L * P + l * P + pt;
        const SampleGeom sg = sample_setup(loc[2 * s], loc[2 * s + 1], H, W);
        if (!sg.in_range) continue;
        const float a = attn[s];
        const ptrdiff_t o1 = (ptrdiff_t)vl + ((ptrdiff_t)sg.h_low * W + sg.w_low) * cs;
        const float v1 = (sg.mask & 1) ? value[o1] : 0.0f, v2 = (sg.mask & 2) ? value[o1 + cs] : 0.0f;
        const float v3 = (sg.mask & 4) ? value[o1 + rs] : 0.0f, v4 = (sg.mask & 8) ? value[o1 + rs + cs] : 0.0f;
        const float lh = sg.lh, lw = sg.lw, hh = 1.0f - lh, hw = 1.0f - lw;
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const float t = go * a;
        if (sg.mask & 1) atomicAdd(grad_value + o1, w1 * t);
        if (sg.mask & 2) atomicAdd(grad_value + o1 + cs, w2 * t);
        if (sg.mask & 4) atomicAdd(grad_value + o1 + rs, w3 * t);
        if (sg.mask & 8) atomicAdd(grad_value + o1 + rs + cs, w4 * t);
        atomicAdd(grad_attn + s, go * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4));
        atomicAdd(grad_loc + 2 * s, (float)W * (-hh * v1 + hh * v2 - lh * v3 + lh * v4) * t);
        atomicAdd(grad_loc + 2 * s + 1, (float)H * (-hw * v1 - lw * v2 + hw * v3 + lw * v4) * t);
      }
    }
  }
}

int launch_backward_f32(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                        const float* attn, const float* grad_out, int N, int S, int M, int D, int L, int Lq, int P,
                        float* grad_value, float* grad_loc, float* grad_attn, cudaStream_t stream) {
  const long long units = (long long)N * Lq * M;
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool vec_ok = aligned16(value) && aligned16(grad_out) && aligned16(grad_value) &&
                      (reinterpret_cast<uintptr_t>(loc) & 7u) == 0 && (reinterpret_cast<uintptr_t>(grad_loc) & 7u) == 0;
  if ((D == 32 || D == 64) && L <= kMaxLevels && vec_ok) {
    constexpr int NW = 8;
    const int upw = 32 / (D / 4);
    long long blocks = (units + (long long)NW * upw - 1) / ((long long)NW * upw);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    if (D == 32)
      msda_bwd_tiled_kernel<32, NW><<<(int)blocks, NW * 32, 0, stream>>>(value, shapes, lsi, loc, attn, grad_out, N, S, M, L,
                                                                         Lq, P, grad_value, grad_loc, grad_attn);
    else
      msda_bwd_tiled_kernel<64, NW><<<(int)blocks, NW * 32, 0, stream>>>(value, shapes, lsi, loc, attn, grad_out, N, S, M, L,
                                                                         Lq, P, grad_value, grad_loc, grad_attn);
    return (int)cudaGetLastError();
  }
  cudaError_t e = cudaMemsetAsync(grad_loc, 0, sizeof(float) * 2 * (size_t)units * L * P, stream);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(grad_attn, 0, sizeof(float) * (size_t)units * L * P, stream);
  if (e != cudaSuccess) return (int)e;
  const long long total = units * D;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  msda_bwd_generic_kernel<<<(int)(blocks < 1 ? 1 : blocks), 256, 0, stream>>>(value, shapes, lsi, loc, attn, grad_out, N, S,
                                                                             M, D, L, Lq, P, grad_value, grad_loc,
                                                                             grad_attn);
  return (int)cudaGetLastError();
}

}  // namespace msda
