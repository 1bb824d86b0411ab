// Forward kernels of the B200 MSDeformAttn library (sm_100a).
//
// Replaces ms_deformable_im2col_gpu_kernel (third_party/adet/layers/csrc/DeformAttn/
// ms_deform_im2col_cuda.cuh:237-299) and, in the fused form, the eager glue of MSDeformAttn.forward
// (third_party/adet/layers/ms_deform_attn.py:137-147).  Design (DESIGN.md "kernels"):
//
//   * a "unit" is one (batch b, query q, head m): 1 output row of D channels, L*P samples x 4 corners.
//   * a value row (one pixel, one head) is D*sizeof(T) bytes = 128 B (fp32, D=32).  LPR = D*sizeof(T)/16
//     lanes read one row with ONE 16-byte load each; a warp therefore gathers 32/LPR rows per LDG.128
//     (4 for fp32, 8 for bf16) -- each a full cache line, so every L1 wavefront delivers only useful bytes.
//   * phase 1 (cooperative): the lanes of a unit split its L*P samples, do the location/index/weight
//     arithmetic ONCE per sample and park a 16-byte record {byte offset | corner mask, lh, lw, attn} in
//     shared memory.  In the fused kernel phase 1 also does the softmax over L*P (warp shuffles) and the
//     offset -> location arithmetic, so loc/attn never exist in HBM.
//   * phase 2 (gather): every lane walks the unit's samples in the reference order (l outer, p inner),
//     reads the record with one broadcast LDS.128, issues the 4 predicated corner loads and accumulates
//     its 4 (fp32) or 8 (bf16) channels in registers with the reference's exact FMA chain.  No
//     cross-lane reduction, no atomics, the output row is written once with a streaming 16-byte store.
//   * persistent grid (k CTAs per SM) over tiles; one tile = one head x a block of queries.  In pyramid
//     mode (encoder self-attention, Lq == S) a tile is a TH x TW pixel block of one level, so the value
//     rows the tile gathers (its own neighbourhood at every level) stay L1-resident across its queries.
//     Tile geometry is derived in-kernel from the DEVICE shapes tensor: no host copy, no sync.
#include "msda_device.cuh"
#include "msda_launch.h"
#include "../../include/msda_b200.h"

namespace msda {

__host__ __device__ constexpr int next_pow2(int x) { int r = 1; while (r < x) r <<= 1; return r; }

// -----------------------------------------------------------------------------------------------
// Phase-1 helpers
// -----------------------------------------------------------------------------------------------
// Fused glue for one unit, LP compile-time.  Lane k of the unit's LPR lanes owns samples s = i*LPR + k.
// Softmax reproduces the operation order of PyTorch's persistent warp softmax (softmax_warp_forward:
// element e on virtual lane e % WS, per-lane sequential sum, xor-butterfly WS/2..1), so the weights are
// the ones the eager reference computes.
template <int LPR, int LP>
struct FusedGlue {
  static constexpr int SPL = LP / LPR;
  static constexpr int NP2 = next_pow2(LP);
  static constexpr int WS = NP2 < 32 ? NP2 : 32;
  static constexpr int R = WS / LPR;
  static_assert(LP % LPR == 0 && R >= 1, "unsupported L*P for this lane layout");

  __device__ static __forceinline__ void softmax(const float* __restrict__ logits_unit, int k, float (&a)[SPL]) {
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
      a[i] = ld_stream_f1(logits_unit + i * LPR + k);
      mx = fmaxf(mx, a[i]);
    }
#pragma unroll
    for (int off = LPR / 2; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float vs[R];
#pragma unroll
    for (int j = 0; j < R; ++j) vs[j] = 0.0f;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
      a[i] = expf(__fsub_rn(a[i], mx));
      vs[i % R] = __fadd_rn(vs[i % R], a[i]);
    }
#pragma unroll
    for (int h = R / 2; h >= 1; h >>= 1) {
#pragma unroll
      for (int j = 0; j < h; ++j) vs[j] = __fadd_rn(vs[j], vs[j + h]);
    }
    float sum = vs[0];
#pragma unroll
    for (int off = LPR / 2; off >= 1; off >>= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, off));
#pragma unroll
    for (int i = 0; i < SPL; ++i) a[i] = __fdiv_rn(a[i], sum);
  }
};

// -----------------------------------------------------------------------------------------------
// Tiled forward kernel
//   T      float | __nv_bfloat16 (storage of value/out; arithmetic is fp32)
//   D      channels per head (compile-time: fixes the lane layout)
//   NW     warps per CTA
//   UNROLL samples in flight per lane in the gather loop
//   LPT    0: core operator (loc/attn given, L*P runtime) ; >0: fused operator with L*P == LPT
// -----------------------------------------------------------------------------------------------
template <typename T, int D, int NW, int UNROLL, int LPT, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) msda_fwd_tiled_kernel(const FwdParams p) {
  constexpr int VEC = Elem<T>::kVec;
  constexpr int LPR = D / VEC;           // lanes per value row
  constexpr int UPW = 32 / LPR;          // units per warp iteration
  constexpr bool FUSED = LPT > 0;
  static_assert(D % VEC == 0 && 32 % LPR == 0 && LPR >= 2, "bad D for this layout");

  __shared__ int sH[kMaxLevels], sW[kMaxLevels], sStart[kMaxLevels], sTileCum[kMaxLevels + 1];
  __shared__ unsigned char sLvl[kMaxSamples];
  extern __shared__ float4 sRecAll[];    // [NW][LP][UPW] records

  const int L = p.L, P = p.P, M = p.M, Lq = p.Lq;
  const int LP = FUSED ? LPT : L * P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane / LPR, k = lane % LPR;

  if (tid < L) {
    sH[tid] = (int)p.shapes[2 * tid];
    sW[tid] = (int)p.shapes[2 * tid + 1];
    sStart[tid] = (int)p.lsi[tid];
  }
  if (tid < LP) sLvl[tid] = (unsigned char)(tid / P);
  __syncthreads();
  const int tw_log2 = p.tile_w_log2, TH = p.tile_h;
  if (p.mode == kModePyramid && tid == 0) {
    int cum = 0;
    for (int l = 0; l < L; ++l) {
      sTileCum[l] = cum;
      cum += ((sH[l] + TH - 1) / TH) * ((sW[l] + (1 << tw_log2) - 1) >> tw_log2);
    }
    sTileCum[L] = cum;
  }
  __syncthreads();

  const int tiles_per_bm = (p.mode == kModePyramid) ? sTileCum[L] : (Lq + p.tile_q - 1) / p.tile_q;
  const long long total_tiles = (long long)p.N * tiles_per_bm * M;
  float4* sRec = sRecAll + (size_t)warp * LP * UPW;
  const int cstride = M * D * (int)sizeof(T);   // bytes between horizontally adjacent pixels
  const float inv_p = 1.0f / (float)P;

  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    // heads vary fastest so CTAs running side by side share the tile's loc/attn DRAM pages
    const int m = (int)(tile % M);
    const long long r = tile / M;
    const int t = (int)(r % tiles_per_bm);
    const int b = (int)(r / tiles_per_bm);
    int lvl = 0, ty = 0, tx = 0;
    if (p.mode == kModePyramid) {
      while (lvl + 1 < L && t >= sTileCum[lvl + 1]) ++lvl;
      const int tt = t - sTileCum[lvl];
      const int ntx = (sW[lvl] + (1 << tw_log2) - 1) >> tw_log2;
      ty = tt / ntx;
      tx = tt - ty * ntx;
    }
    const char* vbase = reinterpret_cast<const char*>(p.value) +
                        ((size_t)b * p.S * M * D + (size_t)m * D + (size_t)k * VEC) * sizeof(T);

    for (int j0 = warp * UPW; j0 < p.tile_q; j0 += NW * UPW) {
      const int j = j0 + g;
      int q;
      bool valid;
      if (p.mode == kModePyramid) {
        const int y = ty * TH + (j >> tw_log2), x = (tx << tw_log2) + (j & ((1 << tw_log2) - 1));
        valid = (y < sH[lvl]) && (x < sW[lvl]);
        q = sStart[lvl] + y * sW[lvl] + x;
      } else {
        q = t * p.tile_q + j;
        valid = q < Lq;
      }
      if (!__any_sync(0xffffffffu, valid)) continue;   // warp-uniform
      const size_t bq = (size_t)b * Lq + (valid ? q : 0);
      const size_t unit = bq * M + m;

      // ---------------- phase 1: one record per sample, computed once ----------------
      if constexpr (FUSED) {
        constexpr int SPL = LPT / LPR;
        float a[SPL];
        FusedGlue<LPR, LPT>::softmax(p.logits + unit * LPT, k, a);
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          const int s = i * LPR + k;
          const int l = sLvl[s];
          const int H = sH[l], W = sW[l];
          const float2 off = ld_stream_f2(p.offsets + (unit * LPT + s) * 2);
          const float* rp = p.ref + (bq * L + l) * p.ref_dim;
          const float r0 = __ldg(rp), r1 = __ldg(rp + 1);
          float r2 = 0.0f, r3 = 0.0f;
          if (p.ref_dim == 4) { r2 = __ldg(rp + 2); r3 = __ldg(rp + 3); }
          const float lx = location_from_offset(r0, r2, off.x, (float)W, inv_p, p.ref_dim);
          const float ly = location_from_offset(r1, r3, off.y, (float)H, inv_p, p.ref_dim);
          const SampleGeom sg = sample_setup(lx, ly, H, W);
          int packed = 0;
          if (sg.in_range && valid)
            packed = (((sStart[l] + sg.h_low * W + sg.w_low) * M * D) * (int)sizeof(T)) | sg.mask;
          sRec[s * UPW + g] = make_float4(__int_as_float(packed), sg.lh, sg.lw, (sg.in_range && valid) ? a[i] : 0.0f);
        }
      } else {
        const float* locp = p.loc + unit * LP * 2;
        const float* attp = p.attn + unit * LP;
        for (int s = k; s < LP; s += LPR) {
          const int l = sLvl[s];
          const int H = sH[l], W = sW[l];
          const float2 xy = ld_stream_f2(locp + 2 * s);
          const float a = ld_stream_f1(attp + s);
          const SampleGeom sg = sample_setup(xy.x, xy.y, H, W);
          int packed = 0;
          if (sg.in_range && valid)
            packed = (((sStart[l] + sg.h_low * W + sg.w_low) * M * D) * (int)sizeof(T)) | sg.mask;
          sRec[s * UPW + g] = make_float4(__int_as_float(packed), sg.lh, sg.lw, (sg.in_range && valid) ? a : 0.0f);
        }
      }
      __syncwarp();

      // ---------------- phase 2: gather + weighted reduction ----------------
      float acc[VEC];
#pragma unroll
      for (int c = 0; c < VEC; ++c) acc[c] = 0.0f;
      const float4* rec = sRec + g;
      for (int l = 0; l < L; ++l) {
        const int rstride = sW[l] * cstride;      // bytes between vertically adjacent pixels
#pragma unroll UNROLL
        for (int pt = 0; pt < P; ++pt) {
          const float4 rc = *rec;
          rec += UPW;
          const int packed = __float_as_int(rc.x);
          const int mask = packed & 15;
          const char* c1 = vbase + (ptrdiff_t)(packed & ~15);
          const char* c3 = c1 + rstride;
          uint4 u1 = make_uint4(0, 0, 0, 0), u2 = u1, u3 = u1, u4 = u1;
          if (mask & 1) u1 = ld_value16(c1);
          if (mask & 2) u2 = ld_value16(c1 + cstride);
          if (mask & 4) u3 = ld_value16(c3);
          if (mask & 8) u4 = ld_value16(c3 + cstride);
          float w1, w2, w3, w4;
          bilinear_weights(rc.y, rc.z, w1, w2, w3, w4);
          float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
          Elem<T>::unpack(u1, v1);
          Elem<T>::unpack(u2, v2);
          Elem<T>::unpack(u3, v3);
          Elem<T>::unpack(u4, v4);
#pragma unroll
          for (int c = 0; c < VEC; ++c) acc[c] = corner_accumulate(acc[c], rc.w, w1, w2, w3, w4, v1[c], v2[c], v3[c], v4[c]);
        }
      }
      if (valid) {
        char* op = reinterpret_cast<char*>(p.out) + (unit * D + (size_t)k * VEC) * sizeof(T);
        st_stream16(op, Elem<T>::pack(acc));
      }
      __syncwarp();   // records are overwritten by the next iteration's phase 1
    }
  }
}

// -----------------------------------------------------------------------------------------------
// Generic forward kernel: any D, L, P.  One thread per output element, same device functions.
// (Slow path for shapes the tiled layout does not cover; the DeepSolo configs never take it.)
// -----------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) msda_fwd_generic_kernel(const FwdParams p) {
  const long long total = (long long)p.N * p.Lq * p.M * p.D;
  const int M = p.M, D = p.D, L = p.L, P = p.P;
  const T* value = reinterpret_cast<const T*>(p.value);
  T* out = reinterpret_cast<T*>(p.out);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    const long long unit = idx / D;
    const int m = (int)(unit % M);
    const long long b = unit / ((long long)M * p.Lq);
    const T* vb = value + (size_t)b * p.S * M * D + (size_t)m * D + c;
    const float* locp = p.loc + (size_t)unit * L * P * 2;
    const float* attp = p.attn + (size_t)unit * L * P;
    float acc = 0.0f;
    for (int l = 0; l < L; ++l) {
      const int H = (int)p.shapes[2 * l], W = (int)p.shapes[2 * l + 1];
      const T* vl = vb + (size_t)((int)p.lsi[l]) * M * D;
      for (int pt = 0; pt < P; ++pt) {
        const int s = l * P + pt;
        const SampleGeom sg = sample_setup(locp[2 * s], locp[2 * s + 1], H, W);
        if (!sg.in_range) continue;
        const float a = attp[s];
        const ptrdiff_t o1 = ((ptrdiff_t)sg.h_low * W + sg.w_low) * M * D;
        const ptrdiff_t cs = (ptrdiff_t)M * D, rs = (ptrdiff_t)W * M * D;
        const float v1 = (sg.mask & 1) ? Elem<T>::load1(vl + o1) : 0.0f;
        const float v2 = (sg.mask & 2) ? Elem<T>::load1(vl + o1 + cs) : 0.0f;
        const float v3 = (sg.mask & 4) ? Elem<T>::load1(vl + o1 + rs) : 0.0f;
        const float v4 = (sg.mask & 8) ? Elem<T>::load1(vl + o1 + rs + cs) : 0.0f;
        float w1, w2, w3, w4;
        bilinear_weights(sg.lh, sg.lw, w1, w2, w3, w4);
        acc = corner_accumulate(acc, a, w1, w2, w3, w4, v1, v2, v3, v4);
      }
    }
    Elem<T>::store1(out + idx, acc);
  }
}

// -----------------------------------------------------------------------------------------------
// float64 forward: the reference instantiates its kernel for double too (AT_DISPATCH_FLOATING_TYPES,
// ms_deform_attn_cuda.cu:64).  Same loop as the generic kernel in double arithmetic with the contraction nvcc
// applies to the reference's double expressions (DFMA), written out explicitly.  Not a performance path.
// -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msda_fwd_generic_f64_kernel(const double* __restrict__ value,
                                                                   const int64_t* __restrict__ shapes,
                                                                   const int64_t* __restrict__ lsi,
                                                                   const double* __restrict__ loc,
                                                                   const double* __restrict__ attn, int N, int S, int M,
                                                                   int D, int L, int Lq, int P, double* __restrict__ out) {
  const long long total = (long long)N * Lq * M * D;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    const long long unit = idx / D;
    const int m = (int)(unit % M);
    const long long b = unit / ((long long)M * Lq);
    const double* vb = value + (size_t)b * S * M * D + (size_t)m * D + c;
    const double* locp = loc + (size_t)unit * L * P * 2;
    const double* attp = attn + (size_t)unit * L * P;
    double acc = 0.0;
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const double* vl = vb + (size_t)((int)lsi[l]) * M * D;
      const ptrdiff_t cs = (ptrdiff_t)M * D, rs = (ptrdiff_t)W * M * D;
      for (int pt = 0; pt < P; ++pt) {
        const int s = l * P + pt;
        const double h_im = __fma_rn(locp[2 * s + 1], (double)H, -0.5);
        const double w_im = __fma_rn(locp[2 * s], (double)W, -0.5);
        if (!(h_im > -1 && w_im > -1 && h_im < H && w_im < W)) continue;
        const double hf = floor(h_im), wf = floor(w_im);
        const int h_low = (int)hf, w_low = (int)wf;
        const double lh = __dsub_rn(h_im, hf), lw = __dsub_rn(w_im, wf);
        const double hh = __dsub_rn(1.0, lh), hw = __dsub_rn(1.0, lw);
        const double w1 = __dmul_rn(hh, hw), w2 = __dmul_rn(hh, lw), w3 = __dmul_rn(lh, hw), w4 = __dmul_rn(lh, lw);
        const ptrdiff_t o1 = ((ptrdiff_t)h_low * W + w_low) * cs;
        const bool t = h_low >= 0, bt = h_low + 1 <= H - 1, lf = w_low >= 0, rt = w_low + 1 <= W - 1;
        const double v1 = (t && lf) ? vl[o1] : 0.0, v2 = (t && rt) ? vl[o1 + cs] : 0.0;
        const double v3 = (bt && lf) ? vl[o1 + rs] : 0.0, v4 = (bt && rt) ? vl[o1 + rs + cs] : 0.0;
        double tt = __dmul_rn(w2, v2);
        tt = __fma_rn(w1, v1, tt);
        tt = __fma_rn(w3, v3, tt);
        tt = __fma_rn(w4, v4, tt);
        acc = __fma_rn(attp[s], tt, acc);
      }
    }
    out[idx] = acc;
  }
}

int launch_forward_f64(const double* value, const int64_t* shapes, const int64_t* lsi, const double* loc,
                       const double* attn, int N, int S, int M, int D, int L, int Lq, int P, double* out,
                       cudaStream_t stream) {
  const long long total = (long long)N * Lq * M * D;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  msda_fwd_generic_f64_kernel<<<(int)(blocks < 1 ? 1 : blocks), 256, 0, stream>>>(value, shapes, lsi, loc, attn, N, S, M, D,
                                                                                 L, Lq, P, out);
  return (int)cudaGetLastError();
}

// -----------------------------------------------------------------------------------------------
// Sampling-index dump (test/inspection): record layout == msda_b200_index_t == oracle's
// -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msda_sample_index_kernel(const float* __restrict__ loc,
                                                                const int64_t* __restrict__ shapes,
                                                                const int64_t* __restrict__ lsi, long long total,
                                                                int M, int D, int L, int P,
                                                                msda_b200_index_t* __restrict__ out) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < total;
       s += (long long)gridDim.x * blockDim.x) {
    const int l = (int)((s / P) % L);
    const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
    const SampleGeom sg = sample_setup(loc[2 * s], loc[2 * s + 1], H, W);
    msda_b200_index_t r;
    r.h_low = sg.h_low; r.w_low = sg.w_low; r.in_range = sg.in_range ? 1 : 0; r.corner_mask = sg.mask;
    r.level_offset = (int64_t)((int)lsi[l]) * M * D;
    out[s] = r;
  }
}

// -----------------------------------------------------------------------------------------------
// Glue on its own: the fused kernel's phase-1 arithmetic written out to HBM (tests/inspection)
// -----------------------------------------------------------------------------------------------
template <int LPR, int LP>
__global__ void __launch_bounds__(256) msda_glue_kernel(const int64_t* __restrict__ shapes, const float* __restrict__ ref,
                                                        int ref_dim, const float* __restrict__ offsets,
                                                        const float* __restrict__ logits, long long units, int M, int L,
                                                        int P, float* __restrict__ loc_out, float* __restrict__ attn_out) {
  constexpr int UPW = 32 / LPR, SPL = LP / LPR;
  const int lane = threadIdx.x & 31, g = lane / LPR, k = lane % LPR;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float inv_p = 1.0f / (float)P;
  for (long long u0 = warp_global * UPW; u0 < units; u0 += nwarps * UPW) {
    const long long unit_raw = u0 + g;
    const bool valid = unit_raw < units;
    const size_t unit = valid ? (size_t)unit_raw : 0;
    const size_t bq = unit / M;
    float a[SPL];
    FusedGlue<LPR, LP>::softmax(logits + unit * LP, k, a);
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
      const int s = i * LPR + k;
      const int l = s / P;
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const float2 off = ld_stream_f2(offsets + (unit * LP + s) * 2);
      const float* rp = ref + (bq * L + l) * ref_dim;
      const float r0 = rp[0], r1 = rp[1];
      float r2 = 0.0f, r3 = 0.0f;
      if (ref_dim == 4) { r2 = rp[2]; r3 = rp[3]; }
      const float lx = location_from_offset(r0, r2, off.x, (float)W, inv_p, ref_dim);
      const float ly = location_from_offset(r1, r3, off.y, (float)H, inv_p, ref_dim);
      if (valid) {
        if (loc_out) *reinterpret_cast<float2*>(loc_out + (unit * LP + s) * 2) = make_float2(lx, ly);
        if (attn_out) attn_out[unit * LP + s] = a[i];
      }
    }
  }
}

// -----------------------------------------------------------------------------------------------
// Launchers
// -----------------------------------------------------------------------------------------------
namespace {

template <typename T, int D, int NW, int UNROLL, int LPT, int MINB>
int launch_tiled(const FwdParams& p, cudaStream_t stream) {
  constexpr int LPR = D / Elem<T>::kVec, UPW = 32 / LPR;
  const int LP = LPT > 0 ? LPT : p.L * p.P;
  const size_t smem = (size_t)NW * LP * UPW * sizeof(float4);
  auto kern = msda_fwd_tiled_kernel<T, D, NW, UNROLL, LPT, MINB>;
  static PerDeviceOnce configured;   // function attributes are per device
  if (configured.need()) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    // the gather lives on L1 hits: give L1 everything the records do not need
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 15);
  }
  if (smem > 64 * 1024) return MSDA_E_UNSUPPORTED;
  kern<<<p.grid, NW * 32, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

// variant table: (NW, UNROLL, MINB) instantiations tried by bench/tests; 0 is the default
template <typename T, int D, int LPT>
int dispatch_variant(const FwdParams& p, cudaStream_t stream) {
  switch (p.variant) {
    default:
    case 0: return launch_tiled<T, D, 8, 4, LPT, 2>(p, stream);
    case 1: return launch_tiled<T, D, 8, 2, LPT, 3>(p, stream);
    case 2: return launch_tiled<T, D, 4, 4, LPT, 4>(p, stream);
    case 3: return launch_tiled<T, D, 16, 4, LPT, 1>(p, stream);
    case 4: return launch_tiled<T, D, 8, 1, LPT, 4>(p, stream);
    case 5: return launch_tiled<T, D, 4, 2, LPT, 6>(p, stream);
    case 6: return launch_tiled<T, D, 8, 1, LPT, 3>(p, stream);
    case 7: return launch_tiled<T, D, 4, 1, LPT, 8>(p, stream);
    case 8: return launch_tiled<T, D, 16, 1, LPT, 2>(p, stream);
    case 9: return launch_tiled<T, D, 16, 2, LPT, 1>(p, stream);
  }
}

template <typename T>
int launch_forward(const FwdParams& p, cudaStream_t stream) {
  const bool fused = p.loc == nullptr;
  if (p.mode == kModeGeneric) {
    if (fused) return MSDA_E_UNSUPPORTED;
    const long long total = (long long)p.N * p.Lq * p.M * p.D;
    const int grid = (int)((total + 255) / 256 < (long long)p.grid * 8 ? (total + 255) / 256 : (long long)p.grid * 8);
    msda_fwd_generic_kernel<T><<<grid > 0 ? grid : 1, 256, 0, stream>>>(p);
    return (int)cudaGetLastError();
  }
  const int LP = p.L * p.P;
  if (p.mode == kModePipelined) {
    long long hw[4][2], lsi[4];
    if (sizeof(T) == 4 && staged_supported(p) && staged_get_host_shapes(hw, lsi)) {
      FwdParams lv0 = p;
      lv0.grid = p.grid / 4;                              // one 24-warp CTA per SM
      int rc = shape_guard_acquire(&lv0.shape_flag, &lv0.shape_report, &lv0.shape_epoch, stream);
      if (rc == 0) rc = launch_forward_pipelined_f32(lv0, hw, lsi, stream);
      if (rc != MSDA_E_UNSUPPORTED) {
        if (rc) return rc;
        FwdParams rest = p;                               // query levels 1..3: register-gather kernel
        rest.mode = p.tile_h > 0 ? kModePyramid : kModeLinear;     // 2-D tiles for the coarse levels when asked (tuning tile_h / tile_w)
        rest.tile_q = rest.mode == kModePyramid ? p.tile_h * (1 << p.tile_w_log2) : p.tile_q;
        rest.q_level_begin = 1;
        rest.variant = 3;
        rest.shape_flag = lv0.shape_flag;                 // ... or every query, if the host geometry was wrong
        rest.shape_epoch = lv0.shape_epoch;
        return launch_forward_fast_f32(rest, stream);
      }
    }
    FwdParams all = p;                                    // no host geometry: everything on the register-gather kernel
    all.mode = kModeLinear;
    all.variant = 3;
    return sizeof(T) == 4 ? launch_forward_fast_f32(all, stream) : launch_forward_fast_bf16(all, stream);
  }
  if (p.mode == kModeStaged) {
    if (sizeof(T) != 4 || !staged_supported(p)) return MSDA_E_UNSUPPORTED;
    int rc = launch_forward_staged_f32(p, stream);      // queries of pyramid levels 0 .. staged_levels-1
    if (rc || p.staged_levels >= p.L) return rc;
    FwdParams rest = p;                                  // the coarser query levels: register-gather kernel
    rest.mode = kModeLinear;
    rest.q_level_begin = p.staged_levels;
    rest.grid = p.grid * 2;                              // 4 CTAs per SM
    rest.variant = 3;
    return launch_forward_fast_f32(rest, stream);
  }
  if (fast_supported(p) && !p.force_v1)
    return sizeof(T) == 4 ? launch_forward_fast_f32(p, stream) : launch_forward_fast_bf16(p, stream);
  if (p.D == 32) {
    if (!fused) return dispatch_variant<T, 32, 0>(p, stream);
    if (LP == 16) return dispatch_variant<T, 32, 16>(p, stream);
    if (LP == 8) return dispatch_variant<T, 32, 8>(p, stream);
    if (LP == 32) return dispatch_variant<T, 32, 32>(p, stream);
    return MSDA_E_UNSUPPORTED;
  }
  if (p.D == 64) {
    if (!fused) return dispatch_variant<T, 64, 0>(p, stream);
    if (LP == 16) return dispatch_variant<T, 64, 16>(p, stream);
    return MSDA_E_UNSUPPORTED;
  }
  return MSDA_E_UNSUPPORTED;
}

}  // namespace

int forward_variant_count() { return fast_variant_count(); }

bool tiled_supported(int elem_bytes, int D, int L, int P, bool fused) {
  if (D != 32 && D != 64) return false;
  const int lpr = D * elem_bytes / 16, LP = L * P;
  if (L > kMaxLevels || LP > kMaxSamples || LP % lpr != 0) return false;
  if (fused) return D == 32 ? (LP == 16 || LP == 8 || LP == 32) : LP == 16;
  return true;
}

int launch_forward_f32(const FwdParams& p, cudaStream_t stream) { return launch_forward<float>(p, stream); }
int launch_forward_bf16(const FwdParams& p, cudaStream_t stream) { return launch_forward<__nv_bfloat16>(p, stream); }

int launch_sample_index(const float* loc, const int64_t* shapes, const int64_t* lsi, int N, int Lq, int M, int D,
                        int L, int P, void* out_records, cudaStream_t stream) {
  const long long total = (long long)N * Lq * M * L * P;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  msda_sample_index_kernel<<<grid > 0 ? grid : 1, 256, 0, stream>>>(loc, shapes, lsi, total, M, D, L, P,
                                                                    reinterpret_cast<msda_b200_index_t*>(out_records));
  return (int)cudaGetLastError();
}

int launch_locations_softmax(const int64_t* shapes, const float* ref, int ref_dim, const float* offsets,
                             const float* logits, int N, int M, int L, int Lq, int P, int lanes_per_unit,
                             float* loc_out, float* attn_out, cudaStream_t stream) {
  const long long units = (long long)N * Lq * M;
  const int LP = L * P;
  const int upw = 32 / lanes_per_unit;
  const long long warps = (units + upw - 1) / upw;
  long long blocks = (warps + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
#define MSDA_GLUE(LPR_, LP_)                                                                                    \
  if (lanes_per_unit == LPR_ && LP == LP_) {                                                                    \
    msda_glue_kernel<LPR_, LP_><<<(int)blocks, 256, 0, stream>>>(shapes, ref, ref_dim, offsets, logits, units, M, L, P, \
                                                                 loc_out, attn_out);                            \
    return (int)cudaGetLastError();                                                                             \
  }
  MSDA_GLUE(8, 16) MSDA_GLUE(4, 16) MSDA_GLUE(8, 8) MSDA_GLUE(4, 8) MSDA_GLUE(8, 32) MSDA_GLUE(4, 32)
  MSDA_GLUE(16, 16)
#undef MSDA_GLUE
  return MSDA_E_UNSUPPORTED;
}

}  // namespace msda
