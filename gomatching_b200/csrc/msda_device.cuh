// Shared device functions of the B200 MSDeformAttn kernels.
//
// Every kernel in this library (tiled forward, generic forward, fused forward, index dump, backward)
// derives sampling indices through `sample_setup()` below, so the index-parity tests exercise exactly
// the arithmetic the product kernels run.
//
// Reference being matched (third_party/adet/layers/csrc/DeformAttn/ms_deform_im2col_cuda.cuh):
//   :285-286  h_im = loc_h * spatial_h - 0.5 ; w_im = loc_w * spatial_w - 0.5
//             -> as compiled by nvcc 12.9 for sm_100a this is ONE FFMA (size, loc, -0.5); the fused
//                rounding is the contract, written here as an explicit __fmaf_rn.
//   :288      sample contributes iff h_im > -1 && w_im > -1 && h_im < H && w_im < W
//   :39-45    h_low = floor(h_im) ; lh = h_im - h_low ; hh = 1 - lh
//   :56-78    corner validity (zero padding), :80-82 weights and the weighted corner sum
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace msda {

constexpr int kMaxLevels = 16;      // levels the tiled kernels cache in shared memory
constexpr int kMaxSamples = 64;     // L*P the tiled kernels support (larger -> generic kernel)

struct SampleGeom {
  float lh, lw;        // fractional parts (cuh:43-44)
  int h_low, w_low;    // floor(h_im), floor(w_im): the "sampling indices"
  int mask;            // bit0 (h_low,w_low) bit1 (h_low,w_high) bit2 (h_high,w_low) bit3 (h_high,w_high)
  bool in_range;       // cuh:288
};

// Level sizes are passed both as int and as the (exact) float the reference converts them to.
__device__ __forceinline__ SampleGeom sample_setup_f(float loc_w, float loc_h, float Hf, float Wf, int H, int W) {
  SampleGeom g;
  const float h_im = __fmaf_rn(loc_h, Hf, -0.5f);
  const float w_im = __fmaf_rn(loc_w, Wf, -0.5f);
  g.in_range = (h_im > -1.0f) && (w_im > -1.0f) && (h_im < Hf) && (w_im < Wf);
  g.h_low = 0; g.w_low = 0; g.mask = 0; g.lh = 0.0f; g.lw = 0.0f;
  if (g.in_range) {
    const float hf = floorf(h_im), wf = floorf(w_im);
    g.h_low = (int)hf;
    g.w_low = (int)wf;
    g.lh = __fsub_rn(h_im, hf);
    g.lw = __fsub_rn(w_im, wf);
    const bool t = g.h_low >= 0, b = g.h_low + 1 <= H - 1;
    const bool l = g.w_low >= 0, r = g.w_low + 1 <= W - 1;
    g.mask = (int)(t && l) | ((int)(t && r) << 1) | ((int)(b && l) << 2) | ((int)(b && r) << 3);
  }
  return g;
}

__device__ __forceinline__ SampleGeom sample_setup(float loc_w, float loc_h, int H, int W) {
  return sample_setup_f(loc_w, loc_h, (float)H, (float)W, H, W);
}

// The four bilinear weights in the reference's order (cuh:80): w1=hh*hw w2=hh*lw w3=lh*hw w4=lh*lw
__device__ __forceinline__ void bilinear_weights(float lh, float lw, float& w1, float& w2, float& w3, float& w4) {
  const float hh = __fsub_rn(1.0f, lh), hw = __fsub_rn(1.0f, lw);
  w1 = __fmul_rn(hh, hw);
  w2 = __fmul_rn(hh, lw);
  w3 = __fmul_rn(lh, hw);
  w4 = __fmul_rn(lh, lw);
}

// val = w1*v1 + w2*v2 + w3*v3 + w4*v4 (cuh:82) in the operation order of the reference SASS:
//   FMUL(w2,v2) -> FFMA(w1,v1,.) -> FFMA(w3,v3,.) -> FFMA(w4,v4,.) ; then col = FFMA(attn, val, col) (cuh:290)
__device__ __forceinline__ float corner_accumulate(float acc, float a, float w1, float w2, float w3, float w4,
                                                   float v1, float v2, float v3, float v4) {
  float t = __fmul_rn(w2, v2);
  t = __fmaf_rn(w1, v1, t);
  t = __fmaf_rn(w3, v3, t);
  t = __fmaf_rn(w4, v4, t);
  return __fmaf_rn(a, t, acc);
}

// offsets -> locations, ms_deform_attn.py:141-147, each eager op rounded separately.
//   ref_dim 2: ref + off / (float)size         (true division, then add)
//   ref_dim 4: ref + ((off * (1/P)) * ref_wh) * 0.5   (CUDA `tensor / python_scalar` multiplies by the reciprocal)
__device__ __forceinline__ float location_from_offset(float ref, float ref_wh, float off, float size_f, float inv_p,
                                                      int ref_dim) {
  if (ref_dim == 2) return __fadd_rn(ref, __fdiv_rn(off, size_f));
  return __fadd_rn(ref, __fmul_rn(__fmul_rn(__fmul_rn(off, inv_p), ref_wh), 0.5f));
}

// ---- 16-byte loads/stores with cache hints ---------------------------------------------------------
// value rows: read-only path, keep in L1 (they are re-read by neighbouring queries)
__device__ __forceinline__ uint4 ld_value16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// streamed-once operands (loc / attn / offsets / logits): do not displace value rows in L1, and tell L2 to evict
// them first -- ncu showed 1.24x the algorithmic DRAM bytes per encoder launch because the streams pushed value
// rows (re-read ~64x each) out of L2.  The policy is a constant; the non-volatile asm lets nvcc hoist it.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float2 ld_stream_f2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
               : "=f"(r.x), "=f"(r.y) : "l"(p), "l"(l2_evict_first_policy()));
  return r;
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(l2_evict_first_policy()));
  return r;
}
// 256-bit global load (sm_100a LDG.E.256): 4 lanes cover one 128-byte fp32 row, a warp gathers 8 rows per instruction
struct __align__(32) Vec32B { uint4 lo, hi; };
__device__ __forceinline__ void ld_value32(const void* p, Vec32B& r) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "+r"(r.lo.x), "+r"(r.lo.y), "+r"(r.lo.z), "+r"(r.lo.w), "+r"(r.hi.x), "+r"(r.hi.y), "+r"(r.hi.z), "+r"(r.hi.w)
               : "l"(p));
}
__device__ __forceinline__ void ld_value16_keep(const void* p, uint4& r) {
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w) : "l"(p));
}
__device__ __forceinline__ void st_stream32(void* p, const uint4& lo, const uint4& hi) {
  asm volatile("st.global.cs.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w),
               "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
}
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- element traits: how 16 bytes of a value row turn into fp32 lanes ------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int kVec = 4;  // elements per 16-byte load
  __device__ static __forceinline__ void unpack(const uint4& u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  __device__ static __forceinline__ uint4 pack(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
  __device__ static __forceinline__ float load1(const float* p) { return __ldg(p); }
  __device__ static __forceinline__ void store1(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int kVec = 8;
  __device__ static __forceinline__ void unpack(const uint4& u, float (&f)[8]) {
    // bf16 -> fp32 is a 16-bit shift: low half = element 0, high half = element 1
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
  }
  __device__ static __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);   // round-to-nearest-even
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  __device__ static __forceinline__ uint4 pack(const float (&f)[8]) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
  __device__ static __forceinline__ float load1(const __nv_bfloat16* p) {
    return __uint_as_float(((uint32_t) * reinterpret_cast<const unsigned short*>(p)) << 16);
  }
  __device__ static __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

}  // namespace msda
