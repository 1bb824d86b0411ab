// Device helpers shared by the specialised forward kernels (msda_forward_fast.cu, msda_forward_staged.cu):
// packed fp32x2 arithmetic, register images of value-row slices, streaming vector loads, and the reference's
// per-sample arithmetic two channels at a time.
#pragma once
#include "msda_device.cuh"
#include "msda_launch.h"

namespace msda {

__host__ __device__ constexpr int fast_next_pow2(int x) { int r = 1; while (r < x) r <<= 1; return r; }

// ---- packed fp32x2 arithmetic (one instruction, two IEEE-rounded results) ------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ---- a VB-byte slice of a value row held in registers ------------------------------------------------
template <int VB> struct RowVec;
template <> struct RowVec<16> {
  uint4 a;
  __device__ __forceinline__ void load(const void* p) {
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
  }
  // loads only if pred != 0; otherwise the registers keep their (finite) contents and nothing is written back
  __device__ __forceinline__ void load_if(const void* p, uint32_t pred) {
    asm volatile("{\n .reg .pred pq;\n setp.ne.u32 pq, %5, 0;\n @pq ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n}"
                 : "+r"(a.x), "+r"(a.y), "+r"(a.z), "+r"(a.w) : "l"(p), "r"(pred));
  }
  __device__ __forceinline__ void zero() { a = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ uint32_t word(int i) const { return i == 0 ? a.x : i == 1 ? a.y : i == 2 ? a.z : a.w; }
};
template <> struct RowVec<32> {
  uint4 lo, hi;
  __device__ __forceinline__ void load(const void* p) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
  }
  __device__ __forceinline__ void load_if(const void* p, uint32_t pred) {
    asm volatile("{\n .reg .pred pq;\n setp.ne.u32 pq, %9, 0;\n @pq ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n}"
                 : "+r"(lo.x), "+r"(lo.y), "+r"(lo.z), "+r"(lo.w), "+r"(hi.x), "+r"(hi.y), "+r"(hi.z), "+r"(hi.w)
                 : "l"(p), "r"(pred));
  }
  __device__ __forceinline__ void zero() { lo = make_uint4(0, 0, 0, 0); hi = lo; }
  __device__ __forceinline__ uint32_t word(int i) const {
    return i == 0 ? lo.x : i == 1 ? lo.y : i == 2 ? lo.z : i == 3 ? lo.w : i == 4 ? hi.x : i == 5 ? hi.y : i == 6 ? hi.z : hi.w;
  }
};

// channel pair j (channels 2j, 2j+1) of a row slice as packed fp32x2
template <typename T, int VB>
__device__ __forceinline__ f32x2 chan_pair(const RowVec<VB>& r, int j) {
  if constexpr (sizeof(T) == 4) {
    return pk(__uint_as_float(r.word(2 * j)), __uint_as_float(r.word(2 * j + 1)));
  } else {   // one 32-bit word holds two bf16: low half = even channel
    const uint32_t w = r.word(j);
    return pk(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
  }
}

template <typename T, int NP>
__device__ __forceinline__ void store_row(void* p, const f32x2 (&acc)[NP]) {
  float f[2 * NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) upk(acc[j], f[2 * j], f[2 * j + 1]);
  if constexpr (sizeof(T) == 4) {
    if constexpr (NP == 2) {
      st_stream16(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
    } else {
      st_stream32(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])),
                  make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
    }
  } else {
    using E = Elem<__nv_bfloat16>;
    if constexpr (NP == 4) {
      st_stream16(p, make_uint4(E::pack2(f[0], f[1]), E::pack2(f[2], f[3]), E::pack2(f[4], f[5]), E::pack2(f[6], f[7])));
    } else {
      st_stream32(p, make_uint4(E::pack2(f[0], f[1]), E::pack2(f[2], f[3]), E::pack2(f[4], f[5]), E::pack2(f[6], f[7])),
                  make_uint4(E::pack2(f[8], f[9]), E::pack2(f[10], f[11]), E::pack2(f[12], f[13]), E::pack2(f[14], f[15])));
    }
  }
}

// N consecutive floats of a streamed-once operand (N = 2, 4, 8), one vector load, no L1 allocation, L2 evict-first
template <int N>
__device__ __forceinline__ void ld_stream_vec(const float* p, float (&f)[N]) {
  static_assert(N == 1 || N == 2 || N == 4 || N == 8 || N == 16, "vector width");
  if constexpr (N == 1) {
    f[0] = ld_stream_f1(p);
  } else if constexpr (N == 2) {
    const float2 r = ld_stream_f2(p);
    f[0] = r.x; f[1] = r.y;
  } else if constexpr (N == 4) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "l"(p), "l"(l2_evict_first_policy()));
  } else if constexpr (N == 8) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7])
                 : "l"(p), "l"(l2_evict_first_policy()));
  } else {
    float lo[8], hi[8];
    ld_stream_vec<8>(p, lo);
    ld_stream_vec<8>(p + 8, hi);
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = lo[i]; f[8 + i] = hi[i]; }
  }
}

// The reference's arithmetic for one sample, two channels at a time (cuh:80-82, :290):
//   val = FMUL(w2,v2) -> FFMA(w1,v1,.) -> FFMA(w3,v3,.) -> FFMA(w4,v4,.) ; acc = FFMA(attn, val, acc).
// Each packed instruction rounds its two elements exactly like the scalar one.
template <typename T, int VB, int NP>
__device__ __forceinline__ void accumulate_sample(f32x2 (&acc)[NP], const float4& w, float attn, const RowVec<VB>& q1,
                                                  const RowVec<VB>& q2, const RowVec<VB>& q3, const RowVec<VB>& q4) {
  const f32x2 W1 = pk(w.x, w.x), W2 = pk(w.y, w.y), W3 = pk(w.z, w.z), W4 = pk(w.w, w.w), A = pk(attn, attn);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    f32x2 t = mul2(W2, chan_pair<T, VB>(q2, j));
    t = fma2(W1, chan_pair<T, VB>(q1, j), t);
    t = fma2(W3, chan_pair<T, VB>(q3, j), t);
    t = fma2(W4, chan_pair<T, VB>(q4, j), t);
    acc[j] = fma2(A, t, acc[j]);
  }
}

// Softmax over the LPT logits of one unit, spread over LPR lanes with SPL = LPT / LPR consecutive elements per lane
// (element e = k * SPL + i lives in register i of lane k), in the operation order of PyTorch's persistent warp softmax
// for <= 32 elements: one element per virtual lane, xor butterfly over the element index LPT/2 .. 1.  Index bits
// >= log2(SPL) are lane bits (shuffle), the rest are register bits (in-lane pairs); fp32 addition is commutative, so
// both partners of a step get the same sum.  Returns the un-normalised exponentials in e[] and their sum -- the caller
// divides (IEEE), like the reference's softmax (ms_deform_attn.py:138-139).  All 32 lanes must call it.
template <int SPL, int LPR, int LPT>
__device__ __forceinline__ float unit_softmax_terms(const float (&lg)[SPL], float (&e)[SPL]) {
  static_assert(fast_next_pow2(LPT) == LPT && LPT <= 32 && SPL * LPR == LPT, "L*P must be a power of two <= 32");
  float mx = lg[0];
#pragma unroll
  for (int i = 1; i < SPL; ++i) mx = fmaxf(mx, lg[i]);
#pragma unroll
  for (int off = LPR / 2; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float v[SPL];
#pragma unroll
  for (int i = 0; i < SPL; ++i) { e[i] = expf(__fsub_rn(lg[i], mx)); v[i] = e[i]; }
#pragma unroll
  for (int o = LPT / 2; o >= 1; o >>= 1) {
    if (o >= SPL) {
#pragma unroll
      for (int i = 0; i < SPL; ++i) v[i] = __fadd_rn(v[i], __shfl_xor_sync(0xffffffffu, v[i], o / SPL));
    } else {
      float t[SPL];
#pragma unroll
      for (int i = 0; i < SPL; ++i) t[i] = __fadd_rn(v[i], v[i ^ o]);
#pragma unroll
      for (int i = 0; i < SPL; ++i) v[i] = t[i];
    }
  }
  return v[0];
}

// Raw operands of one lane's SPL samples between the prefetch and phase 1.
template <int SPL, int NLV, bool FUSED> struct Prefetched;
template <int SPL, int NLV> struct Prefetched<SPL, NLV, false> {
  float xy[2 * SPL];
  float a[SPL];
};
template <int SPL, int NLV> struct Prefetched<SPL, NLV, true> {
  float off[2 * SPL];
  float lg[SPL];
  float4 ref[NLV];
};

// Issue the streaming loads of one lane's phase-1 operands: samples first .. first+SPL-1 of unit (bq, m), where
// bq = batch * Lq + query.  Core entry: sampling locations and attention weights; fused entry: raw offsets, logits and
// the reference point(s) of the NLV levels the samples span (first level lvl0).  NL levels, LPT = L * P samples.
template <int SPL, int NLV, bool FUSED, int LPT, int NL>
__device__ __forceinline__ void load_unit_operands(Prefetched<SPL, NLV, FUSED>& pf, const FwdParams& p, size_t bq, int m,
                                                   int first, int lvl0) {
  if constexpr (FUSED) {
    // rows of offsets / logits may be slices of one merged projection output (pitch > dense row length)
    ld_stream_vec<SPL>(p.logits + bq * p.logit_pitch + m * LPT + first, pf.lg);
    ld_stream_vec<2 * SPL>(p.offsets + bq * p.off_pitch + (m * LPT + first) * 2, pf.off);
#pragma unroll
    for (int j = 0; j < NLV; ++j) {
      const float* rp = p.ref + (bq * NL + lvl0 + j) * p.ref_dim;
      if (p.ref_dim == 4) {
        pf.ref[j] = __ldg(reinterpret_cast<const float4*>(rp));
      } else {
        const float2 r2 = __ldg(reinterpret_cast<const float2*>(rp));
        pf.ref[j] = make_float4(r2.x, r2.y, 0.0f, 0.0f);
      }
    }
  } else {
    const size_t unit = bq * p.M + m;
    ld_stream_vec<2 * SPL>(p.loc + (unit * LPT + first) * 2, pf.xy);
    ld_stream_vec<SPL>(p.attn + unit * LPT + first, pf.a);
  }
}

}  // namespace msda
