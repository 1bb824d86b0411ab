// Fast forward kernels for the DeepSolo configuration (L*P = 16, P = 4 known at compile time, D = 32).
//
// Same two-phase scheme as msda_forward.cu (phase 1: per-sample records computed once per unit and parked in
// shared memory; phase 2: gather + weighted reduction), re-cut from what ncu and the gather microbenchmark
// (tools/ubench/gather_bw.cu, profiles/r01_ubench_gather_v2.log) showed on B200:
//
//   * 128-bit accesses -- LDG.128 and LDS.128 alike -- top out at ~62 B/clk/SM; LDG.E.256 (new on sm_100a)
//     reaches 122 B/clk/SM on L1 hits and 73 B/clk/SM from L2.  So the gather uses 256-bit loads (VB = 32: 4
//     lanes cover a 128-byte fp32 row, a warp gathers 8 rows per instruction) and the value rows are NOT staged
//     in shared memory: an LDS.128 gather would be half as fast as LDG.256 through L1.
//   * with every sample out of range (no gather at all) the previous revision still took 60 % of its normal
//     time: it was instruction-issue bound.  Hence (a) packed fp32x2 arithmetic (FFMA2/FMUL2/FADD2, sm_100a;
//     same IEEE rounding per element, so results stay bit-identical to the reference's FMA chain) -- 20 FP
//     instructions per 8-unit step instead of 40; (b) a warp-uniform three-way split per sample step:
//       all 4 corners of all units valid  -> plain loads, no zero-fill, no predicates            (common)
//       every unit out of range           -> step skipped                                          (borders)
//       otherwise                         -> zero-filled, predicated loads                         (rare)
//     (c) L*P, P and the pixel pitch M*D*sizeof(T) are template constants: unrolled sample loop, the
//     horizontal-neighbour offset is an immediate of the load.
//   * the next step's loc/attn (or offsets/logits/reference points) are prefetched into registers before the
//     current step's gather starts, so their DRAM latency hides behind the gather.
//   * LDG.256 destination registers are pure asm outputs: a read-write ("+r") operand made nvcc wrap the load
//     in MOVs that wait for it (ncu source view), serialising the gather.
#include "msda_device.cuh"
#include "msda_launch.h"
#include "../../include/msda_b200.h"

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
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ---- a VB-byte slice of a value row held in registers ------------------------------------------------
template <int VB> struct RowVec;
template <> struct RowVec<16> {
  uint4 a;
  __device__ __forceinline__ void load(const void* p) {
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
  }
  __device__ __forceinline__ void load_or_zero(const void* p, bool pred) {
    asm volatile(
        "{\n .reg .pred pq;\n setp.ne.u32 pq, %5, 0;\n"
        " mov.u32 %0, 0; mov.u32 %1, 0; mov.u32 %2, 0; mov.u32 %3, 0;\n"
        " @pq ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n}"
        : "=&r"(a.x), "=&r"(a.y), "=&r"(a.z), "=&r"(a.w)
        : "l"(p), "r"((uint32_t)pred));
  }
  // pred: this lane loads; nofill (warp-uniform): every lane of the warp loads, so the zero-fill is branched over
  __device__ __forceinline__ void load_sel(const void* p, bool pred, bool nofill) {
    asm volatile(
        "{\n .reg .pred pq, pf;\n setp.ne.u32 pq, %5, 0;\n setp.ne.u32 pf, %6, 0;\n"
        " @pf bra.uni FILLED;\n"
        " mov.u32 %0, 0; mov.u32 %1, 0; mov.u32 %2, 0; mov.u32 %3, 0;\n"
        "FILLED:\n"
        " @pq ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n}"
        : "=&r"(a.x), "=&r"(a.y), "=&r"(a.z), "=&r"(a.w)
        : "l"(p), "r"((uint32_t)pred), "r"((uint32_t)nofill));
  }
  __device__ __forceinline__ uint32_t word(int i) const { return i == 0 ? a.x : i == 1 ? a.y : i == 2 ? a.z : a.w; }
};
template <> struct RowVec<32> {
  uint4 lo, hi;
  __device__ __forceinline__ void load(const void* p) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
  }
  __device__ __forceinline__ void load_or_zero(const void* p, bool pred) {
    asm volatile(
        "{\n .reg .pred pq;\n setp.ne.u32 pq, %9, 0;\n"
        " mov.u32 %0, 0; mov.u32 %1, 0; mov.u32 %2, 0; mov.u32 %3, 0;\n"
        " mov.u32 %4, 0; mov.u32 %5, 0; mov.u32 %6, 0; mov.u32 %7, 0;\n"
        " @pq ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n}"
        : "=&r"(lo.x), "=&r"(lo.y), "=&r"(lo.z), "=&r"(lo.w), "=&r"(hi.x), "=&r"(hi.y), "=&r"(hi.z), "=&r"(hi.w)
        : "l"(p), "r"((uint32_t)pred));
  }
  __device__ __forceinline__ void load_sel(const void* p, bool pred, bool nofill) {
    asm volatile(
        "{\n .reg .pred pq, pf;\n setp.ne.u32 pq, %9, 0;\n setp.ne.u32 pf, %10, 0;\n"
        " @pf bra.uni FILLED;\n"
        " mov.u32 %0, 0; mov.u32 %1, 0; mov.u32 %2, 0; mov.u32 %3, 0;\n"
        " mov.u32 %4, 0; mov.u32 %5, 0; mov.u32 %6, 0; mov.u32 %7, 0;\n"
        "FILLED:\n"
        " @pq ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n}"
        : "=&r"(lo.x), "=&r"(lo.y), "=&r"(lo.z), "=&r"(lo.w), "=&r"(hi.x), "=&r"(hi.y), "=&r"(hi.z), "=&r"(hi.w)
        : "l"(p), "r"((uint32_t)pred), "r"((uint32_t)nofill));
  }
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

// The reference's arithmetic for one sample, two channels at a time (cuh:80-82, :290):
//   w1=hh*hw w2=hh*lw w3=lh*hw w4=lh*lw ; val = FMUL(w2,v2) -> FFMA(w1,v1,.) -> FFMA(w3,v3,.) -> FFMA(w4,v4,.)
//   acc = FFMA(attn, val, acc).  Each packed instruction rounds its two elements exactly like the scalar one.
template <typename T, int VB, int NP>
__device__ __forceinline__ void accumulate_sample(f32x2 (&acc)[NP], const float4& rc, const RowVec<VB>& q1,
                                                  const RowVec<VB>& q2, const RowVec<VB>& q3, const RowVec<VB>& q4) {
  const float lh = rc.y, lw = rc.z;
  const f32x2 one = pk(1.0f, 1.0f);
  const f32x2 h2 = sub2(one, pk(lh, lw));          // {hh, hw}
  float hh, hw;
  upk(h2, hh, hw);
  const f32x2 cw = pk(hw, lw);                     // {hw, lw}
  const f32x2 wtop = mul2(pk(hh, hh), cw);         // {w1, w2}
  const f32x2 wbot = mul2(pk(lh, lh), cw);         // {w3, w4}
  float w1, w2, w3, w4;
  upk(wtop, w1, w2);
  upk(wbot, w3, w4);
  const f32x2 W1 = pk(w1, w1), W2 = pk(w2, w2), W3 = pk(w3, w3), W4 = pk(w4, w4), A = pk(rc.w, rc.w);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    f32x2 t = mul2(W2, chan_pair<T, VB>(q2, j));
    t = fma2(W1, chan_pair<T, VB>(q1, j), t);
    t = fma2(W3, chan_pair<T, VB>(q3, j), t);
    t = fma2(W4, chan_pair<T, VB>(q4, j), t);
    acc[j] = fma2(A, t, acc[j]);
  }
}

// Raw operands of one unit's samples held by one lane between the prefetch and phase 1.
template <int SPL, bool FUSED> struct Prefetched;
template <int SPL> struct Prefetched<SPL, false> {
  float2 xy[SPL];
  float a[SPL];
};
template <int SPL> struct Prefetched<SPL, true> {
  float2 off[SPL];
  float lg[SPL];
  float4 ref[SPL];
};

//   T     float | __nv_bfloat16 storage of value/out (arithmetic fp32)
//   VB    bytes of a value row one lane loads (32: LDG.E.256, 16: LDG.E.128)
//   LPT   L*P, PT = P (compile time) ; CSB = M*D*sizeof(T) if known at compile time else 0
//   NW    warps per CTA ; MINB min CTAs per SM (register budget)
//   PD    gather pipeline depth: the corner loads of sample s+PD-1 are issued before sample s is consumed, so a
//         warp keeps 4*(PD-1)..4*PD row loads in flight (ncu: with PD=1 the kernel is latency-bound -- almost
//         every step waits for an L2 round trip because one of its 32 rows misses L1)
template <typename T, int D, int VB, int LPT, int PT, int CSB, bool FUSED, int NW, int MINB, int PD>
__global__ void __launch_bounds__(NW * 32, MINB) msda_fwd_fast_kernel(const FwdParams p) {
  constexpr int EB = (int)sizeof(T);
  constexpr int VEC = VB / EB;          // channels per lane
  constexpr int NP = VEC / 2;           // channel pairs per lane
  constexpr int LPR = D / VEC;          // lanes per value row
  constexpr int UPW = 32 / LPR;         // units per warp step
  constexpr int SPL = LPT / LPR;        // samples per lane in phase 1
  constexpr int NL = LPT / PT;          // levels
  static_assert(D % VEC == 0 && LPR >= 1 && 32 % LPR == 0, "bad lane layout");
  static_assert(LPT % LPR == 0 && LPT % PT == 0 && SPL >= 1, "bad sample layout");
  // Record slot of (sample s, lane group g): s*UPW + (g ^ swz(s)).  The XOR swizzle makes the phase-1 STS.128
  // of a quarter-warp (8 lanes = 8 different (s,g)) hit 8 different 16-byte bank groups; without it lanes with
  // equal g collide 4-way (ncu: 64 instead of 16 wavefronts per step).  swz(s) takes only 4 values, so phase 2
  // keeps 4 pre-swizzled base pointers and every record read is LDS.128 [base_c + immediate].
  auto swz = [](int s) -> int { return (LPR == 8 ? (s >> 1) : s * (8 / LPR)) & (UPW - 1); };
  auto slot = [&](int s, int g) -> int { return s * UPW + (g ^ swz(s)); };
  constexpr int kSwzStep = LPR == 8 ? 1 : 8 / LPR;      // swz(s) in {0,1,2,3} * kSwzStep

  __shared__ int sH[NL], sW[NL], sStart[NL], sTileCum[NL + 1];
  __shared__ float sHf[NL], sWf[NL];
  extern __shared__ float4 sRecAll[];   // [NW][LPT][UPW]

  const int M = p.M, Lq = p.Lq;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane / LPR, k = lane % LPR;
  if (tid < NL) {
    sH[tid] = (int)p.shapes[2 * tid];
    sW[tid] = (int)p.shapes[2 * tid + 1];
    sStart[tid] = (int)p.lsi[tid];
    sHf[tid] = (float)sH[tid];
    sWf[tid] = (float)sW[tid];
  }
  __syncthreads();
  const int tw_log2 = p.tile_w_log2, TH = p.tile_h;
  const bool pyramid = p.mode == kModePyramid;
  if (pyramid && tid == 0) {
    int cum = 0;
    for (int l = 0; l < NL; ++l) {
      sTileCum[l] = cum;
      cum += ((sH[l] + TH - 1) / TH) * ((sW[l] + (1 << tw_log2) - 1) >> tw_log2);
    }
    sTileCum[NL] = cum;
  }
  __syncthreads();

  const int cstride = CSB ? CSB : M * D * EB;      // bytes between horizontally adjacent pixels
  int rstride[NL];                                  // bytes between vertically adjacent pixels, per level
#pragma unroll
  for (int l = 0; l < NL; ++l) rstride[l] = sW[l] * cstride;

  const int tiles_per_bm = pyramid ? sTileCum[NL] : (Lq + p.tile_q - 1) / p.tile_q;
  const long long total_tiles = (long long)p.N * tiles_per_bm * M;
  const int chunks_per_warp = (p.tile_q + NW * UPW - 1) / (NW * UPW);   // warp steps per tile
  float4* sRec = sRecAll + (size_t)warp * LPT * UPW;
  const float4* sRecG[4];                           // this lane group's record column under each swizzle value
#pragma unroll
  for (int c = 0; c < 4; ++c) sRecG[c] = sRec + (g ^ ((c * kSwzStep) & (UPW - 1)));
  const float inv_p = 1.0f / (float)PT;

  // ---- work cursor: (tile, chunk) pairs of this warp, flattened so the next step can be prefetched ----
  long long tile = blockIdx.x;
  int chunk = 0;
  int t_m = 0, t_b = 0, t_t = 0, t_lvl = 0, t_ty = 0, t_tx = 0;
  auto decode_tile = [&]() {
    t_m = (int)(tile % M);
    const long long r = tile / M;
    t_t = (int)(r % tiles_per_bm);
    t_b = (int)(r / tiles_per_bm);
    if (pyramid) {
      t_lvl = 0;
      while (t_lvl + 1 < NL && t_t >= sTileCum[t_lvl + 1]) ++t_lvl;
      const int tt = t_t - sTileCum[t_lvl];
      const int ntx = (sW[t_lvl] + (1 << tw_log2) - 1) >> tw_log2;
      t_ty = tt / ntx;
      t_tx = tt - t_ty * ntx;
    }
  };
  auto locate = [&](bool& valid, size_t& bq, size_t& unit) {
    const int j = (chunk * NW + warp) * UPW + g;
    int qi;
    if (pyramid) {
      const int y = t_ty * TH + (j >> tw_log2), x = (t_tx << tw_log2) + (j & ((1 << tw_log2) - 1));
      valid = (j < p.tile_q) && (y < sH[t_lvl]) && (x < sW[t_lvl]);
      qi = sStart[t_lvl] + y * sW[t_lvl] + x;
    } else {
      qi = t_t * p.tile_q + j;
      valid = (j < p.tile_q) && (qi < Lq);
    }
    bq = (size_t)t_b * Lq + (valid ? qi : 0);
    unit = bq * M + t_m;
  };
  auto prefetch = [&](Prefetched<SPL, FUSED>& pf, size_t bq, size_t unit) {
    if constexpr (FUSED) {
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        const int s = i * LPR + k;
        // rows of offsets / logits may be slices of one merged projection output (pitch > dense row length)
        pf.lg[i] = ld_stream_f1(p.logits + bq * p.logit_pitch + t_m * LPT + s);
        pf.off[i] = ld_stream_f2(p.offsets + bq * p.off_pitch + (t_m * LPT + s) * 2);
        const float* rp = p.ref + (bq * NL + s / PT) * p.ref_dim;
        if (p.ref_dim == 4) {
          pf.ref[i] = __ldg(reinterpret_cast<const float4*>(rp));
        } else {
          const float2 r2 = __ldg(reinterpret_cast<const float2*>(rp));
          pf.ref[i] = make_float4(r2.x, r2.y, 0.0f, 0.0f);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        const int s = i * LPR + k;
        pf.xy[i] = ld_stream_f2(p.loc + (unit * LPT + s) * 2);
        pf.a[i] = ld_stream_f1(p.attn + unit * LPT + s);
      }
    }
  };

  bool have = tile < total_tiles;
  bool n_valid = false;
  size_t n_bq = 0, n_unit = 0;
  const char* n_vbase = nullptr;
  Prefetched<SPL, FUSED> pf;
  if (have) {
    decode_tile();
    locate(n_valid, n_bq, n_unit);
    n_vbase = reinterpret_cast<const char*>(p.value) + ((size_t)t_b * p.S * M * D + (size_t)t_m * D + (size_t)k * VEC) * EB;
    prefetch(pf, n_bq, n_unit);
  }

  while (have) {
    const bool valid = n_valid;
    const size_t unit = n_unit;
    const char* vbase = n_vbase;

    // ---------------- phase 1: records from the prefetched operands ----------------
    {
      float a[SPL], lx[SPL], ly[SPL];
      if constexpr (FUSED) {
        // softmax over the unit's LPT logits in the operation order of PyTorch's persistent warp softmax
        // (element e on virtual lane e % WS, per-lane sequential sum, xor butterfly WS/2..1)
        constexpr int NP2 = fast_next_pow2(LPT), WS = NP2 < 32 ? NP2 : 32, R = WS / LPR;
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < SPL; ++i) mx = fmaxf(mx, pf.lg[i]);
#pragma unroll
        for (int off = LPR / 2; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        float vs[R];
#pragma unroll
        for (int j = 0; j < R; ++j) vs[j] = 0.0f;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          a[i] = expf(__fsub_rn(pf.lg[i], mx));
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
        for (int i = 0; i < SPL; ++i) {
          const int l = (i * LPR + k) / PT;
          a[i] = __fdiv_rn(a[i], sum);
          lx[i] = location_from_offset(pf.ref[i].x, pf.ref[i].z, pf.off[i].x, sWf[l], inv_p, p.ref_dim);
          ly[i] = location_from_offset(pf.ref[i].y, pf.ref[i].w, pf.off[i].y, sHf[l], inv_p, p.ref_dim);
        }
      } else {
#pragma unroll
        for (int i = 0; i < SPL; ++i) { lx[i] = pf.xy[i].x; ly[i] = pf.xy[i].y; a[i] = pf.a[i]; }
      }
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        const int s = i * LPR + k;
        const int l = s / PT;
        const int Wl = sW[l];
        const SampleGeom sg = sample_setup_f(lx[i], ly[i], sHf[l], sWf[l], sH[l], Wl);
        const bool inr = valid && sg.in_range;
        int packed = 0;
        if (inr) packed = (((sStart[l] + sg.h_low * Wl + sg.w_low) * M * D) * EB) | sg.mask;
        sRec[slot(s, g)] = make_float4(__int_as_float(packed), sg.lh, sg.lw, inr ? a[i] : 0.0f);
      }
    }
    __syncwarp();

    // ---------------- advance the cursor and prefetch the next step's operands ----------------
    ++chunk;
    if (chunk == chunks_per_warp) {
      chunk = 0;
      tile += gridDim.x;
      have = tile < total_tiles;
      if (have) {
        decode_tile();
        n_vbase = reinterpret_cast<const char*>(p.value) +
                  ((size_t)t_b * p.S * M * D + (size_t)t_m * D + (size_t)k * VEC) * EB;
      }
    }
    if (have) {
      locate(n_valid, n_bq, n_unit);
      prefetch(pf, n_bq, n_unit);
    }

    // ---------------- phase 2: gather + weighted reduction ----------------
    f32x2 acc[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) acc[j] = 0ull;
    auto rec_at = [&](int s) -> float4 { return sRecG[(swz(s) / kSwzStep) & 3][s * UPW]; };
    // warp-uniform three-way split per sample:
    //   all 4 corners of all units valid -> plain loads, no zero-fill, no predicates                  (common)
    //   every unit out of range          -> nothing issued, nothing accumulated (cuh:288)              (borders)
    //   otherwise                        -> zero-filled, predicated loads = zero padding               (rare)
    if constexpr (PD == 1) {
#pragma unroll
      for (int s = 0; s < LPT; ++s) {
        const float4 rc = rec_at(s);
        const int packed = __float_as_int(rc.x);
        const int mk = packed & 15;
        const char* c1 = vbase + (ptrdiff_t)(packed & ~15);
        const char* c3 = c1 + rstride[s / PT];
        RowVec<VB> q1, q2, q3, q4;
        if (__all_sync(0xffffffffu, mk == 15)) {
          q1.load(c1);
          q2.load(c1 + cstride);
          q3.load(c3);
          q4.load(c3 + cstride);
          accumulate_sample<T, VB, NP>(acc, rc, q1, q2, q3, q4);
        } else if (__any_sync(0xffffffffu, mk != 0)) {
          q1.load_or_zero(c1, mk & 1);
          q2.load_or_zero(c1 + cstride, mk & 2);
          q3.load_or_zero(c3, mk & 4);
          q4.load_or_zero(c3 + cstride, mk & 8);
          accumulate_sample<T, VB, NP>(acc, rc, q1, q2, q3, q4);
        }
      }
    } else {
      RowVec<VB> q[PD][4];
      float4 rc[PD];
      bool live[PD];            // warp-uniform: at least one unit takes this sample
      auto issue = [&](int s, int d) {
        rc[d] = rec_at(s);
        const int packed = __float_as_int(rc[d].x);
        const int mk = packed & 15;
        const bool all_in = __all_sync(0xffffffffu, mk == 15);
        live[d] = all_in || __any_sync(0xffffffffu, mk != 0);
        if (live[d]) {
          const char* c1 = vbase + (ptrdiff_t)(packed & ~15);
          const char* c3 = c1 + rstride[s / PT];
          q[d][0].load_sel(c1, mk & 1, all_in);          // zero-fill branched over inside the asm when all_in
          q[d][1].load_sel(c1 + cstride, mk & 2, all_in);
          q[d][2].load_sel(c3, mk & 4, all_in);
          q[d][3].load_sel(c3 + cstride, mk & 8, all_in);
        }
      };
#pragma unroll
      for (int s = 0; s < PD - 1; ++s) issue(s, s);
#pragma unroll
      for (int s = 0; s < LPT; ++s) {
        if (s + PD - 1 < LPT) issue(s + PD - 1, (s + PD - 1) % PD);
        const int d = s % PD;
        if (live[d]) accumulate_sample<T, VB, NP>(acc, rc[d], q[d][0], q[d][1], q[d][2], q[d][3]);
      }
    }
    if (valid) {
      char* op = reinterpret_cast<char*>(p.out) + (unit * D + (size_t)k * VEC) * EB;
      store_row<T, NP>(op, acc);
    }
    __syncwarp();   // records are rewritten by the next step's phase 1
  }
}

// -----------------------------------------------------------------------------------------------
namespace {

template <typename T, int D, int VB, int LPT, int PT, int CSB, bool FUSED, int NW, int MINB, int PD>
int launch_fast(const FwdParams& p, cudaStream_t stream) {
  constexpr int LPR = D / (VB / (int)sizeof(T)), UPW = 32 / LPR;
  const size_t smem = (size_t)NW * LPT * UPW * sizeof(float4);
  auto kern = msda_fwd_fast_kernel<T, D, VB, LPT, PT, CSB, FUSED, NW, MINB, PD>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    // the gather lives on L1 hits: shared memory only holds the records
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 20);
    configured = true;
  }
  kern<<<p.grid, NW * 32, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

template <typename T, int LPT, int PT, int CSB, bool FUSED>
int dispatch_fast_variant(const FwdParams& p, cudaStream_t stream) {
  switch (p.variant) {   //                  D  VB                      NW MINB PD
    default:
    case 0: return launch_fast<T, 32, 16, LPT, PT, CSB, FUSED, 16, 2, 1>(p, stream);
    case 1: return launch_fast<T, 32, 16, LPT, PT, CSB, FUSED, 8, 4, 1>(p, stream);
    case 2: return launch_fast<T, 32, 32, LPT, PT, CSB, FUSED, 8, 2, 1>(p, stream);
    case 3: return launch_fast<T, 32, 32, LPT, PT, CSB, FUSED, 8, 2, 2>(p, stream);
    case 4: return launch_fast<T, 32, 32, LPT, PT, CSB, FUSED, 16, 1, 2>(p, stream);
    case 5: return launch_fast<T, 32, 32, LPT, PT, CSB, FUSED, 4, 4, 2>(p, stream);
    case 6: return launch_fast<T, 32, 32, LPT, PT, CSB, FUSED, 16, 1, 1>(p, stream);
    case 7: return launch_fast<T, 32, 16, LPT, PT, CSB, FUSED, 8, 3, 2>(p, stream);
    case 8: return launch_fast<T, 32, 16, LPT, PT, CSB, FUSED, 4, 8, 1>(p, stream);
    case 9: return launch_fast<T, 32, 32, LPT, PT, CSB, FUSED, 8, 3, 1>(p, stream);
  }
}

template <typename T>
int dispatch_fast(const FwdParams& p, cudaStream_t stream) {
  const bool fused = p.loc == nullptr;
  constexpr int EB = (int)sizeof(T);
  const bool csb = (p.M * p.D == 256);   // M*D = 256: the pixel pitch becomes an immediate
  if (fused) {
    return csb ? dispatch_fast_variant<T, 16, 4, 256 * EB, true>(p, stream)
               : dispatch_fast_variant<T, 16, 4, 0, true>(p, stream);
  }
  return csb ? dispatch_fast_variant<T, 16, 4, 256 * EB, false>(p, stream)
             : dispatch_fast_variant<T, 16, 4, 0, false>(p, stream);
}

}  // namespace

bool fast_supported(int D, int L, int P) { return D == 32 && L == 4 && P == 4; }
int launch_forward_fast_f32(const FwdParams& p, cudaStream_t stream) { return dispatch_fast<float>(p, stream); }
int launch_forward_fast_bf16(const FwdParams& p, cudaStream_t stream) { return dispatch_fast<__nv_bfloat16>(p, stream); }

}  // namespace msda
