// Fast forward kernels for the DeepSolo configuration (L*P = 16, P = 4 known at compile time, D = 32).
//
// Same two-phase scheme as msda_forward.cu (phase 1: per-sample records computed once per unit and parked in
// shared memory; phase 2: gather + weighted reduction), re-cut from what the B200 measurements showed
// (tools/ubench/*.cu, profiles/r01_ubench_*.log, profiles/r01_diag_*.log):
//
//   * the gather is bounded by the SM's L1 data path at ONE 128-byte wavefront per clock (LDG.128, LDG.256,
//     LDS.128 and LDSM all top out at 0.97-1.0 rows/clk/SM) -- 66 k clocks per 720p encoder call -- and the
//     previous revision spent as many clocks just ISSUING instructions: 329 warp instructions per unit, 203 us of
//     its 500 us per 8-frame launch with every sample out of range (no gather at all).  This revision is on an
//     instruction diet:
//       - phase 1 does everything that depends only on the sample: the four bilinear weights (rows / columns
//         that fall outside the map get weight 0, exactly the reference's zero padding), the byte offset of the
//         top-left corner clamped into the map, and the distances dx / dy to the right / lower neighbour (0 at
//         the border, so a clamped corner re-reads a valid pixel with weight 0).  A record is two float4:
//         {offset, dy, attention, dx} and {w1, w2, w3, w4};
//       - phase 2 is straight-line: per sample 2 LDS.128, 4 pointer adds, 4 unconditional loads and the
//         reference's FMA chain in packed fp32x2 -- no masks, no predicates, no votes, no branches, so ptxas is
//         free to keep the loads of the next samples in flight across the arithmetic of the current one
//         (an L1 miss costs ~800 clocks under load; the gather needs ~1000 rows in flight per SM);
//       - a lane owns CONTIGUOUS samples (lane k of a unit: samples k*SPL .. k*SPL+SPL-1, i.e. one level), so its
//         phase-1 operands are one 256-bit + one 128-bit streaming load instead of eight scalar ones, and the
//         fused kernel needs one reference point per lane.
//   * packed fp32x2 arithmetic (FFMA2/FMUL2): same IEEE rounding per element as the reference's scalar chain,
//     so outputs stay bit-identical to the reference kernel for finite inputs.  (0 * v stands in for the
//     reference's skipped corner: identical unless v is Inf/NaN at a clamped border pixel.)
//   * the next step's loc/attn (or offsets/logits/reference points) are prefetched into registers before the
//     current step's gather starts, so their DRAM latency hides behind the gather.
#include "msda_fast_common.cuh"
#include "msda_launch.h"
#include "../../include/msda_b200.h"

namespace msda {

//   T     float | __nv_bfloat16 storage of value/out (arithmetic fp32)
//   VB    bytes of a value row one lane loads (32: LDG.E.256, 16: LDG.E.128)
//   LPT   L*P, PT = P (compile time)
//   NW    warps per CTA ; MINB min CTAs per SM (register budget)
//   PD    gather pipeline depth in SOURCE order: the corner loads of sample s+PD-1 are issued before sample s is
//         consumed (ptxas may hoist further: the sample loop is branch-free)
//   (An L1 prefetch of the corner rows 1-3 samples ahead -- prefetch.global.L1 from four lanes per unit -- was
//   measured and removed: 545 vs 502 us per 8-frame encoder launch, profiles/r01_s6_sweep_core.log.  So were a
//   64-register two-record variant (weights precomputed in phase 1: 25 % fewer instructions, one more LDS.128 per
//   sample: 547 vs 500 us) and a 48-register / 40-warp one (spills: 555 us), profiles/r01_s9_sweep_core.log; and a
//   WINDOW prefetch -- one prefetch.global.L1 per row of the tile's level-(l+1) window, issued by the whole CTA a level
//   pass ahead, one 8x16 tile per 32-warp CTA: 682 vs 543 us for the same shape without it,
//   profiles/r01_s18_sweep_window_prefetch.log.  L1 prefetches are not free on this part.)
template <typename T, int D, int VB, int LPT, int PT, bool FUSED, int NW, int MINB, int PD, bool REC16>
__global__ void __launch_bounds__(NW * 32, MINB) msda_fwd_fast_kernel(const FwdParams p) {
  constexpr int EB = (int)sizeof(T);
  constexpr int VEC = VB / EB;          // channels per lane
  constexpr int NP = VEC / 2;           // channel pairs per lane
  constexpr int LPR = D / VEC;          // lanes per value row = lanes per unit
  constexpr int UPW = 32 / LPR;         // units per warp step
  constexpr int SPL = LPT / LPR;        // samples per lane in phase 1 (contiguous: k*SPL ..)
  constexpr int NL = LPT / PT;          // levels
  constexpr int NLV = SPL > PT ? SPL / PT : 1;   // levels one lane's samples span
  static_assert(D % VEC == 0 && LPR >= 2 && LPR <= 8 && 32 % LPR == 0, "bad lane layout");
  static_assert(LPT % LPR == 0 && LPT % PT == 0 && SPL >= 1, "bad sample layout");
  static_assert(SPL % PT == 0 || PT % SPL == 0, "a lane's samples must not straddle levels unevenly");
  // Record slot of (sample s, lane group g): s*UPW + (g ^ swz(s)), swz(s) = (s / SPL) * (8 / LPR): the phase-1
  // STS.128 of a quarter-warp (8 lanes = 8/LPR groups x LPR sample owners) then hits 8 different 16-byte bank
  // groups (needs UPW >= 8; with UPW = 4 the store is 2-way conflicted).  swz(s) takes LPR values.
  auto swz = [](int s) -> int { return ((s / SPL) * (8 / LPR)) & (UPW - 1); };

  __shared__ int sH[NL], sW[NL], sStart[NL], sTileCum[NL + 1];
  __shared__ float sHf[NL], sWf[NL];
  extern __shared__ float4 sRecAll[];   // [NW][2][LPT][UPW]

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
  // first query level served; the whole range if the window kernel ahead of this launch reported a shape mismatch
  const bool all_q = p.shape_flag != nullptr && *reinterpret_cast<const volatile int*>(p.shape_flag) == p.shape_epoch;
  const int lvl_begin = all_q ? 0 : p.q_level_begin;
  if (pyramid && tid == 0) {
    int cum = 0;
    for (int l = 0; l < NL; ++l) {
      sTileCum[l] = cum;
      if (l >= lvl_begin) cum += ((sH[l] + TH - 1) / TH) * ((sW[l] + (1 << tw_log2) - 1) >> tw_log2);
    }
    sTileCum[NL] = cum;
  }
  __syncthreads();

  const int cstride = M * D * EB;                   // bytes between horizontally adjacent pixels
  const int q0 = (!pyramid && lvl_begin > 0) ? sStart[lvl_begin] : 0;          // first query of the range served (linear tiles)
  const int tiles_per_bm = pyramid ? sTileCum[NL] : (Lq - q0 + p.tile_q - 1) / p.tile_q;
  const long long total_tiles = (long long)p.N * tiles_per_bm * M;
  const int chunks_per_warp = (p.tile_q + NW * UPW - 1) / (NW * UPW);   // warp steps per tile
  // REC16: one record {offset | corner mask, lh, lw, attn}, weights rebuilt in phase 2 (half the record wavefronts on
  // the L1 data pipe, ~18 more instructions per sample step); else two: {offset, dy, attn, dx | mask} {w1, w2, w3, w4}
  constexpr int RECS = REC16 ? 1 : 2;
  float4* sRecA = sRecAll + (size_t)warp * RECS * LPT * UPW;
  float4* sRecB = sRecA + (REC16 ? 0 : LPT * UPW);
  int rstride[NL];                                  // bytes between vertically adjacent pixels, per level
#pragma unroll
  for (int l = 0; l < NL; ++l) rstride[l] = sW[l] * cstride;
  const float inv_p = 1.0f / (float)PT;
  const int lvl0 = (k * SPL) / PT;                  // first level of this lane's samples

  // ---- work cursor: (tile, chunk) pairs of this warp, flattened so the next step can be prefetched ----
  // walk 0 (default): tile = blockIdx.x + i * gridDim.x with heads varying fastest: all CTAs work on the same frame
  //   (its value map stays in L2) and the 8 heads of a query block stream the same loc/attn DRAM pages together.
  // walk 1 (diagnostic): a CTA owns a CONTIGUOUS range of tiles, spatial index fastest inside one (batch, head),
  //   pyramid rows boustrophedon.  Measured: no L1 gain (a tile's own working set already exceeds L1, hit rate
  //   81.2 % either way) and 1.65x the DRAM reads at 8 frames (8 value maps in flight > L2): profiles/r01_walk_ncu.txt.
  const bool raster = p.walk == 1;
  const long long tiles_per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
  long long tile = raster ? (long long)blockIdx.x * tiles_per_cta : (long long)blockIdx.x;
  const long long tile_end = raster ? (tile + tiles_per_cta < total_tiles ? tile + tiles_per_cta : total_tiles) : total_tiles;
  const long long tile_step = raster ? 1 : (long long)gridDim.x;
  int chunk = 0;
  int t_m = 0, t_b = 0, t_t = 0, t_lvl = 0, t_ty = 0, t_tx = 0;
  auto decode_tile = [&]() {
    if (raster) {
      t_t = (int)(tile % tiles_per_bm);
      const long long r = tile / tiles_per_bm;
      t_m = (int)(r % M);
      t_b = (int)(r / M);
    } else {
      t_m = (int)(tile % M);
      const long long r = tile / M;
      t_t = (int)(r % tiles_per_bm);
      t_b = (int)(r / tiles_per_bm);
    }
    if (pyramid) {
      t_lvl = 0;
      while (t_lvl + 1 < NL && t_t >= sTileCum[t_lvl + 1]) ++t_lvl;
      const int tt = t_t - sTileCum[t_lvl];
      const int ntx = (sW[t_lvl] + (1 << tw_log2) - 1) >> tw_log2;
      t_ty = tt / ntx;
      t_tx = tt - t_ty * ntx;
      if (raster && (t_ty & 1)) t_tx = ntx - 1 - t_tx;
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
      qi = q0 + t_t * p.tile_q + j;
      valid = (j < p.tile_q) && (qi < Lq);
    }
    bq = (size_t)t_b * Lq + (valid ? qi : 0);
    unit = bq * M + t_m;
  };
  auto prefetch = [&](Prefetched<SPL, NLV, FUSED>& pf, size_t bq, size_t /*unit*/) {
    load_unit_operands<SPL, NLV, FUSED, LPT, NL>(pf, p, bq, t_m, k * SPL, lvl0);
  };

  bool have = tile < tile_end;
  bool n_valid = false;
  size_t n_bq = 0, n_unit = 0;
  const char* n_vbase = nullptr;
  Prefetched<SPL, NLV, FUSED> pf;
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
        // softmax over the unit's LPT logits in PyTorch's operation order (msda_fast_common.cuh)
        const float sum = unit_softmax_terms<SPL, LPR, LPT>(pf.lg, a);
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          const int l = lvl0 + i / PT;
          const float4 rf = pf.ref[i / PT];
          a[i] = __fdiv_rn(a[i], sum);
          lx[i] = location_from_offset(rf.x, rf.z, pf.off[2 * i], sWf[l], inv_p, p.ref_dim);
          ly[i] = location_from_offset(rf.y, rf.w, pf.off[2 * i + 1], sHf[l], inv_p, p.ref_dim);
        }
      } else {
#pragma unroll
        for (int i = 0; i < SPL; ++i) { lx[i] = pf.xy[2 * i]; ly[i] = pf.xy[2 * i + 1]; a[i] = pf.a[i]; }
      }
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        const int s = k * SPL + i;
        const int l = lvl0 + i / PT;
        const int Hl = sH[l], Wl = sW[l];
        const float Hf = sHf[l], Wf = sWf[l];
        // cuh:285-288 (one FFMA each), then cuh:39-45, :56-78 with the zero padding moved into the weights
        const float h_im = __fmaf_rn(ly[i], Hf, -0.5f);
        const float w_im = __fmaf_rn(lx[i], Wf, -0.5f);
        const bool inr = valid && (h_im > -1.0f) && (w_im > -1.0f) && (h_im < Hf) && (w_im < Wf);
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h_low = inr ? (int)hf : 0, w_low = inr ? (int)wf : 0;
        const float lh = __fsub_rn(h_im, hf), lw = __fsub_rn(w_im, wf);
        const float hh = __fsub_rn(1.0f, lh), hw = __fsub_rn(1.0f, lw);
        const bool top = inr && h_low >= 0, bot = inr && h_low + 1 <= Hl - 1;
        const bool lef = inr && w_low >= 0, rig = inr && w_low + 1 <= Wl - 1;
        const float fh0 = top ? hh : 0.0f, fh1 = bot ? lh : 0.0f;     // row factors (upper, lower)
        const float fw0 = lef ? hw : 0.0f, fw1 = rig ? lw : 0.0f;     // column factors (left, right)
        const int hb = top ? h_low : 0, wb = lef ? w_low : 0;         // h_low / w_low = -1 -> the valid neighbour 0
        int off = ((sStart[l] + hb * Wl + wb) * M * D) * EB;
        int dy = (top && bot) ? Wl * cstride : 0;
        int dx = (lef && rig) ? cstride : 0;
        // corner-live bits (cuh:56-78) ride in the low bits of dx (a multiple of 64): a corner that is outside the
        // map, or belongs to a skipped sample (cuh:288), is not loaded at all -- every loaded byte costs L1 write-back
        // bandwidth, the kernel's limiter -- its weight is 0 and its registers keep finite contents.
        dx |= (int)(top && lef) | ((int)(top && rig) << 1) | ((int)(bot && lef) << 2) | ((int)(bot && rig) << 3);
        const int sl = s * UPW + (g ^ swz(s));
        if constexpr (REC16) {
          const int cmask = dx & 15;
          const int off0 = inr ? (((sStart[l] + h_low * Wl + w_low) * M * D) * EB) | cmask : 0;   // unclamped: masked corners are never dereferenced
          sRecA[sl] = make_float4(__int_as_float(off0), lh, lw, inr ? a[i] : 0.0f);
        } else {
          sRecA[sl] = make_float4(__int_as_float(off), __int_as_float(dy), inr ? a[i] : 0.0f, __int_as_float(dx));
          sRecB[sl] = make_float4(__fmul_rn(fh0, fw0), __fmul_rn(fh0, fw1), __fmul_rn(fh1, fw0), __fmul_rn(fh1, fw1));
        }
      }
    }
    __syncwarp();

    // ---------------- advance the cursor and prefetch the next step's operands ----------------
    ++chunk;
    if (chunk == chunks_per_warp) {
      chunk = 0;
      tile += tile_step;
      have = tile < tile_end;
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

    // ---------------- phase 2: gather + weighted reduction (branch-free) ----------------
    f32x2 acc[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) acc[j] = 0ull;
    RowVec<VB> q[PD][4];
    float4 wq[PD];
    float at[PD];
    auto issue = [&](int s, int d) {
      const int sl = s * UPW + (g ^ swz(s));
      const float4 ra = sRecA[sl];
      if constexpr (REC16) {
        const int packed = __float_as_int(ra.x);
        const uint32_t m = (uint32_t)packed & 15u;
        const float lh = ra.y, lw = ra.z;
        const float hh = __fsub_rn(1.0f, lh), hw = __fsub_rn(1.0f, lw);
        const float fh0 = (m & 3u) ? hh : 0.0f, fh1 = (m & 12u) ? lh : 0.0f;
        const float fw0 = (m & 5u) ? hw : 0.0f, fw1 = (m & 10u) ? lw : 0.0f;
        wq[d] = make_float4(__fmul_rn(fh0, fw0), __fmul_rn(fh0, fw1), __fmul_rn(fh1, fw0), __fmul_rn(fh1, fw1));
        at[d] = ra.w;
        const char* c1 = vbase + (ptrdiff_t)(packed & ~15);
        const char* c3 = c1 + rstride[s / PT];
        q[d][0].load_if(c1, m & 1u);
        q[d][1].load_if(c1 + cstride, m & 2u);
        q[d][2].load_if(c3, m & 4u);
        q[d][3].load_if(c3 + cstride, m & 8u);
      } else {
        wq[d] = sRecB[sl];
        at[d] = ra.z;
        const char* c1 = vbase + (uint32_t)__float_as_int(ra.x);
        const char* c3 = c1 + (uint32_t)__float_as_int(ra.y);
        const uint32_t dxm = (uint32_t)__float_as_int(ra.w), dx = dxm & ~15u;
        q[d][0].load_if(c1, dxm & 1u);
        q[d][1].load_if(c1 + dx, dxm & 2u);
        q[d][2].load_if(c3, dxm & 4u);
        q[d][3].load_if(c3 + dx, dxm & 8u);
      }
    };
    // registers of never-loaded corners must hold finite values (their weight is 0): clear them once per step, so
    // a non-finite value can only reach outputs the reference also makes non-finite (same unit, same channels)
#pragma unroll
    for (int d = 0; d < PD; ++d)
#pragma unroll
      for (int c = 0; c < 4; ++c) q[d][c].zero();
#pragma unroll
    for (int s = 0; s < PD - 1; ++s) issue(s, s);
#pragma unroll
    for (int s = 0; s < LPT; ++s) {
      if (s + PD - 1 < LPT) issue(s + PD - 1, (s + PD - 1) % PD);
      const int d = s % PD;
      accumulate_sample<T, VB, NP>(acc, wq[d], at[d], q[d][0], q[d][1], q[d][2], q[d][3]);
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

template <typename T, int D, int VB, int LPT, int PT, bool FUSED, int NW, int MINB, int PD, bool REC16>
int launch_fast(const FwdParams& p, cudaStream_t stream) {
  constexpr int LPR = D / (VB / (int)sizeof(T)), UPW = 32 / LPR;
  const size_t smem = (size_t)NW * (REC16 ? 1 : 2) * LPT * UPW * sizeof(float4);
  auto kern = msda_fwd_fast_kernel<T, D, VB, LPT, PT, FUSED, NW, MINB, PD, REC16>;
  static PerDeviceOnce configured;   // function attributes are per device
  if (configured.need()) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    // the gather lives on L1 hits: shared memory only holds the records.  Ask for just enough carve-out that MINB
    // CTAs fit (ncu: with too small a hint only ONE CTA was resident per SM).
    const int need_kb = (int)((MINB * (smem + 1024 + 256) + 1023) / 1024);
    int pct = (need_kb * 100 + 227) / 228 + 2;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
  }
  kern<<<p.grid, NW * 32, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

template <typename T, bool FUSED>
int dispatch_fast_variant(const FwdParams& p, cudaStream_t stream) {
  constexpr int LPT = 16, PT = 4;
  switch (p.variant) {   //                  D  VB                 NW MINB PD REC16
    default:
    case 0: return launch_fast<T, 32, 32, LPT, PT, FUSED, 8, 2, 1, false>(p, stream);
    case 1: return launch_fast<T, 32, 32, LPT, PT, FUSED, 8, 2, 1, true>(p, stream);
    case 2: return launch_fast<T, 32, 32, LPT, PT, FUSED, 8, 3, 1, true>(p, stream);
    case 3: return launch_fast<T, 32, 16, LPT, PT, FUSED, 8, 4, 1, true>(p, stream);
    case 4: return launch_fast<T, 32, 16, LPT, PT, FUSED, 8, 3, 2, true>(p, stream);
    case 5: return launch_fast<T, 32, 16, LPT, PT, FUSED, 8, 3, 2, false>(p, stream);
    case 6: return launch_fast<T, 32, 16, LPT, PT, FUSED, 16, 2, 1, true>(p, stream);
    case 7: return launch_fast<T, 32, 32, LPT, PT, FUSED, 8, 2, 2, true>(p, stream);
  }
}

template <typename T>
int dispatch_fast(const FwdParams& p, cudaStream_t stream) {
  return p.loc == nullptr ? dispatch_fast_variant<T, true>(p, stream) : dispatch_fast_variant<T, false>(p, stream);
}

inline bool aligned_to(const void* ptr, uintptr_t a) { return (reinterpret_cast<uintptr_t>(ptr) & (a - 1)) == 0; }

}  // namespace

int fast_variant_count() { return 8; }

// The fast kernels read a lane's phase-1 operands with one vector load: the operand rows must be 32-byte aligned.
bool fast_supported(const FwdParams& p) {
  if (!(p.D == 32 && p.L == 4 && p.P == 4)) return false;
  if (!aligned_to(p.value, 32) || !aligned_to(p.out, 32)) return false;
  if (p.loc) return aligned_to(p.loc, 32) && aligned_to(p.attn, 32);
  return aligned_to(p.offsets, 32) && aligned_to(p.logits, 32) && (p.off_pitch % 8) == 0 && (p.logit_pitch % 8) == 0 &&
         aligned_to(p.ref, p.ref_dim == 4 ? 16 : 8);
}
bool fast_shape_supported(int D, int L, int P) { return D == 32 && L == 4 && P == 4; }
int launch_forward_fast_f32(const FwdParams& p, cudaStream_t stream) { return dispatch_fast<float>(p, stream); }
int launch_forward_fast_bf16(const FwdParams& p, cudaStream_t stream) { return dispatch_fast<__nv_bfloat16>(p, stream); }

}  // namespace msda
