// Staged forward kernel for encoder self-attention (Lq == S, D = 32, L = 4, P = 4, fp32 storage).
//
// Why it exists (profiles/r01_s6_*: ncu of the register-gather kernels): with the value rows gathered straight from
// global memory every variant of msda_forward_fast.cu lands on the same plateau (487-545 us per 8-frame encoder
// launch) whatever its L1 hit rate (47 % .. 84 %), vector width or occupancy.  The SM's L1 returns loads IN ORDER, so
// one miss among the 16 rows of a sample step holds the whole step for an L2 round trip (~500 clk); the rows a warp
// can keep in flight are bounded by its registers; the data pipe idles at 62-64 %.  The way out is the one the brief
// names: put the rows a query tile will touch into shared memory first, asynchronously, and gather from there --
// LDS has no misses, ~30 clk latency, and the same 128 B/clk pipe.
//
// Scheme, per work item = (frame b, head m, tile of 8 x 8 queries of pyramid level 0; (8 >> ql) squared for level ql):
//   1. phase 1 (as in the fast kernel): every lane turns the operands of its 4 samples (one level per lane) into
//      h_low / w_low / lh / lw / attention; the CTA averages h_low, w_low per sampled level (shared atomics);
//   2. a window of 20 / 16 / 14 / 13 pixels squared of this head (128-byte rows) centred on that average is copied per
//      level with cp.async.cg (16 B per thread, L2 -> shared memory, no register, no L1 allocation); pixels outside the
//      map are ZERO-FILLED (src-size 0), which is exactly the reference's zero padding (cuh:56-78) -- the gather needs no
//      corner masks.  Two ring slots: A = level 0 then 2, B = level 1 then 3; a slot is refilled while the other's
//      pass runs.  ~98 KB per CTA, two CTAs per SM;
//   3. records {window byte offset, lh, lw, attention} go to shared memory; a sample whose 2x2 footprint is not inside
//      the window keeps the global byte offset + corner mask instead and takes the predicated-LDG path (bit 4 of the
//      offset word tells which) -- results never depend on where the window sits;
//   4. four level passes: per sample one LDS.128 record, eight LDS.128 row halves (4 lanes per unit, immediate offsets
//      for the four corners), the reference's FMUL/FFMA chain in packed fp32x2, level-major = the reference's
//      accumulation order (cuh:272-296) -> bit-identical outputs;
//   5. outputs stored with streaming stores.  The next item's operands are prefetched into registers before the passes.
// Only query levels whose tiles are worth a window run here (FwdParams::staged_levels); the remaining queries are
// served by the register-gather kernel in the same stream (msda_forward.cu, launch_forward).
//
// Window fills: with the level shapes known on the host (msda_b200_staged_set_host_shapes; the Python wrapper reads
// them once per shapes tensor) each window is ONE 5-D TMA tile load (value viewed as (N, H_l, W_l, M, D), box
// {D, 1 head, WW, WH, 1 frame}, out-of-map pixels zero-filled by the TMA unit) completing on the slot's mbarrier --
// off the LSU pipe and off the threads; otherwise cp.async as described above.
//
// Measured (profiles/r01_s16_*, r01_s17_*, r01_s21_staged_tma_breakdown.log, r01_s22_*): bit-exact; the gather passes run
// AT the pipe's limit (~1 wavefront/clk/SM; 201 us per 8-frame launch for the 59 M rows of the level-0 queries), the
// TMA fills hide almost completely (+34 us; cp.async fills cost +111 us and 14.7 M wavefronts on the same pipe), but
// phase 1 + barriers + records (133 us) and the divergent fallback branch (33 us) do not: 390 us for the level-0
// queries vs ~370 us in the fast kernel, 515 vs 494 us per launch (571 vs 543 fused).  Opt-in (tuning.mode = 4).
// (A head start of 1-8 us for one of the two co-resident CTAs -- in case they ran in lockstep -- changes nothing.)
#include <cuda.h>
#include <type_traits>
#include "msda_fast_common.cuh"
#include "tma_common.cuh"
#include "msda_launch.h"
#include <string.h>
#include "../../include/msda_b200.h"

namespace msda {

namespace {

constexpr int kSgWarps = 8, kSgThreads = kSgWarps * 32, kSgCtasPerSm = 2;
constexpr int kSgLPR = 4, kSgUPW = 8, kSgUnits = kSgWarps * kSgUPW;   // 64 units per tile
constexpr int kSgTHlog2 = 3, kSgTWlog2 = 3;                            // level-0 tile: 8 x 8 queries
constexpr int kSgL = 4, kSgP = 4, kSgLPT = 16, kSgD = 32;
constexpr int kRowB = kSgD * 4;                                        // 128-byte value rows

// window extents per sampled level: tile extent at that level + 12 pixels (grid bias up to 4 px + noise, both sides).
// Two ring slots: A holds level 0, then level 2; B holds level 1, then level 3 -- a window is refilled as soon as its
// pass is over, while the other slot's pass runs.  ~98 KB per CTA, two CTAs per SM: one CTA's phase 1 / fill latency /
// barriers overlap the other's gather (measured one CTA per SM: overhead + fill + gather simply add up,
// profiles/r01_s16_staged_breakdown.log).
__host__ __device__ constexpr int sg_wh(int l) { return l == 0 ? 20 : l == 1 ? 16 : l == 2 ? 14 : 13; }
__host__ __device__ constexpr int sg_ww(int l) { return sg_wh(l); }
__host__ __device__ constexpr int sg_rows(int l) { return sg_wh(l) * sg_ww(l); }
constexpr int kSgSlotA = (sg_rows(0) > sg_rows(2) ? sg_rows(0) : sg_rows(2)) * kRowB;   // 51200 B
constexpr int kSgSlotB = (sg_rows(1) > sg_rows(3) ? sg_rows(1) : sg_rows(3)) * kRowB;   // 32768 B
__host__ __device__ constexpr int sg_woff(int l) { return (l & 1) ? kSgSlotA : 0; }
constexpr int kSgRecBytes = kSgWarps * kSgLPT * kSgUPW * 16;           // 16 KB
constexpr int kSgWinBytes = kSgSlotA + kSgSlotB;
constexpr int kSgSmem = kSgRecBytes + kSgWinBytes;                     // 100352 B

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// ---- TMA window fill (used when the level shapes are known on the host: msda_b200_staged_set_host_shapes) ----
struct StagedMaps { CUtensorMap lv[kSgL]; };
// value viewed as (N, H_l, W_l, M, D): box = {D, 1 head, WW, WH, 1 frame}; out-of-map pixels arrive as zeros
__device__ __forceinline__ void tma_window(uint32_t dst, const CUtensorMap* map, uint32_t bar, int m, int w0, int h0, int b) {
  tma_load_5d(dst, map, bar, 0, m, w0, h0, b);
}
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int OFF> __device__ __forceinline__ uint4 lds128_at(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a), "n"(OFF));
  return r;
}

// DBG != 0: diagnostic builds that drop the fallback path / the window fill / the gather passes (wrong results) to
// attribute the time of a work item (profiles/r01_s16_staged_breakdown.log).
template <bool FUSED, int DBG, bool TMA>
__global__ void __launch_bounds__(kSgThreads, kSgCtasPerSm) msda_fwd_staged_kernel(const FwdParams p, const __grid_constant__ StagedMaps maps) {
  constexpr bool NOFB = (DBG & 1) != 0, NOSTAGE = (DBG & 2) != 0, NOGATHER = (DBG & 4) != 0;   // diagnostics only
  constexpr int NL = kSgL, PT = kSgP, LPT = kSgLPT, SPL = 4, D = kSgD;
  extern __shared__ __align__(128) unsigned char sg_smem[];
  __shared__ int sH[NL], sW[NL], sStart[NL], sTileCum[NL + 1];
  __shared__ float sHf[NL], sWf[NL];
  __shared__ int sAcc[2][NL][4];
  __shared__ __align__(8) unsigned long long sBar[2];   // TMA: one mbarrier per ring slot, two fills per item each          // per sampled level: sum h_low, sum w_low, count (double-buffered by item parity)

  const int M = p.M, Lq = p.Lq;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, k = lane & 3;   // unit slot in the warp, lane in the unit = sampled level of its 4 samples
  if (tid < NL) {
    sH[tid] = (int)p.shapes[2 * tid];
    sW[tid] = (int)p.shapes[2 * tid + 1];
    sStart[tid] = (int)p.lsi[tid];
    sHf[tid] = (float)sH[tid];
    sWf[tid] = (float)sW[tid];
  }
  if (tid < 2 * NL * 4) (&sAcc[0][0][0])[tid] = 0;
  if (TMA && tid == 0) {
    mbar_init((uint32_t)__cvta_generic_to_shared(&sBar[0]), 1);
    mbar_init((uint32_t)__cvta_generic_to_shared(&sBar[1]), 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int QL = p.staged_levels;          // query levels served here (tiles of level ql: (8 >> ql) x (16 >> ql))
  if (tid == 0) {
    int cum = 0;
    for (int l = 0; l < NL; ++l) {
      sTileCum[l] = cum;
      if (l < QL) {
        const int th = 1 << (kSgTHlog2 - l), tw = 1 << (kSgTWlog2 - l);
        cum += ((sH[l] + th - 1) / th) * ((sW[l] + tw - 1) / tw);
      }
    }
    sTileCum[NL] = cum;
  }
  __syncthreads();

  float4* sRec = reinterpret_cast<float4*>(sg_smem) + (size_t)warp * LPT * kSgUPW;
  const uint32_t sWinBase = (uint32_t)__cvta_generic_to_shared(sg_smem + kSgRecBytes);
  const int cstride = M * kRowB;                      // bytes between horizontally adjacent pixels
  const int tiles_per_bm = sTileCum[NL];
  const int total = p.N * tiles_per_bm * M;           // checked on the host to fit in int
  const float inv_p = 1.0f / (float)PT;
  const int c0 = k * 16 + (g & 1) * 64;               // this lane's first 16-byte chunk of a row (second: c0 ^ 64)
  const int dhi = (g & 1) ? -64 : 64;                 // ... as a signed distance: no alignment assumption on the rows

  // ---- work item decode + operand prefetch ----
  struct Item { int b, m, lvl, ty, tx, units; };
  auto decode = [&](int item) {
    Item it;
    it.m = item % M;
    const int r = item / M;
    const int t = r % tiles_per_bm;
    it.b = r / tiles_per_bm;
    it.lvl = 0;
    while (it.lvl + 1 < NL && t >= sTileCum[it.lvl + 1]) ++it.lvl;
    const int tt = t - sTileCum[it.lvl];
    const int tw_log2 = kSgTWlog2 - it.lvl;
    const int ntx = (sW[it.lvl] + (1 << tw_log2) - 1) >> tw_log2;
    it.ty = tt / ntx;
    it.tx = tt - it.ty * ntx;
    it.units = kSgUnits >> (2 * it.lvl);
    return it;
  };
  auto locate = [&](const Item& it, bool& valid, size_t& bq) {
    const int j = warp * kSgUPW + g;
    const int tw_log2 = kSgTWlog2 - it.lvl, th_log2 = kSgTHlog2 - it.lvl;
    const int y = (it.ty << th_log2) + (j >> tw_log2), x = (it.tx << tw_log2) + (j & ((1 << tw_log2) - 1));
    valid = (j < it.units) && (y < sH[it.lvl]) && (x < sW[it.lvl]);
    const int qi = sStart[it.lvl] + y * sW[it.lvl] + x;
    bq = (size_t)it.b * Lq + (valid ? qi : 0);
  };
  auto prefetch = [&](Prefetched<SPL, 1, FUSED>& pf, size_t bq, int m) {
    load_unit_operands<SPL, 1, FUSED, LPT, NL>(pf, p, bq, m, k * SPL, k);
  };

  int item = blockIdx.x;
  int parity = 0;
  Item cur{}, nxt{};
  bool n_valid = false;
  size_t n_bq = 0;
  Prefetched<SPL, 1, FUSED> pf;
  if (item < total) {
    nxt = decode(item);
    locate(nxt, n_valid, n_bq);
    prefetch(pf, n_bq, nxt.m);
  }

  while (item < total) {
    cur = nxt;
    const bool valid = n_valid;
    const size_t unit = n_bq * M + cur.m;
    const bool warp_active = warp * kSgUPW < cur.units;
    const char* vhead = reinterpret_cast<const char*>(p.value) + ((size_t)cur.b * p.S * M + cur.m) * kRowB;

    // ---------------- phase 1a: sample geometry of this lane's 4 samples (all of sampled level k) ----------------
    float a[SPL], lhv[SPL], lwv[SPL];
    int hl[SPL], wl[SPL];
    int inr_bits = 0;
    const int Hl = sH[k], Wl = sW[k];
    const float Hf = sHf[k], Wf = sWf[k];
    {
      float lx[SPL], ly[SPL];
      if constexpr (FUSED) {
        // softmax over the unit's 16 logits in the operation order of PyTorch's persistent warp softmax (see
        // msda_forward_fast.cu): element e = 4k + i lives in register i of lane k
        const float sum = unit_softmax_terms<SPL, kSgLPR, LPT>(pf.lg, a);
        const float4 rf = pf.ref[0];
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          a[i] = __fdiv_rn(a[i], sum);
          lx[i] = location_from_offset(rf.x, rf.z, pf.off[2 * i], Wf, inv_p, p.ref_dim);
          ly[i] = location_from_offset(rf.y, rf.w, pf.off[2 * i + 1], Hf, inv_p, p.ref_dim);
        }
      } else {
#pragma unroll
        for (int i = 0; i < SPL; ++i) { lx[i] = pf.xy[2 * i]; ly[i] = pf.xy[2 * i + 1]; a[i] = pf.a[i]; }
      }
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        // cuh:285-288 (one FFMA each), cuh:39-45
        const float h_im = __fmaf_rn(ly[i], Hf, -0.5f);
        const float w_im = __fmaf_rn(lx[i], Wf, -0.5f);
        const bool inr = valid && (h_im > -1.0f) && (w_im > -1.0f) && (h_im < Hf) && (w_im < Wf);
        const float hf = floorf(h_im), wf = floorf(w_im);
        hl[i] = inr ? (int)hf : 0;
        wl[i] = inr ? (int)wf : 0;
        lhv[i] = __fsub_rn(h_im, hf);
        lwv[i] = __fsub_rn(w_im, wf);
        if (!inr) a[i] = 0.0f;
        inr_bits |= (int)inr << i;
      }
    }
    // ---------------- window centre per sampled level: CTA-wide mean of h_low / w_low ----------------
    {
      int sh = 0, sw = 0, cnt = 0;
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        if ((inr_bits >> i) & 1) { sh += hl[i]; sw += wl[i]; ++cnt; }
      }
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        sh += __shfl_xor_sync(0xffffffffu, sh, o);
        sw += __shfl_xor_sync(0xffffffffu, sw, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      }
      if (g == 0 && cnt > 0) {
        atomicAdd(&sAcc[parity][k][0], sh);
        atomicAdd(&sAcc[parity][k][1], sw);
        atomicAdd(&sAcc[parity][k][2], cnt);
      }
    }
    __syncthreads();   // S1: sums complete; previous item's windows and records are free (S6 of the previous item)
    int h0[NL], w0[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int cnt = sAcc[parity][l][2];
      int ch = -1, cw = -1;
      if (cnt > 0) {
        const float inv = 1.0f / (float)cnt;
        ch = __float2int_rn((float)sAcc[parity][l][0] * inv) - (sg_wh(l) - 2) / 2;
        cw = __float2int_rn((float)sAcc[parity][l][1] * inv) - (sg_ww(l) - 2) / 2;
      }
      // keep the window on the map (one zero row / column of padding on each side is all a sample can touch)
      ch = min(ch, sH[l] + 1 - sg_wh(l));
      cw = min(cw, sW[l] + 1 - sg_ww(l));
      h0[l] = max(ch, -1);
      w0[l] = max(cw, -1);
    }
    if (tid < NL * 4) (&sAcc[parity ^ 1][0][0])[tid] = 0;   // next item's sums (its atomics come after this item's S2)

    // ---------------- window fill: cp.async.cg 16 B per thread, zero-fill outside the map ----------------
    auto stage = [&](auto level_c) {
      constexpr int l = decltype(level_c)::value;
      if constexpr (TMA) {
        if (tid == 0 && !NOSTAGE) {
          const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&sBar[l & 1]);
          mbar_expect_tx(bar, sg_rows(l) * kRowB);
          tma_window(sWinBase + sg_woff(l), &maps.lv[l], bar, cur.m, w0[l], h0[l], cur.b);
        }
        return;
      }
      const int c = tid & 7;
      const int Hs = sH[l], Ws = sW[l], st = sStart[l];
      const uint32_t dst0 = sWinBase + sg_woff(l) + c * 16;
      for (int row = tid >> 3; row < (NOSTAGE ? 0 : sg_rows(l)); row += kSgThreads / 8) {
        const int i = row / sg_ww(l), j = row - i * sg_ww(l);
        const int h = h0[l] + i, w = w0[l] + j;
        const bool ok = ((unsigned)h < (unsigned)Hs) && ((unsigned)w < (unsigned)Ws);
        const char* src = vhead + (ok ? (size_t)(unsigned)((st + h * Ws + w) * cstride) + c * 16 : 0);
        cp_async16(dst0 + row * kRowB, src, ok ? 16 : 0);
      }
      cp_async_commit();
    };
    stage(std::integral_constant<int, 0>{});   // slot A
    stage(std::integral_constant<int, 1>{});   // slot B

    // ---------------- records ----------------
    {
      const int hk = k == 0 ? h0[0] : k == 1 ? h0[1] : k == 2 ? h0[2] : h0[3];
      const int wk = k == 0 ? w0[0] : k == 1 ? w0[1] : k == 2 ? w0[2] : w0[3];
      const int whk = k == 0 ? sg_wh(0) : k == 1 ? sg_wh(1) : k == 2 ? sg_wh(2) : sg_wh(3);
      const int wwk = k == 0 ? sg_ww(0) : k == 1 ? sg_ww(1) : k == 2 ? sg_ww(2) : sg_ww(3);
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        const int s = k * SPL + i;
        const bool inr = (inr_bits >> i) & 1;
        const int dh = hl[i] - hk, dw = wl[i] - wk;
        const bool inwin = ((unsigned)dh <= (unsigned)(whk - 2)) && ((unsigned)dw <= (unsigned)(wwk - 2));
        int x = 0;
        if (inr) {
          if (inwin) {
            x = (dh * wwk + dw) * kRowB;
          } else if (!NOFB) {
            const bool top = hl[i] >= 0, bot = hl[i] + 1 <= Hl - 1, lef = wl[i] >= 0, rig = wl[i] + 1 <= Wl - 1;
            const int cmask = (int)(top && lef) | ((int)(top && rig) << 1) | ((int)(bot && lef) << 2) | ((int)(bot && rig) << 3);
            x = ((sStart[k] + hl[i] * Wl + wl[i]) * cstride) | 16 | cmask;   // masked corners are never dereferenced
          }
        }
        // slot swizzle: the STS.128 of a quarter-warp (2 units x 4 sample owners) covers 8 different 16-byte bank groups
        sRec[s * kSgUPW + (g ^ (2 * k))] = make_float4(__int_as_float(x), lhv[i], lwv[i], a[i]);
      }
    }

    // ---------------- next item: decode + operand prefetch (latency hides behind the passes) ----------------
    const int next_item = item + gridDim.x;
    if (next_item < total) {
      nxt = decode(next_item);
      locate(nxt, n_valid, n_bq);
      prefetch(pf, n_bq, nxt.m);
    }

    // ---------------- four level passes ----------------
    f32x2 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = 0ull;
    auto level_pass = [&](auto level_c) {
      constexpr int l = decltype(level_c)::value;
      if (warp_active && !NOGATHER) {
        const uint32_t wbase = sWinBase + sg_woff(l) + c0;
        const int rs = sW[l] * cstride;             // fallback: bytes between vertically adjacent pixels
#pragma unroll
        for (int pt = 0; pt < PT; ++pt) {
          const int s = l * PT + pt;
          const float4 r = sRec[s * kSgUPW + (g ^ (2 * l))];
          const int x = __float_as_int(r.x);
          RowVec<32> q[4];
          if (NOFB || !(x & 16)) {
            const uint32_t a0 = wbase + (uint32_t)x, a1 = a0 + (uint32_t)dhi;
            q[0].lo = lds128_at<0>(a0);                             q[0].hi = lds128_at<0>(a1);
            q[1].lo = lds128_at<kRowB>(a0);                         q[1].hi = lds128_at<kRowB>(a1);
            q[2].lo = lds128_at<sg_ww(l) * kRowB>(a0);              q[2].hi = lds128_at<sg_ww(l) * kRowB>(a1);
            q[3].lo = lds128_at<sg_ww(l) * kRowB + kRowB>(a0);      q[3].hi = lds128_at<sg_ww(l) * kRowB + kRowB>(a1);
          } else {
            // footprint outside the window: predicated global loads, skipped corners read as zero (cuh:56-78)
            const char* c1 = vhead + (ptrdiff_t)(x & ~31) + c0;
            const char* c3 = c1 + rs;
#pragma unroll
            for (int c = 0; c < 4; ++c) q[c].zero();
            if (x & 1) { q[0].lo = ld_value16(c1);               q[0].hi = ld_value16(c1 + dhi); }
            if (x & 2) { q[1].lo = ld_value16(c1 + cstride);     q[1].hi = ld_value16(c1 + cstride + dhi); }
            if (x & 4) { q[2].lo = ld_value16(c3);               q[2].hi = ld_value16(c3 + dhi); }
            if (x & 8) { q[3].lo = ld_value16(c3 + cstride);     q[3].hi = ld_value16(c3 + cstride + dhi); }
          }
          float w1, w2, w3, w4;
          bilinear_weights(r.y, r.z, w1, w2, w3, w4);
          accumulate_sample<float, 32, 4>(acc, make_float4(w1, w2, w3, w4), r.w, q[0], q[1], q[2], q[3]);
        }
      }
    };
    // landed(slot, n): the n-th fill (0, 1) of this item into ring slot A (0) / B (1) is complete and visible
    auto landed = [&](int slot, uint32_t nth) {
      if constexpr (TMA) {
        if (!NOSTAGE) mbar_wait((uint32_t)__cvta_generic_to_shared(&sBar[slot]), nth);
      }
    };
    if constexpr (TMA) {
      __syncwarp();      // records were written by this warp
      landed(0, 0);
      level_pass(std::integral_constant<int, 0>{});
      __syncthreads();   // S3: pass 0 is over everywhere (slot A free)
      stage(std::integral_constant<int, 2>{});
      landed(1, 0);
      level_pass(std::integral_constant<int, 1>{});
      __syncthreads();   // S4: slot B free
      stage(std::integral_constant<int, 3>{});
      landed(0, 1);
      level_pass(std::integral_constant<int, 2>{});
      landed(1, 1);
      level_pass(std::integral_constant<int, 3>{});
    } else {
      cp_async_wait<1>();
      __syncthreads();   // S2: level 0 has landed (every thread's share); records visible
      level_pass(std::integral_constant<int, 0>{});
      cp_async_wait<0>();
      __syncthreads();   // S3: pass 0 is over everywhere (slot A free) and level 1 has landed
      stage(std::integral_constant<int, 2>{});   // slot A, fills while pass 1 runs
      level_pass(std::integral_constant<int, 1>{});
      cp_async_wait<0>();
      __syncthreads();   // S4: slot B free, level 2 landed
      stage(std::integral_constant<int, 3>{});   // slot B, fills while pass 2 runs
      level_pass(std::integral_constant<int, 2>{});
      cp_async_wait<0>();
      __syncthreads();   // S5: level 3 landed
      level_pass(std::integral_constant<int, 3>{});
    }
    if (valid) {
      char* op = reinterpret_cast<char*>(p.out) + unit * (size_t)(D * 4);
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) upk(acc[j], f[2 * j], f[2 * j + 1]);
      st_stream16(op + c0, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
      st_stream16(op + c0 + dhi, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
    }
    __syncthreads();   // S6: slots and records may be overwritten
    item = next_item;
    parity ^= 1;
  }
}

// ---- host side ------------------------------------------------------------------------------------------
struct HostShapes { bool valid = false; long long hw[kSgL][2]; long long lsi[kSgL]; };
// per calling thread: a thread's set_host_shapes + launch sequence cannot be disturbed by another thread's pyramid
thread_local HostShapes g_host_shapes;

// One 5-D map per sampled level over value (N, S, M, D) fp32: (D, M, W_l, H_l, N), base at the level's first pixel.
bool make_window_maps(const FwdParams& p, StagedMaps& maps) {
  const HostShapes hs = g_host_shapes;
  EncodeTiledFn fn = tensor_map_encode_fn();
  if (!hs.valid || !fn) return false;
  long long total = 0;
  for (int l = 0; l < kSgL; ++l) total += hs.hw[l][0] * hs.hw[l][1];
  if (total != p.S) return false;                          // the hint belongs to another pyramid
  const cuuint64_t row = (cuuint64_t)p.D * 4, px = (cuuint64_t)p.M * row;
  for (int l = 0; l < kSgL; ++l) {
    const cuuint64_t H = (cuuint64_t)hs.hw[l][0], W = (cuuint64_t)hs.hw[l][1];
    cuuint64_t gdim[5] = {(cuuint64_t)p.D, (cuuint64_t)p.M, W, H, (cuuint64_t)p.N};
    cuuint64_t gstride[4] = {row, px, W * px, (cuuint64_t)p.S * px};
    cuuint32_t box[5] = {(cuuint32_t)p.D, 1, (cuuint32_t)sg_ww(l), (cuuint32_t)sg_wh(l), 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void* base = const_cast<char*>(reinterpret_cast<const char*>(p.value) + (size_t)hs.lsi[l] * px);
    if (fn(&maps.lv[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

template <bool FUSED, int DBG, bool TMA>
int launch_staged_impl(const FwdParams& p, const StagedMaps& maps, cudaStream_t stream) {
  auto kern = msda_fwd_staged_kernel<FUSED, DBG, TMA>;
  static PerDeviceOnce configured;   // function attributes are per device
  if (configured.need()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSgSmem);
    if (e != cudaSuccess) return (int)e;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);   // two ~98 KB CTAs per SM
  }
  kern<<<p.grid, kSgThreads, kSgSmem, stream>>>(p, maps);
  return (int)cudaGetLastError();
}

template <bool FUSED, int DBG>
int launch_staged(const FwdParams& p, cudaStream_t stream) {
  StagedMaps maps;
  if (p.walk != 1 && make_window_maps(p, maps)) return launch_staged_impl<FUSED, DBG, true>(p, maps, stream);   // walk = 1: force cp.async fills
  memset(&maps, 0, sizeof(maps));
  return launch_staged_impl<FUSED, DBG, false>(p, maps, stream);
}

}  // namespace

bool staged_supported(const FwdParams& p) {
  return p.Lq == p.S && p.D == 32 && p.L == 4 && p.P == 4 && fast_supported(p) &&
         (long long)p.N * p.M * ((long long)p.S / 4 + 64) < (1ll << 31);
}

void staged_set_host_shapes(const int64_t* shapes_host, const int64_t* lsi_host, int L) {
  g_host_shapes.valid = false;
  if (!shapes_host || !lsi_host || L != kSgL) return;
  for (int l = 0; l < kSgL; ++l) {
    g_host_shapes.hw[l][0] = shapes_host[2 * l];
    g_host_shapes.hw[l][1] = shapes_host[2 * l + 1];
    g_host_shapes.lsi[l] = lsi_host[l];
  }
  g_host_shapes.valid = true;
}

bool staged_get_host_shapes(long long (*hw)[2], long long* lsi) {
  if (!g_host_shapes.valid) return false;
  for (int l = 0; l < kSgL; ++l) {
    hw[l][0] = g_host_shapes.hw[l][0];
    hw[l][1] = g_host_shapes.hw[l][1];
    lsi[l] = g_host_shapes.lsi[l];
  }
  return true;
}

int launch_forward_staged_f32(const FwdParams& p, cudaStream_t stream) {
#ifdef MSDA_DIAG
  // Time-attribution builds that produce WRONG results (tools/sweep.py --staged-diag).  Compiled only with -DMSDA_DIAG
  // (python -m gomatching_b200.build --diag): in the product library `variant` never changes results in any mode.
  if (p.loc != nullptr) {
    if (p.variant == 1) return launch_staged<false, 1>(p, stream);    // no fallback path
    if (p.variant == 2) return launch_staged<false, 3>(p, stream);    // ... and no window fill
    if (p.variant == 3) return launch_staged<false, 5>(p, stream);    // ... no gather passes (fill only)
    if (p.variant == 4) return launch_staged<false, 7>(p, stream);    // ... neither: phase 1, records, barriers, stores
  }
#endif
  return p.loc == nullptr ? launch_staged<true, 0>(p, stream) : launch_staged<false, 0>(p, stream);
}

}  // namespace msda
