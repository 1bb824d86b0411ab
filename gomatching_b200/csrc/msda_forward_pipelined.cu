// Pipelined shared-memory-window kernel for fp32 encoder self-attention (Lq == S, D = 32, L = 4, P = 4): the staged
// kernel of msda_forward_staged.cu re-cut as a producer / consumer pipeline without any CTA-wide barrier.
//
// What the staged kernel measured (profiles/r01_s21_*, r01_s22_*): its gather passes run at the L1 pipe's limit and TMA
// window fills hide, but the per-item front end (phase 1, window placement by a CTA-wide mean, six __syncthreads) does
// not overlap with anything: 133 us of 390.  Here
//   * one CTA per SM: 16 consumer (gather) warps + 7 front-end (phase 1) warps + 1 producer warp.  A work item is (frame, head, 8 x 16 tile of level-0 queries);
//     a CTA owns a CONTIGUOUS range of items (same head, raster-adjacent tiles);
//   * four window slots in shared memory, one per sampled level (20x28, 16x20, 14x16, 13x14 pixels of 128-byte rows),
//     each with a `full` mbarrier (TMA transaction bytes) and an `empty` mbarrier (one arrival per consumer warp).  The
//     producer refills slot l for item n+1 as soon as every consumer warp has finished pass l of item n -- three passes
//     ahead of its next use;
//   * the front-end warps run phase 1 (operand loads, softmax / offset->location in the fused entry, sample geometry) ONE
//     ITEM AHEAD of the gather and publish records {h_low|w_low, lh, lw, attention} in double-buffered shared-memory
//     strips, one strip per consumer warp, each with its own full / empty mbarrier pair;
//   * a consumer warp owns 8 units of the tile: wait for its strip, then for each level: wait `full`, turn h_low / w_low
//     into window offsets, gather with LDS.128, arrive on `empty`.  Samples outside the window take predicated global
//     loads (same arithmetic);
//   * window placement: the tile's geometric image in the sampled level plus the offset the PREVIOUS item of the same
//     head measured between its samples' mean and its own geometric centre (the heads' directional bias), left in shared
//     memory by the front end and read by the producer without synchronisation (a stale value only moves a window).
// Results are bit-identical to every other kernel: level-major accumulation order, the reference's FMUL/FFMA chain in
// packed fp32x2, zero padding by the TMA unit's out-of-bounds fill, and window placement only decides which of two
// equivalent load paths a sample takes.  Needs the level geometry on the host (msda_b200_staged_set_host_shapes).
#include <type_traits>
#include <string.h>
#include "msda_fast_common.cuh"
#include "tma_common.cuh"
#include "msda_launch.h"
#include "../../include/msda_b200.h"

namespace msda {

namespace {

constexpr int kPlCons = 16;                              // consumer (gather) warps
// front-end (phase 1) warps.  The register file is split per scheduler: 16 + 7 + 1 = 24 warps put 6 on each and leave every
// thread 80 registers; 8 front-end warps (25 warps, 7 on one scheduler) leave 72 and the gather spills (543 us fused),
// 4 are too few for the fused entry's phase 1 (softmax, twelve IEEE divisions: ~490 instructions per strip; 706 us),
// 3 too few even for the core entry (482 vs 471 us).  profiles/r01_s3*_pipelined*.log
constexpr int kPlFrontMax = 8;
__host__ __device__ constexpr int pl_front(bool) { return 7; }
__host__ __device__ constexpr int pl_threads(bool fused) { return (kPlCons + pl_front(fused) + 1) * 32; }   // + 1 producer warp
constexpr int kPlUPW = 8;                                // units per consumer warp: 128 units per tile
constexpr int kPlTH = 8, kPlTWlog2 = 4;                  // 8 x 16 level-0 queries
constexpr int kPlL = 4, kPlP = 4, kPlLPT = 16;
constexpr int kPlRowB = 128;
constexpr int kPlMaxList = 160;                          // longest per-CTA item list kept in shared memory (frame-group walk)

__host__ __device__ constexpr int pl_wh(int l) { return l == 0 ? 20 : l == 1 ? 16 : l == 2 ? 14 : 13; }
__host__ __device__ constexpr int pl_ww(int l) { return l == 0 ? 28 : l == 1 ? 20 : l == 2 ? 16 : 14; }
__host__ __device__ constexpr int pl_rows(int l) { return pl_wh(l) * pl_ww(l); }
__host__ __device__ constexpr int pl_woff(int l) { return l == 0 ? 0 : pl_woff(l - 1) + pl_rows(l - 1) * kPlRowB; }
constexpr int kPlRecStrip = kPlLPT * kPlUPW;                           // float4 records of one consumer warp's 8 units
constexpr int kPlRecBuf = kPlCons * kPlRecStrip * 16;                  // 32 KB: the records of one item
constexpr int kPlRecBytes = 2 * kPlRecBuf;                             // double-buffered: the front end runs one item ahead
constexpr int kPlWinBytes = pl_woff(3) + pl_rows(3) * kPlRowB;          // 164608 B
constexpr int kPlSmem = kPlRecBytes + kPlWinBytes;

struct PipeGeom {
  CUtensorMap lv[kPlL];
  int H[kPlL], W[kPlL], start[kPlL];
  int group;                      // frames per group of the item walk (>= 1; >= N: one group)
};

template <int OFF> __device__ __forceinline__ uint4 pl_lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a), "n"(OFF));
  return r;
}

template <bool FUSED>
__global__ void __launch_bounds__(pl_threads(FUSED), 1) msda_fwd_pipelined_kernel(const FwdParams p, const __grid_constant__ PipeGeom geo) {
  constexpr int NL = kPlL, PT = kPlP, LPT = kPlLPT, SPL = 4, D = 32;
  constexpr int kPlFront = pl_front(FUSED);
  extern __shared__ __align__(128) unsigned char pl_smem[];
  __shared__ __align__(8) unsigned long long sFull[NL], sEmpty[NL];                        // window slots
  __shared__ __align__(8) unsigned long long sRecFull[2][kPlCons], sRecEmpty[2][kPlCons];   // record strips
  // per item parity, per front-end warp, per sampled level: sum h_low, sum w_low, count of its in-range samples.
  // Plain stores read by the producer one item later; a stale or torn value only moves a window (never a result).
  __shared__ int sPart[2][kPlFrontMax][NL][3];
  __shared__ int sOrg[2][NL][2];      // per item parity, per sampled level: window origin (h0, w0) chosen by the producer

  __shared__ int sMismatch;
  __shared__ int sItems[kPlMaxList];
  const int M = p.M, Lq = p.Lq;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 2 * kPlFrontMax * NL * 3) (&sPart[0][0][0][0])[tid] = 0;
  if (tid == 0) {
    // shape guard: the tensor maps and tile geometry come from a HOST copy of the level shapes; the operator's contract
    // is the DEVICE tensors (ms_deform_attn_cuda.cu:20-80 reads them in-kernel).  If they disagree this kernel does
    // nothing and tells the register-gather launch that follows to serve every query.
    bool bad = false;
    for (int l = 0; l < NL; ++l)
      bad = bad || p.shapes[2 * l] != (int64_t)geo.H[l] || p.shapes[2 * l + 1] != (int64_t)geo.W[l] || p.lsi[l] != (int64_t)geo.start[l];
    sMismatch = bad ? 1 : 0;
    if (bad && p.shape_flag) {
      *p.shape_flag = p.shape_epoch;
      if (p.shape_report) *reinterpret_cast<volatile int*>(p.shape_report) = p.shape_epoch;
    }
  }
  if (tid == 0) {
    for (int l = 0; l < NL; ++l) {
      mbar_init(smem_u32(&sFull[l]), 1);
      mbar_init(smem_u32(&sEmpty[l]), kPlCons);
    }
    for (int i = 0; i < 2 * kPlCons; ++i) {
      mbar_init(smem_u32(&sRecFull[0][0] + i), 1);
      mbar_init(smem_u32(&sRecEmpty[0][0] + i), 1);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (sMismatch) return;

  const uint32_t sWinBase = smem_u32(pl_smem + kPlRecBytes);
  const int cstride = M * kPlRowB;
  const int ntx = (geo.W[0] + (1 << kPlTWlog2) - 1) >> kPlTWlog2, nty = (geo.H[0] + kPlTH - 1) / kPlTH;
  const int tiles = ntx * nty;
  const int total = p.N * M * tiles;                       // item = (b * M + m) * tiles + tile
  // One group (geo.group >= N): a CTA owns ONE contiguous run of items (same head, raster-adjacent tiles) of the whole
  // launch.  With 8 frames in flight (157 MB of value maps > L2) that reads 657 MB from DRAM per launch, 3x the maps
  // (profiles/r02_bench_ncu_full_summary.csv).  Frame groups: the frames are walked geo.group at a time -- inside a
  // group every CTA owns a contiguous run, and all CTAs reach the next group together -- so only one group's maps are
  // in flight (381 MB read).  The per-CTA item sequence is tabulated in shared memory once (sItems) so that the three
  // roles' loops stay `for n < count`: at its 80-register cap the kernel lost 10 % to any in-loop walk arithmetic.
  // A group needs >= ~12 items per CTA or the ragged split costs more than the re-reads (round 1's frame-by-frame
  // split, 6.5 items per CTA, ran 11 % slower: profiles/r01_s43_*).
  int first, count;
  bool listed = false;
  {
    const int G = geo.group;
    const int items_g = G * M * tiles, full_groups = G < p.N ? p.N / G : 0;
    const int items_t = total - full_groups * items_g;                               // the tail group (or the whole launch)
    const int per_g = (items_g + gridDim.x - 1) / gridDim.x, per_t = (items_t + gridDim.x - 1) / gridDim.x;
    const int cnt_g = max(0, min(per_g, items_g - (int)blockIdx.x * per_g));
    const int cnt_t = max(0, min(per_t, items_t - (int)blockIdx.x * per_t));
    count = full_groups * cnt_g + cnt_t;
    first = full_groups * items_g + (int)blockIdx.x * per_t;                          // one group: first item of the run
    if (full_groups > 0 && count <= kPlMaxList) {
      listed = true;
      for (int n = tid; n < count; n += blockDim.x) {
        const int g = cnt_g > 0 ? min(n / cnt_g, full_groups) : full_groups;
        sItems[n] = g < full_groups ? g * items_g + (int)blockIdx.x * per_g + (n - g * cnt_g)
                                    : first + (n - full_groups * cnt_g);
      }
      first = 0;
    } else if (full_groups > 0) {                                                     // list too long: one contiguous run
      const int per_cta = (total + gridDim.x - 1) / gridDim.x;
      first = blockIdx.x * per_cta;
      count = max(0, min(per_cta, total - first));
    }
  }
  __syncthreads();
  auto item_at = [&](int n) -> int { return listed ? sItems[n] : first + n; };

  // geometric image of tile (ty, tx) in sampled level l: top-left window corner that centres the image
  auto geo_origin = [&](int l, int ty, int tx, int& gh, int& gw) {
    const float sy = (float)geo.H[l] / (float)geo.H[0], sx = (float)geo.W[l] / (float)geo.W[0];
    const float cy = ((float)(ty * kPlTH) + 0.5f * (float)kPlTH) * sy - 0.5f;      // image of the tile centre
    const float cx = ((float)(tx << kPlTWlog2) + 0.5f * (float)(1 << kPlTWlog2)) * sx - 0.5f;
    gh = __float2int_rd(cy) - (pl_wh(l) - 2) / 2;
    gw = __float2int_rd(cx) - (pl_ww(l) - 2) / 2;
  };

  if (warp == kPlCons + kPlFront) {
    // ============================== producer ==============================
    if (lane == 0) {
      int prev_m = -1, prev_gh[NL], prev_gw[NL], prev_h0[NL], prev_w0[NL];
      for (int n = 0; n < count; ++n) {
        const int item = item_at(n);
        const int tile = item % tiles, bm = item / tiles;
        const int m = bm % M, b = bm / M;
        const int ty = tile / ntx, tx = tile - ty * ntx;
        const uint32_t prev_par = (uint32_t)(n - 1) & 1u;
        if (n >= 1) mbar_wait(smem_u32(&sEmpty[0]), prev_par);        // pass 0 of item n-1 is over: its phase-1 sums are complete
        int h0[NL], w0[NL];
#pragma unroll
        for (int l = 0; l < NL; ++l) {
          int gh, gw;
          geo_origin(l, ty, tx, gh, gw);
          int dh = 0, dw = 0;
          if (n >= 1 && m == prev_m) {
            int sh = 0, sw = 0, cnt = 0;
#pragma unroll
            for (int f = 0; f < kPlFront; ++f) {
              sh += sPart[(n - 1) & 1][f][l][0]; sw += sPart[(n - 1) & 1][f][l][1]; cnt += sPart[(n - 1) & 1][f][l][2];
            }
            if (cnt > 0) {     // where item n-1 would have wanted its window, relative to its geometric origin
              const float inv = 1.0f / (float)cnt;
              dh = __float2int_rn((float)sh * inv) - (pl_wh(l) - 2) / 2 - prev_gh[l];
              dw = __float2int_rn((float)sw * inv) - (pl_ww(l) - 2) / 2 - prev_gw[l];
              dh = max(-16, min(16, dh));      // a torn read of the statistics must not throw the window off the tile
              dw = max(-16, min(16, dw));
            } else {
              dh = prev_h0[l] - prev_gh[l];
              dw = prev_w0[l] - prev_gw[l];
            }
          }
          prev_gh[l] = gh; prev_gw[l] = gw;
          int ch = gh + dh, cw = gw + dw;
          ch = min(ch, geo.H[l] + 1 - pl_wh(l));
          cw = min(cw, geo.W[l] + 1 - pl_ww(l));
          h0[l] = max(ch, -1);
          w0[l] = max(cw, -1);
          prev_h0[l] = h0[l]; prev_w0[l] = w0[l];
          sOrg[n & 1][l][0] = h0[l];
          sOrg[n & 1][l][1] = w0[l];
        }
        prev_m = m;
#pragma unroll
        for (int l = 0; l < NL; ++l) {
          if (n >= 1 && l >= 1) mbar_wait(smem_u32(&sEmpty[l]), prev_par);
          const uint32_t bar = smem_u32(&sFull[l]);
          mbar_expect_tx(bar, pl_rows(l) * kPlRowB);      // release: the origins above are visible to whoever sees the phase complete
          tma_load_5d(sWinBase + pl_woff(l), &geo.lv[l], bar, 0, m, w0[l], h0[l], b);
        }
      }
    }
    return;
  }

  const int g = lane >> 2, k = lane & 3;                 // unit slot in the warp; lane in the unit = sampled level of its 4 samples
  // an item is (frame b, head m, tile (ty, tx)); decoded once per item (integer divisions), not once per strip
  struct ItemPos { int b, m, y0, x0; };
  auto decode = [&](int item) {
    ItemPos it;
    const int tile = item % tiles, bm = item / tiles;
    it.m = bm % M; it.b = bm / M;
    const int ty = tile / ntx;
    it.y0 = ty * kPlTH; it.x0 = (tile - ty * ntx) << kPlTWlog2;
    return it;
  };
  // unit (strip w, slot g) of an item: validity and flattened (batch, query) index
  auto locate = [&](const ItemPos& it, int w, bool& valid, size_t& bq) {
    const int j = w * kPlUPW + g;
    const int y = it.y0 + (j >> kPlTWlog2), x = it.x0 + (j & ((1 << kPlTWlog2) - 1));
    valid = (y < geo.H[0]) && (x < geo.W[0]);
    bq = (size_t)it.b * Lq + (valid ? geo.start[0] + y * geo.W[0] + x : 0);
  };

  if (warp >= kPlCons) {
    // ============================== front end: phase 1, one item ahead of the gather ==============================
    const int f = warp - kPlCons;
    const float inv_p = 1.0f / (float)PT;
    const float Hf = (float)geo.H[k], Wf = (float)geo.W[k];
    auto prefetch = [&](Prefetched<SPL, 1, FUSED>& pf, size_t bq, int m) {
      load_unit_operands<SPL, 1, FUSED, LPT, NL>(pf, p, bq, m, k * SPL, k);
    };
    bool n_valid = false;
    size_t n_bq = 0;
    ItemPos cur_it{}, nxt_it{};
    Prefetched<SPL, 1, FUSED> pf;
    if (count > 0) {
      cur_it = decode(item_at(0));
      locate(cur_it, f, n_valid, n_bq);
      prefetch(pf, n_bq, cur_it.m);
    }
    for (int n = 0; n < count; ++n) {
      const int buf = n & 1, j = n >> 1;
      int sh = 0, sw = 0, cnt = 0;
      if (n + 1 < count) nxt_it = decode(item_at(n + 1));
#pragma unroll 1
      for (int w = f; w < kPlCons; w += kPlFront) {       // this warp's strips of the item
        const bool valid = n_valid;
        float a[SPL], lx[SPL], ly[SPL];
        if constexpr (FUSED) {
          // softmax over the unit's 16 logits in the operation order of PyTorch's persistent warp softmax
          // (msda_forward_fast.cu): element e = 4k + i lives in register i of lane k
          const float sum = unit_softmax_terms<SPL, 4, LPT>(pf.lg, a);
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
        // operands of the next strip (or of the next item's first strip) are in flight while this one is finished
        {
          const bool same = w + kPlFront < kPlCons;
          if (same || n + 1 < count) {
            const ItemPos& it = same ? cur_it : nxt_it;
            locate(it, same ? w + kPlFront : f, n_valid, n_bq);
            prefetch(pf, n_bq, it.m);
          }
        }
        if (n >= 2) mbar_wait(smem_u32(&sRecEmpty[buf][w]), (uint32_t)(j - 1) & 1u);   // the gather of item n-2 released this strip
        float4* sRec = reinterpret_cast<float4*>(pl_smem + (size_t)buf * kPlRecBuf) + (size_t)w * kPlRecStrip;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
          // cuh:285-288 (one FFMA each), cuh:39-45
          const float h_im = __fmaf_rn(ly[i], Hf, -0.5f);
          const float w_im = __fmaf_rn(lx[i], Wf, -0.5f);
          const bool inr = valid && (h_im > -1.0f) && (w_im > -1.0f) && (h_im < Hf) && (w_im < Wf);
          const float hf = floorf(h_im), wf = floorf(w_im);
          const int hl = inr ? (int)hf : 0, wl = inr ? (int)wf : 0;
          if (inr) { sh += hl; sw += wl; ++cnt; }
          // h_low, w_low >= -1: stored + 1 in 16 bits each; 0xffffffff marks a skipped sample (cuh:288)
          const uint32_t hw = inr ? (((uint32_t)(hl + 1) << 16) | (uint32_t)(wl + 1)) : 0xffffffffu;
          sRec[(k * SPL + i) * kPlUPW + (g ^ (2 * k))] =
              make_float4(__uint_as_float(hw), __fsub_rn(h_im, hf), __fsub_rn(w_im, wf), inr ? a[i] : 0.0f);
        }
        if (w + kPlFront >= kPlCons) {   // placement statistics of this item, before the last strip is published
#pragma unroll
          for (int o = 4; o <= 16; o <<= 1) {
            sh += __shfl_xor_sync(0xffffffffu, sh, o);
            sw += __shfl_xor_sync(0xffffffffu, sw, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
          }
          if (g == 0) { sPart[buf][f][k][0] = sh; sPart[buf][f][k][1] = sw; sPart[buf][f][k][2] = cnt; }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sRecFull[buf][w]));
      }
      cur_it = nxt_it;
    }
    return;
  }

  // ============================== consumers: the gather ==============================
  const int c0 = k * 16 + (g & 1) * 64;                  // first 16-byte chunk of a row this lane owns; the second is c0 +- 64
  const int dhi = (g & 1) ? -64 : 64;
  for (int n = 0; n < count; ++n) {
    bool valid;
    size_t bq;
    const ItemPos it = decode(item_at(n));
    const int m = it.m, b = it.b;
    locate(it, warp, valid, bq);
    const size_t unit = bq * M + m;
    const uint32_t par = (uint32_t)n & 1u;
    const int buf = n & 1;
    const char* vhead = reinterpret_cast<const char*>(p.value) + ((size_t)b * p.S * M + m) * kPlRowB;
    const float4* sRec = reinterpret_cast<const float4*>(pl_smem + (size_t)buf * kPlRecBuf) + (size_t)warp * kPlRecStrip;
    mbar_wait(smem_u32(&sRecFull[buf][warp]), (uint32_t)(n >> 1) & 1u);

    // ---------------- four level passes ----------------
    f32x2 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = 0ull;
    auto level_pass = [&](auto level_c) {
      constexpr int l = decltype(level_c)::value;
      mbar_wait(smem_u32(&sFull[l]), par);
      const int h0 = sOrg[par][l][0], w0 = sOrg[par][l][1];
      const uint32_t wbase = sWinBase + pl_woff(l) + c0;
      const int Hl = geo.H[l], Wl = geo.W[l];
#pragma unroll
      for (int pt = 0; pt < PT; ++pt) {
        const int s = l * PT + pt;
        const float4 r = sRec[s * kPlUPW + (g ^ (2 * l))];
        const uint32_t hw = __float_as_uint(r.x);
        const int hl = (int)(hw >> 16) - 1, wl = (int)(hw & 0xffffu) - 1;
        const int dh = hl - h0, dw = wl - w0;
        const bool skipped = hw == 0xffffffffu;
        const bool inwin = ((unsigned)dh <= (unsigned)(pl_wh(l) - 2)) && ((unsigned)dw <= (unsigned)(pl_ww(l) - 2));
        RowVec<32> q[4];
        if (inwin || skipped) {
          // a skipped sample (attention 0) reads the window's first pixel: finite data, no effect on the sum.  (Not
          // loading at all -- a third branch that zeroes the registers -- was measured: 525 vs 476 us.)
          const uint32_t a0 = wbase + (skipped ? 0u : (uint32_t)((dh * pl_ww(l) + dw) * kPlRowB)), a1 = a0 + (uint32_t)dhi;
          q[0].lo = pl_lds128<0>(a0);                              q[0].hi = pl_lds128<0>(a1);
          q[1].lo = pl_lds128<kPlRowB>(a0);                        q[1].hi = pl_lds128<kPlRowB>(a1);
          q[2].lo = pl_lds128<pl_ww(l) * kPlRowB>(a0);             q[2].hi = pl_lds128<pl_ww(l) * kPlRowB>(a1);
          q[3].lo = pl_lds128<pl_ww(l) * kPlRowB + kPlRowB>(a0);   q[3].hi = pl_lds128<pl_ww(l) * kPlRowB + kPlRowB>(a1);
        } else {
          // footprint outside the window: predicated global loads, skipped corners read as zero (cuh:56-78)
          const bool top = hl >= 0, bot = hl + 1 <= Hl - 1, lef = wl >= 0, rig = wl + 1 <= Wl - 1;
          const char* c1 = vhead + (ptrdiff_t)((geo.start[l] + hl * Wl + wl) * cstride) + c0;
          const char* c3 = c1 + Wl * cstride;
#pragma unroll
          for (int c = 0; c < 4; ++c) q[c].zero();
          if (top && lef) { q[0].lo = ld_value16(c1);            q[0].hi = ld_value16(c1 + dhi); }
          if (top && rig) { q[1].lo = ld_value16(c1 + cstride);  q[1].hi = ld_value16(c1 + cstride + dhi); }
          if (bot && lef) { q[2].lo = ld_value16(c3);            q[2].hi = ld_value16(c3 + dhi); }
          if (bot && rig) { q[3].lo = ld_value16(c3 + cstride);  q[3].hi = ld_value16(c3 + cstride + dhi); }
        }
        float w1, w2, w3, w4;
        bilinear_weights(r.y, r.z, w1, w2, w3, w4);
        accumulate_sample<float, 32, 4>(acc, make_float4(w1, w2, w3, w4), r.w, q[0], q[1], q[2], q[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sEmpty[l]));      // this warp is done with window slot l for this item
    };
    level_pass(std::integral_constant<int, 0>{});
    level_pass(std::integral_constant<int, 1>{});
    level_pass(std::integral_constant<int, 2>{});
    level_pass(std::integral_constant<int, 3>{});
    if (lane == 0) mbar_arrive(smem_u32(&sRecEmpty[buf][warp]));   // (after the __syncwarp of pass 3) the strip may be rewritten
    if (valid) {
      char* op = reinterpret_cast<char*>(p.out) + unit * (size_t)(D * 4);
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) upk(acc[j], f[2 * j], f[2 * j + 1]);
      st_stream16(op + c0, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
      st_stream16(op + c0 + dhi, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
    }
  }
}

template <bool FUSED>
int launch_pipelined(const FwdParams& p, const PipeGeom& geo, cudaStream_t stream) {
  auto kern = msda_fwd_pipelined_kernel<FUSED>;
  static PerDeviceOnce configured;   // function attributes are per device
  if (configured.need()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPlSmem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<p.grid, pl_threads(FUSED), kPlSmem, stream>>>(p, geo);
  return (int)cudaGetLastError();
}

}  // namespace

// Returns MSDA_E_UNSUPPORTED when the host-side level geometry is missing or does not match (the caller then uses
// another kernel).  Serves the level-0 queries only; the caller runs the remaining levels on the register-gather kernel.
int launch_forward_pipelined_f32(const FwdParams& p, const long long (*hw)[2], const long long* lsi, cudaStream_t stream) {
  EncodeTiledFn fn = tensor_map_encode_fn();
  if (!fn || !hw || !lsi) return MSDA_E_UNSUPPORTED;
  PipeGeom geo;
  memset(&geo, 0, sizeof(geo));
  long long total = 0;
  for (int l = 0; l < kPlL; ++l) total += hw[l][0] * hw[l][1];
  if (total != p.S || hw[0][0] >= 32768 || hw[0][1] >= 32768) return MSDA_E_UNSUPPORTED;
  const cuuint64_t row = (cuuint64_t)p.D * 4, px = (cuuint64_t)p.M * row;
  for (int l = 0; l < kPlL; ++l) {
    const cuuint64_t H = (cuuint64_t)hw[l][0], W = (cuuint64_t)hw[l][1];
    geo.H[l] = (int)H; geo.W[l] = (int)W; geo.start[l] = (int)lsi[l];
    cuuint64_t gdim[5] = {(cuuint64_t)p.D, (cuuint64_t)p.M, W, H, (cuuint64_t)p.N};
    cuuint64_t gstride[4] = {row, px, W * px, (cuuint64_t)p.S * px};
    cuuint32_t box[5] = {(cuuint32_t)p.D, 1, (cuuint32_t)pl_ww(l), (cuuint32_t)pl_wh(l), 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void* base = const_cast<char*>(reinterpret_cast<const char*>(p.value) + (size_t)lsi[l] * px);
    if (fn(&geo.lv[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MSDA_E_UNSUPPORTED;
  }
  // frames per group: as many as keep the group's value maps well inside L2 (48 MB of the 126), but at least enough
  // for ~12 items per CTA; p.walk = 1 (diagnostic) restores the single group of round 1
  {
    const long long map_bytes = (long long)p.S * px;
    const long long tiles = ((hw[0][0] + kPlTH - 1) / kPlTH) * ((hw[0][1] + (1 << kPlTWlog2) - 1) >> kPlTWlog2);
    long long g = (48ll << 20) / (map_bytes > 0 ? map_bytes : 1);
    if (g < 1) g = 1;
    while (g < p.N && g * p.M * tiles < 12ll * p.grid) ++g;
    if (g > p.N || p.walk == 1) g = p.N;
    geo.group = (int)g;
  }
  return p.loc == nullptr ? launch_pipelined<true>(p, geo, stream) : launch_pipelined<false>(p, geo, stream);
}

}  // namespace msda
