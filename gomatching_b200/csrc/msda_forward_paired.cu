// Neighbour-paired bf16 value layout and its sampler (BASELINE.json config 3: bf16 storage, 2e-2 bar) -- an explicit
// operator MODE, not the fp32 default: it changes the value format and the summation order.
//
// Why: the gather is bound by the SM's L1 data path at one 128-byte wavefront per clock.  In the
// operator's layout (N, S, M, D) a bf16 pixel-head row is 64 B but still costs a whole wavefront, and the four corners of
// a bilinear sample sit in four different 128-byte lines: 4 wavefronts per sample, exactly like fp32.  Here
//     paired[b][m][p] = { value[b][p][m][0..31] , value[b][p+1][m][0..31] }   (bf16, 128 B; the second half is zero at the
//                                                                           last pixel of an image row)
// so ONE line holds both horizontal neighbours: a sample is two wavefronts (upper pair, lower pair), and a head's map is
// contiguous (neighbouring pixels 128 B apart instead of M*D*2 = 512 B).  Same bytes as the fp32 tensor (every pixel is
// stored twice), half the wavefronts of either unpaired layout.  Written by msda_pair_value_kernel below (or, in a
// module, by whoever produces `value`); S, level_start_index and spatial_shapes are unchanged -- pixel p keeps its index.
//
// Sampler: a unit (batch, query, head) is 8 lanes: lanes 0-3 own the LEFT pixel of a pair (channels 8c .. 8c+7 each),
// lanes 4-7 the RIGHT one.  Per sample a lane loads 16 B of the upper and 16 B of the lower pair line (two LDG.128 per
// warp = 8 wavefronts for 4 unit-samples), accumulates attn * (row weight) * (its column weight) * v in fp32, and the two
// halves are added once per unit with a shuffle.  w_low = -1 (left corner outside the map): the right corner v(y, 0) is
// the FIRST half of pair 0, so the right-hand lanes read offset 0 instead of 64; rows outside the map read the valid
// neighbour row with weight 0.  Phase 1 (softmax, offset -> location, floor, range test) is the fast kernels' code, so
// sampling indices are the same bits; outputs differ from the fp32 chain only by bf16 storage and summation order.
#include "msda_fast_common.cuh"
#include "msda_launch.h"
#include "../../include/msda_b200.h"

namespace msda {

namespace {

constexpr int kPrD = 32, kPrL = 4, kPrP = 4, kPrLPT = 16;
constexpr int kPrLanes = 8;                    // lanes per unit
constexpr int kPrUPW = 32 / kPrLanes;          // 4 units per warp step
constexpr int kPrSPL = kPrLPT / kPrLanes;      // 2 samples per lane in phase 1
constexpr int kPrLine = 128;                   // bytes of one pair

// ---- layout conversion: (N, S, M, 32) fp32 | bf16  ->  (N, M, S, 2, 32) bf16 --------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) msda_pair_value_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                                                              const int64_t* __restrict__ lsi, int N, int S, int M, int L,
                                                              uint4* __restrict__ paired) {
  __shared__ int sW[kMaxLevels], sStart[kMaxLevels + 1];
  if (threadIdx.x < L) {
    sW[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
    sStart[threadIdx.x] = (int)lsi[threadIdx.x];
  }
  if (threadIdx.x == 0) sStart[L] = S;
  __syncthreads();
  // one thread = one 16-byte chunk (8 channels) of one half of one pair
  const long long total = (long long)N * M * S * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 3), half = (int)((i >> 2) & 1);
    long long r = i >> 3;
    const int p = (int)(r % S);
    r /= S;
    const int m = (int)(r % M), b = (int)(r / M);
    int l = 0;
    while (l + 1 < L && p >= sStart[l + 1]) ++l;
    const int x = (p - sStart[l]) % sW[l];
    uint4 out = make_uint4(0, 0, 0, 0);
    if (!(half && x == sW[l] - 1)) {
      const T* src = value + (((size_t)b * S + p + half) * M + m) * kPrD + c * 8;
      if constexpr (sizeof(T) == 2) {
        out = *reinterpret_cast<const uint4*>(src);
      } else {
        const float4 a = *reinterpret_cast<const float4*>(src), d = *reinterpret_cast<const float4*>(src + 4);
        using E = Elem<__nv_bfloat16>;
        out = make_uint4(E::pack2(a.x, a.y), E::pack2(a.z, a.w), E::pack2(d.x, d.y), E::pack2(d.z, d.w));
      }
    }
    paired[i] = out;
  }
}

// ---- sampler ------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int rec_slot(int s) { return s ^ ((s >> 3) & 1); }     // sample s = 2k + i, k = lane in unit
__host__ __device__ constexpr int rec_swz(int s) { return (s >> 1) & 3; }

template <bool FUSED, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) msda_fwd_paired_kernel(const FwdParams p) {
  constexpr int NL = kPrL, PT = kPrP, LPT = kPrLPT, SPL = kPrSPL, UPW = kPrUPW;
  __shared__ int sH[NL], sW[NL], sStart[NL];
  __shared__ float sHf[NL], sWf[NL];
  extern __shared__ float4 sRecAll[];            // [NW][LPT][UPW]

  const int M = p.M, Lq = p.Lq;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 3, k = lane & 7;         // unit slot in the warp, lane in the unit
  const int half = k >> 2, chunk16 = (k & 3) * 16;
  if (tid < NL) {
    sH[tid] = (int)p.shapes[2 * tid];
    sW[tid] = (int)p.shapes[2 * tid + 1];
    sStart[tid] = (int)p.lsi[tid];
    sHf[tid] = (float)sH[tid];
    sWf[tid] = (float)sW[tid];
  }
  __syncthreads();
  float4* sRec = sRecAll + (size_t)warp * LPT * UPW;
  int rstride[NL];                               // bytes between vertically adjacent pairs, per level
#pragma unroll
  for (int l = 0; l < NL; ++l) rstride[l] = sW[l] * kPrLine;
  const float inv_p = 1.0f / (float)PT;
  const int lvl = (k * SPL) / PT;                // level of this lane's two phase-1 samples
  const int Hl = sH[lvl], Wl = sW[lvl];
  const float Hf = sHf[lvl], Wf = sWf[lvl];

  // work: tile = (batch, block of tile_q queries, head), heads fastest (all CTAs stay inside one frame's value map)
  const int tiles_per_bm = (Lq + p.tile_q - 1) / p.tile_q;
  const long long total_tiles = (long long)p.N * tiles_per_bm * M;
  const int chunks = (p.tile_q + NW * UPW - 1) / (NW * UPW);
  long long tile = blockIdx.x;
  int chunk = 0, t_m = 0, t_b = 0, t_t = 0;
  auto decode_tile = [&]() {
    t_m = (int)(tile % M);
    const long long r = tile / M;
    t_t = (int)(r % tiles_per_bm);
    t_b = (int)(r / tiles_per_bm);
  };
  auto locate = [&](bool& valid, size_t& bq) {
    const int j = (chunk * NW + warp) * UPW + g;
    const int qi = t_t * p.tile_q + j;
    valid = (j < p.tile_q) && (qi < Lq);
    bq = (size_t)t_b * Lq + (valid ? qi : 0);
  };
  bool have = tile < total_tiles;
  bool n_valid = false;
  size_t n_bq = 0;
  int n_m = 0;
  const char* n_vbase = nullptr;
  Prefetched<SPL, 1, FUSED> pf;
  if (have) {
    decode_tile();
    locate(n_valid, n_bq);
    n_m = t_m;
    n_vbase = reinterpret_cast<const char*>(p.value) + ((size_t)t_b * M + t_m) * (size_t)p.S * kPrLine + chunk16;
    load_unit_operands<SPL, 1, FUSED, LPT, NL>(pf, p, n_bq, n_m, k * SPL, lvl);
  }

  while (have) {
    const bool valid = n_valid;
    const size_t bq = n_bq;
    const int m = n_m;
    const char* vbase = n_vbase;

    // ---------------- phase 1: one 16-byte record per sample ----------------
    {
      float a[SPL], lx[SPL], ly[SPL];
      if constexpr (FUSED) {
        const float sum = unit_softmax_terms<SPL, kPrLanes, LPT>(pf.lg, a);
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
        const int s = k * SPL + i;
        // cuh:285-288 (one FFMA each), :39-45 -- the index contract of every kernel in this library
        const float h_im = __fmaf_rn(ly[i], Hf, -0.5f);
        const float w_im = __fmaf_rn(lx[i], Wf, -0.5f);
        const bool inr = valid && (h_im > -1.0f) && (w_im > -1.0f) && (h_im < Hf) && (w_im < Wf);
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h_low = inr ? (int)hf : 0, w_low = inr ? (int)wf : 0;
        const bool top = h_low >= 0, bot = h_low + 1 <= Hl - 1, lef = w_low >= 0, rig = w_low + 1 <= Wl - 1;
        const int row = top ? h_low : 0, col = lef ? w_low : 0;      // -1 -> the valid neighbour 0 (weight 0 / first half)
        // byte offset of the upper pair inside the head's map, flags in the low bits (a pair is 128 B)
        const int off = ((sStart[lvl] + row * Wl + col) * kPrLine) | (int)top | ((int)bot << 1) | ((int)lef << 2) | ((int)rig << 3);
        // slot swizzle: the quarter-warp STS.128 (8 lanes of one unit, samples 2k + i) then covers all eight 16-byte bank
        // groups; readers apply the same compile-time permutation
        sRec[rec_slot(s) * UPW + (g ^ rec_swz(s))] =
            make_float4(__int_as_float(off), __fsub_rn(h_im, hf), __fsub_rn(w_im, wf), inr ? a[i] : 0.0f);
      }
    }
    __syncwarp();

    // ---------------- advance, prefetch the next step's operands ----------------
    ++chunk;
    if (chunk == chunks) {
      chunk = 0;
      tile += gridDim.x;
      have = tile < total_tiles;
      if (have) {
        decode_tile();
        n_vbase = reinterpret_cast<const char*>(p.value) + ((size_t)t_b * M + t_m) * (size_t)p.S * kPrLine + chunk16;
      }
    }
    if (have) {
      locate(n_valid, n_bq);
      n_m = t_m;
      load_unit_operands<SPL, 1, FUSED, LPT, NL>(pf, p, n_bq, n_m, k * SPL, lvl);
    }

    // ---------------- phase 2: two pair lines per sample ----------------
    f32x2 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = 0ull;
#pragma unroll
    for (int s = 0; s < LPT; ++s) {
      const float4 r = sRec[rec_slot(s) * UPW + (g ^ rec_swz(s))];
      const int packed = __float_as_int(r.x);
      const bool top = packed & 1, bot = packed & 2, lef = packed & 4, rig = packed & 8;
      const float lh = r.y, lw = r.z;
      // column weight of this lane's half, attention folded in; rows outside the map get weight 0
      const float colw = half ? (rig ? lw : 0.0f) : (lef ? __fsub_rn(1.0f, lw) : 0.0f);
      const float t = __fmul_rn(colw, r.w);
      const float wt = top ? __fmul_rn(__fsub_rn(1.0f, lh), t) : 0.0f;
      const float wb = bot ? __fmul_rn(lh, t) : 0.0f;
      const char* up = vbase + (packed & ~15) + ((half && lef) ? 64 : 0);
      const char* dn = up + ((top && bot) ? rstride[s / PT] : 0);
      const uint4 u = ld_value16(up), d = ld_value16(dn);
      const f32x2 WT = pk(wt, wt), WB = pk(wb, wb);
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j] = fma2(WT, pk(__uint_as_float(uw[j] << 16), __uint_as_float(uw[j] & 0xffff0000u)), acc[j]);
        acc[j] = fma2(WB, pk(__uint_as_float(dw[j] << 16), __uint_as_float(dw[j] & 0xffff0000u)), acc[j]);
      }
    }
    // left + right halves, then lanes 0-3 of the unit store 8 bf16 channels each
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo, hi;
      upk(acc[j], lo, hi);
      f[2 * j] = __fadd_rn(lo, __shfl_xor_sync(0xffffffffu, lo, 4));
      f[2 * j + 1] = __fadd_rn(hi, __shfl_xor_sync(0xffffffffu, hi, 4));
    }
    if (valid && half == 0) {
      using E = Elem<__nv_bfloat16>;
      char* op = reinterpret_cast<char*>(p.out) + ((bq * M + m) * kPrD) * 2 + chunk16;
      st_stream16(op, make_uint4(E::pack2(f[0], f[1]), E::pack2(f[2], f[3]), E::pack2(f[4], f[5]), E::pack2(f[6], f[7])));
    }
    __syncwarp();
  }
}

template <bool FUSED>
int launch_paired(const FwdParams& p, cudaStream_t stream) {
  constexpr int NW = 8, MINB = 4;
  const size_t smem = (size_t)NW * kPrLPT * kPrUPW * sizeof(float4);      // 8 KB
  auto kern = msda_fwd_paired_kernel<FUSED, NW, MINB>;
  static PerDeviceOnce configured;
  if (configured.need()) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 20);
  kern<<<p.grid, NW * 32, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace

int launch_pair_value(const void* value, int value_is_bf16, const int64_t* shapes, const int64_t* lsi, int N, int S, int M,
                      int L, void* paired, int sms, cudaStream_t stream) {
  const long long total = (long long)N * M * S * 8;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
  if (blocks < 1) blocks = 1;
  if (value_is_bf16)
    msda_pair_value_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(value), shapes, lsi,
                                                                          N, S, M, L, reinterpret_cast<uint4*>(paired));
  else
    msda_pair_value_kernel<float><<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(value), shapes, lsi, N, S, M, L,
                                                                  reinterpret_cast<uint4*>(paired));
  return (int)cudaGetLastError();
}

int launch_forward_paired_bf16(const FwdParams& p, cudaStream_t stream) {
  return p.loc == nullptr ? launch_paired<true>(p, stream) : launch_paired<false>(p, stream);
}

}  // namespace msda
