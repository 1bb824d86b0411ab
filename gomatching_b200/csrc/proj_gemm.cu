// Bordering dense projections of MSDeformAttn on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   y[M, N] = x[M, K] . w[N, K]^T + bias[N]        (nn.Linear: value_proj, sampling_offsets || attention_weights,
//                                                    output_proj -- third_party/adet/layers/ms_deform_attn.py:133-153)
//
// The reference runs these as fp32 cuBLAS GEMMs.  A single-pass TF32 tensor-core GEMM (10-bit mantissa) is ~1e-3
// off and would break the 1e-4 output bar -- and, through the sampling offsets, move sampling indices -- so the
// kernel computes the 3xTF32 product: x = xh + xl, w = wh + wl (each part exactly representable in TF32),
//   y = xh.wh + xh.wl + xl.wh      (xl.wl ~ 2^-22 is dropped), fp32 accumulation in TMEM,
// which is fp32-grade (measured max-norm error vs float64 in tests/test_proj_gemm_gpu.py) at tensor-core speed.
//
// One CTA = one 128-row tile of x times one N tile (<= 256 columns) of w; K is walked in blocks of 16 floats:
//   warp 0     TMA producer   x block (128 x 64 B), wh and wl blocks (BN x 64 B), 64-byte swizzle, mbarrier tx
//   warps 2-5  splitter       xh = x & ~0x1fff (in place), xl = tf32(x - xh) into a second buffer; layout-agnostic
//                             (element-wise on 16-byte chunks, so the TMA swizzle is preserved)
//   warp 1     MMA issuer     one elected lane: 2 K-steps x 3 tcgen05.mma.kind::tf32 per block, tcgen05.commit
//                             releases the stage / publishes the accumulator
//   warps 2-5  epilogue       tcgen05.ld 32 lanes x 32 columns, + bias, optional row zeroing (padding mask),
//                             128-byte-swizzled staging in the drained pipeline buffers, TMA store per warp
// Two CTAs share an SM (2 x ~100 KB shared memory, 2 x 256 TMEM columns), so one CTA's epilogue overlaps the
// other's main loop.  Rows beyond M are zero-filled by the TMA loads and clipped by the TMA stores.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/msda_b200.h"
#include "tma_common.cuh"
#include "msda_launch.h"

namespace msda {
namespace {

constexpr int kSubM = 128;                  // rows per MMA (UMMA M)
constexpr int kBlockK = 16;                 // floats per K block = one 64-byte swizzle row
constexpr int kUmmaK = 8;                   // tf32: 32 bytes per MMA K step
constexpr int kThreads = 192;
constexpr int kSubBytes = kSubM * kBlockK * 4;          // 8 KB: one 128-row x 64-byte operand slice

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile in shared memory, rows of 64 bytes, 64-byte swizzle: 8-row groups are 512 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);          // start address  [0,14)
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(512 >> 4) << 32;                   // stride byte offset: next 8-row group [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                            // layout type: SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// hi = x rounded to TF32 (round half away in magnitude on the bit pattern; low 13 mantissa bits then zero, so the
// tensor core sees it exactly whether it truncates or rounds), lo = TF32 head of the exact remainder x - hi
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& hi, uint32_t& lo) {
  hi = (x + 0x1000u) & 0xffffe000u;
  if ((x & 0x7f800000u) == 0x7f800000u) hi = x & 0xffffe000u;      // Inf / NaN: do not carry into the exponent
  lo = __float_as_uint(__fsub_rn(__uint_as_float(x), __uint_as_float(hi))) & 0xffffe000u;
}

struct GemmParams {
  const float* w_hi;             // [K/16][N][16] TF32 head of the weight, 64-byte-swizzle image (split_weight_kernel)
  const float* w_lo;             // same layout, TF32 head of the remainder
  float* y;                      // [M][ldy]
  int ldy;
  const float* bias;             // [N] or nullptr
  const unsigned char* row_zero; // [M] or nullptr: rows written as 0 (padding mask, ms_deform_attn.py:135)
  int relu;                      // epilogue: y = max(y, 0) after the bias (encoder FFN linear1 + activation, deformable_transformer.py:249)
  int M, N, K;
  int block_n;                   // columns per CTA (multiple of 32, <= 256)
  int tmem_cols;                 // power of two >= the accumulator columns of one CTA
  int diag;                      // -DMSDA_DIAG builds only: 1 = epilogue skips its work, 2 = splitter skips the split, 4 = no MMAs,
                                 // 16 = epilogue converts but does not store
  int split_groups;              // persistent kernel: 1 or 2 groups of four splitter warps
  int stages;                    // persistent kernel: pipeline depth (the tile-per-CTA kernels fix it at compile time)
  long long* trace;              // optional [grid][8] clock64 stamps (MSDA_GEMM_TRACE), else nullptr
};

// measurement switches of the persistent kernel (tools/gemm_trace.py): compiled in only with -DMSDA_DIAG (build.py --diag)
__device__ __forceinline__ int diag_on(const GemmParams& p) {
#ifdef MSDA_DIAG
  return p.diag;
#else
  return 0;
#endif
}

// MSUB 128-row sub-tiles per CTA share every weight block: MSUB = 2 halves the bytes the TMA has to pull from L2 per
// output row (the weights are re-streamed for every CTA tile; at MSUB = 1 the kernel is bound by that stream --
// 10240 64-byte TMA rows per tile, ~3.7 clk each -- not by the tensor pipe: profiles/r01_gemm_ncu_v1.txt).
template <int MSUB, int STAGES>
__global__ void __launch_bounds__(kThreads, MSUB == 1 ? 2 : 1)
proj_gemm_3xtf32_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                        const GemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  constexpr int kStages = STAGES;
  constexpr int kBlockM = kSubM * MSUB;
  constexpr uint32_t kABytes = kSubBytes * MSUB;
  __shared__ __align__(8) unsigned long long s_bar[3 * kStages + 1];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float s_bias[256];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.block_n;
  const int m0 = blockIdx.x * kBlockM, n0 = blockIdx.y * BN;
  const int num_kb = p.K / kBlockK;
  const uint32_t b_bytes = (uint32_t)BN * kBlockK * 4;
  const uint32_t stage_bytes = 2 * kABytes + 2 * b_bytes;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto stage_x = [&](int s) { return smem_base + (uint32_t)s * stage_bytes; };
  auto stage_xl = [&](int s) { return stage_x(s) + kABytes; };
  auto stage_wh = [&](int s) { return stage_x(s) + 2 * kABytes; };
  auto stage_wl = [&](int s) { return stage_wh(s) + b_bytes; };
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto bar_split = [&](int s) { return bar0 + 8u * (uint32_t)(kStages + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (uint32_t)(2 * kStages + s); };
  const uint32_t bar_accum = bar0 + 8u * (uint32_t)(3 * kStages);

  for (int i = tid; i < BN; i += kThreads) s_bias[i] = p.bias ? p.bias[n0 + i] : 0.0f;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_split(s), 4);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_accum, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  long long* trace = p.trace ? p.trace + 8 * (size_t)(blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  if (trace && tid == 0) trace[0] = clock64();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
        mbar_wait(bar_empty(s), ph ^ 1u);
        mbar_expect_tx(bar_full(s), kABytes + 2 * b_bytes);
        tma_load_2d(stage_x(s), &tm_x, bar_full(s), kb * kBlockK, m0);
        const size_t woff = ((size_t)kb * p.N + n0) * kBlockK;                 // k-blocked, pre-swizzled weight image
        bulk_load(stage_wh(s), p.w_hi + woff, b_bytes, bar_full(s));
        bulk_load(stage_wl(s), p.w_lo + woff, b_bytes, bar_full(s));
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kSubM >> 4) << 24);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
      mbar_wait(bar_full(s), ph);
      if (trace && lane == 0 && kb == 0) trace[1] = clock64();
      mbar_wait(bar_split(s), ph);
      if (trace && lane == 0 && kb == 0) trace[2] = clock64();
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          const uint32_t koff = (uint32_t)k * kUmmaK * 4;
          const uint64_t wh = umma_desc_sw64(stage_wh(s) + koff), wl = umma_desc_sw64(stage_wl(s) + koff);
#pragma unroll
          for (int h = 0; h < MSUB; ++h) {
            const uint64_t xh = umma_desc_sw64(stage_x(s) + h * kSubBytes + koff), xl = umma_desc_sw64(stage_xl(s) + h * kSubBytes + koff);
            const uint32_t acc = tmem_base + (uint32_t)(h * BN);
            umma_tf32(acc, xl, wh, idesc, (kb | k) ? 1u : 0u);     // small terms first
            umma_tf32(acc, xh, wl, idesc, 1u);
            umma_tf32(acc, xh, wh, idesc, 1u);
          }
        }
        umma_commit(bar_empty(s));                  // stage reusable once these MMAs have read it
        if (kb == num_kb - 1) umma_commit(bar_accum);
        if (trace && kb == num_kb - 1) trace[3] = clock64();
      }
      __syncwarp();
    }
  } else {
    // ===================== splitter (main loop) =====================
    const int t = tid - 64;                          // 0..127
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
      mbar_wait(bar_full(s), ph);
      const uint32_t xa = stage_x(s) + (uint32_t)t * 16u, la = stage_xl(s) + (uint32_t)t * 16u;
#pragma unroll
      for (int c = 0; c < (int)(kABytes / (128 * 16)); ++c) {       // consecutive threads take consecutive 16-byte chunks
        uint32_t v0, v1, v2, v3, h0, h1, h2, h3, l0, l1, l2, l3;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(xa + 2048u * c));
        split_tf32(v0, h0, l0); split_tf32(v1, h1, l1); split_tf32(v2, h2, l2); split_tf32(v3, h3, l3);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(xa + 2048u * c), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(la + 2048u * c), "r"(l0), "r"(l1), "r"(l2), "r"(l3) : "memory");
      }
      fence_proxy_async();                           // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_split(s));      // one arrival per splitter warp
    }
    // ===================== epilogue =====================
    mbar_wait(bar_accum, 0);
    if (trace && tid == 64) trace[4] = clock64();
    tc_fence_after();
    const int q = warp & 3;                          // TMEM lane quadrant this warp may read
    // staging in the drained pipeline buffers: each warp owns G slabs of 4 KB (32 rows x 32 columns, 128-byte swizzle,
    // 1024-byte aligned) used round-robin; chunk i is converted into slab i % G and handed to one TMA store, and a
    // slab is only waited for when it comes round again (cp.async.bulk.wait_group.read G-1)
    constexpr int G = MSUB == 2 ? 8 : 4;
    const uint32_t slab0 = smem_base + (uint32_t)(warp - 2) * (uint32_t)(G * 4096);
    const int chunks = BN / 32;
    constexpr int Q = 4;                             // chunks converted per proxy fence / store batch
    int it = 0;
#pragma unroll 1
    for (int h = 0; h < MSUB; ++h) {
      const int row0 = m0 + h * kSubM + 32 * q;
      const bool zero_row = p.row_zero != nullptr && row0 + lane < p.M && p.row_zero[row0 + lane] != 0;
#pragma unroll 1
      for (int c0 = 0; c0 < chunks; c0 += Q, it += Q) {
        const int qn = chunks - c0 < Q ? chunks - c0 : Q;
        if (it >= G) {                               // the batch that last used these slabs must be done reading them
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(G / Q - 1) : "memory");
          __syncwarp();
        }
#pragma unroll 1
        for (int g = 0; g < qn; ++g) {
          const int c = c0 + g;
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(h * BN + c * 32), v);
          const uint32_t slab = slab0 + (uint32_t)((it + g) % G) * 4096u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bj = *reinterpret_cast<const float4*>(&s_bias[c * 32 + 4 * j]);
            float f[4] = {__fadd_rn(__uint_as_float(v[4 * j + 0]), bj.x), __fadd_rn(__uint_as_float(v[4 * j + 1]), bj.y),
                          __fadd_rn(__uint_as_float(v[4 * j + 2]), bj.z), __fadd_rn(__uint_as_float(v[4 * j + 3]), bj.w)};
            if (p.relu) {                         // x < 0 ? 0 : x keeps NaN like torch.relu
#pragma unroll
              for (int e = 0; e < 4; ++e) f[e] = f[e] < 0.0f ? 0.0f : f[e];
            }
            if (zero_row) f[0] = f[1] = f[2] = f[3] = 0.0f;
            const uint32_t dst = slab + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) * 16);
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]) : "memory");
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          for (int g = 0; g < qn; ++g)
            tma_store_2d(&tm_y, slab0 + (uint32_t)((it + g) % G) * 4096u, n0 + (c0 + g) * 32, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    if (trace && tid == 64) trace[5] = clock64();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ---- persistent variant --------------------------------------------------------------------------------------------
// One CTA per SM walks the output tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (N tiles fastest, so CTAs that run
// at the same time share x tiles in L2).  What the tile-per-CTA kernel above leaves on the table at 8 frames
// (profiles/r01_gemm_trace.txt: main loop 98 % tensor-bound, but 3.6 k clk of prologue per tile, a 12.5 k clk epilogue
// that overlaps nothing, and 599 tiles on 148 SMs = 5 rounds for 4.05 rounds of work) is what this one removes:
//   * the accumulator is DOUBLE-BUFFERED in TMEM (2 x BN <= 512 columns): the MMA warp starts tile i+1 in the other
//     buffer while four dedicated epilogue warps drain tile i (tcgen05.ld -> +bias -> swizzled staging -> TMA store);
//   * barrier init, TMEM allocation, tensor-map fetch and the bias load happen once per CTA; the TMA producer runs ahead
//     across tile boundaries, so the pipeline never drains between tiles;
//   * tiles are half the size (128 x BN), so the last round wastes half as much.
//   * when everything fits one round (tiles <= SMs: the decoder's 2 500-token GEMMs) a CTA's life is one serial
//     TMA -> split -> MMA chain of K blocks and the splitter's ~700 clk per block is what it waits for; a second group
//     of four splitter warps then takes every other block (no gain, slightly worse, when CTAs run many tiles).
//   warp 0 TMA producer | warp 1 MMA issuer | warps 2 .. 2+4G-1 splitter (G = 1 or 2 groups) | next 4 warps epilogue
//   (TMEM lane quadrant = warp & 3)
constexpr int kPMaxThreads = 448;
constexpr int kPMaxStages = 6;
constexpr int kPSlabs = 2;                   // 4 KB staging slabs per epilogue warp
constexpr int kPBatch = 1;                   // chunks converted per proxy fence / store batch

__global__ void __launch_bounds__(kPMaxThreads, 1)
proj_gemm_3xtf32_persistent_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                                   const GemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int MSUB = 1;                      // two sub-tiles per CTA (as in the kernel above) never won here: the tile
                                               // would need BN <= 128 and the MMAs of a 128-column tile read twice the
                                               // operand bytes per flop
  constexpr int kBlockM = kSubM * MSUB;
  constexpr uint32_t kABytes = kSubBytes * MSUB;
  __shared__ __align__(8) unsigned long long s_bar[3 * kPMaxStages + 4];
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.block_n, stages = p.stages;
  const int n_tiles = p.N / BN;
  const int num_tiles = ((p.M + kBlockM - 1) / kBlockM) * n_tiles;
  const int num_kb = p.K / kBlockK;
  const uint32_t b_bytes = (uint32_t)BN * kBlockK * 4;
  const uint32_t stage_bytes = 2 * kABytes + 2 * b_bytes;
  const uint32_t acc_cols = (uint32_t)(MSUB * BN);           // columns of one accumulator buffer

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto stage_x = [&](int s) { return smem_base + (uint32_t)s * stage_bytes; };
  auto stage_xl = [&](int s) { return stage_x(s) + kABytes; };
  auto stage_wh = [&](int s) { return stage_x(s) + 2 * kABytes; };
  auto stage_wl = [&](int s) { return stage_wh(s) + b_bytes; };
  const uint32_t slab_base = smem_base + (uint32_t)stages * stage_bytes;
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto bar_split = [&](int s) { return bar0 + 8u * (uint32_t)(kPMaxStages + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (uint32_t)(2 * kPMaxStages + s); };
  auto bar_acc_full = [&](int b) { return bar0 + 8u * (uint32_t)(3 * kPMaxStages + b); };
  auto bar_acc_empty = [&](int b) { return bar0 + 8u * (uint32_t)(3 * kPMaxStages + 2 + b); };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_y) : "memory");
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_split(s), 4);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  long long* trace = p.trace ? p.trace + 8 * (size_t)blockIdx.x : nullptr;
  if (trace && tid == 0) trace[0] = clock64();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / n_tiles) * kBlockM, n0 = (t % n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_empty(s), ph ^ 1u);
          mbar_expect_tx(bar_full(s), kABytes + 2 * b_bytes);
          tma_load_2d(stage_x(s), &tm_x, bar_full(s), kb * kBlockK, m0);
          const size_t woff = ((size_t)kb * p.N + n0) * kBlockK;
          bulk_load(stage_wh(s), p.w_hi + woff, b_bytes, bar_full(s));
          bulk_load(stage_wl(s), p.w_lo + woff, b_bytes, bar_full(s));
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One thread; the loop is kept lean on purpose: with 6 MMAs of ~125 clk per K block the ~130 dependent
    // uniform-datapath instructions of a naive issue loop (descriptor arithmetic, two barrier polls) took about as long
    // as the MMAs themselves (ncu: tensor pipe 71 % busy, the issuer stalled on fixed latencies, never on the MMA queue).
    // Descriptors are therefore built once and stepped by adding (bytes >> 4) to the start-address field, and only the
    // splitter's barrier is polled -- it completes after the stage's TMA barrier, which the splitter waited for.
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kSubM >> 4) << 24);
      const uint64_t d0 = umma_desc_sw64(smem_base);                       // stage 0, byte 0
      const uint64_t d_step = (uint64_t)(stage_bytes >> 4);
      const uint64_t o_xl = (uint64_t)(kABytes >> 4), o_wh = (uint64_t)((2 * kABytes) >> 4),
                     o_wl = (uint64_t)((2 * kABytes + b_bytes) >> 4);
      constexpr uint64_t o_k = (uint64_t)((kUmmaK * 4) >> 4), o_sub = (uint64_t)(kSubBytes >> 4);
      uint64_t dst = d0;
      int s = 0, i = 0;
      uint32_t ph = 0;
      long long w_split = 0, w_acc = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
        const int buf = i & 1;
        long long c0 = trace ? clock64() : 0;
        mbar_wait(bar_acc_empty(buf), (((uint32_t)i >> 1) & 1u) ^ 1u);      // the epilogue has drained this buffer
        if (trace) w_acc += clock64() - c0;
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)buf * acc_cols;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (trace) c0 = clock64();
          mbar_wait(bar_split(s), ph);
          if (trace) w_split += clock64() - c0;
          tc_fence_after();
          if (!(diag_on(p) & 4)) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t wh = dst + o_wh + k * o_k, wl = dst + o_wl + k * o_k;
#pragma unroll
              for (int h = 0; h < MSUB; ++h) {
                const uint64_t xh = dst + h * o_sub + k * o_k, xl = xh + o_xl;
                const uint32_t acc = acc0 + (uint32_t)(h * BN);
                umma_tf32(acc, xl, wh, idesc, (kb | k) ? 1u : 0u);     // small terms first
                umma_tf32(acc, xh, wl, idesc, 1u);
                umma_tf32(acc, xh, wh, idesc, 1u);
              }
            }
          }
          umma_commit(bar_empty(s));
          if (kb == num_kb - 1) umma_commit(bar_acc_full(buf));
          if (++s == stages) { s = 0; ph ^= 1u; dst = d0; } else dst += d_step;
        }
      }
      if (trace) { trace[1] = 0; trace[2] = w_split; trace[3] = w_acc; trace[4] = clock64(); }
    }
  } else if (warp < 2 + 4 * p.split_groups) {
    // ===================== splitter =====================
    const int t4 = (tid - 64) & 127;                 // 0..127 within the group
    const int grp = (warp - 2) >> 2, ngrp = p.split_groups;   // group g takes the K blocks with running index % G == g
    int s = 0, jb = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb, ++jb) {
        if (ngrp == 2 && (jb & 1) != grp) {
          if (++s == stages) { s = 0; ph ^= 1u; }
          continue;
        }
        mbar_wait(bar_full(s), ph);
        const uint32_t xa = stage_x(s) + (uint32_t)t4 * 16u, la = stage_xl(s) + (uint32_t)t4 * 16u;
#pragma unroll
        for (int c = 0; c < (int)(kABytes / (128 * 16)); ++c) {
          if (diag_on(p) & 2) break;
          uint32_t v0, v1, v2, v3, h0, h1, h2, h3, l0, l1, l2, l3;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(xa + 2048u * c));
          split_tf32(v0, h0, l0); split_tf32(v1, h1, l1); split_tf32(v2, h2, l2); split_tf32(v3, h3, l3);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(xa + 2048u * c), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(la + 2048u * c), "r"(l0), "r"(l1), "r"(l2), "r"(l3) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_split(s));
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                          // TMEM lane quadrant this warp may read
    const uint32_t slab0 = slab_base + (uint32_t)(warp - 2 - 4 * p.split_groups) * (uint32_t)(kPSlabs * 4096);
    const int chunks = BN / 32;
    int it = 0, i = 0;
    long long w_ready = 0, busy = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
      const int buf = i & 1;
      const int m0 = (t / n_tiles) * kBlockM, n0 = (t % n_tiles) * BN;
      long long c0 = trace ? clock64() : 0;
      mbar_wait(bar_acc_full(buf), ((uint32_t)i >> 1) & 1u);
      long long c1 = trace ? clock64() : 0;
      tc_fence_after();
      if (diag_on(p) & 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty(buf));
        continue;
      }
#pragma unroll 1
      for (int h = 0; h < MSUB; ++h) {
        const int row0 = m0 + h * kSubM + 32 * q;
        const bool zero_row = p.row_zero != nullptr && row0 + lane < p.M && p.row_zero[row0 + lane] != 0;
#pragma unroll 1
        for (int c0i = 0; c0i < chunks; c0i += kPBatch, it += kPBatch) {
          const int qn = chunks - c0i < kPBatch ? chunks - c0i : kPBatch;
          if (it >= kPSlabs) {                       // the batch that last used these slabs must be done reading them
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPSlabs / kPBatch - 1) : "memory");
            __syncwarp();
          }
#pragma unroll 1
          for (int g = 0; g < qn; ++g) {
            const int c = c0i + g;
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)buf * acc_cols + (uint32_t)(h * BN + c * 32), v);
            const uint32_t slab = slab0 + (uint32_t)((it + g) % kPSlabs) * 4096u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bj = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c * 32 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
              float f[4] = {__fadd_rn(__uint_as_float(v[4 * j + 0]), bj.x), __fadd_rn(__uint_as_float(v[4 * j + 1]), bj.y),
                            __fadd_rn(__uint_as_float(v[4 * j + 2]), bj.z), __fadd_rn(__uint_as_float(v[4 * j + 3]), bj.w)};
              if (p.relu) {
#pragma unroll
                for (int e = 0; e < 4; ++e) f[e] = f[e] < 0.0f ? 0.0f : f[e];
              }
              if (zero_row) f[0] = f[1] = f[2] = f[3] = 0.0f;
              const uint32_t dst = slab + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) * 16);
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]) : "memory");
            }
          }
          if (h == MSUB - 1 && c0i + kPBatch >= chunks) {   // the accumulator is in registers / shared memory: hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(buf));
          }
          if (diag_on(p) & 16) continue;
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            for (int g = 0; g < qn; ++g)
              tma_store_2d(&tm_y, slab0 + (uint32_t)((it + g) % kPSlabs) * 4096u, n0 + (c0i + g) * 32, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      if (trace) { w_ready += c1 - c0; busy += clock64() - c1; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    if (trace && (warp & 3) == 2 && lane == 0) { trace[5] = clock64(); trace[6] = w_ready; trace[7] = busy; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// w [N][K] -> TF32 head / remainder in K-BLOCKED order [K/16][N][16], so the B tile of one K block is one contiguous
// BN x 64-byte chunk
__global__ void split_weight_kernel(const float* __restrict__ w, float* __restrict__ wh, float* __restrict__ wl, int N, int K) {
  const long long n = (long long)N * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / K), col = (int)(i % K);
    uint32_t h, l;
    split_tf32(__float_as_uint(w[i]), h, l);
    // position inside the 64-byte row of the K block, with the 16-byte chunk index XOR-swizzled exactly as the TMA /
    // UMMA 64-byte swizzle places it in shared memory (chunk ^= (row >> 1) & 3; tiles start at multiples of 8 rows),
    // so a K block of any N tile is ONE contiguous bulk copy instead of BN 64-byte TMA box rows
    const int c = col % kBlockK;
    const int chunk = (c >> 2) ^ ((row >> 1) & 3);
    const long long o = ((long long)(col / kBlockK) * N + row) * kBlockK + chunk * 4 + (c & 3);
    wh[o] = __uint_as_float(h);
    wl[o] = __uint_as_float(l);
  }
}

// ---- host side -------------------------------------------------------------------------------------
// 2-D fp32 row-major tensor [rows, cols] with a row pitch in floats; box = box_cols x box_rows
int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long pitch_floats, int box_cols,
             int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = tensor_map_encode_fn();
  if (!fn) return MSDA_E_NOCUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch_floats * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fprintf(stderr, "msda_b200: cuTensorMapEncodeTiled failed (%d): base %p rows %lld cols %lld pitch %lld box %d x %d swizzle %d\n",
            (int)r, base, rows, cols, pitch_floats, box_cols, box_rows, (int)swz);
  return r == CUDA_SUCCESS ? 0 : MSDA_E_DIMS;
}

}  // namespace
}  // namespace msda

namespace { long long* g_trace = nullptr; }

extern "C" {

// diagnostics: per-CTA clock64 stamps of the next launches go to buf ([grid][8] int64 on the device); NULL disables
void msda_b200_linear_set_trace(long long* buf) { g_trace = buf; }

int msda_b200_linear_split_weight_f32(const float* w, int N, int K, float* w_hi, float* w_lo, void* stream) {
  if (!w || !w_hi || !w_lo) return MSDA_E_NULLPTR;
  if (N <= 0 || K <= 0) return MSDA_E_DIMS;
  const long long n = (long long)N * K;
  const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  msda::split_weight_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, w_hi, w_lo, N, K);
  return (int)cudaGetLastError();
}

static int linear_impl(const float* x, int ldx, const float* w_hi, const float* w_lo, const float* bias,
                       const unsigned char* row_zero, int M, int N, int K, float* y, int ldy, void* stream, int relu) {
  using namespace msda;
  if (!x || !w_hi || !w_lo || !y) return MSDA_E_NULLPTR;
  if (M <= 0 || N <= 0 || K <= 0) return MSDA_E_DIMS;
  if (K % kBlockK != 0 || N % 32 != 0 || N > 1024) return MSDA_E_UNSUPPORTED;
  if (ldx < K || ldy < N || (ldx & 3) || (ldy & 3)) return MSDA_E_DIMS;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_hi) | reinterpret_cast<uintptr_t>(w_lo) |
       reinterpret_cast<uintptr_t>(y)) & 15u)
    return MSDA_E_ALIGN;
  const int sms = msda_b200_sm_count();
  if (sms < 0) return sms;
  const char* pe = getenv("MSDA_GEMM_PERSISTENT");            // 0: the tile-per-CTA kernels below (A/B measurements)
  if (!pe || atoi(pe) != 0) {
    if (bias && (reinterpret_cast<uintptr_t>(bias) & 15u)) return MSDA_E_ALIGN;      // the epilogue reads it as float4
    // Tile = 128 rows x BN columns, BN a multiple of 32 that divides N, 2 x BN <= 512 TMEM columns.  The widest tile
    // is the cheapest per flop (a K block costs ~4 BN + 90 clk of MMAs but never less than the ~450 clk the
    // TMA -> split -> MMA chain needs, and the weights are re-read once per tile), so take it whenever it still makes more
    // tiles than SMs; otherwise everything fits one round and the narrowest tile that still does spreads it over the
    // most SMs (tools/gemm_sweep.py, profiles/r02_gemm_persistent.txt).
    const long long row_tiles = (M + 127) / 128;
    int BN = 0;
    const char* forced = getenv("MSDA_GEMM_BN");               // measurements and tests only
    if (forced && atoi(forced) >= 32 && atoi(forced) <= 256 && atoi(forced) % 32 == 0 && N % atoi(forced) == 0) {
      BN = atoi(forced);
    } else {
      for (int bn = 32; bn <= 256 && bn <= N; bn += 32)
        if (N % bn == 0) BN = bn;                                // widest
      if (row_tiles * (N / BN) <= sms)
        for (int bn = 32; bn < BN; bn += 32)
          if (N % bn == 0 && row_tiles * (N / bn) <= sms) { BN = bn; break; }
    }
    int tmem_cols = 32;
    while (tmem_cols < 2 * BN) tmem_cols <<= 1;
    const size_t stage_bytes = 2 * (size_t)kSubBytes + 2 * (size_t)BN * kBlockK * 4;
    const size_t slabs = (size_t)4 * kPSlabs * 4096;
    int stages = (int)((226 * 1024 - 1024 - slabs) / stage_bytes);
    if (stages > kPMaxStages) stages = kPMaxStages;
    const size_t smem = stages * stage_bytes + slabs + 1024;
    CUtensorMap tm_x, tm_y;
    int rc;
    if ((rc = make_map(&tm_x, x, M, K, ldx, kBlockK, 128, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_map(&tm_y, y, M, N, ldy, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    GemmParams p;
    p.trace = g_trace;
    p.w_hi = w_hi; p.w_lo = w_lo; p.y = y; p.ldy = ldy; p.bias = bias; p.row_zero = row_zero; p.M = M; p.N = N; p.K = K;
    p.block_n = BN; p.tmem_cols = tmem_cols; p.relu = relu; p.stages = stages;
    p.split_groups = row_tiles * (N / BN) <= sms ? 2 : 1;
    { const char* g = getenv("MSDA_GEMM_SPLIT_GROUPS"); if (g && (atoi(g) == 1 || atoi(g) == 2)) p.split_groups = atoi(g); }
    { const char* d = getenv("MSDA_GEMM_DIAG"); p.diag = d ? atoi(d) : 0; }
    static msda::PerDeviceOnce configured_p;
    if (configured_p.need())
      cudaFuncSetAttribute(proj_gemm_3xtf32_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    const long long tiles = row_tiles * (N / BN);
    const int grid = (int)(tiles < sms ? tiles : sms);
    proj_gemm_3xtf32_persistent_kernel<<<grid, 64 + 128 * p.split_groups + 128, smem, (cudaStream_t)stream>>>(tm_x, tm_y, p);
    return (int)cudaGetLastError();
  }

  // N tiles of at most 256 columns, all equal and a multiple of 32 (384 -> 2 x 192)
  int n_tiles = (N + 255) / 256;
  while (N % n_tiles != 0 || (N / n_tiles) % 32 != 0) ++n_tiles;
  // Small M (one 720p frame = 150 row tiles for 148 SMs): with one CTA per SM nothing overlaps that CTA's prologue and
  // epilogue, and the 2 left-over tiles make a second wave.  Halve the N tile (down to 64 columns) until two CTAs per SM
  // are resident, so one's epilogue runs under the other's main loop; the x tile is then read twice, from L2.
  static const int split_n = []() { const char* e = getenv("MSDA_GEMM_SPLIT_N"); return e ? atoi(e) : 1; }();
  if (split_n) {
    const long long tiles_m = (M + 127) / 128;
    while (tiles_m * n_tiles < 2LL * sms && N % (2 * n_tiles) == 0 && (N / (2 * n_tiles)) % 32 == 0 && N / (2 * n_tiles) >= 64)
      n_tiles *= 2;
  }
  const int BN = N / n_tiles;

  // two 128-row sub-tiles per CTA when there are enough rows to fill the machine that way
  const bool two = (long long)((M + 255) / 256) * n_tiles >= 2LL * sms && 2 * BN <= 512;
  const int block_m = two ? 256 : 128;
  int tmem_cols = 32;
  while (tmem_cols < (two ? 2 : 1) * BN) tmem_cols <<= 1;

  CUtensorMap tm_x, tm_y;
  int rc;
  if ((rc = make_map(&tm_x, x, M, K, ldx, kBlockK, block_m, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;

  if ((rc = make_map(&tm_y, y, M, N, ldy, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;

  GemmParams p;
  p.trace = g_trace;
  p.w_hi = w_hi; p.w_lo = w_lo; p.y = y; p.ldy = ldy; p.bias = bias; p.row_zero = row_zero; p.M = M; p.N = N; p.K = K; p.block_n = BN; p.tmem_cols = tmem_cols; p.relu = relu; p.stages = 0; p.diag = 0; p.split_groups = 1;
  const int stages = two ? 3 : 2;
  const size_t stage_bytes = 2 * (size_t)kSubBytes * (two ? 2 : 1) + 2 * (size_t)BN * kBlockK * 4;
  size_t smem = stages * stage_bytes;
  const size_t stage_need = (size_t)4 * (two ? 8 : 4) * 4096;   // epilogue staging: 4 warps x G slabs x 4 KB
  if (smem < stage_need) smem = stage_need;
  smem += 1024;                                  // alignment slack
  static msda::PerDeviceOnce configured;   // function attributes are per device
  if (configured.need()) {
    cudaFuncSetAttribute(proj_gemm_3xtf32_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    cudaFuncSetAttribute(proj_gemm_3xtf32_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  dim3 grid((M + block_m - 1) / block_m, n_tiles);
  if (two)
    proj_gemm_3xtf32_kernel<2, 3><<<grid, kThreads, smem, (cudaStream_t)stream>>>(tm_x, tm_y, p);
  else
    proj_gemm_3xtf32_kernel<1, 2><<<grid, kThreads, smem, (cudaStream_t)stream>>>(tm_x, tm_y, p);
  return (int)cudaGetLastError();
}

int msda_b200_linear_f32(const float* x, int ldx, const float* w_hi, const float* w_lo, const float* bias,
                         const unsigned char* row_zero, int M, int N, int K, float* y, int ldy, void* stream) {
  return linear_impl(x, ldx, w_hi, w_lo, bias, row_zero, M, N, K, y, ldy, stream, 0);
}

int msda_b200_linear_relu_f32(const float* x, int ldx, const float* w_hi, const float* w_lo, const float* bias,
                              int M, int N, int K, float* y, int ldy, void* stream) {
  return linear_impl(x, ldx, w_hi, w_lo, bias, nullptr, M, N, K, y, ldy, stream, 1);
}

}  // extern "C"
