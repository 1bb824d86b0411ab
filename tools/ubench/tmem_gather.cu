// Microbenchmark 6: is tensor memory a second gather path beside the L1 / shared-memory data pipe?
// Layout under test: TMEM lane = channel (32 lanes of a warp's sub-partition = the 32 channels of one head),
// TMEM column = pixel.  One tcgen05.ld.32x32b.x1 with a dynamic, warp-uniform column then returns one 128-byte
// "row" (pixel, head) with channel c on lane c — the same lane layout as the register-gather kernel's rows.
// Measures, per SM: (A) LDTM rows/clk for x1/x2/x4 at 4..32 warps, random columns; (B) STTM rows/clk (the fill
// path from registers); (C) LDTM and LDS.128 row gathers running side by side (shared pipe or not);
// (D) tcgen05.cp smem -> TMEM fill rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_gather tmem_gather.cu && ./tmem_gather
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int ITERS = 2048;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int XN>
__device__ __forceinline__ void ldtm(uint32_t taddr, uint32_t (&v)[4]) {
  if (XN == 1) asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(taddr));
  else if (XN == 2) asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr));
  else asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
template <int XN>
__device__ __forceinline__ void sttm(uint32_t taddr, const uint32_t (&v)[4]) {
  if (XN == 1) asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v[0]));
  else if (XN == 2) asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(v[0]), "r"(v[1]));
  else asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]));
}

// MODE 0: LDTM gather, 1: STTM fill, 2: mix (odd warps LDS.128 row gather, even warps LDTM), 3: LDS only on the odd
// warps (even warps idle: the reference point for the mix), 4: LDTM only on the even warps
template <int MODE, int XN, int DEPTH>
__global__ void __launch_bounds__(1024) k_tmem(long long* cyc, float* out, int* check) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint32_t s_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (MODE >= 2) {
    for (int i = threadIdx.x * 16; i < 128 * 128; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(i, i + 1, i + 2, i + 3);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = s_base + ((uint32_t)(warp & 3) * 32u << 16);
  // fill: column c of lane l holds c * 32 + l (so a gather can be checked)
  if (warp < 4) {
    for (int c = 0; c < 512; ++c) {
      uint32_t v[4] = {(uint32_t)(c * 32 + lane), 0, 0, 0};
      sttm<1>(tbase + c, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  uint32_t s = (blockIdx.x * 32 + warp) * 2654435761u + 12345u;
  uint32_t acc = 0;
  int bad = 0;
  const uint32_t sbase = smem_u32(smem);
  const bool lds_warp = (MODE == 2 || MODE == 3) && (warp & 1);
  const bool idle = (MODE == 3 && !(warp & 1)) || (MODE == 4 && (warp & 1));
  const long long t0 = clock64();
  if (!idle) {
#pragma unroll 1
    for (int it = 0; it < ITERS; it += DEPTH) {
      if (lds_warp) {
        uint32_t v[DEPTH][4];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
          s = s * 1664525u + 1013904223u;
          const uint32_t r = ((s >> 9) + (lane >> 3) * 37u) & 127u;      // 4 random rows per instruction
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3]) : "r"(sbase + r * 128u + (lane & 7) * 16u));
        }
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) acc += v[u][0] ^ v[u][1] ^ v[u][2] ^ v[u][3];
      } else if (MODE == 1) {
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
          s = s * 1664525u + 1013904223u;
          const uint32_t c = (s >> 9) % (512u - XN + 1);
          uint32_t v[4] = {s, s + 1, s + 2, s + 3};
          sttm<XN>(tbase + c, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else {
        uint32_t v[DEPTH][4];
        uint32_t cs[DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
          s = s * 1664525u + 1013904223u;
          cs[u] = (s >> 9) % (512u - XN + 1);
          ldtm<XN>(tbase + cs[u], v[u]);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
#pragma unroll
          for (int x = 0; x < XN; ++x) {
            acc += v[u][x];
            if (MODE == 0 && it == 0) bad += v[u][x] != (cs[u] + x) * 32 + lane;
          }
        }
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) out[0] = 1.f;
  if (bad) atomicAdd(check, bad);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "r"(512u) : "memory");
}

// (D) tcgen05.cp: one thread copies 128 lanes x 256 bits (8 columns) per instruction from shared memory; NCP
// instructions fill all 512 columns; commit to an mbarrier, wait, repeat.
__global__ void __launch_bounds__(128) k_cp(long long* cyc, int reps) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint32_t s_base;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x * 16; i < 64 * 1024; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(i, i + 1, i + 2, i + 3);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r) {
      for (int i = 0; i < 64; ++i) {
        // no-swizzle K-major descriptor: 128 rows x 32 bytes; core matrices (8 rows x 16 B = 128 B) contiguous,
        // LBO (next 16-byte column chunk) = 128 B * 16 row groups = 2048, SBO (next 8-row group) = 128
        const uint32_t saddr = smem_u32(smem) + (uint32_t)(i & 15) * 4096u;
        uint64_t d = 0;
        d |= (uint64_t)((saddr & 0x3ffff) >> 4);
        d |= (uint64_t)(2048 >> 4) << 16;
        d |= (uint64_t)(128 >> 4) << 32;
        d |= (uint64_t)1 << 46;
        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(s_base + (uint32_t)i * 8u), "l"(d) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t done = 0;
      while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
      }
      phase ^= 1;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "r"(512u) : "memory");
}

static long long* d_cyc;
static float* d_out;
static int* d_check;
static int g_sms, g_khz;

template <int MODE, int XN, int DEPTH>
void run(const char* name, int warps) {
  cudaFuncSetAttribute(k_tmem<MODE, XN, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  cudaMemset(d_check, 0, 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_tmem<MODE, XN, DEPTH><<<g_sms, warps * 32, 16 * 1024>>>(d_cyc, d_out, d_check);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k_tmem<MODE, XN, DEPTH><<<g_sms, warps * 32, 16 * 1024>>>(d_cyc, d_out, d_check);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  std::vector<long long> h(g_sms);
  cudaMemcpy(h.data(), d_cyc, g_sms * sizeof(long long), cudaMemcpyDeviceToHost);
  int bad = 0; cudaMemcpy(&bad, d_check, 4, cudaMemcpyDeviceToHost);
  double mean = 0; for (auto c : h) mean += (double)c; mean /= g_sms;
  double tm_warps = warps, lds_warps = 0;
  if (MODE == 2) { tm_warps = (warps + 1) / 2; lds_warps = warps / 2; }
  if (MODE == 3) { tm_warps = 0; lds_warps = warps / 2; }
  if (MODE == 4) { tm_warps = (warps + 1) / 2; }
  const double tm_rows = tm_warps * ITERS * XN, lds_rows = lds_warps * ITERS * 4;
  printf("%-40s warps/SM=%2d depth=%2d : TMEM %6.3f rows/clk/SM  LDS %6.3f rows/clk/SM (clock64, %8.0f clk; events %.1f us) mismatches=%d %s\n",
         name, warps, DEPTH, tm_rows / mean, lds_rows / mean, mean, ms * 1e3, bad, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&g_khz, cudaDevAttrClockRate, 0);
  cudaMalloc(&d_out, 4); cudaMalloc(&d_check, 4); cudaMalloc(&d_cyc, sizeof(long long) * g_sms);
  for (int w : {4, 8, 16, 32}) {
    if (w == 4) {
      run<0, 1, 8>("LDTM 32x32b.x1 (1 row/instr)", 4);
      run<0, 2, 8>("LDTM 32x32b.x2 (2 adjacent cols)", 4);
      run<0, 4, 8>("LDTM 32x32b.x4", 4);
      run<0, 1, 16>("LDTM 32x32b.x1", 4);
      run<1, 1, 8>("STTM 32x32b.x1", 4);
      run<1, 4, 8>("STTM 32x32b.x4", 4);
    } else if (w == 8) {
      run<0, 1, 8>("LDTM 32x32b.x1", 8);
      run<0, 2, 8>("LDTM 32x32b.x2", 8);
      run<0, 4, 8>("LDTM 32x32b.x4", 8);
      run<1, 1, 8>("STTM 32x32b.x1", 8);
    } else if (w == 16) {
      run<0, 1, 8>("LDTM 32x32b.x1", 16);
      run<0, 2, 8>("LDTM 32x32b.x2", 16);
      run<0, 4, 8>("LDTM 32x32b.x4", 16);
      run<0, 1, 16>("LDTM 32x32b.x1", 16);
      run<1, 1, 8>("STTM 32x32b.x1", 16);
      run<1, 4, 8>("STTM 32x32b.x4", 16);
      run<3, 1, 8>("LDS.128 alone on odd warps", 16);
      run<4, 1, 8>("LDTM.x1 alone on even warps", 16);
      run<2, 1, 8>("mix: odd LDS.128, even LDTM.x1", 16);
      run<4, 2, 8>("LDTM.x2 alone on even warps", 16);
      run<2, 2, 8>("mix: odd LDS.128, even LDTM.x2", 16);
    } else {
      run<0, 1, 8>("LDTM 32x32b.x1", 32);
      run<0, 2, 8>("LDTM 32x32b.x2", 32);
      run<0, 4, 8>("LDTM 32x32b.x4", 32);
      run<3, 1, 8>("LDS.128 alone on odd warps", 32);
      run<4, 1, 8>("LDTM.x1 alone on even warps", 32);
      run<2, 1, 8>("mix: odd LDS.128, even LDTM.x1", 32);
      run<2, 2, 8>("mix: odd LDS.128, even LDTM.x2", 32);
      run<2, 4, 8>("mix: odd LDS.128, even LDTM.x4", 32);
    }
  }
  // (D) smem -> TMEM copy engine
  {
    cudaFuncSetAttribute(k_cp, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int reps = 64;
    k_cp<<<g_sms, 128, 64 * 1024>>>(d_cyc, reps);
    cudaDeviceSynchronize();
    k_cp<<<g_sms, 128, 64 * 1024>>>(d_cyc, reps);
    cudaDeviceSynchronize();
    std::vector<long long> h(g_sms);
    cudaMemcpy(h.data(), d_cyc, g_sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0; for (auto c : h) mean += (double)c; mean /= g_sms;
    const double bytes = (double)reps * 64 * 128 * 32;
    printf("tcgen05.cp 128x256b smem->TMEM: %.1f B/clk/SM (%.3f 128-B rows/clk/SM), %.0f clk per 256 KB fill  %s\n", bytes / mean,
           bytes / 128 / mean, mean / reps, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
