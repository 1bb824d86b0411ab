// Microbenchmark 3: peak rate of gathering random 128-byte rows from (a) an L1-resident global table and (b) shared
// memory, per access width, with minimal ALU overhead (one LCG step yields two row indices).  Reports rows/clk/SM
// from clock64() AND from CUDA events at the current SM clock, so a broken cycle count shows up.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o row_rate row_rate.cu && ./row_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <algorithm>

constexpr int ROW = 128;
constexpr int ITERS = 4096;

// SPACE 0 global (L1-resident table), 1 shared.  W = bytes per lane (4, 8, 16, 32; 32 global only).
// W = 17: LDSM.x4 (lanes 8j..8j+7 address the eight 16-byte chunks of row j), W = 18: LDSM.x2, W = 19: LDSM.x1
// PAT 0: every lane group picks its own random row; 1: "bilinear": the instruction's row groups are the corners
// (r, r+1, r+WIN, r+WIN+1, ...) of one random sample position (2 or 4 or 8 rows per instruction)
template <int SPACE, int W, int PAT>
__global__ void __launch_bounds__(1024) k_rows(const char* __restrict__ table, uint32_t row_mask, long long* cyc, float* out) {
  extern __shared__ __align__(128) char smem[];
  const int lane = threadIdx.x & 31;
  constexpr int WB = (W >= 17 && W <= 19) ? 16 : W;            // bytes per lane
  constexpr int LPR = ROW / WB;                    // lanes per row
  constexpr int RPI = W == 18 ? 2 : W == 19 ? 1 : 32 / LPR;   // rows per instruction
  if (SPACE == 1) {
    for (int i = threadIdx.x * 16; i < (int)(row_mask + 1) * ROW; i += blockDim.x * 16)
      *reinterpret_cast<uint4*>(smem + i) = make_uint4(i, i + 1, i + 2, i + 3);
  }
  __syncthreads();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t tid_g = blockIdx.x * 1024 + threadIdx.x;
  const int grp = (lane / LPR) % RPI;
  // PAT 0: seed per lane group; PAT 1: seed per warp (all groups derive their row from the same sample position)
  uint32_t s = (PAT == 0 ? (tid_g / LPR) : (tid_g >> 5)) * 2654435761u + 12345u;
  const uint32_t corner = PAT == 1 ? ((grp & 1) + ((grp >> 1) & 1) * 24 + (grp >> 2) * 7) : 0;   // WIN = 24 rows per image row
  const uint32_t lane_off = (lane % LPR) * WB;
  float acc = 0.f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it += 4) {
    uint32_t v[4][2];
#pragma unroll
    for (int u = 0; u < 4; u += 2) {
      s = s * 1664525u + 1013904223u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t r = (((h ? (s >> 19) : (s >> 6)) + corner) & row_mask);
        const uint32_t off = r * ROW + lane_off;
        uint32_t a = 0, b = 0, c = 0, d = 0, e = 0, f = 0, g = 0, hh = 0;
        if (SPACE == 0) {
          const char* p = table + off;
          if (W == 4) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(a) : "l"(p));
          else if (W == 8) asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
          else if (W == 16) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(c), "=r"(d), "=r"(b) : "l"(p));
          else asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(hh), "=r"(b) : "l"(p));
        } else {
          const uint32_t p = sbase + off;
          if (W == 4) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(p));
          else if (W == 8) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(p));
          else if (W == 16) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(c), "=r"(d), "=r"(b) : "r"(p));
          else if (W == 17) asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(c), "=r"(d), "=r"(b) : "r"(p));
          else if (W == 18) asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(p));
          else asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];" : "=r"(a) : "r"(p));
        }
        v[u + h][0] = a ^ c ^ d; v[u + h][1] = b ^ e ^ f ^ g ^ hh;   // every loaded word is consumed (ptxas narrows partly-used vector loads)
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += __uint_as_float(v[u][0]) + __uint_as_float(v[u][1]);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) out[0] = acc;
}

static long long* d_cyc;
static float* d_out;
static char* d_table;
static int g_sms, g_khz;

template <int SPACE, int W, int PAT>
void run(const char* name, uint32_t rows, int bps) {
  constexpr int WB = (W >= 17 && W <= 19) ? 16 : W;
  constexpr int RPI = W == 18 ? 2 : W == 19 ? 1 : 32 / (ROW / WB);
  const int blocks = g_sms * bps;   // 1024-thread CTAs: 1 or 2 per SM, so the block scheduler cannot spread them unevenly
  const size_t smem = SPACE == 1 ? (size_t)rows * ROW : 0;
  cudaFuncSetAttribute(k_rows<SPACE, W, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_rows<SPACE, W, PAT><<<blocks, 1024, smem>>>(d_table, rows - 1, d_cyc, d_out);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k_rows<SPACE, W, PAT><<<blocks, 1024, smem>>>(d_table, rows - 1, d_cyc, d_out);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  std::vector<long long> h(blocks);
  cudaMemcpy(h.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0; for (auto c : h) mean += (double)c; mean /= blocks;
  const double rows_per_block = 32.0 * ITERS * RPI;
  const double rpc_clk = rows_per_block * bps / mean;
  const double rpc_evt = rows_per_block * bps / (ms * 1e-3 * g_khz * 1e3);
  printf("%-46s rows=%4u warps/SM=%2d : %6.3f rows/clk/SM by clock64 (%6.1f B/clk) | %6.3f by events@%dMHz | %5.2f clk/instr  %s\n",
         name, rows, bps * 32, rpc_clk, rpc_clk * ROW, rpc_evt, g_khz / 1000, RPI / rpc_clk, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&g_khz, cudaDevAttrClockRate, 0);
  cudaMalloc(&d_table, 1 << 20); cudaMemset(d_table, 0, 1 << 20);
  cudaMalloc(&d_out, 4); cudaMalloc(&d_cyc, sizeof(long long) * g_sms * 8);
  for (int bps : {1, 2}) {
    {
      run<0, 4, 0>("LDG.32   L1, 1 row/instr", 256, bps);
      run<0, 8, 0>("LDG.64   L1, 2 rows/instr", 256, bps);
      run<0, 16, 0>("LDG.128  L1, 4 rows/instr", 256, bps);
      run<0, 32, 0>("LDG.256  L1, 8 rows/instr", 256, bps);
      run<0, 16, 1>("LDG.128  L1, 4 corner rows of one sample", 256, bps);
      run<0, 32, 1>("LDG.256  L1, 8 rows = corners of 2 samples", 256, bps);
    }
    run<1, 4, 0>("LDS.32   1 row/instr", 128, bps);
    run<1, 8, 0>("LDS.64   2 rows/instr", 128, bps);
    run<1, 16, 0>("LDS.128  4 rows/instr", 128, bps);
    run<1, 17, 0>("LDSM.x4  4 rows/instr", 128, bps);
    run<1, 18, 0>("LDSM.x2  2 rows/instr", 128, bps);
    run<1, 19, 0>("LDSM.x1  1 row/instr", 128, bps);
    run<1, 8, 1>("LDS.64   2 adjacent rows (r, r+1)", 128, bps);
    run<1, 16, 1>("LDS.128  4 corner rows of one sample", 128, bps);
    run<1, 17, 1>("LDSM.x4  4 corner rows of one sample", 128, bps);
  }
  return 0;
}
