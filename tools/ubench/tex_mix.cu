// Microbenchmark 4: can the TEX data pipe add gather bandwidth on top of the LSU pipe's one 128-byte row per clock?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tex_mix tex_mix.cu && ./tex_mix
// Every iteration a warp gathers NL x 4 random 128-byte rows with LDG.128 (8 lanes per row) and NT x 4 rows with
// tex1Dfetch<float4> (same lane layout) from the same L1-resident table.  If the two pipes were independent the mixed
// cases would exceed the LDG-only row rate.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int ROW = 128;
constexpr int ITERS = 2048;

template <int NL, int NT>
__global__ void __launch_bounds__(1024) k_mix(const char* __restrict__ table, cudaTextureObject_t tex, uint32_t row_mask,
                                              long long* cyc, float* out) {
  const int lane = threadIdx.x & 31;
  const uint32_t tid_g = blockIdx.x * 1024 + threadIdx.x;
  uint32_t s = (tid_g >> 3) * 2654435761u + 12345u;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    uint4 v[NL > 0 ? NL : 1];
    float4 t[NT > 0 ? NT : 1];
#pragma unroll
    for (int u = 0; u < NL; ++u) {
      s = s * 1664525u + 1013904223u;
      const char* p = table + (size_t)((s >> 9) & row_mask) * ROW + (lane & 7) * 16;
      asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p));
    }
#pragma unroll
    for (int u = 0; u < NT; ++u) {
      s = s * 1664525u + 1013904223u;
      t[u] = tex1Dfetch<float4>(tex, (int)(((s >> 9) & row_mask) * 8 + (lane & 7)));
    }
#pragma unroll
    for (int u = 0; u < NL; ++u) acc += (__uint_as_float(v[u].x) + __uint_as_float(v[u].y)) + (__uint_as_float(v[u].z) + __uint_as_float(v[u].w));
#pragma unroll
    for (int u = 0; u < NT; ++u) acc += (t[u].x + t[u].y) + (t[u].z + t[u].w);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) out[0] = acc;
}

static int g_sms, g_khz;
static long long* d_cyc;
static float* d_out;

template <int NL, int NT>
void run(const char* name, const char* table, cudaTextureObject_t tex, uint32_t rows, int bps) {
  const int blocks = g_sms * bps;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_mix<NL, NT><<<blocks, 1024>>>(table, tex, rows - 1, d_cyc, d_out);
  cudaEventRecord(a);
  k_mix<NL, NT><<<blocks, 1024>>>(table, tex, rows - 1, d_cyc, d_out);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  std::vector<long long> h(blocks);
  cudaMemcpy(h.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0; for (auto c : h) mean += (double)c; mean /= blocks;
  const double rows_per_block = 32.0 * 4 * (NL + NT) * ITERS;   // 32 warps x 4 rows per instruction
  const double rpc = rows_per_block * bps / mean;
  const double rpc_ev = rows_per_block * blocks / g_sms / (ms * 1e-3 * g_khz * 1e3);
  printf("%-34s rows=%4u warps/SM=%2d : %6.3f rows/clk/SM by clock64 | %6.3f by events@%dMHz  %s\n", name, rows, bps * 32, rpc,
         rpc_ev, g_khz / 1000, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&g_khz, cudaDevAttrClockRate, 0);
  char* table;
  const uint32_t rows = 256;   // 32 KB: L1-resident
  cudaMalloc(&table, 1 << 20); cudaMemset(table, 0, 1 << 20);
  cudaMalloc(&d_out, 4); cudaMalloc(&d_cyc, sizeof(long long) * g_sms * 8);
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypeLinear;
  rd.res.linear.devPtr = table;
  rd.res.linear.desc = cudaCreateChannelDesc<float4>();
  rd.res.linear.sizeInBytes = 1 << 20;
  cudaTextureDesc td = {};
  td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex = 0;
  cudaError_t e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
  printf("texture object: %s\n", cudaGetErrorString(e));
  for (int bps : {1, 2}) {
    run<4, 0>("4 LDG.128 + 0 TEX", table, tex, rows, bps);
    run<0, 4>("0 LDG.128 + 4 TEX", table, tex, rows, bps);
    run<3, 1>("3 LDG.128 + 1 TEX", table, tex, rows, bps);
    run<2, 2>("2 LDG.128 + 2 TEX", table, tex, rows, bps);
    run<4, 1>("4 LDG.128 + 1 TEX", table, tex, rows, bps);
    run<4, 2>("4 LDG.128 + 2 TEX", table, tex, rows, bps);
    run<8, 0>("8 LDG.128 + 0 TEX", table, tex, rows, bps);
    run<6, 2>("6 LDG.128 + 2 TEX", table, tex, rows, bps);
  }
  return 0;
}
