// Microbenchmark 2: what limits a gather whose rows mostly hit L1 but sometimes miss to L2?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_mix gather_mix.cu && ./gather_mix
// Every test reports 128-byte rows per clock per SM, measured with clock64() inside the kernel (independent of
// the SM clock).  1024 threads per SM (4 CTAs x 256), like the MSDeformAttn kernels.
//   mix     LDG.256 (4 lanes/row, 8 rows/instr): each row comes from a COLD table (64 MB, L2-resident) with
//           probability f, else from a HOT table (32 KB, L1-resident).  If L1 returned hits independently of
//           misses, rows/clk would fall gently with f; if the L1 return path is in-order, a few % of misses cost
//           as much as 100 %.
//   mixw    same, but a whole warp instruction is hot or cold.
//   pf      all rows cold; each row is prefetched (prefetch.global.L1) K iterations before its demand load.
//   pfl2    same with prefetch.global.L2 (control).
//   depth   all rows cold; U independent loads in flight per thread (U = 1, 2, 4) at 4 and 8 warps per SM sub-partition.
//   lds     shared-memory row gathers: LDS.128 random rows / sequential / broadcast, LDS.64, LDSM.x4.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <algorithm>

constexpr int ROW = 128;
constexpr int ITERS = 1024;

__device__ __forceinline__ uint32_t lcg(uint32_t x) { return x * 1664525u + 1013904223u; }
__device__ __forceinline__ uint32_t mixbits(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

struct V8 { uint32_t a, b, c, d, e, f, g, h; };
__device__ __forceinline__ V8 ldg256(const void* p) {
  V8 v;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v.a), "=r"(v.b), "=r"(v.c), "=r"(v.d), "=r"(v.e), "=r"(v.f), "=r"(v.g), "=r"(v.h) : "l"(p));
  return v;
}

// MODE 0: row-granular mix; 1: warp-granular mix; 2: prefetch L1; 3: prefetch L2; 4: depth test (U loads per iter)
template <int MODE, int U>
__global__ void __launch_bounds__(256) k_mix(const char* __restrict__ hot, uint32_t hot_mask, const char* __restrict__ cold,
                                             uint32_t cold_mask, uint32_t thresh, int K, long long* cyc, float* out) {
  const int lane = threadIdx.x & 31;
  const uint32_t grp = (blockIdx.x * 256 + threadIdx.x) >> 2;       // 4 lanes share a row
  const uint32_t wrp = (blockIdx.x * 256 + threadIdx.x) >> 5;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0 || MODE == 1) {
    uint32_t s = mixbits(grp * 2654435761u + 17u), sw = mixbits(wrp * 2654435761u + 99u);
    for (int it = 0; it < ITERS; it += 4) {
      V8 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s = lcg(s); sw = lcg(sw);
        const uint32_t h = mixbits(MODE == 0 ? s : sw);
        const bool is_cold = (h & 0xffffu) < thresh;
        const uint32_t r = mixbits(s + 0x9e3779b9u);
        const char* p = is_cold ? cold + (size_t)(r & cold_mask) * ROW : hot + (size_t)(r & hot_mask) * ROW;
        v[u] = ldg256(p + (lane & 3) * 32);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += __uint_as_float(v[u].a) + __uint_as_float(v[u].h);
    }
  } else if (MODE == 2 || MODE == 3) {
    uint32_t s = mixbits(grp * 2654435761u + 17u), sp = s;
    for (int i = 0; i < K; ++i) {
      sp = lcg(sp);
      const char* pp = cold + (size_t)(mixbits(sp) & cold_mask) * ROW;
      if ((lane & 3) == 0) {
        if (MODE == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(pp));
        else asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
      }
    }
    for (int it = 0; it < ITERS; it += 4) {
      V8 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        sp = lcg(sp);
        const char* pp = cold + (size_t)(mixbits(sp) & cold_mask) * ROW;
        if ((lane & 3) == 0) {
          if (MODE == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(pp));
          else asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s = lcg(s);
        const char* p = cold + (size_t)(mixbits(s) & cold_mask) * ROW;
        v[u] = ldg256(p + (lane & 3) * 32);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += __uint_as_float(v[u].a) + __uint_as_float(v[u].h);
    }
  } else {
    uint32_t s = mixbits(grp * 2654435761u + 17u);
    for (int it = 0; it < ITERS; it += U) {
      V8 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        s = lcg(s);
        const char* p = cold + (size_t)(mixbits(s) & cold_mask) * ROW;
        v[u] = ldg256(p + (lane & 3) * 32);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc += __uint_as_float(v[u].a) + __uint_as_float(v[u].h);
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) out[0] = acc;
}

// shared-memory gathers.  MODE 0: LDS.128 random rows (8 lanes/row); 1: LDS.128 sequential (lane*16 + it*512);
// 2: LDS.128 all four quarter-warps read the SAME row (broadcast across quarters); 3: LDS.64 16 lanes/row;
// 4: LDSM.x4 -- lanes 8j..8j+7 give the eight 16-byte chunks of random row j (4 rows per instruction);
// 5: LDS.128, 8 lanes/row, rows of one instruction forced into DIFFERENT 512-byte-bank phases (row index = 4*q + j)
template <int MODE>
__global__ void __launch_bounds__(256) k_lds(uint32_t row_mask, long long* cyc, float* out) {
  extern __shared__ __align__(128) char smem[];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x * 16; i < (int)(row_mask + 1) * ROW; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + i) = make_uint4(i, i + 1, i + 2, i + 3);
  __syncthreads();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t tid_g = blockIdx.x * 256 + threadIdx.x;
  uint32_t s8 = mixbits((tid_g >> 3) * 2654435761u + 5u), s32 = mixbits((tid_g >> 5) * 2654435761u + 7u),
           s16 = mixbits((tid_g >> 4) * 2654435761u + 9u);
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; it += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s8 = lcg(s8); s32 = lcg(s32); s16 = lcg(s16);
      uint32_t addr;
      if (MODE == 0) addr = ((s8 >> 8) & row_mask) * ROW + (lane & 7) * 16;
      else if (MODE == 1) addr = (((s32 >> 8) & row_mask & ~3u) * ROW + lane * 16);
      else if (MODE == 2) addr = ((s32 >> 8) & row_mask) * ROW + (lane & 7) * 16;
      else if (MODE == 3) addr = ((s16 >> 8) & row_mask) * ROW + (lane & 15) * 8;
      else if (MODE == 4) addr = ((s8 >> 8) & row_mask) * ROW + (lane & 7) * 16;
      else addr = ((s8 >> 8) & row_mask) * ROW + (lane & 7) * 16;
      if (MODE == 3) {
        uint2 w;
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w.x), "=r"(w.y) : "r"(sbase + addr));
        v[u] = make_uint4(w.x, w.y, 0, 0);
      } else if (MODE == 4) {
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "r"(sbase + addr));
      } else {
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "r"(sbase + addr));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += __uint_as_float(v[u].x) + __uint_as_float(v[u].w);
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) out[0] = acc;
}

static long long* d_cyc;
static float* d_out;
static int g_sms;

static double finish(int blocks, double rows_per_block, int blocks_per_sm) {
  cudaDeviceSynchronize();
  std::vector<long long> h(blocks);
  cudaMemcpy(h.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0; long long mx = 0;
  for (auto c : h) { mean += (double)c; mx = std::max(mx, c); }
  mean /= blocks;
  (void)mx;
  return rows_per_block * blocks_per_sm / mean;   // rows per clock per SM (all CTAs of an SM run concurrently)
}

template <int MODE, int U>
void run_mix(const char* name, const char* hot, uint32_t hot_rows, const char* cold, uint32_t cold_rows, double f, int K, int bps = 4) {
  const int blocks = g_sms * bps;
  const uint32_t thresh = (uint32_t)(f * 65536.0);
  for (int rep = 0; rep < 2; ++rep)
    k_mix<MODE, U><<<blocks, 256>>>(hot, hot_rows - 1, cold, cold_rows - 1, thresh, K, d_cyc, d_out);
  const double rpc = finish(blocks, 256.0 / 4 * ITERS, bps);
  printf("%-10s f=%6.4f K=%2d U=%d warps/SM=%2d : %6.3f rows/clk/SM  (%6.1f B/clk/SM)  %s\n", name, f, K, U, bps * 8, rpc, rpc * ROW,
         cudaGetErrorString(cudaGetLastError()));
}

template <int MODE>
void run_lds(const char* name, uint32_t rows, double bytes_per_instr) {
  const int blocks = g_sms * 4;
  cudaFuncSetAttribute(k_lds<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  for (int rep = 0; rep < 2; ++rep) k_lds<MODE><<<blocks, 256, rows * ROW>>>(rows - 1, d_cyc, d_out);
  const double rpc = finish(blocks, 8.0 * ITERS * bytes_per_instr / ROW, 4);
  printf("%-44s : %6.3f rows/clk/SM  (%6.1f B/clk/SM)  %s\n", name, rpc, rpc * ROW, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t cold_rows = 1u << 19;   // 64 MB
  const uint32_t hot_rows = 256;         // 32 KB
  char *cold, *hot;
  cudaMalloc(&cold, (size_t)cold_rows * ROW); cudaMalloc(&hot, (size_t)hot_rows * ROW);
  cudaMalloc(&d_out, 4); cudaMalloc(&d_cyc, sizeof(long long) * g_sms * 8);
  cudaMemset(cold, 0, (size_t)cold_rows * ROW); cudaMemset(hot, 0, (size_t)hot_rows * ROW);

  for (double f : {0.0, 0.01, 0.03, 0.0625, 0.125, 0.25, 0.5, 1.0}) run_mix<0, 1>("mix", hot, hot_rows, cold, cold_rows, f, 0);
  for (double f : {0.03, 0.125, 0.25}) run_mix<1, 1>("mixw", hot, hot_rows, cold, cold_rows, f, 0);
  for (double f : {0.125, 0.25}) run_mix<0, 1>("mix", hot, hot_rows, cold, cold_rows, f, 0, 2);
  for (double f : {0.125, 0.25}) run_mix<0, 1>("mix", hot, hot_rows, cold, cold_rows, f, 0, 8);
  for (int K : {4, 8, 16, 32}) run_mix<2, 1>("pfL1", hot, hot_rows, cold, cold_rows, 1.0, K);
  for (int K : {8}) run_mix<3, 1>("pfL2", hot, hot_rows, cold, cold_rows, 1.0, K);
  for (int bps : {1, 2, 4, 8}) {
    run_mix<4, 1>("depth", hot, hot_rows, cold, cold_rows, 1.0, 0, bps);
    run_mix<4, 2>("depth", hot, hot_rows, cold, cold_rows, 1.0, 0, bps);
    run_mix<4, 4>("depth", hot, hot_rows, cold, cold_rows, 1.0, 0, bps);
    run_mix<4, 8>("depth", hot, hot_rows, cold, cold_rows, 1.0, 0, bps);
  }
  run_lds<0>("LDS.128 random rows, 8 lanes/row", 256, 512);
  run_lds<1>("LDS.128 sequential 512 B", 256, 512);
  run_lds<2>("LDS.128 same row in all quarters", 256, 512);
  run_lds<3>("LDS.64 random rows, 16 lanes/row", 256, 256);
  run_lds<4>("LDSM.x4 random rows (4 rows/instr)", 256, 512);
  return 0;
}
