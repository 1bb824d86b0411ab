// Microbenchmark: how fast can one SM gather 128-byte rows through (a) the L1 cache path with various
// load widths / rows-per-instruction and (b) shared memory?  Decides the design of the MSDeformAttn gather.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu && ./gather_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

constexpr int ROW = 128;            // bytes per row
constexpr int ITERS = 2048;

__device__ __forceinline__ uint32_t next_row(uint32_t x, uint32_t mask) { return (x * 1664525u + 1013904223u) & mask; }

template <int MODE>
__global__ void __launch_bounds__(256) k_gather(const char* __restrict__ table, uint32_t row_mask, float* out, int stride_rows) {
  extern __shared__ __align__(128) char smem[];
  const int lane = threadIdx.x & 31;
  if (MODE == 3 || MODE == 6) {
    for (int i = threadIdx.x * 16; i < (int)(row_mask + 1) * ROW; i += blockDim.x * 16)
      *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(table + i);
    __syncthreads();
  }
  const uint32_t tid_g = blockIdx.x * 256 + threadIdx.x;
  uint32_t seed = 0, s32 = (tid_g >> 5) * 2654435761u + 12345u, s16 = (tid_g >> 4) * 2654435761u + 12345u, s8 = (tid_g >> 3) * 2654435761u + 12345u, s4 = (tid_g >> 2) * 2654435761u + 12345u, s2 = (tid_g >> 1) * 2654435761u + 12345u;
  float acc = 0.f;
  for (int it = 0; it < ITERS; it += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s32 = s32 * 1664525u + 1013904223u; s16 = s16 * 1664525u + 1013904223u; s8 = s8 * 1664525u + 1013904223u; s4 = s4 * 1664525u + 1013904223u; s2 = s2 * 1664525u + 1013904223u; (void)seed;
      if (MODE == 0) {            // LDG.128, 8 lanes per row, 4 random rows per instruction
        uint32_t r = (s8 >> 8) & row_mask;
        uint4 v = __ldg(reinterpret_cast<const uint4*>(table + (size_t)r * ROW + (lane & 7) * 16));
        acc += __uint_as_float(v.x) + __uint_as_float(v.w);
      } else if (MODE == 1) {     // LDG.128, 4 consecutive rows (512 contiguous bytes)
        uint32_t r = ((s32 >> 8) & row_mask & ~3u) + (lane >> 3);
        uint4 v = __ldg(reinterpret_cast<const uint4*>(table + (size_t)r * ROW + (lane & 7) * 16));
        acc += __uint_as_float(v.x) + __uint_as_float(v.w);
      } else if (MODE == 2) {     // LDG.32, one random row per warp instruction (the reference's pattern)
        uint32_t r = (s32 >> 8) & row_mask;
        float v = __ldg(reinterpret_cast<const float*>(table + (size_t)r * ROW + lane * 4));
        acc += v;
      } else if (MODE == 3) {     // LDS.128 from shared memory, 4 random rows per instruction
        uint32_t r = (s8 >> 8) & row_mask;
        uint4 v = *reinterpret_cast<const uint4*>(smem + (size_t)r * ROW + (lane & 7) * 16);
        acc += __uint_as_float(v.x) + __uint_as_float(v.w);
      } else if (MODE == 4) {     // LDG.256, 4 lanes per row, 8 random rows per instruction
        uint32_t r = (s4 >> 8) & row_mask;
        uint32_t a, b, c, d, e, f, g2, h;
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g2), "=r"(h)
                     : "l"(table + (size_t)r * ROW + (lane & 3) * 32));
        acc += __uint_as_float(a) + __uint_as_float(h);
      } else if (MODE == 5) {     // LDG.64, 16 lanes per row, 2 random rows per instruction
        uint32_t r = (s16 >> 8) & row_mask;
        uint2 v = __ldg(reinterpret_cast<const uint2*>(table + (size_t)r * ROW + (lane & 15) * 8));
        acc += __uint_as_float(v.x) + __uint_as_float(v.y);
      } else if (MODE == 6) {     // LDS.32 from shared memory, one random row per warp instruction
        uint32_t r = (s32 >> 8) & row_mask;
        float v = *reinterpret_cast<const float*>(smem + (size_t)r * ROW + lane * 4);
        acc += v;
      } else if (MODE == 8) {     // LDG.256, 2 lanes per 64-byte (bf16) row, 16 random half-line rows per instruction
        uint32_t r = (s2 >> 8) & (row_mask * 2 + 1);
        uint32_t a, b, c, d, e, f, g2, h;
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g2), "=r"(h)
                     : "l"(table + (size_t)r * 64 + (lane & 1) * 32));
        acc += __uint_as_float(a) + __uint_as_float(h);
      } else if (MODE == 9) {     // LDG.128, 4 lanes per 64-byte (bf16) row, 8 random half-line rows per instruction
        uint32_t r = (s4 >> 8) & (row_mask * 2 + 1);
        uint4 v = __ldg(reinterpret_cast<const uint4*>(table + (size_t)r * 64 + (lane & 3) * 16));
        acc += __uint_as_float(v.x) + __uint_as_float(v.w);
      } else if (MODE == 7) {     // LDG.128, 8 lanes per row, 4 rows strided by stride_rows (pixel-neighbour pattern)
        uint32_t r = (((s32 >> 8) & row_mask) + (lane >> 3) * stride_rows) & row_mask;
        uint4 v = __ldg(reinterpret_cast<const uint4*>(table + (size_t)r * ROW + (lane & 7) * 16));
        acc += __uint_as_float(v.x) + __uint_as_float(v.w);
      }
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int MODE>
void run(const char* name, const char* table, uint32_t rows, float* out, int sms, double bytes_per_instr_warp, int stride_rows = 8) {
  const size_t smem = (MODE == 3 || MODE == 6) ? (size_t)rows * ROW : 0;
  cudaFuncSetAttribute(k_gather<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int blocks = sms * 4;
  const int threads = 256;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_gather<MODE><<<blocks, threads, smem>>>(table, rows - 1, out, stride_rows);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k_gather<MODE><<<blocks, threads, smem>>>(table, rows - 1, out, stride_rows);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double warps = (double)blocks * threads / 32;
  const double bytes = warps * ITERS * bytes_per_instr_warp;
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("%-58s rows=%6u  %8.1f us  %8.1f GB/s  %6.1f B/clk/SM (at %d MHz)  err=%s\n", name, rows, ms * 1e3, bytes / ms / 1e6,
         bytes / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t max_rows = 1u << 20;   // 128 MB
  char* table; float* out;
  cudaMalloc(&table, (size_t)max_rows * ROW); cudaMalloc(&out, 4);
  cudaMemset(table, 0, (size_t)max_rows * ROW);
  for (uint32_t rows : {256u, 1024u, 65536u}) {   // 64 KB (L1-resident), 128 KB, 8 MB (L2-resident)
    run<0>("LDG.128  8 lanes/row, 4 random rows/instr", table, rows, out, sms, 512);
    run<1>("LDG.128  4 consecutive rows/instr (coalesced 512 B)", table, rows, out, sms, 512);
    run<7>("LDG.128  4 rows strided 1 KB (pixel neighbours)", table, rows, out, sms, 512, 8);
    run<2>("LDG.32   1 random row/instr", table, rows, out, sms, 128);
    run<5>("LDG.64   16 lanes/row, 2 random rows/instr", table, rows, out, sms, 256);
    run<4>("LDG.256  4 lanes/row, 8 random rows/instr", table, rows, out, sms, 1024);
    run<8>("LDG.256  2 lanes/64B row, 16 random rows/instr (bf16 rows)", table, rows, out, sms, 1024);
    run<9>("LDG.128  4 lanes/64B row, 8 random rows/instr (bf16 rows)", table, rows, out, sms, 512);
    if (rows <= 256) {
      run<3>("LDS.128  8 lanes/row, 4 random rows/instr (smem)", table, rows, out, sms, 512);
      run<6>("LDS.32   1 random row/instr (smem)", table, rows, out, sms, 128);
    }
  }
  return 0;
}
