// Microbenchmark 5: how fast can the TMA engine gather random 128-byte rows into shared memory (no LSU, no registers)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather tma_gather.cu && ./tma_gather
// mode 0: cp.async.bulk.shared::cta.global, one 128-byte row per instruction
// mode 1: cp.async.bulk.tensor.2d ... tile::gather4, four rows of a [rows, 32 float] tensor per instruction
// Every warp keeps BATCH instructions in flight on its own mbarrier (lane 0 issues, the warp waits), 32 warps per SM.
// Reports rows per clock per SM (clock64 and CUDA events).  The table is 16 MB (L2-resident) or 32 KB.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int ROW = 128;
constexpr int ITERS = 256;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(bar),
      "r"(parity) : "memory");
}

template <int MODE, int BATCH>
__global__ void __launch_bounds__(1024) k_tma(const char* __restrict__ table, const __grid_constant__ CUtensorMap tmap, uint32_t row_mask,
                                              long long* cyc) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int ROWS_PER_OP = MODE == 1 ? 4 : 1;
  const uint32_t slot = (uint32_t)__cvta_generic_to_shared(smem) + warp * BATCH * ROWS_PER_OP * ROW;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[warp]);
  if (lane == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint32_t s = (blockIdx.x * 32 + warp) * 2654435761u + 777u;
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (lane == 0) {
      mbar_expect_tx(bar, BATCH * ROWS_PER_OP * ROW);
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const uint32_t dst = slot + b * ROWS_PER_OP * ROW;
        if (MODE == 0) {
          s = s * 1664525u + 1013904223u;
          const char* src = table + (size_t)((s >> 8) & row_mask) * ROW;
          asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(ROW), "r"(bar) : "memory");
        } else {
          uint32_t r[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) { s = s * 1664525u + 1013904223u; r[j] = (s >> 8) & row_mask; }
          asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                       ::"r"(dst), "l"(&tmap), "r"(bar), "r"(0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
        }
      }
    }
    mbar_wait(bar, it & 1);
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int g_sms, g_khz;
static long long* d_cyc;

template <int MODE, int BATCH>
void run(const char* name, const char* table, const CUtensorMap& tmap, uint32_t rows) {
  constexpr int ROWS_PER_OP = MODE == 1 ? 4 : 1;
  const int smem = 32 * BATCH * ROWS_PER_OP * ROW;
  cudaFuncSetAttribute(k_tma<MODE, BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_tma<MODE, BATCH><<<g_sms, 1024, smem>>>(table, tmap, rows - 1, d_cyc);
  cudaEventRecord(a);
  k_tma<MODE, BATCH><<<g_sms, 1024, smem>>>(table, tmap, rows - 1, d_cyc);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  std::vector<long long> h(g_sms);
  cudaMemcpy(h.data(), d_cyc, g_sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0; for (auto c : h) mean += (double)c; mean /= g_sms;
  const double rows_per_block = 32.0 * BATCH * ROWS_PER_OP * ITERS;
  printf("%-44s table=%7u rows, %2d ops/warp in flight: %6.3f rows/clk/SM by clock64 | %6.3f by events  %s\n", name, rows, BATCH,
         rows_per_block / mean, rows_per_block / (ms * 1e-3 * g_khz * 1e3), cudaGetErrorString(e));
}

int main() {
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&g_khz, cudaDevAttrClockRate, 0);
  const uint32_t big = 1u << 17;   // 16 MB
  char* table;
  cudaMalloc(&table, (size_t)big * ROW); cudaMemset(table, 0, (size_t)big * ROW);
  cudaMalloc(&d_cyc, sizeof(long long) * g_sms);
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {32, big};
  cuuint64_t gstride[1] = {ROW};
  cuuint32_t box[2] = {32, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(sym)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, table, gdim, gstride, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("tensor map: %d\n", (int)r);
  for (uint32_t rows : {256u, big}) {
    run<0, 4>("cp.async.bulk 128 B rows", table, tmap, rows);
    run<0, 8>("cp.async.bulk 128 B rows", table, tmap, rows);
    run<0, 16>("cp.async.bulk 128 B rows", table, tmap, rows);
    run<1, 2>("TMA tile::gather4 (4 x 128 B per op)", table, tmap, rows);
    run<1, 4>("TMA tile::gather4 (4 x 128 B per op)", table, tmap, rows);
    run<1, 8>("TMA tile::gather4 (4 x 128 B per op)", table, tmap, rows);
  }
  return 0;
}
