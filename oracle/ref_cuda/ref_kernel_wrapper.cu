// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
//
// C-ABI wrapper that instantiates the UNMODIFIED reference forward launcher
//   ms_deformable_im2col_cuda<float>  (third_party/adet/layers/csrc/DeformAttn/ms_deform_im2col_cuda.cuh:923-954)
// and therefore the unmodified reference kernel
//   ms_deformable_im2col_gpu_kernel   (same file :237-299)
// compiled for sm_100a with the same nvcc that builds the product.  The header is
// #included from where it lies under /root/reference (include dir given on the nvcc
// command line by oracle/Makefile); no reference source is copied into this repo.
// Output goes to oracle/_ref/libmsda_refcuda.so (git-ignored, shipped by gpurun).
//
// Used by: tests/ (bit-exact index/output parity of the new kernels against the real
// reference kernel on the GPU) and bench.py (the "kernel to beat" timing line).
#include "ms_deform_im2col_cuda.cuh"

extern "C" int refcuda_msda_forward_f32(const float* value, const int64_t* spatial_shapes,
                                        const int64_t* level_start_index, const float* sampling_loc,
                                        const float* attn_weight, int batch, int spatial_size,
                                        int num_heads, int channels, int num_levels, int num_query,
                                        int num_point, float* out, void* stream) {
  // the reference host wrapper zero-fills the output first (ms_deform_attn_cuda.cu:54)
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)batch * num_query * num_heads * channels,
                                  (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  ms_deformable_im2col_cuda<float>((cudaStream_t)stream, value, spatial_shapes, level_start_index,
                                   sampling_loc, attn_weight, batch, spatial_size, num_heads, channels,
                                   num_levels, num_query, num_point, out);
  return (int)cudaGetLastError();
}

// same, without the memset (kernel-only timing)
extern "C" int refcuda_msda_forward_f32_nomemset(const float* value, const int64_t* spatial_shapes,
                                                 const int64_t* level_start_index, const float* sampling_loc,
                                                 const float* attn_weight, int batch, int spatial_size,
                                                 int num_heads, int channels, int num_levels, int num_query,
                                                 int num_point, float* out, void* stream) {
  ms_deformable_im2col_cuda<float>((cudaStream_t)stream, value, spatial_shapes, level_start_index,
                                   sampling_loc, attn_weight, batch, spatial_size, num_heads, channels,
                                   num_levels, num_query, num_point, out);
  return (int)cudaGetLastError();
}

extern "C" int refcuda_msda_forward_f64(const double* value, const int64_t* spatial_shapes,
                                        const int64_t* level_start_index, const double* sampling_loc,
                                        const double* attn_weight, int batch, int spatial_size,
                                        int num_heads, int channels, int num_levels, int num_query,
                                        int num_point, double* out, void* stream) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * (size_t)batch * num_query * num_heads * channels,
                                  (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  ms_deformable_im2col_cuda<double>((cudaStream_t)stream, value, spatial_shapes, level_start_index,
                                    sampling_loc, attn_weight, batch, spatial_size, num_heads, channels,
                                    num_levels, num_query, num_point, out);
  return (int)cudaGetLastError();
}
