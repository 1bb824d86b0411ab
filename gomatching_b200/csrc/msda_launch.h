// Internal launch interface between the C-ABI layer (msda_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace msda {

// "once per device" latch for cudaFuncSetAttribute calls: the attributes belong to the (function, device) pair, so a
// process-wide flag would leave a second GPU of the same process without its shared-memory opt-in.  Benign race:
// the guarded calls are idempotent.
struct PerDeviceOnce {
  bool done[64] = {};
  bool need() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

enum TileMode { kModeLinear = 1, kModePyramid = 2, kModeGeneric = 3, kModeStaged = 4, kModePipelined = 5 };

struct FwdParams {
  const void* value;          // (N,S,M,D) fp32 or bf16
  const int64_t* shapes;      // (L,2) device
  const int64_t* lsi;         // (L,)  device
  // core operator inputs
  const float* loc;           // (N,Lq,M,L,P,2)
  const float* attn;          // (N,Lq,M,L,P)
  // fused operator inputs (loc/attn == nullptr)
  const float* ref;           // (N,Lq,L,ref_dim)
  const float* offsets;       // (N,Lq,M,L,P,2)
  const float* logits;        // (N,Lq,M,L*P)
  int ref_dim;
  int off_pitch, logit_pitch; // floats between consecutive (b,q) rows of offsets / logits (M*L*P*2 and M*L*P when dense)
  void* out;                  // (N,Lq,M*D)
  int N, S, M, D, L, Lq, P;
  // tiling (never changes results)
  int mode;                   // TileMode
  int tile_q;                 // queries per tile (linear) = tile_h*tile_w (pyramid)
  int tile_h, tile_w_log2;    // pyramid tile
  int grid;                   // CTAs to launch
  int variant;                // kernel instantiation
  int force_v1;               // use the runtime-L*P tiled kernel even where the specialised one applies
  int walk;                   // fast kernels: 0 strided heads-fastest tile walk, 1 contiguous raster walk per CTA (diagnostic)
  int q_level_begin;          // fast kernels, linear mode: only queries >= level_start_index[q_level_begin] (self-attention)
  int staged_levels;          // staged mode: query levels 0 .. staged_levels-1 run the shared-memory-window kernel
  // shape guard (pipelined mode): the window kernel compares the device-side shapes with the host-side geometry its
  // tensor maps were built from; on a mismatch it writes shape_epoch to *shape_flag (and to *shape_report, pinned host
  // memory) and does nothing, and the register-gather launch that follows serves ALL queries instead of levels 1..3
  int* shape_flag;
  int* shape_report;
  int shape_epoch;
};

// Each returns a cudaError_t cast to int (0 = ok) or MSDA_E_UNSUPPORTED (-5).
int launch_forward_f32(const FwdParams& p, cudaStream_t stream);
int launch_forward_bf16(const FwdParams& p, cudaStream_t stream);
int launch_forward_f64(const double* value, const int64_t* shapes, const int64_t* lsi, const double* loc,
                       const double* attn, int N, int S, int M, int D, int L, int Lq, int P, double* out,
                       cudaStream_t stream);
int forward_variant_count();
// compile-time-specialised kernels for the DeepSolo configuration (msda_forward_fast.cu)
bool fast_shape_supported(int D, int L, int P);
bool fast_supported(const FwdParams& p);   // shape + the 32-byte operand alignment the vector loads need
int fast_variant_count();
int launch_forward_fast_f32(const FwdParams& p, cudaStream_t stream);
int launch_forward_fast_bf16(const FwdParams& p, cudaStream_t stream);
// encoder self-attention, fp32: value windows of one query tile staged in shared memory (msda_forward_staged.cu)
bool staged_supported(const FwdParams& p);
int launch_forward_staged_f32(const FwdParams& p, cudaStream_t stream);
// optional: level shapes on the host let the staged kernel fill its windows with TMA (tensor maps need them)
void staged_set_host_shapes(const int64_t* shapes_host, const int64_t* lsi_host, int L);
bool staged_get_host_shapes(long long (*hw)[2], long long* lsi);   // 4 levels; false if no hint is set
// producer / consumer version of the staged kernel (msda_forward_pipelined.cu): level-0 queries only, needs the hint;
// MSDA_E_UNSUPPORTED if it cannot run (the caller falls back)
int launch_forward_pipelined_f32(const FwdParams& p, const long long (*hw)[2], const long long* lsi, cudaStream_t stream);
// neighbour-paired bf16 value layout (msda_forward_paired.cu): conversion and sampler (D = 32, L = 4, P = 4)
int launch_pair_value(const void* value, int value_is_bf16, const int64_t* shapes, const int64_t* lsi, int N, int S, int M,
                      int L, void* paired, int sms, cudaStream_t stream);
int launch_forward_paired_bf16(const FwdParams& p, cudaStream_t stream);   // p.value = the paired tensor, p.out bf16
// shape guard: a flag slot + a fresh epoch for one pipelined launch pair; MSDA_E_UNSUPPORTED while the stream is being
// captured and the guard's buffers do not exist yet (they are allocated on first use, outside capture)
int shape_guard_acquire(int** flag, int** report, int* epoch, cudaStream_t stream);
int shape_guard_last_mismatch();   // epoch of the most recent mismatch any kernel reported (0: none); a stale read is fine
// true if the tiled kernels can run this problem (else only the generic kernel can)
bool tiled_supported(int elem_bytes, int D, int L, int P, bool fused);

int launch_sample_index(const float* loc, const int64_t* shapes, const int64_t* lsi, int N, int Lq, int M, int D,
                        int L, int P, void* out_records, cudaStream_t stream);
int launch_locations_softmax(const int64_t* shapes, const float* ref, int ref_dim, const float* offsets,
                             const float* logits, int N, int M, int L, int Lq, int P, int lanes_per_unit,
                             float* loc_out, float* attn_out, cudaStream_t stream);
int launch_backward_f32(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                        const float* attn, const float* grad_out, int N, int S, int M, int D, int L, int Lq, int P,
                        float* grad_value, float* grad_loc, float* grad_attn, cudaStream_t stream);

}  // namespace msda
