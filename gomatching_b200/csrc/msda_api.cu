// C-ABI layer of libmsda_b200.so: argument validation, launch heuristics, host-buffer context.
// Declarations and the reference interfaces they replace: include/msda_b200.h
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <new>

#include "../../include/msda_b200.h"
#include "msda_launch.h"

using namespace msda;

namespace {

int g_sm_count[64] = {0};

int sm_count_current() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return MSDA_E_NOCUDA;
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_sm_count[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return MSDA_E_NOCUDA;
    g_sm_count[dev] = n;
  }
  return g_sm_count[dev];
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int ilog2_floor(int x) { int r = 0; while ((1 << (r + 1)) <= x) ++r; return r; }

int check_dims(int N, int S, int M, int D, int L, int Lq, int P, int elem_bytes) {
  if (N <= 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || Lq <= 0 || P <= 0) return MSDA_E_DIMS;
  // int32 byte offsets inside one batch item's value map; int64 everywhere else
  const long long vbytes = (long long)S * M * D * elem_bytes;
  if (vbytes >= (1ll << 31)) return MSDA_E_DIMS;
  if ((long long)L * P > (1 << 20)) return MSDA_E_DIMS;
  return 0;
}

// Fill the tiling fields of FwdParams from the optional tuning block.
int plan(FwdParams& p, int elem_bytes, bool fused, const msda_b200_tuning_t* tn) {
  const int sms = sm_count_current();
  if (sms < 0) return sms;
  const bool can_tile = tiled_supported(elem_bytes, p.D, p.L, p.P, fused);
  int mode = tn ? tn->mode : 0;
  const bool self_attn = p.Lq == p.S;
  // defaults from the B200 sweeps (profiles/r01_s9_sweep_*.log, r01_s12_sweep_*.log): 2-D pyramid tiles only pay for the fp32 core
  // operator in encoder self-attention; everything else runs linear query tiles
  if (mode == 0) mode = !can_tile ? kModeGeneric : ((self_attn && elem_bytes == 4 && !fused) ? kModePyramid : kModeLinear);
  if (mode == kModePyramid && !self_attn) mode = kModeLinear;
  if ((mode == kModeStaged || mode == kModePipelined) &&
      !(self_attn && elem_bytes == 4 && can_tile && fast_shape_supported(p.D, p.L, p.P)))
    mode = kModeLinear;
  if (mode != kModeGeneric && !can_tile) {
    if (fused) return MSDA_E_UNSUPPORTED;
    mode = kModeGeneric;
  }
  if (mode == kModeGeneric && fused) return MSDA_E_UNSUPPORTED;
  p.mode = mode;
  p.variant = (tn && tn->variant >= 0 && (tn->mode || tn->variant)) ? tn->variant : 3;   // VB16, 8 warps, 4 CTAs/SM, 16-B records
  p.force_v1 = (tn && tn->reserved[0] == 1) ? 1 : 0;
  p.walk = (tn && tn->reserved[1] == 1) ? 1 : 0;
  if (p.variant < 0 || p.variant >= forward_variant_count()) p.variant = 3;
  int th = (tn && tn->tile_h > 0) ? tn->tile_h : 16;
  int tw = (tn && tn->tile_w > 0) ? tn->tile_w : 8;
  int tq = (tn && tn->tile_q > 0) ? tn->tile_q : ((self_attn && elem_bytes == 2) ? 256 : 64);
  int cps = (tn && tn->ctas_per_sm > 0) ? tn->ctas_per_sm : (mode == kModePyramid ? 3 : 4);
  if (mode == kModePipelined) {
    // producer / consumer window kernel for the level-0 queries (one CTA per SM = grid / 4), 64-query linear tiles for the rest
    p.staged_levels = 1;
    p.tile_q = tq;
    // tuning tile_h / tile_w > 0: 2-D pyramid tiles for the coarse query levels on the register-gather kernel
    p.tile_h = (tn && tn->tile_h > 0) ? tn->tile_h : 0;
    p.tile_w_log2 = (tn && tn->tile_h > 0) ? ilog2_floor(tw < 4 ? 4 : tw) : 0;
    p.grid = sms * 4;
    return 0;
  }
  if (mode == kModeStaged) {
    // staged kernel: two 256-thread CTAs per SM; the remaining query levels run 64-query linear tiles (variant 3)
    p.staged_levels = (tn && tn->tile_h > 0 && tn->tile_h <= 3) ? tn->tile_h : 1;
    p.tile_q = tq;
    p.tile_h = 0; p.tile_w_log2 = 0;
    p.grid = sms * 2;
    return 0;
  }
  p.tile_w_log2 = ilog2_floor(tw < 4 ? 4 : tw);
  p.tile_h = th;
  p.tile_q = (mode == kModePyramid) ? th * (1 << p.tile_w_log2) : tq;
  long long grid = (long long)sms * cps;
  if (mode == kModeLinear) {
    const long long tiles = (long long)p.N * p.M * ((p.Lq + p.tile_q - 1) / p.tile_q);
    if (tiles < grid) grid = tiles;
  }
  p.grid = (int)(grid < 1 ? 1 : grid);
  return 0;
}

int forward_common(const void* value, const int64_t* shapes, const int64_t* lsi, const float* loc, const float* attn,
                   const float* ref, int ref_dim, const float* offsets, const float* logits, int N, int S, int M, int D,
                   int L, int Lq, int P, void* out, void* stream, const msda_b200_tuning_t* tn, int elem_bytes,
                   int off_pitch = 0, int logit_pitch = 0) {
  const bool fused = (loc == nullptr);
  if (!value || !shapes || !lsi || !out) return MSDA_E_NULLPTR;
  if (fused) {
    if (!ref || !offsets || !logits) return MSDA_E_NULLPTR;
    if (ref_dim != 2 && ref_dim != 4) return MSDA_E_REFDIM;
  } else if (!attn) {
    return MSDA_E_NULLPTR;
  }
  int rc = check_dims(N, S, M, D, L, Lq, P, elem_bytes);
  if (rc) return rc;
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.value = value; p.shapes = shapes; p.lsi = lsi; p.loc = loc; p.attn = attn;
  p.ref = ref; p.offsets = offsets; p.logits = logits; p.ref_dim = ref_dim; p.out = out;
  p.N = N; p.S = S; p.M = M; p.D = D; p.L = L; p.Lq = Lq; p.P = P;
  p.off_pitch = off_pitch > 0 ? off_pitch : M * L * P * 2;
  p.logit_pitch = logit_pitch > 0 ? logit_pitch : M * L * P;
  if (p.off_pitch < M * L * P * 2 || p.logit_pitch < M * L * P || (p.off_pitch & 1)) return MSDA_E_DIMS;
  const bool pitched = fused && (p.off_pitch != M * L * P * 2 || p.logit_pitch != M * L * P);
  if (pitched && !fast_shape_supported(D, L, P)) return MSDA_E_UNSUPPORTED;   // only the specialised kernels take pitches
  rc = plan(p, elem_bytes, fused, tn);
  if (rc) return rc;
  if (pitched && (p.mode == kModeGeneric || p.force_v1 || !fast_supported(p))) return MSDA_E_UNSUPPORTED;
  if (p.mode != kModeGeneric) {
    if (!aligned16(value) || !aligned16(out)) return MSDA_E_ALIGN;
    if (fused ? ((reinterpret_cast<uintptr_t>(offsets) & 7u) != 0) : ((reinterpret_cast<uintptr_t>(loc) & 7u) != 0))
      return MSDA_E_ALIGN;
  }
  return elem_bytes == 4 ? launch_forward_f32(p, (cudaStream_t)stream) : launch_forward_bf16(p, (cudaStream_t)stream);
}

}  // namespace

// ---- shape guard state: per device a ring of flag slots, one pinned host word for all devices ---------------------
namespace msda {
namespace {
constexpr int kGuardSlots = 64, kGuardDevs = 64;
std::mutex g_guard_mu;
int* g_guard_ring[kGuardDevs] = {};
int* g_guard_host = nullptr;          // pinned, mapped: the kernels store the epoch of a mismatch here
std::atomic<int> g_guard_epoch{0};
}  // namespace

int shape_guard_acquire(int** flag, int** report, int* epoch, cudaStream_t stream) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= kGuardDevs) return MSDA_E_UNSUPPORTED;
  if (!g_guard_ring[dev]) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return MSDA_E_UNSUPPORTED;                       // cannot allocate while a graph is being captured
    }
    std::lock_guard<std::mutex> lock(g_guard_mu);
    if (!g_guard_host) {
      e = cudaHostAlloc(reinterpret_cast<void**>(&g_guard_host), sizeof(int), cudaHostAllocPortable | cudaHostAllocMapped);
      if (e != cudaSuccess) return (int)e;
      *g_guard_host = 0;
    }
    if (!g_guard_ring[dev]) {
      int* ring = nullptr;
      e = cudaMalloc(reinterpret_cast<void**>(&ring), kGuardSlots * sizeof(int));
      if (e != cudaSuccess) return (int)e;
      e = cudaMemset(ring, 0, kGuardSlots * sizeof(int));
      if (e != cudaSuccess) return (int)e;
      g_guard_ring[dev] = ring;
    }
  }
  int ep = g_guard_epoch.fetch_add(1) + 1;
  if (ep <= 0) { g_guard_epoch = 1; ep = 1; }
  *flag = g_guard_ring[dev] + (ep % kGuardSlots);
  *report = g_guard_host;                              // portable + unified addressing: valid on every device
  *epoch = ep;
  return 0;
}

int shape_guard_last_mismatch() { return g_guard_host ? *reinterpret_cast<volatile int*>(g_guard_host) : 0; }
}  // namespace msda

extern "C" {

int msda_b200_shape_mismatch_epoch(void) { return shape_guard_last_mismatch(); }

int msda_b200_abi_version(void) { return MSDA_B200_ABI_VERSION; }

const char* msda_b200_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case MSDA_E_NULLPTR: return "msda_b200: a required pointer is NULL";
    case MSDA_E_DIMS: return "msda_b200: invalid dimensions (non-positive, or value map >= 2 GiB per batch item)";
    case MSDA_E_ALIGN: return "msda_b200: tensor pointer is not sufficiently aligned (16 B value/out, 8 B loc/offsets)";
    case MSDA_E_REFDIM: return "msda_b200: last dim of reference_points must be 2 or 4";
    case MSDA_E_UNSUPPORTED: return "msda_b200: no kernel for this dtype/shape combination";
    case MSDA_E_NOCUDA: return "msda_b200: no CUDA device available (this library has no CPU path)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "msda_b200: unknown error";
}

int msda_b200_pair_value_bf16(const void* value, int value_is_bf16, const int64_t* shapes, const int64_t* lsi, int N, int S,
                              int M, int D, int L, void* paired, void* stream) {
  if (!value || !shapes || !lsi || !paired) return MSDA_E_NULLPTR;
  if (N <= 0 || S <= 0 || M <= 0 || L <= 0 || L > 16) return MSDA_E_DIMS;
  if (D != 32) return MSDA_E_UNSUPPORTED;
  if (!aligned16(value) || !aligned16(paired)) return MSDA_E_ALIGN;
  if ((long long)S * 128 >= (1ll << 31)) return MSDA_E_DIMS;           // int32 byte offsets inside one head's map
  const int sms = sm_count_current();
  if (sms < 0) return sms;
  return launch_pair_value(value, value_is_bf16, shapes, lsi, N, S, M, L, paired, sms, (cudaStream_t)stream);
}

static int forward_paired_common(const void* paired, const int64_t* shapes, const int64_t* lsi, const float* loc,
                                 const float* attn, const float* ref, int ref_dim, const float* offsets, const float* logits,
                                 int N, int S, int M, int D, int L, int Lq, int P, void* out, void* stream) {
  const bool fused = loc == nullptr;
  if (!paired || !shapes || !lsi || !out) return MSDA_E_NULLPTR;
  if (fused ? (!ref || !offsets || !logits) : !attn) return MSDA_E_NULLPTR;
  if (fused && ref_dim != 2 && ref_dim != 4) return MSDA_E_REFDIM;
  int rc = check_dims(N, S, M, D, L, Lq, P, 4);
  if (rc) return rc;
  if (!(D == 32 && L == 4 && P == 4)) return MSDA_E_UNSUPPORTED;
  if ((long long)S * 128 >= (1ll << 31)) return MSDA_E_DIMS;
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.value = paired; p.shapes = shapes; p.lsi = lsi; p.loc = loc; p.attn = attn;
  p.ref = ref; p.offsets = offsets; p.logits = logits; p.ref_dim = ref_dim; p.out = out;
  p.N = N; p.S = S; p.M = M; p.D = D; p.L = L; p.Lq = Lq; p.P = P;
  p.off_pitch = M * L * P * 2; p.logit_pitch = M * L * P;
  if (!aligned16(paired) || !aligned16(out)) return MSDA_E_ALIGN;
  const uintptr_t a = fused ? (reinterpret_cast<uintptr_t>(offsets) | reinterpret_cast<uintptr_t>(logits) |
                               reinterpret_cast<uintptr_t>(ref))
                            : (reinterpret_cast<uintptr_t>(loc) | reinterpret_cast<uintptr_t>(attn));
  if (a & 15u) return MSDA_E_ALIGN;
  const int sms = sm_count_current();
  if (sms < 0) return sms;
  p.mode = kModeLinear;
  p.tile_q = 64;
  const long long tiles = (long long)N * M * ((Lq + p.tile_q - 1) / p.tile_q);
  long long grid = (long long)sms * 4;
  if (tiles < grid) grid = tiles;
  p.grid = (int)(grid < 1 ? 1 : grid);
  return launch_forward_paired_bf16(p, (cudaStream_t)stream);
}

int msda_b200_forward_paired_bf16(const void* paired, const int64_t* shapes, const int64_t* lsi, const float* loc,
                                  const float* attn, int N, int S, int M, int D, int L, int Lq, int P, void* out, void* stream) {
  if (!loc) return MSDA_E_NULLPTR;
  return forward_paired_common(paired, shapes, lsi, loc, attn, nullptr, 2, nullptr, nullptr, N, S, M, D, L, Lq, P, out, stream);
}

int msda_b200_forward_fused_paired_bf16(const void* paired, const int64_t* shapes, const int64_t* lsi, const float* ref,
                                        int ref_dim, const float* offsets, const float* logits, int N, int S, int M, int D,
                                        int L, int Lq, int P, void* out, void* stream) {
  return forward_paired_common(paired, shapes, lsi, nullptr, nullptr, ref, ref_dim, offsets, logits, N, S, M, D, L, Lq, P, out,
                               stream);
}

int msda_b200_sm_count(void) { return sm_count_current(); }
int msda_b200_variant_count(void) { return forward_variant_count(); }

void msda_b200_staged_set_host_shapes(const int64_t* shapes_host, const int64_t* lsi_host, int L) {
  staged_set_host_shapes(shapes_host, lsi_host, L);
}

int msda_b200_forward_f32(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                          const float* attn, int N, int S, int M, int D, int L, int Lq, int P, float* out,
                          void* stream) {
  if (!loc) return MSDA_E_NULLPTR;
  return forward_common(value, shapes, lsi, loc, attn, nullptr, 0, nullptr, nullptr, N, S, M, D, L, Lq, P, out, stream,
                        nullptr, 4);
}

int msda_b200_forward_bf16(const void* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                           const float* attn, int N, int S, int M, int D, int L, int Lq, int P, void* out,
                           void* stream) {
  if (!loc) return MSDA_E_NULLPTR;
  return forward_common(value, shapes, lsi, loc, attn, nullptr, 0, nullptr, nullptr, N, S, M, D, L, Lq, P, out, stream,
                        nullptr, 2);
}

int msda_b200_forward_f64(const double* value, const int64_t* shapes, const int64_t* lsi, const double* loc,
                          const double* attn, int N, int S, int M, int D, int L, int Lq, int P, double* out,
                          void* stream) {
  if (!value || !shapes || !lsi || !loc || !attn || !out) return MSDA_E_NULLPTR;
  int rc = check_dims(N, S, M, D, L, Lq, P, 8);
  if (rc) return rc;
  return launch_forward_f64(value, shapes, lsi, loc, attn, N, S, M, D, L, Lq, P, out, (cudaStream_t)stream);
}

int msda_b200_forward_f32_ex(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                             const float* attn, int N, int S, int M, int D, int L, int Lq, int P, float* out,
                             void* stream, const msda_b200_tuning_t* tuning) {
  if (!loc) return MSDA_E_NULLPTR;
  return forward_common(value, shapes, lsi, loc, attn, nullptr, 0, nullptr, nullptr, N, S, M, D, L, Lq, P, out, stream,
                        tuning, 4);
}

int msda_b200_forward_bf16_ex(const void* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                              const float* attn, int N, int S, int M, int D, int L, int Lq, int P, void* out,
                              void* stream, const msda_b200_tuning_t* tuning) {
  if (!loc) return MSDA_E_NULLPTR;
  return forward_common(value, shapes, lsi, loc, attn, nullptr, 0, nullptr, nullptr, N, S, M, D, L, Lq, P, out, stream,
                        tuning, 2);
}

int msda_b200_forward_fused_f32(const float* value, const int64_t* shapes, const int64_t* lsi, const float* ref,
                                int ref_dim, const float* offsets, const float* logits, int N, int S, int M, int D,
                                int L, int Lq, int P, float* out, void* stream, const msda_b200_tuning_t* tuning) {
  return forward_common(value, shapes, lsi, nullptr, nullptr, ref, ref_dim, offsets, logits, N, S, M, D, L, Lq, P, out,
                        stream, tuning, 4);
}

int msda_b200_forward_fused_bf16(const void* value, const int64_t* shapes, const int64_t* lsi, const float* ref,
                                 int ref_dim, const float* offsets, const float* logits, int N, int S, int M, int D,
                                 int L, int Lq, int P, void* out, void* stream, const msda_b200_tuning_t* tuning) {
  return forward_common(value, shapes, lsi, nullptr, nullptr, ref, ref_dim, offsets, logits, N, S, M, D, L, Lq, P, out,
                        stream, tuning, 2);
}

int msda_b200_forward_fused_pitched_f32(const float* value, const int64_t* shapes, const int64_t* lsi, const float* ref,
                                        int ref_dim, const float* offsets, int off_pitch, const float* logits,
                                        int logit_pitch, int N, int S, int M, int D, int L, int Lq, int P, float* out,
                                        void* stream, const msda_b200_tuning_t* tuning) {
  return forward_common(value, shapes, lsi, nullptr, nullptr, ref, ref_dim, offsets, logits, N, S, M, D, L, Lq, P, out,
                        stream, tuning, 4, off_pitch, logit_pitch);
}

int msda_b200_forward_fused_pitched_bf16(const void* value, const int64_t* shapes, const int64_t* lsi, const float* ref,
                                         int ref_dim, const float* offsets, int off_pitch, const float* logits,
                                         int logit_pitch, int N, int S, int M, int D, int L, int Lq, int P, void* out,
                                         void* stream, const msda_b200_tuning_t* tuning) {
  return forward_common(value, shapes, lsi, nullptr, nullptr, ref, ref_dim, offsets, logits, N, S, M, D, L, Lq, P, out,
                        stream, tuning, 2, off_pitch, logit_pitch);
}

int msda_b200_locations_softmax_f32(const int64_t* shapes, const float* ref, int ref_dim, const float* offsets,
                                    const float* logits, int N, int M, int L, int Lq, int P, int lanes_per_unit,
                                    float* loc_out, float* attn_out, void* stream) {
  if (!shapes || !ref || !offsets || !logits) return MSDA_E_NULLPTR;
  if (ref_dim != 2 && ref_dim != 4) return MSDA_E_REFDIM;
  if (N <= 0 || M <= 0 || L <= 0 || Lq <= 0 || P <= 0) return MSDA_E_DIMS;
  if (lanes_per_unit != 4 && lanes_per_unit != 8 && lanes_per_unit != 16) return MSDA_E_UNSUPPORTED;
  return launch_locations_softmax(shapes, ref, ref_dim, offsets, logits, N, M, L, Lq, P, lanes_per_unit, loc_out,
                                  attn_out, (cudaStream_t)stream);
}

int msda_b200_sample_index_f32(const float* loc, const int64_t* shapes, const int64_t* lsi, int N, int Lq, int M, int D,
                               int L, int P, msda_b200_index_t* out, void* stream) {
  if (!loc || !shapes || !lsi || !out) return MSDA_E_NULLPTR;
  if (N <= 0 || M <= 0 || D <= 0 || L <= 0 || Lq <= 0 || P <= 0) return MSDA_E_DIMS;
  return launch_sample_index(loc, shapes, lsi, N, Lq, M, D, L, P, out, (cudaStream_t)stream);
}

int msda_b200_backward_f32(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                           const float* attn, const float* grad_out, int N, int S, int M, int D, int L, int Lq, int P,
                           float* grad_value, float* grad_loc, float* grad_attn, void* stream) {
  if (!value || !shapes || !lsi || !loc || !attn || !grad_out || !grad_value || !grad_loc || !grad_attn)
    return MSDA_E_NULLPTR;
  int rc = check_dims(N, S, M, D, L, Lq, P, 4);
  if (rc) return rc;
  return launch_backward_f32(value, shapes, lsi, loc, attn, grad_out, N, S, M, D, L, Lq, P, grad_value, grad_loc,
                             grad_attn, (cudaStream_t)stream);
}

// ---- host-buffer context ----------------------------------------------------------------------------
struct msda_b200_host_ctx {
  int device;
  cudaStream_t stream;
  void* ws;          // device workspace
  size_t ws_bytes;
};

int msda_b200_host_ctx_create(msda_b200_host_ctx_t** ctx, int device) {
  if (!ctx) return MSDA_E_NULLPTR;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return (int)e;
  msda_b200_host_ctx* c = new (std::nothrow) msda_b200_host_ctx();
  if (!c) return (int)cudaErrorMemoryAllocation;
  c->device = device; c->ws = nullptr; c->ws_bytes = 0;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; return (int)e; }
  *ctx = c;
  return 0;
}

void msda_b200_host_ctx_destroy(msda_b200_host_ctx_t* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->ws) cudaFree(c->ws);
  cudaStreamDestroy(c->stream);
  delete c;
}

int msda_b200_forward_f32_host(msda_b200_host_ctx_t* c, const float* value, const int64_t* shapes, const int64_t* lsi,
                               const float* loc, const float* attn, int N, int S, int M, int D, int L, int Lq, int P,
                               float* out) {
  if (!c || !value || !shapes || !lsi || !loc || !attn || !out) return MSDA_E_NULLPTR;
  int rc = check_dims(N, S, M, D, L, Lq, P, 4);
  if (rc) return rc;
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) return (int)e;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_value = up((size_t)N * S * M * D * 4), b_loc = up((size_t)N * Lq * M * L * P * 8),
               b_attn = up((size_t)N * Lq * M * L * P * 4), b_out = up((size_t)N * Lq * M * D * 4),
               b_shapes = up((size_t)L * 16), b_lsi = up((size_t)L * 8);
  const size_t need = b_value + b_loc + b_attn + b_out + b_shapes + b_lsi;
  if (need > c->ws_bytes) {
    if (c->ws) cudaFree(c->ws);
    c->ws = nullptr; c->ws_bytes = 0;
    e = cudaMalloc(&c->ws, need);
    if (e != cudaSuccess) return (int)e;
    c->ws_bytes = need;
  }
  char* w = static_cast<char*>(c->ws);
  float* d_value = reinterpret_cast<float*>(w); w += b_value;
  float* d_loc = reinterpret_cast<float*>(w); w += b_loc;
  float* d_attn = reinterpret_cast<float*>(w); w += b_attn;
  float* d_out = reinterpret_cast<float*>(w); w += b_out;
  int64_t* d_shapes = reinterpret_cast<int64_t*>(w); w += b_shapes;
  int64_t* d_lsi = reinterpret_cast<int64_t*>(w);
  cudaStream_t st = c->stream;
#define MSDA_TRY(x) do { e = (x); if (e != cudaSuccess) return (int)e; } while (0)
  MSDA_TRY(cudaMemcpyAsync(d_shapes, shapes, (size_t)L * 16, cudaMemcpyHostToDevice, st));
  MSDA_TRY(cudaMemcpyAsync(d_lsi, lsi, (size_t)L * 8, cudaMemcpyHostToDevice, st));
  MSDA_TRY(cudaMemcpyAsync(d_value, value, (size_t)N * S * M * D * 4, cudaMemcpyHostToDevice, st));
  MSDA_TRY(cudaMemcpyAsync(d_loc, loc, (size_t)N * Lq * M * L * P * 8, cudaMemcpyHostToDevice, st));
  MSDA_TRY(cudaMemcpyAsync(d_attn, attn, (size_t)N * Lq * M * L * P * 4, cudaMemcpyHostToDevice, st));
  rc = forward_common(d_value, d_shapes, d_lsi, d_loc, d_attn, nullptr, 0, nullptr, nullptr, N, S, M, D, L, Lq, P, d_out,
                      st, nullptr, 4);
  if (rc) return rc;
  MSDA_TRY(cudaMemcpyAsync(out, d_out, (size_t)N * Lq * M * D * 4, cudaMemcpyDeviceToHost, st));
  MSDA_TRY(cudaStreamSynchronize(st));
#undef MSDA_TRY
  return 0;
}

}  // extern "C"
