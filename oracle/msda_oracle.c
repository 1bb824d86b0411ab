/*
 * TEST INFRASTRUCTURE -- CPU oracle for the MSDeformAttn hot path.  NOT part of the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product path (gomatching_b200/) never links, imports or calls it.
 *
 * Parity status: the reference ships NO golden vectors or tests for this path,
 * so this restatement is pinned against outputs of the reference itself, generated in the build
 * container by importing the reference's own Python oracle `ms_deform_attn_core_pytorch`
 * (tests/golden/make_golden.py -> tests/golden/ npz fixtures) and, on the GPU box, against the
 * unmodified reference CUDA kernel compiled from /root/reference (oracle/_ref/libmsda_refcuda.so).
 *
 * What is restated (reference file:line, all under third_party/adet/layers/):
 *   - csrc/DeformAttn/ms_deform_im2col_cuda.cuh:237-299  ms_deformable_im2col_gpu_kernel
 *         loop order b,q,m,c -> l -> p ; h_im/w_im ; in-range test ; accumulate
 *   - csrc/DeformAttn/ms_deform_im2col_cuda.cuh:33-84    ms_deform_attn_im2col_bilinear
 *         floor, corner validity, weights, weighted corner sum
 *   - ms_deform_attn.py:138-147  softmax over L*P and offset->location (both reference_points forms)
 *
 * Rounding contract.  The arithmetic below reproduces the reference kernel AS COMPILED by
 * nvcc 12.9 for sm_100a (SASS read with cuobjdump; see DESIGN.md "bit-exact contract"):
 *     h_im = FFMA(loc_h, (float)H, -0.5)                 one rounding  (cuh:285)
 *     w_im = FFMA(loc_w, (float)W, -0.5)                 one rounding  (cuh:286)
 *     lh = h_im - (float)floor(h_im) ; hh = 1 - lh  (and lw, hw)        (cuh:43-45)
 *     w1 = hh*hw  w2 = hh*lw  w3 = lh*hw  w4 = lh*lw                    (cuh:80)
 *     val = FFMA(w4,v4, FFMA(w3,v3, FFMA(w1,v1, FMUL(w2,v2))))          (cuh:82)
 *     col = FFMA(attn, val, col)                                        (cuh:290)
 * Compile with -ffp-contract=off so the C compiler adds no contraction of its own; every fused
 * operation is written as an explicit fmaf()/fma().
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define MSDA_ORACLE_VERSION 1

int msda_oracle_version(void) { return MSDA_ORACLE_VERSION; }

int msda_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * Per-sample index record: what the reference kernel derives from one sampling location.
 *   h_low, w_low  : floor(h_im), floor(w_im)                       (cuh:39-40)  "sampling indices"
 *   in_range      : h_im>-1 && w_im>-1 && h_im<H && w_im<W         (cuh:288)
 *   corner_mask   : bit0 (h_low,w_low) bit1 (h_low,w_high) bit2 (h_high,w_low) bit3 (h_high,w_high)
 *                   valid per cuh:56,62,68,74; 0 when !in_range
 *   level_offset  : level_start_index[l] * M * D  element offset    (cuh:274-278) "level offsets"
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t h_low, w_low;
  int32_t in_range;
  int32_t corner_mask;
  int64_t level_offset;
} msda_oracle_index_t;

static inline void sample_index_f32(float loc_w, float loc_h, int H, int W, float *h_im_o, float *w_im_o,
                                    int *h_low_o, int *w_low_o, int *in_range_o, int *mask_o) {
  const float Hf = (float)H, Wf = (float)W;
  const float h_im = fmaf(loc_h, Hf, -0.5f);
  const float w_im = fmaf(loc_w, Wf, -0.5f);
  const int in_range = (h_im > -1.0f && w_im > -1.0f && h_im < Hf && w_im < Wf);
  int h_low = 0, w_low = 0, mask = 0;
  if (in_range) {
    h_low = (int)floorf(h_im);
    w_low = (int)floorf(w_im);
    const int h_high = h_low + 1, w_high = w_low + 1;
    if (h_low >= 0 && w_low >= 0) mask |= 1;
    if (h_low >= 0 && w_high <= W - 1) mask |= 2;
    if (h_high <= H - 1 && w_low >= 0) mask |= 4;
    if (h_high <= H - 1 && w_high <= W - 1) mask |= 8;
  }
  *h_im_o = h_im; *w_im_o = w_im; *h_low_o = h_low; *w_low_o = w_low;
  *in_range_o = in_range; *mask_o = mask;
}

/* loc: (N, Lq, M, L, P, 2) with x(=w) first, y(=h) second (ms_deform_attn.py:143-144, cuh:282-283) */
int msda_oracle_sample_index_f32(const float *loc, const int64_t *shapes, const int64_t *lsi, int N, int Lq,
                                 int M, int D, int L, int P, msda_oracle_index_t *out) {
  const int64_t total = (int64_t)N * Lq * M;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < total; ++i) {
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      for (int p = 0; p < P; ++p) {
        const int64_t s = (i * L + l) * P + p;
        float h_im, w_im; int h_low, w_low, in_range, mask;
        sample_index_f32(loc[2 * s], loc[2 * s + 1], H, W, &h_im, &w_im, &h_low, &w_low, &in_range, &mask);
        out[s].h_low = h_low; out[s].w_low = w_low; out[s].in_range = in_range; out[s].corner_mask = mask;
        out[s].level_offset = (int64_t)((int)lsi[l]) * M * D;
      }
    }
  }
  return 0;
}

/* Literal fp32 restatement of the reference forward kernel (cuh:237-299 + :33-84), one output
 * element per (b,q,m,c).  value: (N,S,M,D)  loc: (N,Lq,M,L,P,2)  attn: (N,Lq,M,L,P)  out: (N,Lq,M*D). */
int msda_oracle_forward_f32(const float *value, const int64_t *shapes, const int64_t *lsi, const float *loc,
                            const float *attn, int N, int S, int M, int D, int L, int Lq, int P, float *out) {
  const int64_t total = (int64_t)N * Lq * M;
  const int qid_stride = M * D;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < total; ++i) {
    const int m = (int)(i % M);
    const int64_t b = i / ((int64_t)M * Lq);
    float *o = out + i * D;
    for (int c = 0; c < D; ++c) o[c] = 0.0f;
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const float *vbase = value + ((size_t)b * S + (size_t)(int)lsi[l]) * qid_stride;
      const int w_stride = qid_stride, h_stride = W * qid_stride;
      for (int p = 0; p < P; ++p) {
        const int64_t s = (i * L + l) * P + p;
        float h_im, w_im; int h_low, w_low, in_range, mask;
        sample_index_f32(loc[2 * s], loc[2 * s + 1], H, W, &h_im, &w_im, &h_low, &w_low, &in_range, &mask);
        if (!in_range) continue;
        const float a = attn[s];
        const float lh = h_im - (float)h_low, lw = w_im - (float)w_low;
        const float hh = 1.0f - lh, hw = 1.0f - lw;
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const ptrdiff_t p1 = (ptrdiff_t)h_low * h_stride + (ptrdiff_t)w_low * w_stride + m * D;
        const float *c1 = vbase + p1, *c2 = c1 + w_stride, *c3 = c1 + h_stride, *c4 = c3 + w_stride;
        for (int c = 0; c < D; ++c) {
          const float v1 = (mask & 1) ? c1[c] : 0.0f;
          const float v2 = (mask & 2) ? c2[c] : 0.0f;
          const float v3 = (mask & 4) ? c3[c] : 0.0f;
          const float v4 = (mask & 8) ? c4[c] : 0.0f;
          float t = w2 * v2;
          t = fmaf(w1, v1, t);
          t = fmaf(w3, v3, t);
          t = fmaf(w4, v4, t);
          o[c] = fmaf(a, t, o[c]);
        }
      }
    }
  }
  return 0;
}

/* Same kernel instantiated for double, as the reference does through AT_DISPATCH_FLOATING_TYPES
 * (ms_deform_attn_cuda.cu:64).  nvcc contracts the double expressions the same way (DFMA). */
int msda_oracle_forward_f64(const double *value, const int64_t *shapes, const int64_t *lsi, const double *loc,
                            const double *attn, int N, int S, int M, int D, int L, int Lq, int P, double *out) {
  const int64_t total = (int64_t)N * Lq * M;
  const int qid_stride = M * D;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < total; ++i) {
    const int m = (int)(i % M);
    const int64_t b = i / ((int64_t)M * Lq);
    double *o = out + i * D;
    for (int c = 0; c < D; ++c) o[c] = 0.0;
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const double *vbase = value + ((size_t)b * S + (size_t)(int)lsi[l]) * qid_stride;
      const int w_stride = qid_stride, h_stride = W * qid_stride;
      for (int p = 0; p < P; ++p) {
        const int64_t s = (i * L + l) * P + p;
        const double h_im = fma(loc[2 * s + 1], (double)H, -0.5);
        const double w_im = fma(loc[2 * s], (double)W, -0.5);
        if (!(h_im > -1 && w_im > -1 && h_im < H && w_im < W)) continue;
        const int h_low = (int)floor(h_im), w_low = (int)floor(w_im);
        const int h_high = h_low + 1, w_high = w_low + 1;
        const double a = attn[s];
        const double lh = h_im - h_low, lw = w_im - w_low, hh = 1 - lh, hw = 1 - lw;
        const double w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const ptrdiff_t p1 = (ptrdiff_t)h_low * h_stride + (ptrdiff_t)w_low * w_stride + m * D;
        const double *c1 = vbase + p1, *c2 = c1 + w_stride, *c3 = c1 + h_stride, *c4 = c3 + w_stride;
        const int m1 = h_low >= 0 && w_low >= 0, m2 = h_low >= 0 && w_high <= W - 1;
        const int m3 = h_high <= H - 1 && w_low >= 0, m4 = h_high <= H - 1 && w_high <= W - 1;
        for (int c = 0; c < D; ++c) {
          const double v1 = m1 ? c1[c] : 0.0, v2 = m2 ? c2[c] : 0.0, v3 = m3 ? c3[c] : 0.0, v4 = m4 ? c4[c] : 0.0;
          double t = w2 * v2;
          t = fma(w1, v1, t);
          t = fma(w3, v3, t);
          t = fma(w4, v4, t);
          o[c] = fma(a, t, o[c]);
        }
      }
    }
  }
  return 0;
}

/* ---- bf16 storage helpers (round-to-nearest-even, same as __float2bfloat16_rn) ---- */
static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = ((uint32_t)h) << 16; float f; memcpy(&f, &u, 4); return f;
}
static inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fffu;          /* NaN -> canonical */
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

/* bf16 value/out, fp32 loc/attn, fp32 arithmetic in the reference kernel's order, output rounded once
 * to bf16.  The reference has no half path (ms_deform_attn_cuda.cu:64); this is the fp32 kernel applied
 * to the up-cast bf16 value, which is how the bf16 config is defined (last bullet). */
int msda_oracle_forward_bf16(const uint16_t *value, const int64_t *shapes, const int64_t *lsi, const float *loc,
                             const float *attn, int N, int S, int M, int D, int L, int Lq, int P, uint16_t *out) {
  const int64_t total = (int64_t)N * Lq * M;
  const int qid_stride = M * D;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < total; ++i) {
    const int m = (int)(i % M);
    const int64_t b = i / ((int64_t)M * Lq);
    float acc[1024];
    if (D > 1024) continue;
    for (int c = 0; c < D; ++c) acc[c] = 0.0f;
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const uint16_t *vbase = value + ((size_t)b * S + (size_t)(int)lsi[l]) * qid_stride;
      const int w_stride = qid_stride, h_stride = W * qid_stride;
      for (int p = 0; p < P; ++p) {
        const int64_t s = (i * L + l) * P + p;
        float h_im, w_im; int h_low, w_low, in_range, mask;
        sample_index_f32(loc[2 * s], loc[2 * s + 1], H, W, &h_im, &w_im, &h_low, &w_low, &in_range, &mask);
        if (!in_range) continue;
        const float a = attn[s];
        const float lh = h_im - (float)h_low, lw = w_im - (float)w_low;
        const float hh = 1.0f - lh, hw = 1.0f - lw;
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const ptrdiff_t p1 = (ptrdiff_t)h_low * h_stride + (ptrdiff_t)w_low * w_stride + m * D;
        const uint16_t *c1 = vbase + p1, *c2 = c1 + w_stride, *c3 = c1 + h_stride, *c4 = c3 + w_stride;
        for (int c = 0; c < D; ++c) {
          const float v1 = (mask & 1) ? bf16_to_f32(c1[c]) : 0.0f;
          const float v2 = (mask & 2) ? bf16_to_f32(c2[c]) : 0.0f;
          const float v3 = (mask & 4) ? bf16_to_f32(c3[c]) : 0.0f;
          const float v4 = (mask & 8) ? bf16_to_f32(c4[c]) : 0.0f;
          float t = w2 * v2;
          t = fmaf(w1, v1, t);
          t = fmaf(w3, v3, t);
          t = fmaf(w4, v4, t);
          acc[c] = fmaf(a, t, acc[c]);
        }
      }
    }
    for (int c = 0; c < D; ++c) out[i * D + c] = f32_to_bf16_rn(acc[c]);
  }
  return (D > 1024) ? -1 : 0;
}

/* Textbook double-precision evaluation from fp32 inputs (two-step h_im, no fusion): the
 * "mathematically exact" answer used to express tolerances (max|a-b|/max|b|). */
int msda_oracle_forward_exact(const float *value, const int64_t *shapes, const int64_t *lsi, const float *loc,
                              const float *attn, int N, int S, int M, int D, int L, int Lq, int P, double *out) {
  const int64_t total = (int64_t)N * Lq * M;
  const int qid_stride = M * D;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < total; ++i) {
    const int m = (int)(i % M);
    const int64_t b = i / ((int64_t)M * Lq);
    double *o = out + i * D;
    for (int c = 0; c < D; ++c) o[c] = 0.0;
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const float *vbase = value + ((size_t)b * S + (size_t)(int)lsi[l]) * qid_stride;
      for (int p = 0; p < P; ++p) {
        const int64_t s = (i * L + l) * P + p;
        const double h_im = (double)loc[2 * s + 1] * H - 0.5, w_im = (double)loc[2 * s] * W - 0.5;
        if (!(h_im > -1 && w_im > -1 && h_im < H && w_im < W)) continue;
        const int h_low = (int)floor(h_im), w_low = (int)floor(w_im);
        const double lh = h_im - h_low, lw = w_im - w_low, a = attn[s];
        for (int dy = 0; dy < 2; ++dy)
          for (int dx = 0; dx < 2; ++dx) {
            const int y = h_low + dy, x = w_low + dx;
            if (y < 0 || y > H - 1 || x < 0 || x > W - 1) continue;
            const double wt = (dy ? lh : 1 - lh) * (dx ? lw : 1 - lw) * a;
            const float *v = vbase + ((size_t)y * W + x) * qid_stride + m * D;
            for (int c = 0; c < D; ++c) o[c] += wt * v[c];
          }
      }
    }
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Module glue (ms_deform_attn.py:138-147), fp32, every operation rounded separately as the eager
 * PyTorch ops are (no contraction across torch kernels):
 *   ref_dim == 2:  loc = ref + off / (float)(W_l, H_l)                                  (:141-144)
 *   ref_dim == 4:  loc = ref[:2] + ((off * (1/P)) * ref[2:]) * 0.5                      (:145-147)
 *       (`tensor / python_int` on CUDA multiplies by the fp32 reciprocal of the scalar; exact for P=2^k)
 * ref: (N, Lq, L, ref_dim)   off: (N, Lq, M, L, P, 2)   loc out: (N, Lq, M, L, P, 2)
 * --------------------------------------------------------------------------------------------- */
int msda_oracle_locations_f32(const float *ref, int ref_dim, const float *off, const int64_t *shapes, int N,
                              int Lq, int M, int L, int P, float *loc) {
  if (ref_dim != 2 && ref_dim != 4) return -1;
  const float invP = 1.0f / (float)P;
  const int64_t NQ = (int64_t)N * Lq;
#pragma omp parallel for schedule(static)
  for (int64_t nq = 0; nq < NQ; ++nq)
    for (int m = 0; m < M; ++m)
      for (int l = 0; l < L; ++l) {
        const float *r = ref + (nq * L + l) * ref_dim;
        const float Wf = (float)shapes[2 * l + 1], Hf = (float)shapes[2 * l];
        for (int p = 0; p < P; ++p) {
          const int64_t s = ((nq * M + m) * L + l) * P + p;
          const float ox = off[2 * s], oy = off[2 * s + 1];
          if (ref_dim == 2) {
            loc[2 * s] = r[0] + ox / Wf;
            loc[2 * s + 1] = r[1] + oy / Hf;
          } else {
            loc[2 * s] = r[0] + ((ox * invP) * r[2]) * 0.5f;
            loc[2 * s + 1] = r[1] + ((oy * invP) * r[3]) * 0.5f;
          }
        }
      }
  return 0;
}

/* softmax over the last dim (cols = L*P) of a (rows, cols) fp32 matrix, ms_deform_attn.py:139.
 * exp(x - max) / sum in fp32 with a sequential sum; PyTorch's reduction order differs, so parity with
 * torch / the CUDA path is tolerance-based (<= 4 ulp), not bit-exact. */
int msda_oracle_softmax_f32(const float *logits, int64_t rows, int cols, float *out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float *x = logits + r * cols; float *y = out + r * cols;
    float mx = x[0];
    for (int c = 1; c < cols; ++c) mx = x[c] > mx ? x[c] : mx;
    float sum = 0.0f;
    for (int c = 0; c < cols; ++c) { y[c] = expf(x[c] - mx); sum += y[c]; }
    for (int c = 0; c < cols; ++c) y[c] = y[c] / sum;
  }
  return 0;
}
