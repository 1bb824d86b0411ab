/*
 * msda_b200.h -- C-ABI of the B200-native (sm_100a) multi-scale deformable attention library.
 *
 * This is the drop-in boundary for the GoMatching / DeepSolo MSDeformAttn hot path.  It replaces,
 * one for one, the native extension `adet._C` of the reference:
 *
 *   reference interface                                            replaced by
 *   ------------------------------------------------------------   -----------------------------------
 *   ms_deformable_im2col_cuda<scalar_t>(stream, value, shapes,      msda_b200_forward_f32
 *     lsi, loc, attn, batch, spatial_size, heads, channels,         msda_b200_forward_bf16
 *     levels, query, point, data_col)
 *     third_party/adet/layers/csrc/DeformAttn/ms_deform_im2col_cuda.cuh:923-954
 *   ms_deform_attn_cuda_forward(value, shapes, lsi, loc, attn,      (same two; the caller allocates the
 *     im2col_step)   .../ms_deform_attn_cuda.cu:20-80                output, no memset, one launch for
 *   adet._C.ms_deform_attn_forward   csrc/vision.cpp:52-53           all N -- im2col_step is accepted and
 *   ms_deform_attn_forward dispatch  .../ms_deform_attn.h:20-39      ignored by the Python layer)
 *   MSDeformAttn.forward glue (softmax over L*P, offsets ->         msda_b200_forward_fused_f32
 *     locations, both reference_points forms)                       msda_b200_forward_fused_bf16
 *     third_party/adet/layers/ms_deform_attn.py:137-152
 *   ms_deformable_col2im_cuda / ms_deform_attn_cuda_backward        msda_b200_backward_f32
 *     .../ms_deform_im2col_cuda.cuh:956-1327, ms_deform_attn_cuda.cu:83-153
 *   adet._C.ms_deform_attn_backward  csrc/vision.cpp:54-55
 *
 * Conventions (identical to the reference launcher unless noted):
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`; all tensors contiguous row-major
 *   - value   (N, S, M, D)          S = sum_l H_l*W_l value tokens, M heads, D channels per head
 *   - shapes  (L, 2) int64 [H, W]   read on the device, never copied to the host (no sync)
 *   - lsi     (L,)   int64          level start index in tokens
 *   - loc     (N, Lq, M, L, P, 2)   float32, x (width) first, y (height) second, normalised to [0,1]
 *   - attn    (N, Lq, M, L, P)      float32
 *   - out     (N, Lq, M*D)          written exactly once per element; need not be zero-initialised
 *   - `stream` is a cudaStream_t passed as void*; the call only enqueues work (no synchronisation,
 *     CUDA-graph capturable) and is re-entrant
 *   - return value: 0 on success, a positive cudaError_t on a CUDA failure, a negative MSDA_E_* code on
 *     an argument error.  msda_b200_error_string() describes either.
 *   - the sampling-index contract (bit-exact with the reference binary built by nvcc 12.9 for sm_100a):
 *       h_im = fmaf(loc_y, (float)H_l, -0.5f), w_im likewise; h_low = floor(h_im); a sample contributes
 *       iff h_im > -1 && w_im > -1 && h_im < H && w_im < W; corners outside the map read as zero.
 *   - there is NO CPU path: the library does arithmetic only on the GPU.
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_B200_ABI_VERSION 4

/* argument errors (negative so they cannot collide with cudaError_t) */
#define MSDA_E_NULLPTR   (-1)   /* a required pointer is NULL                                   */
#define MSDA_E_DIMS      (-2)   /* a dimension is <= 0 or the product overflows 31-bit indexing */
#define MSDA_E_ALIGN     (-3)   /* a pointer is not 16-byte aligned                             */
#define MSDA_E_REFDIM    (-4)   /* reference_points last dim is neither 2 nor 4                 */
#define MSDA_E_UNSUPPORTED (-5) /* combination has no kernel (e.g. bf16 with odd D)             */
#define MSDA_E_NOCUDA    (-6)   /* no CUDA device / driver                                      */

/* Optional launch tuning.  Pass NULL for the built-in heuristics.  Never changes results. */
typedef struct msda_b200_tuning {
  int32_t mode;          /* 0 auto | 1 linear query tiles | 2 pyramid 2-D tiles (needs Lq == S) | 3 generic kernel |
                            4 staged: value windows in shared memory (fp32 encoder self-attention, D=32 L=4 P=4);
                            tile_h then = number of query levels staged (default 1), the rest runs mode 1 |
                            5 pipelined: producer / consumer version of 4 (TMA fills, no CTA-wide barriers); needs
                            msda_b200_staged_set_host_shapes, else runs mode 1 */
  int32_t tile_h;        /* pyramid tile height in pixels (0 = default)                            */
  int32_t tile_w;        /* pyramid tile width in pixels, multiple of 4 (0 = default)              */
  int32_t tile_q;        /* linear tile length in queries (0 = default)                            */
  int32_t ctas_per_sm;   /* persistent grid = SM count * ctas_per_sm (0 = default)                 */
  int32_t variant;       /* kernel instantiation selector, see msda_b200_variant_count (0 = default)*/
  int32_t reserved[2];   /* [0] = 1: force the runtime-L*P tiled kernel; [1] = 1: contiguous raster tile walk (diagnostics) */
} msda_b200_tuning_t;

int         msda_b200_abi_version(void);
const char* msda_b200_error_string(int code);
/* number of SMs of the current device, or a negative error */
int         msda_b200_sm_count(void);
int         msda_b200_variant_count(void);
/* Optional hint for tuning.mode = 4: the (H_l, W_l) pairs and level start indices as HOST arrays (L = 4).  With it the
 * staged kernel fills its shared-memory windows with TMA (the tensor maps are encoded on the host and need the level
 * geometry, which the operator API only hands over as a device tensor: ms_deform_attn.py:117-123); without it, or if
 * the hint does not match S, the windows are filled with cp.async.  Pass NULL to clear.  The hint is stored per calling
 * thread (set it on the thread that launches). */
void        msda_b200_staged_set_host_shapes(const int64_t* shapes_host, const int64_t* lsi_host, int L);

/* ---- core operator: the _MSDeformAttnFunction boundary (ms_deform_attn.py:20-27) ------------------ */
int msda_b200_forward_f32(const float* value, const int64_t* shapes, const int64_t* lsi,
                          const float* loc, const float* attn,
                          int N, int S, int M, int D, int L, int Lq, int P,
                          float* out, void* stream);

/* value/out stored as bf16 (raw uint16 bit patterns); loc/attn stay fp32; fp32 accumulation, one
 * final round-to-nearest-even.  The reference has no half path (ms_deform_attn_cuda.cu:64): results
 * equal the fp32 kernel applied to the up-cast value, rounded once. */
int msda_b200_forward_bf16(const void* value_bf16, const int64_t* shapes, const int64_t* lsi,
                           const float* loc, const float* attn,
                           int N, int S, int M, int D, int L, int Lq, int P,
                           void* out_bf16, void* stream);

/* float64, as the reference's AT_DISPATCH_FLOATING_TYPES also instantiates (ms_deform_attn_cuda.cu:64); a plain
 * one-thread-per-element kernel in double arithmetic -- completeness, not a performance path */
int msda_b200_forward_f64(const double* value, const int64_t* shapes, const int64_t* lsi,
                          const double* loc, const double* attn,
                          int N, int S, int M, int D, int L, int Lq, int P,
                          double* out, void* stream);

/* same two with explicit tuning (used by bench.py and the tests to pin a variant) */
int msda_b200_forward_f32_ex(const float* value, const int64_t* shapes, const int64_t* lsi,
                             const float* loc, const float* attn,
                             int N, int S, int M, int D, int L, int Lq, int P,
                             float* out, void* stream, const msda_b200_tuning_t* tuning);
int msda_b200_forward_bf16_ex(const void* value_bf16, const int64_t* shapes, const int64_t* lsi,
                              const float* loc, const float* attn,
                              int N, int S, int M, int D, int L, int Lq, int P,
                              void* out_bf16, void* stream, const msda_b200_tuning_t* tuning);

/* ---- fused operator: softmax over L*P + offsets->locations + sampler in ONE kernel -----------------
 *   ref      (N, Lq, L, ref_dim)  ref_dim 2: loc = ref + off / (W_l, H_l)
 *                                 ref_dim 4: loc = ref[:2] + off / P * ref[2:] * 0.5
 *   offsets  (N, Lq, M, L, P, 2)  raw output of the sampling_offsets projection
 *   logits   (N, Lq, M, L*P)      raw output of the attention_weights projection (pre-softmax)
 * Locations are formed with the same IEEE operations, in the same order, as the eager reference
 * (true division, then add; no reciprocal, no FMA), so sampling indices stay bit-exact. */
int msda_b200_forward_fused_f32(const float* value, const int64_t* shapes, const int64_t* lsi,
                                const float* ref, int ref_dim, const float* offsets, const float* logits,
                                int N, int S, int M, int D, int L, int Lq, int P,
                                float* out, void* stream, const msda_b200_tuning_t* tuning);
int msda_b200_forward_fused_bf16(const void* value_bf16, const int64_t* shapes, const int64_t* lsi,
                                 const float* ref, int ref_dim, const float* offsets, const float* logits,
                                 int N, int S, int M, int D, int L, int Lq, int P,
                                 void* out_bf16, void* stream, const msda_b200_tuning_t* tuning);

/* same, with offsets / logits given as row-pitched views: row (b,q) of offsets starts at offsets + (b*Lq+q)*off_pitch
 * floats (M*L*P*2 used), of logits at logits + (b*Lq+q)*logit_pitch (M*L*P used).  Lets the caller compute both
 * projections of `query` (ms_deform_attn.py:137-138) as ONE 256->384 GEMM and hand over the two column slices
 * without a copy.  Only the DeepSolo-shape kernels (D=32, L=4, P=4) take pitches; otherwise MSDA_E_UNSUPPORTED. */
int msda_b200_forward_fused_pitched_f32(const float* value, const int64_t* shapes, const int64_t* lsi,
                                        const float* ref, int ref_dim, const float* offsets, int off_pitch,
                                        const float* logits, int logit_pitch,
                                        int N, int S, int M, int D, int L, int Lq, int P,
                                        float* out, void* stream, const msda_b200_tuning_t* tuning);
int msda_b200_forward_fused_pitched_bf16(const void* value_bf16, const int64_t* shapes, const int64_t* lsi,
                                         const float* ref, int ref_dim, const float* offsets, int off_pitch,
                                         const float* logits, int logit_pitch,
                                         int N, int S, int M, int D, int L, int Lq, int P,
                                         void* out_bf16, void* stream, const msda_b200_tuning_t* tuning);

/* ---- glue kernels on their own (what the fused kernel does in registers), for tests/inspection ---- */
/* loc_out (N,Lq,M,L,P,2), attn_out (N,Lq,M,L,P); either output may be NULL.  lanes_per_unit selects the
 * lane layout being exercised: 8 = the fp32 D=32 kernels, 4 = the bf16 D=32 kernels, 16 = fp32 D=64. */
int msda_b200_locations_softmax_f32(const int64_t* shapes, const float* ref, int ref_dim,
                                    const float* offsets, const float* logits,
                                    int N, int M, int L, int Lq, int P, int lanes_per_unit,
                                    float* loc_out, float* attn_out, void* stream);

/* ---- sampling-index dump: the SAME device function the forward kernels use ------------------------
 * One record per sample (N,Lq,M,L,P), layout identical to oracle/msda_oracle.c msda_oracle_index_t:
 *   int32 h_low, w_low, in_range, corner_mask ; int64 level_offset (= lsi[l]*M*D elements)
 * (h_low/w_low/corner_mask are 0 when !in_range).  cuh:33-84, :272-296. */
typedef struct msda_b200_index {
  int32_t h_low, w_low, in_range, corner_mask;
  int64_t level_offset;
} msda_b200_index_t;
int msda_b200_sample_index_f32(const float* loc, const int64_t* shapes, const int64_t* lsi,
                               int N, int Lq, int M, int D, int L, int P,
                               msda_b200_index_t* out, void* stream);

/* ---- backward of the core operator (ms_deform_attn_cuda.cu:83-153; cuh:301-920) --------------------
 * grad_value (N,S,M,D) MUST be zero-initialised by the caller (accumulated with atomics exactly like
 * the reference); grad_loc (N,Lq,M,L,P,2) and grad_attn (N,Lq,M,L,P) are written once per element. */
int msda_b200_backward_f32(const float* value, const int64_t* shapes, const int64_t* lsi,
                           const float* loc, const float* attn, const float* grad_out,
                           int N, int S, int M, int D, int L, int Lq, int P,
                           float* grad_value, float* grad_loc, float* grad_attn, void* stream);

/* ---- bordering projections on the tensor cores (tcgen05 / TMEM / TMA, 3xTF32 = fp32-grade accuracy) ----------
 * Replace the four nn.Linear calls of MSDeformAttn.forward (ms_deform_attn.py:133 value_proj, :137 sampling_offsets,
 * :138 attention_weights, :153 output_proj), which the reference runs as fp32 cuBLAS GEMMs:
 *     y[M, N] (row pitch ldy floats) = x[M, K] (row pitch ldx floats) . w[N, K]^T + bias[N]
 * w is given pre-split into its TF32 head and remainder (msda_b200_linear_split_weight_f32, once per weight).
 * row_zero (M bytes, may be NULL): rows with a non-zero byte are written as zeros -- the padding-mask
 * masked_fill of ms_deform_attn.py:135 fused into the value_proj epilogue.
 * K % 16 == 0, N % 32 == 0, N <= 1024, ldx % 4 == 0, ldy % 4 == 0, 16-byte aligned pointers. */
int msda_b200_linear_split_weight_f32(const float* w, int N, int K, float* w_hi, float* w_lo, void* stream);
int msda_b200_linear_f32(const float* x, int ldx, const float* w_hi, const float* w_lo, const float* bias,
                         const unsigned char* row_zero, int M, int N, int K, float* y, int ldy, void* stream);
/* same GEMM with y = max(y, 0) in the epilogue: linear1 + ReLU of the encoder / decoder feed-forward block
 * (third_party/adet/layers/deformable_transformer.py:248-252 forward_ffn, :406-411) */
int msda_b200_linear_relu_f32(const float* x, int ldx, const float* w_hi, const float* w_lo, const float* bias,
                              int M, int N, int K, float* y, int ldy, void* stream);
/* diagnostics: per-CTA clock64 stamps of the following msda_b200_linear_f32 launches are written to buf
 * (device memory, [CTAs][8] int64: start, first stage full, split done, last MMA issued, accumulator ready,
 * epilogue done); NULL switches tracing off.  tools/gemm_trace.py. */
void msda_b200_linear_set_trace(long long* buf);

/* ---- attention core for the short sequences of DeepSolo's point-query decoder ---------------------------------------
 * attn_intra (25 points of a proposal) and attn_inter (100 proposals at a point index),
 * third_party/adet/layers/deformable_transformer.py:386-404: nn.MultiheadAttention's
 *     out[b,i,h,:] = softmax_j((q[b,i,h,:] * head_dim^-1/2) . k[b,j,h,:]) . v[b,j,h,:]
 * between its input and output projections (those run on msda_b200_linear_f32).  q, k, v: fp32 column slices of a
 * projection output with row pitch ld floats (head h at columns h*head_dim ..); row of (batch b, position i) =
 * b * batch_stride + i * seq_stride, so a strided sequence order needs no transposed copy; out likewise with pitch ldo.
 * head_dim = 32, L <= 128. */
int msda_b200_small_mha_f32(const float* q, const float* k, const float* v, int ld, float* out, int ldo,
                            int B, int L, int H, int head_dim, long long batch_stride, long long seq_stride, void* stream);

/* ---- elementwise glue of the point-query decoder, one launch each -------------------------------------------------
 * point_pos_embed: gen_point_pos_embed(reference_points_input[:, :, :, 0, :], d_model, temp)
 *   (third_party/adet/modeling/model/utils.py:24-37, called at deformable_transformer.py:477):
 *     out[p][xy * half_dim + i] = (i odd ? cos : sin)( ref[p][xy] * ratio[b(p)][xy] * 2 pi / dim_t[i] )
 *   ref: (points, 2) fp32; ratio: valid ratios of level 0, element [b * ratio_stride + xy], b = p / points_per_batch, or
 *   NULL (= 1); dim_t: (half_dim) fp32 = temp ** (2 * (i // 2) / half_dim) as the caller's framework computes it; out:
 *   (points, 2 * half_dim).
 * refine_points: (tmp + inverse_sigmoid(ref)).sigmoid() over n fp32 elements (deformable_transformer.py:483-486,
 *   adet/utils/misc.py:115-119; eps = 1e-5 there).
 * Both repeat the eager fp32 operation order and are bit-identical to it. */
int msda_b200_point_pos_embed_f32(const float* ref, const float* ratio, int ratio_stride, const float* dim_t, long long points,
                                  int points_per_batch, int half_dim, float* out, void* stream);
int msda_b200_refine_points_f32(const float* tmp, const float* ref, long long n, float eps, float* out, void* stream);

/* ---- neighbour-paired bf16 value layout: an explicit operator MODE for the bf16 configuration ---------------------
 * (BASELINE.json config 3; bar 2e-2 relative vs the fp32 reference.)  The reference has no half/bf16 path
 * (AT_DISPATCH_FLOATING_TYPES, ms_deform_attn_cuda.cu:64); this mode stores value as
 *     paired[b][m][p] = { value[b][p][m][0..D-1], value[b][p+1][m][0..D-1] }  bf16, 128 bytes, second half zero at the last
 *                                                                            pixel of an image row
 * i.e. (N, M, S, 2, D) instead of (N, S, M, D): both horizontal neighbours of a bilinear footprint share one 128-byte
 * line (2 instead of 4 L1 wavefronts per sample) and a head's map is contiguous.  S, spatial_shapes and
 * level_start_index keep their meaning.  D = 32 (pair_value); D = 32, L = 4, P = 4 (samplers).  Sampling indices are
 * the same bits as every other kernel of this library; outputs are bf16 (N, Lq, M*D).
 *   msda_b200_pair_value_bf16          builds the layout from value (N,S,M,D) fp32 (value_is_bf16 = 0) or bf16 (1)
 *   msda_b200_forward_paired_bf16      core operator on it   (sampling_loc / attn_weight as in msda_b200_forward_f32)
 *   msda_b200_forward_fused_paired_bf16  softmax + offsets->locations + sampler (as msda_b200_forward_fused_f32) */
int msda_b200_pair_value_bf16(const void* value, int value_is_bf16, const int64_t* shapes, const int64_t* lsi,
                              int N, int S, int M, int D, int L, void* paired, void* stream);
int msda_b200_forward_paired_bf16(const void* paired, const int64_t* shapes, const int64_t* lsi,
                                  const float* sampling_loc, const float* attn_weight,
                                  int N, int S, int M, int D, int L, int Lq, int P, void* out, void* stream);
int msda_b200_forward_fused_paired_bf16(const void* paired, const int64_t* shapes, const int64_t* lsi,
                                        const float* reference_points, int ref_dim,
                                        const float* sampling_offsets, const float* attention_logits,
                                        int N, int S, int M, int D, int L, int Lq, int P, void* out, void* stream);

/* ---- shape guard of the TMA window kernel (tuning mode 5, the default for fp32 encoder self-attention) ---------
 * The window kernel builds its tensor maps from a HOST copy of the level shapes (msda_b200_staged_set_host_shapes);
 * the operator's contract is the DEVICE tensors, which the reference reads in-kernel (ms_deform_attn_cuda.cu:20-80,
 * ms_deform_im2col_cuda.cuh:272-278).  Every launch therefore checks the two against each other on the device; on a
 * mismatch the window kernel does nothing, the register-gather launch that follows it serves every query (results are
 * always those of the device tensors), and the epoch of the launch is stored in pinned host memory.  This returns the
 * most recent such epoch (0: never) without synchronising -- the caller polls it to drop a stale host-side cache. */
int msda_b200_shape_mismatch_epoch(void);

/* ---- host-buffer entry: what a non-PyTorch host (the cgo/JNI/ctypes stub of INTEGRATION.md) calls --
 * All tensor pointers are HOST pointers (pinned for full PCIe speed, pageable works).  Copies the
 * inputs to a device workspace owned by the handle, runs the core forward, copies the result back,
 * and synchronises the handle's stream before returning. */
typedef struct msda_b200_host_ctx msda_b200_host_ctx_t;
int  msda_b200_host_ctx_create(msda_b200_host_ctx_t** ctx, int device);
void msda_b200_host_ctx_destroy(msda_b200_host_ctx_t* ctx);
int  msda_b200_forward_f32_host(msda_b200_host_ctx_t* ctx,
                                const float* value_host, const int64_t* shapes_host, const int64_t* lsi_host,
                                const float* loc_host, const float* attn_host,
                                int N, int S, int M, int D, int L, int Lq, int P,
                                float* out_host);

/* ---- frame batcher: the input side of the video loop -------------------------------------------------
 * Replaces, in one pass over HBM, the host-side x.astype("float32").transpose(2, 0, 1) (+ optional channel flip) of
 * GoMBatchPredictor.__call__ (gomatching/text_track_visualizer.py:313-321) and GoMatching.preprocess_image
 * (gomatching/modeling/meta_arch/gom_lstmatcher.py:159-170: (x - pixel_mean) / pixel_std, then ImageList.from_tensors
 * zero padding to Hp x Wp).  frames: device uint8 (N, H, W, 3); mean3 / std3: HOST float[3] in OUTPUT channel order;
 * out: device float32 (N, 3, Hp, Wp).  Bit-identical to the eager ops (IEEE subtract, then IEEE divide). */
int msda_b200_frames_u8_to_chw_f32(const unsigned char* frames, int N, int H, int W, int flip_channels,
                                   const float* mean3, const float* std3, int Hp, int Wp, float* out, void* stream);

/* ---- JPEG frame decode, bit-identical to Pillow / libjpeg-turbo with default settings --------------------------------
 * Replaces `read_image(path, format="BGR")` of the reference's video loop (eval.py:324-327: detectron2 -> PIL.Image.open
 * -> convert("RGB") -> numpy -> BGR).  The Huffman stage runs on the calling host thread; dequantisation, the 13-bit
 * integer inverse DCT (IJG jidctint.c), fancy chroma upsampling (jdsample.c) and the fixed-point YCbCr -> RGB conversion
 * (jdcolor.c) run on `stream`.  Baseline / extended-sequential Huffman JPEG, 8-bit, grey or 3 components, 4:4:4 / 4:2:2 /
 * 4:2:0, restart intervals; anything else returns MSDA_E_UNSUPPORTED (no fallback).
 * jpeg_info: frame size from the SOF marker (host only).  jpeg_decode: data / len = the file in HOST memory; out = DEVICE
 * uint8 (height, width, 3) contiguous, channels B,G,R if bgr else R,G,B; width / height must equal jpeg_info's.
 * Truncated or corrupt streams return MSDA_E_DIMS.  Scratch memory comes from the stream-ordered allocator. */
int msda_b200_jpeg_info(const unsigned char* data, size_t len, int* width, int* height, int* components);
int msda_b200_jpeg_decode_u8(const unsigned char* data, size_t len, int bgr, unsigned char* out, int width, int height,
                             void* stream);

/* ---- test-time frame resize, bit-identical to Pillow's 8-bit bilinear resample ---------------------------------------
 * Replaces `self.aug.get_transform(x).apply_image(x)` (ResizeShortestEdge -> PIL.Image.resize(BILINEAR)) of the
 * reference's predictors (gomatching/text_track_visualizer.py:283-284, :318-319).  One separable pass per call:
 *     out = clip8((2^21 + sum_{i < count} in[first + i] * k[i]) >> 22)
 * in: device uint8 (N, H, W, C), C in {1, 3, 4}; axis 1 = horizontal (out (N, H, out_size, C)), axis 0 = vertical
 * (out (N, out_size, W, C)); bounds: device int32 (out_size, 2) = [first, count]; coeffs: device int32 (out_size, ksize)
 * 22-bit fixed point -- both built exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc build them
 * (gomatching_b200/video/resize.py).  Pillow runs the horizontal pass first. */
int msda_b200_resample_u8_hwc(const unsigned char* in, int N, int H, int W, int C, int axis, const int* bounds,
                              const int* coeffs, int ksize, int out_size, unsigned char* out, void* stream);

/* ---- residual add + LayerNorm in one pass: out = LayerNorm(x + y) * gamma + beta ------------------------------
 * The two eager steps after every attention / feed-forward block of the transformer at inference
 * (third_party/adet/layers/deformable_transformer.py:251-252, :272-273).  x, y (may be NULL), out: (rows, C) fp32
 * contiguous; gamma, beta: (C) or NULL; biased variance, eps inside the square root, like nn.LayerNorm.
 * C % 128 == 0, C <= 1024, 16-byte aligned pointers. */
int msda_b200_add_layernorm_f32(const float* x, const float* y, const float* gamma, const float* beta, float eps,
                                long long rows, int C, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
