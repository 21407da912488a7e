/*
 * flux_b200.h -- C ABI of libflux_b200.so: the B200 (sm_100a) FP8 Flux-DiT denoise hot path.
 *
 * The reference (aredden/flux-fp8-api) has no FFI layer: its boundary for this path is the
 * nn.Module surface.  Each entry point below names the reference code
 * (file:line under /root/reference) whose device work it replaces; the Python classes in
 * flux-fp8-api_b200/ keep the reference's signatures and call these through ctypes.
 *
 * Conventions
 *   - plain C types only; every pointer is a CUDA device pointer unless stated otherwise
 *   - bf16 tensors are passed as `const void*` to 2-byte elements, fp8 as 1-byte elements
 *   - fp8 formats: FLUXB200_E4M3 (float8_e4m3fn) / FLUXB200_E5M2 (float8_e5m2)
 *   - per-tensor scales are 0-dim fp32 tensors in device memory (as in F8Linear's buffers)
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); calls are stream-ordered,
 *     allocate nothing, keep no global mutable state besides a per-thread error string
 *   - return value: 0 on success, a negative FLUXB200_ERR_* otherwise; never throws
 */
#ifndef FLUX_B200_H_
#define FLUX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLUXB200_VERSION 100 /* 0.1.0 */

enum { FLUXB200_E4M3 = 0, FLUXB200_E5M2 = 1 };

enum {
  FLUXB200_OK = 0,
  FLUXB200_ERR_INVALID = -1,     /* bad argument (shape / alignment / null pointer)        */
  FLUXB200_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed                    */
  FLUXB200_ERR_UNSUPPORTED = -3, /* device is not sm_100 or shape outside the kernel's domain */
};

typedef void* fluxb200_stream_t;

/* Library version (FLUXB200_VERSION). */
int fluxb200_version(void);
/* Message of the last failing call on this thread ("" if none). Pointer valid until the next call. */
const char* fluxb200_last_error(void);
/* 0 when the current CUDA device can run the kernels (compute capability 10.x); fills *sm_count. */
int fluxb200_device_check(int* sm_count);

/* ---------------------------------------------------------------------------------------------
 * F8Linear.quantize_input / to_fp8_saturated   (float8_quantize.py:217-218, 220-246, 274-276)
 *   y = fp8( clamp( bf16(x * scale), -max, max ) )      -- the double rounding is reproduced
 * ------------------------------------------------------------------------------------------- */
int fluxb200_quantize(const void* x_bf16, void* y_fp8, int64_t n, const float* scale, int fmt,
                      fluxb200_stream_t stream);
/* torch.max(torch.abs(x)) (float8_quantize.py:197, 227): *amax = max(*amax, max|x|); caller zeroes. */
int fluxb200_amax(const void* x_bf16, int64_t n, float* amax, fluxb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * F8Linear.forward GEMM  (float8_quantize.py:284-292  torch._scaled_mm)
 *   acc[M,N] = A_fp8[M,K] . W_fp8[N,K]^T   (tcgen05 kind::f8f6f4, fp32 accumulate in TMEM)
 *   y = bf16( acc * (*a_scale_recip) * (*w_scale_recip) + bias )
 * followed by one of the fused epilogues, which reproduce the eager ops that follow the linear
 * in the reference block code (every intermediate rounded to bf16 as eager PyTorch does).
 * ------------------------------------------------------------------------------------------- */
enum {
  /* out[M,ldo] bf16 = y */
  FLUXB200_EPI_PLAIN = 0,
  /* out = bf16( resid + bf16( gate[b] * y ) )        modules/flux_model.py:387-396, 482-484 */
  FLUXB200_EPI_GATE_RESIDUAL = 1,
  /* out fp8 = quantize( gelu_tanh(y), *out_scale )   modules/flux_model.py:301,335,455 + next
     F8Linear's input quantisation (float8_quantize.py:274-276) */
  FLUXB200_EPI_GELU_QUANT = 2,
  /* N = 3*H*128: per-head QK-RMSNorm (fp32, eps 1e-6) + RoPE on q,k; q,k,v written as
     [B,H,seq_total,128] at sequence offset seq_offset.  modules/flux_model.py:353,164,60-65,380-382 */
  FLUXB200_EPI_QKV_ROPE = 3,
  /* SingleStreamBlock.linear1: columns [0,3*H*128) as QKV_ROPE, the remaining mlp columns as
     GELU_QUANT written to out at column out_col_offset.  modules/flux_model.py:471-480 */
  FLUXB200_EPI_LINEAR1 = 4,
};

typedef struct fluxb200_gemm_args {
  const void* a;              /* fp8 [M,K] row-major, 16-byte aligned, K % 16 == 0            */
  const void* w;              /* fp8 [N,K] row-major (F8Linear.float8_data)                    */
  const void* bias;           /* bf16 [N] or NULL                                              */
  const float* a_scale_recip; /* F8Linear.input_scale_reciprocal                               */
  const float* w_scale_recip; /* F8Linear.scale_reciprocal                                     */
  int32_t M, N, K;
  int32_t a_fmt, w_fmt;   /* FLUXB200_E4M3 / FLUXB200_E5M2                                     */
  int32_t epilogue;       /* FLUXB200_EPI_*                                                    */
  int32_t rows_per_batch; /* row r belongs to sample r / rows_per_batch (0: one sample)        */
  /* PLAIN, GATE_RESIDUAL (bf16) and GELU_QUANT / LINEAR1 (fp8) output */
  void* out;
  int64_t ldo; /* elements */
  /* GATE_RESIDUAL */
  const void* resid; /* bf16 [M,ldr]; may alias out */
  int64_t ldr;
  const void* gate; /* bf16, gate[b*gate_batch_stride + n] */
  int64_t gate_batch_stride;
  /* GELU_QUANT / LINEAR1 */
  const float* out_scale; /* next F8Linear.input_scale */
  int32_t out_fmt;
  int32_t out_col_offset;
  /* QKV_ROPE / LINEAR1 */
  void* q; /* bf16 [B,H,seq_total,128] */
  void* k;
  void* v;
  int32_t num_heads;
  int32_t seq_total;
  int32_t seq_offset;
  int32_t _pad0;
  const float* q_norm_w; /* fp32 [128]  QKNorm.query_norm.scale */
  const float* k_norm_w; /* fp32 [128]  QKNorm.key_norm.scale   */
  const void* rope_cos;  /* bf16 [*, seq_total, 64]  pe[...,0,0] (EmbedND output, bf16-rounded) */
  const void* rope_sin;  /* bf16 [*, seq_total, 64]  pe[...,1,0]                               */
  int64_t rope_batch_stride; /* elements between samples (0: shared)                           */
} fluxb200_gemm_args;

int fluxb200_f8_gemm(const fluxb200_gemm_args* args, fluxb200_stream_t stream);

/* `count` (1 or 2) problems that share N, K, fp8 formats and epilogue in ONE persistent launch: the txt and img
 * streams of a DoubleStreamBlock apply different weights to different row counts (modules/flux_model.py:369/376,
 * 387/393, 388-396), and the txt problem alone (M = 512) cannot fill 148 SMs. */
int fluxb200_f8_gemm_grouped(const fluxb200_gemm_args* args, int count, fluxb200_stream_t stream);

/* Skinny-M variant for Modulation.lin / MLPEmbedder (M = batch <= 16): weight-streaming GEMV.
 *   out[m, n] = bf16( (sum_k A[m,k]*W[n,k]) * sa * sw + bias[n] )   modules/flux_model.py:252 */
int fluxb200_f8_gemv(const void* a_fp8, int a_fmt, const void* w_fp8, int w_fmt, const void* bias_bf16,
                     const float* a_scale_recip, const float* w_scale_recip, void* out_bf16, int M, int N,
                     int K, fluxb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * All Modulation.lin of one denoise step in two launches (they share the input `vec`):
 *   for every layer l:  out[b, out_offset_l + n] = bf16( (q_l(silu(vec[b])) . W_l[n]) * sa_l * sw_l + bias_l[n] )
 * with q_l the layer's own input quantisation.  Replaces 76 x (SiLU + 3-kernel quantise + _scaled_mm)
 * per step (modules/flux_model.py:251-257 called at :363-364, :468).  `layers` is a device array.
 * aq_workspace: fp8 [num_layers, B, K] scratch.  Blocks of 64 output columns: layer l owns blocks
 * [block_start_l, block_start_l + ceil(N_l / 64)); total_blocks = sum of those.
 * ------------------------------------------------------------------------------------------- */
typedef struct fluxb200_gemv_layer {
  const void* w;              /* fp8 [N, K]                                   */
  const void* bias;           /* bf16 [N] or NULL                             */
  const float* in_qscale;     /* multiplier used to quantise silu(vec)        */
  const float* a_scale_recip; /* F8Linear.input_scale_reciprocal              */
  const float* w_scale_recip; /* F8Linear.scale_reciprocal                    */
  int32_t N;
  int32_t out_offset;
  int32_t block_start;
  int32_t _pad;
} fluxb200_gemv_layer;

int fluxb200_modulation_batched(const void* vec_bf16, const fluxb200_gemv_layer* layers, int num_layers,
                                int total_blocks, void* aq_workspace, void* out_bf16, int64_t ld_out, int B,
                                int K, int a_fmt, int w_fmt, fluxb200_stream_t stream);

/* The same for Modulation.lin layers LEFT IN bf16 (quantize_modulation = false: float8_quantize.py:346 skips the
 * swap, so modules/flux_model.py:252 runs nn.Linear): one launch for every layer of the step,
 *   out[b, out_offset_l + n] = bf16( sum_k bf16(silu(vec[b,k])) * W_l[n,k] + bias_l[n] )      (fp32 accumulate)
 * `layers[l].w` points to bf16 [N,K] row-major weights; in_qscale / a_scale_recip / w_scale_recip are ignored.
 * K % 32 == 0, K <= 4096, B <= 16; no workspace (silu is applied while vec is staged into shared memory). */
int fluxb200_modulation_batched_bf16(const void* vec_bf16, const fluxb200_gemv_layer* layers, int num_layers,
                                     int total_blocks, void* out_bf16, int64_t ld_out, int B, int K,
                                     fluxb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Per-step extras around the block stack: the skinny bf16 linears of MLPEmbedder / LastLayer and the
 * elementwise glue of Flux.forward and the Euler loop, so that a step launches no eager torch elementwise kernels.
 *
 * fluxb200_bf16_gemv:  out[b, n] = bf16( bf16( bf16( sum_k f(x[b,k]) W[n,k] + bias[n] ) + add0[b,n] ) + add1[b,n] )
 *   with f = bf16(silu(.)) when silu_input else identity; add0 / add1 optional (NULL).  B <= 16, K % 32 == 0, K <= 4096.
 *   MLPEmbedder.forward (modules/flux_model.py:154-155: in_layer, then out_layer with silu_input),
 *   LastLayer.adaLN_modulation (:495-497, 500) and `vec = time_in(..) + guidance_in(..) + vector_in(y)` (:687-697: the two
 *   bf16 additions ride on the out_layer launch).
 * fluxb200_timestep_embedding:  out[b, :] = bf16([cos | sin](float(bf16(time_factor * t[b])) * freqs))  (:95-116);
 *   freqs = exp(-ln(max_period) * arange(dim/2) / (dim/2)) as fp32 [dim/2], computed once by the caller.
 * fluxb200_euler_update:  out = bf16( img + bf16( (*dt) * pred ) )    flux_pipeline.py:651  (in place allowed).
 * ------------------------------------------------------------------------------------------- */
int fluxb200_bf16_gemv(const void* x_bf16, const void* w_bf16, const void* bias_bf16, const void* add0_bf16,
                       const void* add1_bf16, int64_t ld_add, void* out_bf16, int64_t ld_out, int B, int N, int K,
                       int silu_input, fluxb200_stream_t stream);
int fluxb200_timestep_embedding(const void* t_bf16, const float* freqs, void* out_bf16, int B, int dim,
                                float time_factor, fluxb200_stream_t stream);
int fluxb200_euler_update(const void* img_bf16, const void* pred_bf16, const float* dt, void* out_bf16, int64_t n,
                          fluxb200_stream_t stream);
/* The two un-quantised linears around the block stack, as one small mma.sync kernel (a step then launches no library
 * GEMM):  out[m,n] = bf16( sum_k x[m,k] W[n,k] + bias[n] )  -- Flux.img_in (modules/flux_model.py:686, K = 64) and
 * LastLayer.linear (:502, N = 64).  With euler_img / euler_dt the Euler update of flux_pipeline.py:651 is applied to
 * the result in the same launch:  out = bf16( euler_img[m,n] + bf16( (*euler_dt) * out[m,n] ) )  (euler_img has row
 * stride ldo too).  K % 32 == 0, N even, x / W 16-byte aligned. */
int fluxb200_bf16_gemm_small(const void* x_bf16, int64_t ldx, const void* w_bf16, const void* bias_bf16, void* out_bf16,
                             int64_t ldo, const void* euler_img_bf16, const float* euler_dt, int M, int N, int K,
                             fluxb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * VAE decode: modules/autoencoder.py:203-283 `Decoder.forward`, as flux_pipeline.py:423-438 runs it
 * (under torch.autocast(bfloat16): convolutions and attention in bf16 with fp32 accumulation, GroupNorm and swish in
 * fp32).  Activations are NHWC bf16 between these calls; the Python host (`autoencoder.py`) mirrors the reference's
 * module tree and state-dict keys and packs the Conv2d weights once.
 *
 * fluxb200_conv2d_nhwc: nn.Conv2d(kernel 3 stride 1 padding 1) (taps = 9) or kernel 1 (taps = 1) as an implicit GEMM on
 *   tcgen05 (autoencoder.py:33-36 q/k/v/proj_out, :65-79 conv1/conv2/nin_shortcut, :115-117 Upsample.conv, :216 conv_in,
 *   :249 conv_out).  x: bf16 [B, H, W, Cin] (pixel stride ldx elements, 0 = Cin; Cin % 64 == 0, zero-pad otherwise),
 *   w: bf16 [N][taps * Cin] with w[n][(ky * 3 + kx) * Cin + c] = weight[n, c, ky, kx] (row stride ldw, 0 = taps * Cin).
 *   out_mode 0: out bf16 [B*H*W][ldo] = bf16( residual + bf16( bf16(acc) + bias ) )   (bias / residual optional: the
 *               roundings are cuDNN's bf16 output, at::_convolution's bias add, ResnetBlock's `x + h` :94);
 *   out_mode 1: out fp32 [B*H*W][ldo] = alpha * acc       (attention scores q k^T / sqrt(C), AttnBlock.attention :47);
 *   out_mode 2: out bf16 NCHW [B, N, ldo >= H*W] = bf16( bf16(acc) + bias )   (the decoder's image; v^T of the attention,
 *               whose rows are zero-padded to a multiple of 64 positions by the caller).
 *   With H = 1, taps = 1 this is the dense GEMM out[W x N] = x[W x Cin] w[N x Cin]^T.
 * fluxb200_group_norm_nhwc: nn.GroupNorm(32, C, eps, affine) (:27-29, :62-70, :245-247) then (swish != 0) x*sigmoid(x)
 *   (:18-19); fp32 arithmetic on fp64-accumulated statistics, one rounding to bf16.  stats_ws: 64 * B doubles.
 * fluxb200_upsample2x_nhwc: F.interpolate(scale_factor=2.0, mode="nearest") (:121).
 * fluxb200_softmax_rows: p[r, :] = bf16(softmax(scores[r, :])) for the single-head attention of the mid block (:47).
 * fluxb200_vae_latent_prep: AutoEncoder.decode's `z / scale_factor + shift_factor` (:331-332) on the fp32 NCHW latent,
 *   written as bf16 NHWC with the channel dimension zero-padded to Cpad.
 * ------------------------------------------------------------------------------------------- */
typedef struct fluxb200_conv_args {
  const void* x;        /* bf16 NHWC */
  const void* w;        /* bf16 packed weights */
  const void* bias;     /* bf16 [N] or NULL */
  const void* residual; /* bf16 [B*H*W][ld_res] or NULL (out_mode 0) */
  void* out;
  int64_t ldx, ldw, ld_res, ldo;
  int32_t B, H, W, Cin, N, taps, out_mode;
  float alpha;          /* out_mode 1 */
  int32_t stride;       /* 0 / 1: stride 1, padding 1 all round.  2 (taps = 9): Downsample.forward (autoencoder.py:97-110) =
                           F.pad(x, (0, 1, 0, 1)) then 3x3 stride 2 padding 0; output [B, (H-2)/2+1, (W-2)/2+1, N] */
  double* gn_stats;     /* optional (out_mode 0, B <= 4, N % 64 == 0): receives the GroupNorm(32) statistics of the STORED
                           output, [B][32]{sum, sum of squares} -- pass it to fluxb200_group_norm_nhwc with stats_ready = 1
                           and the normalisation of this tensor skips its own reduction pass */
} fluxb200_conv_args;
int fluxb200_conv2d_nhwc(const fluxb200_conv_args* args, fluxb200_stream_t stream);
int fluxb200_group_norm_nhwc(const void* x_bf16, const void* gamma_bf16, const void* beta_bf16, void* y_bf16,
                             double* stats_ws, int stats_ready, int B, int64_t HW, int C, float eps, int swish,
                             fluxb200_stream_t stream);
int fluxb200_upsample2x_nhwc(const void* x_bf16, void* y_bf16, int B, int H, int W, int C, fluxb200_stream_t stream);
int fluxb200_softmax_rows(const float* scores, int64_t lds, void* p_bf16, int64_t ldp, int rows, int n,
                          fluxb200_stream_t stream);
int fluxb200_vae_latent_prep(const float* z_nchw, void* y_nhwc_bf16, int B, int C, int64_t HW, int Cpad,
                             float scale_factor, float shift_factor, fluxb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Text encoders (second half): modules/conditioner.py:HFEmbedder wraps Hugging Face `T5EncoderModel`
 * (t5-v1_1-xxl) and `CLIPTextModel` (clip-vit-large-patch14) (:80-92) and calls them with attention_mask=None
 * (:101-117).  The arithmetic therefore lives in the third-party `transformers` package (5.5.0 in this image; the
 * reference does not pin it): models/t5/modeling_t5.py (T5LayerNorm, T5Attention, T5DenseGatedActDense, T5Block) and
 * models/clip/modeling_clip.py (CLIPTextTransformer).  Dense layers run on fluxb200_conv2d_nhwc in its dense form
 * (H = 1, taps = 1; bias and residual in the epilogue); the rest:
 *
 * fluxb200_rows_norm:  bias == NULL: T5LayerNorm  y = bf16( w * bf16( x * rsqrt(mean(x^2) + eps) ) )
 *                      else nn.LayerNorm(affine)  y = bf16( (x - mean) * rsqrt(var + eps) * w + b ), fp32 statistics.
 * fluxb200_gated_act:  mode 0: out = bf16( bf16(gelu_new(in[:, :F])) * in[:, F:2F] )   (T5DenseGatedActDense, wi_0 | wi_1
 *                      concatenated along N);  mode 1: out = bf16( x * sigmoid(1.702 x) )   (CLIP quick_gelu).
 * fluxb200_attention_d64:  multi-head attention, head dim 64, q / k / v rows `ld` elements apart (head h at columns
 *                      [64 h, 64 h + 64): three pointers into one fused QKV buffer work), optional additive bias bf16
 *                      [H, S, S] shared by the batch (T5's relative position bias), softmax scale (1 for T5, 1/8 for
 *                      CLIP), optional causal mask (CLIP).  scores are rounded to bf16 after the product, after the scale
 *                      and after the bias as the eager modules do; softmax in fp32; out bf16 [B*S, ldo].
 * ------------------------------------------------------------------------------------------- */
int fluxb200_rows_norm(const void* x_bf16, int64_t ldx, const void* weight_bf16, const void* bias_bf16, void* y_bf16,
                       int64_t ldy, int rows, int D, float eps, fluxb200_stream_t stream);
int fluxb200_gated_act(const void* in_bf16, int64_t ld_in, void* out_bf16, int64_t ld_out, int rows, int F, int mode,
                       fluxb200_stream_t stream);
int fluxb200_attention_d64(const void* q_bf16, const void* k_bf16, const void* v_bf16, int64_t ld, const void* bias_bf16,
                           void* out_bf16, int64_t ldo, int B, int H, int S, float scale, int causal,
                           fluxb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Modulation prologue: y = quantize( bf16(silu(x)), scale )   (modules/flux_model.py:249,252 +
 * float8_quantize.py:274-276).  y_bf16 (optional) receives bf16(silu(x)) for unquantised lins.
 * ------------------------------------------------------------------------------------------- */
int fluxb200_silu_quant(const void* x_bf16, void* y_fp8, void* y_bf16, int64_t n, const float* scale, int fmt,
                        fluxb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm(eps, no affine) -> (1+scale)*x + shift -> next F8Linear's input quantisation
 *   modules/flux_model.py:367-368, 374-375, 389, 395, 469-470 + float8_quantize.py:274-276
 * x bf16 [B*L, D] (row stride ldx); shift/scale bf16 [B][D] (sample stride mod_batch_stride).
 * y_fp8 [B*L, ldy] and/or y_bf16 [B*L, ldy_bf16] (either may be NULL).
 * ------------------------------------------------------------------------------------------- */
int fluxb200_ln_mod_quant(const void* x, int64_t ldx, const void* shift, const void* scale,
                          int64_t mod_batch_stride, void* y_fp8, int64_t ldy, void* y_bf16, int64_t ldy_bf16,
                          const float* in_scale, int fmt, int B, int L, int D, float eps,
                          fluxb200_stream_t stream);

/* The same for up to two independent row sets in ONE launch (the txt and img streams of a DoubleStreamBlock,
 * modules/flux_model.py:367-368 / 374-375 and :389 / :395): fp8 output only, same D / fmt / eps. */
typedef struct fluxb200_ln_args {
  const void* x;       /* bf16 [B*L, D], row stride ldx */
  const void* shift;   /* bf16 [B][D], sample stride mod_batch_stride */
  const void* scale;
  void* y_fp8;         /* [B*L, ldy] */
  const float* in_scale;
  int64_t ldx, ldy, mod_batch_stride;
  int32_t B, L;
} fluxb200_ln_args;
int fluxb200_ln_mod_quant_grouped(const fluxb200_ln_args* args, int count, int fmt, int D, float eps,
                                  fluxb200_stream_t stream);

/* LayerNorm-modulate-quantise FUSED INTO THE CONSUMING GEMM LAUNCH (modules/flux_model.py:367-371, 374-378, 389-390,
 * 395-396, 469-471): every warp of the persistent GEMM grid first turns rows of ln[i].x into ln[i].y_fp8 -- which
 * must be the A operand(s) of `args` -- exactly as fluxb200_ln_mod_quant_grouped would, the grid synchronises, and the
 * GEMM(s) of fluxb200_f8_gemm_grouped run.  One launch instead of two: in the captured step a stand-alone LayerNorm
 * kernel between two persistent GEMMs costs ~45 us (drain and refill of every SM on both sides), four times its own
 * run time.  D must be 3072 (FLUXB200_ERR_UNSUPPORTED otherwise: use the two separate calls).
 * grid_barrier_ws: two 32-bit words in device memory, zeroed ONCE by the caller and then only ever passed to this
 * entry point on launches of one stream (the kernels re-arm it themselves; CUDA-graph replays need no reset). */
int fluxb200_f8_gemm_ln(const fluxb200_gemm_args* args, int count, const fluxb200_ln_args* ln, int ln_count, int fmt,
                        int D, float eps, void* grid_barrier_ws, fluxb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stand-alone QKNorm + apply_rope on [B,H,S,128] tensors (modules/flux_model.py:164, 60-65).
 * norm_w may be NULL (skip RMSNorm); cos/sin may be NULL (skip RoPE).  In-place allowed.
 * ------------------------------------------------------------------------------------------- */
int fluxb200_qknorm_rope(const void* x, void* y, const float* norm_w, const void* rope_cos,
                         const void* rope_sin, int64_t rope_batch_stride, int B, int H, int S, float eps,
                         fluxb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * attention(): F.scaled_dot_product_attention(q,k,v) + transpose/reshape to [B,S,H*128]
 *   modules/flux_model.py:41-45 (RoPE is applied by the producer of q,k).
 * q,k,v bf16 [B,H,S,128] contiguous.  out[b, s, h*128 + d]: row stride ldo elements.
 * out_kind 0: bf16.  out_kind 1: fp8 = quantize(bf16(o), scale) where rows s < split_row use
 * *out_scale0 and the others *out_scale1 (txt / img proj input scales; equal for single blocks).
 * ------------------------------------------------------------------------------------------- */
typedef struct fluxb200_attention_args {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  int64_t ldo;
  int64_t out_batch_stride; /* elements between samples of out */
  int32_t B, H, S;
  float softmax_scale; /* 1/sqrt(128) */
  int32_t out_kind;
  int32_t out_fmt;
  int32_t split_row;
  int32_t variant; /* 0 = the library chooses between its two kernels (bit-identical results): 17 forces the single-CTA
                      form, 16 the cta_group::2 pair form (A/B measurements, tests); other values exist only in
                      -DFLUXB200_ATTN_EXPERIMENTS builds */
  const float* out_scale0;
  const float* out_scale1;
  /* Optional second destination: when out1 != NULL, rows s >= split_row are written to
     out1[b*out1_batch_stride + (s - split_row)*ldo1 + h*128 + d] instead of `out`
     (DoubleStreamBlock: txt rows feed txt_attn.proj, img rows img_attn.proj). */
  void* out1;
  int64_t ldo1;
  int64_t out1_batch_stride;
} fluxb200_attention_args;

int fluxb200_attention(const fluxb200_attention_args* args, fluxb200_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * LoRA fuse / unfuse into one quantised layer, on the device.  Replaces the per-layer body of
 * apply_lora_to_model / remove_lora_from_module (lora_loading.py:679-687, 742-749):
 *   extract_weight_from_linear (:615-626)            W  = float(w_fp8) * (*w_scale_recip)
 *   calculate_lora_weight (:509-547)                 D  = sum_c coeff * (lora_down @ lora_up[c*R:(c+1)*R])     (fp32)
 *   apply_lora_weight_to_module (:566-577) / unfuse (:549-563)   W' = bf16(W + D)  /  bf16(W - D)
 *   first half of F8Linear.quantize_weight (float8_quantize.py:196-198)   *amax_out = max|W'|
 * lora_down = lora_B as fp32 [N, R]; lora_up = lora_A as fp32 [chunks*R, K], already multiplied by alpha/rank by the
 * caller when alpha != rank (:530-531); chunks > 1 is the reference's "uneven rank" case (:532-541); coeff = lora_scale.
 * The caller turns *amax_out into the new scale (amax_to_scale) and requantises w_out_bf16 with fluxb200_quantize --
 * typically over the old w_fp8 buffer, which this call has finished reading.  amax_out is zeroed by the call. */
int fluxb200_lora_fuse(const void* w_fp8, int w_fmt, const float* w_scale_recip, const float* lora_down,
                       const float* lora_up, int N, int K, int R, int chunks, float coeff, int unfuse,
                       void* w_out_bf16, float* amax_out, fluxb200_stream_t stream);

/* Diagnostics (measurement only; outputs of probe-mode launches are garbage): bit mask OR-ed into every following
 * fluxb200_f8_gemm* launch of this process until reset with 0.  1 = operands stay resident in shared memory after the
 * first fill (no TMA traffic), 2 = no MMAs, 4 = no epilogue, 8 = epilogue TMEM loads only.  Mode 1|4 times the bare
 * tcgen05 kind::f8f6f4 issue rate of the kernel's tiling: bench.py measures the FP8 tensor-pipe ceiling of the box it
 * runs on with it (roofline.peak), instead of inferring it from the bf16 figure. */
int fluxb200_gemm_probe_mode(int mode);

/* Measurement: launches the bare tcgen05.mma.kind::f8f6f4 loop of the product GEMM tiling (cta_group::2, M = 256, N = 256,
 * K = 32; operands resident in shared memory: no TMA, no epilogue, no global traffic) on every SM pair, `tiles_per_pair`
 * accumulator tiles of 24 K-slabs each; *flops_out (host pointer, may be NULL) receives the FLOPs the launch performs.
 * Timed with CUDA events by the caller (bench.py) it gives the FP8 tensor-pipe ceiling of the box under its power cap. */
int fluxb200_fp8_mma_probe(int tiles_per_pair, double* flops_out, fluxb200_stream_t stream);

/* Tiling override for A/B measurements and tests (0, 0 = the library's own choice): cta_group 1 = one CTA per 128 x BN
 * tile, 2 = one CTA pair per 256 x 256 tile; pairs_per_cluster 2 = two pairs per four-CTA cluster sharing their A rows
 * by TMA multicast (needs cta_group 2 and an even number of N tiles; ignored otherwise).  Results are identical
 * bit for bit across tilings (same MMA shapes and K order per output element). */
int fluxb200_gemm_force_tiling(int cta_group, int pairs_per_cluster);

/* Diagnostics: cycle counters of the last attention launch's CTA 0 (host pointer to 16 x uint64):
 * [0..5] softmax warp: wait-S, tmem load, max, exp, wait-O, store-P; [6] half-steps;
 * [8..10] MMA issuer: wait-P, wait-KV, issue.  Synchronises the device. */
int fluxb200_debug_counters(unsigned long long* host_out16);

#ifdef __cplusplus
}
#endif
#endif /* FLUX_B200_H_ */
