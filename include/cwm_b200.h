/*
 * cwm_b200.h -- C ABI of libcwm_b200.so: the B200 (sm_100a) implementation of the CWM masked
 * video-autoencoder (VMAE) forward path.
 *
 * This header is the drop-in boundary.  The reference (neuroailab/CounterfactualWorldModels) is pure
 * Python/PyTorch and has no FFI of its own; the entry points below are what a ctypes binding inside the
 * reference's `cwm/models/VideoMAE/vmae.py` would call.  Every function
 *   - takes raw DEVICE pointers, plain ints and a cudaStream_t (as void*), no torch types;
 *   - is asynchronous on `stream` and never synchronises the host;
 *   - returns 0 on success or a negative cwm_status; the message is available from cwm_last_error();
 *   - never throws across the ABI.
 * All matrices are row-major.  "f16" buffers are IEEE binary16 stored as uint16_t.
 *
 * Reference citations are relative to /root/reference (file:line).
 */
#ifndef CWM_B200_H_
#define CWM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CWM_B200_ABI_VERSION 7

typedef void* cwm_stream_t; /* cudaStream_t */

enum cwm_status {
  CWM_OK = 0,
  CWM_ERR_INVALID = -1,     /* bad argument (shape, alignment, null pointer)        */
  CWM_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed                    */
  CWM_ERR_UNSUPPORTED = -3, /* valid in the reference but not implemented here      */
  CWM_ERR_WORKSPACE = -4,   /* caller-provided workspace too small                  */
  CWM_ERR_ARCH = -5         /* device is not sm_100                                 */
};

/* ---- library ------------------------------------------------------------------------------------- */
int cwm_abi_version(void);
/* Operand type of this build: 0 = IEEE f16 (libcwm_b200.so, the parity default), 1 = bf16 (libcwm_b200_bf16.so: the same
 * sources compiled with -DCWM_ACT_BF16; every `uint16_t*` activation / weight tensor of this header then holds bf16). */
int cwm_act_dtype(void);
const char* cwm_last_error(void);
/* 0 when the current device is compute capability 10.x (B200), CWM_ERR_ARCH otherwise. */
int cwm_device_check(void);

/* ---- a4: visible-token compaction ---------------------------------------------------------------
 * Replaces `x[~mask].reshape(B,-1,C)` (cwm/models/VideoMAE/vmae.py:166-167) and the two boolean gathers
 * `expand_pos_embed[~mask]`, `expand_pos_embed[mask]` (vmae.py:555-556).
 *   mask       [B, Ntot] bytes, non-zero = masked (torch.bool storage)
 *   perm       [B, Ntot] int32 out: first n_visible[b] entries = visible token indices ascending, the
 *              rest = masked token indices ascending.  This is exactly the decoder token order of
 *              vmae.py:557 (`cat([x_vis, mask_tokens])`).
 *   inv_perm   [B, Ntot] int32 out (may be NULL): inv_perm[b, perm[b, j]] = j
 *   n_visible  [B] int32 out
 * Integer work: bit-exact with torch.nonzero. */
int cwm_compact_mask(const uint8_t* mask, int B, int Ntot, int32_t* perm, int32_t* inv_perm,
                     int32_t* n_visible, cwm_stream_t stream);

/* ---- a1+a2 (input side): normalise + gather visible patches --------------------------------------
 * Replaces the im2col half of `PatchEmbed.forward` (cwm/models/VideoMAE/utils.py:178-198) for the visible
 * tokens only, optionally fused with `imagenet_normalize` (cwm/models/utils.py:15-21).
 *   x          fp32 video, logical layout [B, C, T, H, W] with arbitrary element strides xs[5]
 *              (the reference hands a transposed view, cwm/models/prediction.py:304-312)
 *   perm       as produced by cwm_compact_mask (row stride Ntot); rows_per_sample = Nvis entries are used
 *   mean/stdv  NULL, or per-channel HOST arrays [C] fp32 (they are constants of the caller, utils.py:12-13):
 *              value = (x - mean[c]) / stdv[c], same operation order as the reference
 *   out        f16 [B*rows_per_sample, K], K = C*pt*ph*pw ordered (c, kt, kh, kw) = Conv3d weight order
 * Token ids >= (T/pt)*(H/ph)*(W/pw) are padding positions (conjoined_vmae.py:130-133): their rows are zero. */
int cwm_patch_gather(const float* x, const int64_t xs[5], int B, int C, int T, int H, int W, int pt,
                     int ph, int pw, const int32_t* perm, int Ntot, int rows_per_sample,
                     const float* mean, const float* stdv, uint16_t* out, cwm_stream_t stream);

/* ---- LayerNorm (a5/a8/a11) -----------------------------------------------------------------------
 * nn.LayerNorm(C, eps) over the last dim (cwm/models/VideoMAE/utils.py:130,136; vmae.py:84,206), fp32
 * statistics, f16 output (the operand of the following GEMM).  C % 4 == 0, C <= 1024.
 * Row mapping: output row m reads input row (m / grp_rows) * grp_stride + grp_offset + m % grp_rows when
 * grp_rows > 0 (used for `x[:, -return_token_num:]`, vmae.py:250-251), else row m. */
int cwm_layernorm_f16(const float* x, int M, int C, const float* gamma, const float* beta, float eps,
                      int grp_rows, int grp_stride, int grp_offset, uint16_t* out, cwm_stream_t stream);

/* First LayerNorm of a stream when LayerNorm is fused into the consumer GEMM (see cwm_gemm_epilogue): x16 = f16(x)
 * [M, C] and stats[m] = (sum, sum of squares) of row m (one partial plane, ln_parts = 1). */
int cwm_rowstats_f16(const float* x, int M, int C, uint16_t* x16, float* stats, cwm_stream_t stream);

/* ---- GEMM with fused epilogue (a2, a6, a7, a9, a11) ---------------------------------------------
 * Y[M,N] = epilogue(A[M,K] . W[N,K]^T), A and W f16 (K contiguous, like nn.Linear.weight), fp32
 * accumulation in tensor memory (tcgen05.mma kind::f16).  K % 16 == 0 and K >= 16; lda = K, ldw = K.  */
enum cwm_epilogue {
  CWM_EPI_F16 = 0,      /* out_f16 = (acc + bias) * (col < scale_cols ? scale : 1)      (qkv: utils.py:89-97)  */
  CWM_EPI_GELU_F16 = 1, /* out_f16 = gelu_erf(acc + bias)                               (fc1: utils.py:48-49)  */
  CWM_EPI_RES_F32 = 2,  /* out_f32 = res_f32 + acc + bias                               (proj/fc2: utils.py:148-149) */
  CWM_EPI_F32 = 3       /* out_f32 = acc + bias                                         (head: vmae.py:251)    */
};

typedef struct cwm_gemm_epilogue {
  int32_t mode;          /* enum cwm_epilogue */
  const float* bias;     /* [N] or NULL */
  float scale;           /* CWM_EPI_F16 only */
  int32_t scale_cols;    /* CWM_EPI_F16 only: columns [0, scale_cols) are multiplied by scale */
  const float* res;      /* CWM_EPI_RES_F32: residual, leading dimension ldr */
  int32_t ldr;
  const int32_t* res_gather; /* optional: residual row = res_gather[(m / grp_rows) * gather_stride + m % grp_rows]
                                (pos-embed add for visible tokens only: vmae.py:162-167, :555-557) */
  int32_t gather_stride;
  int32_t grp_rows;      /* > 0: output row = (m / grp_rows) * grp_out_stride + m % grp_rows (writes x_vis
                            straight into the decoder sequence, vmae.py:557) */
  int32_t grp_out_stride;
  void* out;             /* f16 or fp32, leading dimension ldo */
  int32_t ldo;
  /* ---- LayerNorm fusion (all optional, zero-initialise to switch off) -----------------------------------------
   * `x + f(x)` followed by `LayerNorm` (cwm/models/VideoMAE/utils.py:146-153) without a LayerNorm kernel:
   *  producer (CWM_EPI_RES_F32, no row remap, N % 32 == 0): additionally writes ln_x16 = f16(out) [M, ln_ldx16] and,
   *    per row, partial (sum, sum of squares) of its columns: ln_stats_out[p * M + m], p < cwm_gemm_ln_parts(N);
   *  consumer (CWM_EPI_F16 / CWM_EPI_GELU_F16) with A = ln_x16 and W = weight * gamma[None, :]:
   *    out = act(rstd_m * (acc - mean_m * ln_colsum[n]) + bias[n]), mean / rstd over ln_width elements with eps ln_eps,
   *    ln_colsum[n] = sum_k W[n, k] (of the f16 values), bias[n] = (weight @ beta)[n] + linear bias[n]. */
  uint16_t* ln_x16;
  int32_t ln_ldx16;
  float* ln_stats_out;       /* [cwm_gemm_ln_parts(N), M, 2] fp32 */
  const float* ln_stats_in;  /* [ln_parts, M, 2] fp32 */
  int32_t ln_parts;
  const float* ln_colsum;    /* [N] */
  int32_t ln_width;
  float ln_eps;
} cwm_gemm_epilogue;

/* Number of partial-statistics planes a LayerNorm-producer GEMM with N output columns writes. */
int cwm_gemm_ln_parts(int N);

int cwm_gemm_f16(const uint16_t* A, const uint16_t* W, int M, int N, int K, const cwm_gemm_epilogue* epi,
                 cwm_stream_t stream);

/* ---- attention (a6) --------------------------------------------------------------------------------
 * softmax(q k^T) v per (sample, head), unmasked (cwm/models/VideoMAE/utils.py:108-113); q is already scaled
 * (utils.py:97 is fused into the qkv GEMM epilogue).  head_dim 64 (every RGB stream) runs the tcgen05 kernel;
 * 32 / 96 / 128 / 192 are forwarded to cwm_attention_generic_f16.
 *   qkv  f16 [B, N, 3, H, d]  (the layout `F.linear(...).reshape(B,N,3,H,-1)` has, utils.py:93-94)
 *   out  f16 [B, N, H*d]      (= `x.transpose(1,2).reshape(B,N,-1)`, utils.py:118)               */
int cwm_attention_f16(const uint16_t* qkv, int B, int N, int H, int head_dim, uint16_t* out,
                      cwm_stream_t stream);

/* ---- a10: mask-token rows of the decoder input ---------------------------------------------------
 * x_full[b, Nvis + j, :] = mask_token + pos[perm[b, Nvis + j], :]  (vmae.py:556-557).  The visible rows are
 * written by the encoder_to_decoder GEMM epilogue. */
int cwm_fill_mask_tokens(const float* mask_token, const float* pos, const int32_t* perm, int B, int Ntot,
                         int Nvis, int C, float* x_full, cwm_stream_t stream);

/* ---- a12: scatter predictions + unpatchify ------------------------------------------------------
 * `pred_patches_to_video` (cwm/models/prediction.py:245-259) with `Patchify` (cwm/models/patches.py:67-109):
 * out[b,t,c,y,x] = x_raw[...] where the patch is visible, else y[b, rank, ((kt*ph+kh)*pw+kw)*C + c].
 *   y        fp32 [B, Nmask, D]
 *   x_raw    fp32, logical [B, T, C, H, W] with element strides xs[5]
 *   inv_perm from cwm_compact_mask
 *   out      fp32 contiguous [B, T, C, H, W] */
int cwm_unpatchify_scatter(const float* y, const float* x_raw, const int64_t xs[5], const int32_t* inv_perm,
                           int B, int T, int C, int H, int W, int pt, int ph, int pw, int Nvis, float* out,
                           cwm_stream_t stream);

/* ---- whole forward (a1..a11) ---------------------------------------------------------------------
 * `PretrainVisionTransformer.forward(x, mask)` (cwm/models/VideoMAE/vmae.py:539-560).                */
typedef struct cwm_block_weights {
  const float *ln1_g, *ln1_b;       /* norm1.{weight,bias}              */
  const uint16_t* w_qkv;            /* attn.qkv.weight   f16 [3C, C]    */
  const float* b_qkv;               /* cat(q_bias, 0, v_bias) [3C] or NULL (utils.py:89-91) */
  const uint16_t* w_proj;           /* attn.proj.weight  f16 [C, C]     */
  const float* b_proj;
  const float *ln2_g, *ln2_b;
  const uint16_t* w_fc1;            /* mlp.fc1.weight    f16 [hidden, C] */
  const float* b_fc1;
  const uint16_t* w_fc2;            /* mlp.fc2.weight    f16 [C, hidden] */
  const float* b_fc2;
  /* Optional LayerNorm-folded copies (all six set, or all NULL = separate LayerNorm kernels).  With them the block
   * runs without LayerNorm kernels: the residual GEMMs emit an f16 copy + row statistics, the qkv / fc1 GEMMs apply
   * the normalisation in their epilogue (see cwm_gemm_epilogue):
   *   w_qkv_ln = f16(qkv.weight * norm1.weight[None, :]);  s_qkv[n] = sum_k w_qkv_ln[n, k];
   *   c_qkv    = qkv.weight @ norm1.bias + cat(q_bias, 0, v_bias);      likewise fc1 with norm2. */
  const uint16_t* w_qkv_ln;
  const float* s_qkv;
  const float* c_qkv;
  const uint16_t* w_fc1_ln;
  const float* s_fc1;
  const float* c_fc1;
} cwm_block_weights;

typedef struct cwm_vmae_model {
  /* geometry */
  int32_t in_chans, num_frames, img_h, img_w, pt, ph, pw;
  int32_t enc_dim, enc_depth, enc_heads, enc_hidden;
  int32_t dec_dim, dec_depth, dec_heads, dec_hidden;
  int32_t out_dim; /* D = decoder_num_classes */
  float ln_eps;
  float enc_qk_scale, dec_qk_scale; /* Attention.scale = qk_scale or head_dim ** -0.5 (utils.py:67) */
  /* weights (device pointers, owned by the caller) */
  const uint16_t* w_patch;    /* encoder.patch_embed.proj.weight f16 [Ce, C*pt*ph*pw] */
  const float* b_patch;       /* [Ce] */
  const float* pos_enc;       /* sinusoid table fp32 [Ntot, Ce] (utils.py:251-268) */
  const cwm_block_weights* enc_blocks; /* HOST array, enc_depth entries */
  const float *enc_norm_g, *enc_norm_b;
  const uint16_t* w_e2d;      /* encoder_to_decoder.weight f16 [Cd, Ce], no bias (vmae.py:355) */
  const float* mask_token;    /* [Cd] */
  const float* pos_dec;       /* fp32 [Ntot, Cd] (vmae.py:366) */
  const cwm_block_weights* dec_blocks; /* HOST array, dec_depth entries */
  const float *dec_norm_g, *dec_norm_b;
  const uint16_t* w_head;     /* decoder.head.weight f16 [D, Cd] */
  const float* b_head;        /* [D] */
} cwm_vmae_model;

/* Bytes of scratch the forward needs for a batch of B samples with Nvis visible tokens each. */
size_t cwm_vmae_workspace_bytes(const cwm_vmae_model* model, int B, int Nvis);

/*   x         fp32 logical [B, C, T, H, W], element strides xs[5]
 *   norm_mean/norm_std  NULL, or HOST arrays [C]: fuse imagenet_normalize (prediction.py:309-310) into the gather
 *   perm      [B, Ntot] from cwm_compact_mask; every row must have exactly Nvis visible tokens (the reference
 *             raises at vmae.py:167 otherwise; the host wrapper checks n_visible)
 *   y         fp32 out [B, Ntot - Nvis, D]
 *   workspace device scratch of at least cwm_vmae_workspace_bytes(model, B, Nvis) bytes, 1024-byte aligned */
int cwm_vmae_forward(const cwm_vmae_model* model, const float* x, const int64_t xs[5], int B,
                     const float* norm_mean, const float* norm_std, const int32_t* perm, int Nvis, float* y,
                     void* workspace, size_t workspace_bytes, cwm_stream_t stream);

/* Same forward with the input video given as a counterfactual descriptor (declared below, section 8(f) rank 1):
 * equivalent to cwm_cf_build_videos + cwm_vmae_forward without materialising the videos. */
struct cwm_cf_source;
int cwm_vmae_forward_cf(const cwm_vmae_model* model, const struct cwm_cf_source* src, int S, const float* norm_mean,
                        const float* norm_std, const int32_t* perm, int Nvis, float* y, void* workspace,
                        size_t workspace_bytes, cwm_stream_t stream);

/* ---- generic small attention (a6 for the context stream, a15 cross-attention) ------------------------
 * softmax(q k^T) v for head dims 32 / 64 / 96 / 128 / 192 with independently strided q, k, v:
 *   q row (b, i), head h starts at q + (b*Nq + i)*ldq + h*q_head_stride;  k, v likewise over Nk rows.
 *   out[(b*Nq + i)*ldo + h*head_dim + c], f16.   q must be pre-scaled.
 * Used for (i) self-attention of the 25..50 IMU tokens, head dim 32 (cwm/models/VideoMAE/utils.py:87-121 as
 * instantiated by conjoined_vmae.py:1198-1216) and (ii) both directions of BidirectionalCrossAttention
 * (cwm/models/transformer.py:358-372): N x M ("trg") and M x N ("src", key axis split internally; needs workspace). */
size_t cwm_attention_generic_workspace_bytes(int B, int Nq, int Nk, int H, int head_dim);
int cwm_attention_generic_f16(const uint16_t* q, const uint16_t* k, const uint16_t* v, int ldq, int ldk, int ldv,
                              int q_head_stride, int k_head_stride, int v_head_stride, int B, int Nq, int Nk,
                              int H, int head_dim, uint16_t* out, int ldo, void* workspace, size_t workspace_bytes,
                              cwm_stream_t stream);

/* ---- a13: null (padding) tokens ------------------------------------------------------------------------
 * PaddedVisionTransformer (cwm/models/VideoMAE/conjoined_vmae.py:125-165, :207-208): token ids >= first_pad_token
 * in `perm` are padding positions.  For every sample b and every j in [0, rows) with
 * perm[b*perm_stride + perm_offset + j] >= first_pad_token:   x[b, j, :] = value  (all zeros when value == NULL).
 *   encoder input: value = null_token_enc (pad_and_mask_input, :130-133); output rows: value = NULL (:207-208). */
int cwm_fill_pad_rows(float* x, int B, int rows, int C, const int32_t* perm, int perm_stride, int perm_offset,
                      int first_pad_token, const float* value, cwm_stream_t stream);

/* ---- one transformer block (a5-a7) ---------------------------------------------------------------------
 * `Block.forward` (cwm/models/VideoMAE/utils.py:146-153, gamma_* = None) in place on the fp32 residual stream
 * x [B*N, C]: x += proj(Attn(LN1(x))); x += fc2(GELU(fc1(LN2(x)))).  heads*head_dim is the inner attention width
 * (== C for every shipped model).  head_dim 64 runs the tcgen05 attention kernel, 32/96/128/192 the generic one.
 * The conjoined models interleave these with cross blocks (conjoined_vmae.py:543-576, :688-720). */
size_t cwm_block_workspace_bytes(int B, int N, int C, int heads, int head_dim, int hidden);
int cwm_block_forward(const cwm_block_weights* w, float* x, int B, int N, int C, int heads, int head_dim, int hidden,
                      float ln_eps, float qk_scale, void* workspace, size_t workspace_bytes, cwm_stream_t stream);

/* ---- a14/a15: one conjoining (cross-attention) block ----------------------------------------------------
 * `CrossAttentionTransformerBlock.forward` with with_self_attention=False (cwm/models/transformer.py:559-583;
 * gamma_1 = 0, norm1 = Identity, shortcut = Identity) around `BidirectionalCrossAttention.forward`
 * (:314-378, shared_similarity=False, no qkv bias), in place on both fp32 residual streams:
 *   qk, v = LN1c(x) W;  qk_s, v_s = LN1sc(src) W_s
 *   x   += projection    (softmax(scale * qk[..., :hd] qk_s[..., :hd]^T) v_s)        [N x M attention per head]
 *   src += projection_src(softmax(scale * qk_s[..., hd:] qk[..., hd:]^T) v)          [M x N attention per head]
 *   x   += fc2(GELU(fc1(LN2(x))));  src += fc2_s(GELU(fc1_s(LN2s(src))))
 * x [B*N, C], src [B*M, Cs]; inner width D = heads*head_dim.  f16 weights are [out, in] row-major. */
typedef struct cwm_cross_block_weights {
  const float *ln1_g, *ln1_b;         /* norm1_cross           [C]  */
  const float *ln1s_g, *ln1s_b;       /* norm1_src_cross       [Cs] */
  const uint16_t* w_qkv;              /* cat(cross_attention.qk.weight, cross_attention.v.weight)         f16 [3D, C]  */
  const uint16_t* w_qkv_s;            /* cat(cross_attention.qk_src.weight, cross_attention.v_src.weight) f16 [3D, Cs] */
  const uint16_t* w_proj;             /* cross_attention.projection.weight      f16 [C, D]  */
  const float* b_proj;
  const uint16_t* w_proj_s;           /* cross_attention.projection_src.weight  f16 [Cs, D] */
  const float* b_proj_s;
  const float *ln2_g, *ln2_b;         /* norm2      [C]  */
  const float *ln2s_g, *ln2s_b;       /* norm2_src  [Cs] */
  const uint16_t* w_fc1;              /* mlp.trg.layers.0.weight f16 [hidden, C] */
  const float* b_fc1;
  const uint16_t* w_fc2;              /* mlp.trg.layers.2.weight f16 [C, hidden] */
  const float* b_fc2;
  const uint16_t* w_fc1_s;            /* mlp.src.layers.0.weight f16 [hidden_s, Cs] */
  const float* b_fc1_s;
  const uint16_t* w_fc2_s;            /* mlp.src.layers.2.weight f16 [Cs, hidden_s] */
  const float* b_fc2_s;
} cwm_cross_block_weights;

size_t cwm_cross_block_workspace_bytes(int B, int N, int M, int C, int Cs, int heads, int head_dim, int hidden,
                                       int hidden_s);
int cwm_cross_block_forward(const cwm_cross_block_weights* w, float* x, float* src, int B, int N, int M, int C,
                            int Cs, int heads, int head_dim, int hidden, int hidden_s, float ln_eps, float scale,
                            void* workspace, size_t workspace_bytes, cwm_stream_t stream);

/* ==== SURVEY.md section 8(f) rank 1: batched motion-counterfactual construction ===========================
 * Replaces the per-sample Python loop `FlowGenerator.create_motion_counterfactuals`
 * (cwm/models/segmentation.py:321-338) -> `PatchPerturbation.forward` (cwm/models/perturbation.py:99-112) ->
 * `ShiftPatchesAndMask.perturb` (perturbation.py:245-289), and `MakeStatic.perturb` (perturbation.py:129-150).
 * Conventions of the reference: masks are bytes with non-zero = masked; `active` is a mask whose ZERO entries are
 * the patches to move; tokens are ordered (t, h, w); pt == 1; videos are logical [B, T, C, H, W].
 * Integer work (masks) is bit-exact; pixels are bit-exact too: the blend `x_shift*(1-m) + x*m` (perturbation.py:
 * 278-282) is evaluated literally in fp32 (one rounding per multiply and add, no fused multiply-add), for FINITE
 * pixel values (signed zeros included; an inf/NaN pixel that the reference would spread through `0 * inf` is not). */

/* Masks of S samples in one launch.
 *   passive, active  [S, T*n_h*n_w] bytes   (`masks`, `active_patches` after 'b n s -> (b s) n', segmentation.py:313-314)
 *   mask_shift       [S, 2] int32 (my, mx): shift of the mask in patch units as applied by the mask padding
 *                    (perturbation.py:234-243)
 *   frame            the target frame (already reduced modulo T)
 *   shifted_active   out [S, n_h*n_w]: the target frame of the shifted perturbation mask (perturbation.py:268-269,
 *                    padding value 1); selects which patches of the target frame show the shifted image
 *   mask_out         out [S, T*n_h*n_w]: `minimum(mask, mask_perturbed)` (perturbation.py:106-110), i.e.
 *                    (passive | !active) & (t == frame ? shifted_active : active)                                */
int cwm_cf_shift_masks(const uint8_t* passive, const uint8_t* active, const int32_t* mask_shift, int S, int T,
                       int n_h, int n_w, int frame, uint8_t* shifted_active, uint8_t* mask_out, cwm_stream_t stream);

/* The "virtual" counterfactual video of sample i (never materialised on the fused path):
 *   v[i, t, c, y, x] = img[t', c, y, x]                                                  for t != frame
 *   v[i, frame, c, y, x] = sh * (1 - m) + img[frame', c, y, x] * m,   m = shifted_active[i, y/ph, x/pw],
 *   sh = img[frame', c, y - sy, x - sx] inside the image, 0 outside (perturbation.py:257-258), img = x[sample_image[i]],
 *   t' = static_frame if static_frame >= 0 (`make_static_movie`, cwm/models/prediction.py:731-740) else t.       */
typedef struct cwm_cf_source {
  const float* x;               /* fp32 raw frames, logical [B_img, T, C, H, W] with element strides xs */
  int64_t xs[5];
  const int32_t* sample_image;  /* [S] image index of every sample (`sample_tile`, prediction.py:484-487) or NULL = 0 */
  const int32_t* shift_px;      /* [S, 2] (sy, sx) pixel shift */
  const uint8_t* shifted_active;/* [S, n_h*n_w] from cwm_cf_shift_masks */
  int32_t frame;
  int32_t static_frame;         /* -1, or the source frame every frame reads */
} cwm_cf_source;

/* Materialises the S counterfactual videos: out fp32 contiguous [S, T, C, H, W] (= `x_shift` of
 * segmentation.py:339).  HBM-bound: 4*T*C*H*W bytes written per sample; the source image stays in L2. */
int cwm_cf_build_videos(const cwm_cf_source* src, int S, int T, int C, int H, int W, int ph, int pw, float* out,
                        cwm_stream_t stream);

/* `MakeStatic.perturb` (perturbation.py:129-150): out[b,t] = (1 - m) * x[b,0] + m * x[b,t], m = mask[b,t,y/ph,x/pw]. */
int cwm_cf_make_static(const float* x, const int64_t xs[5], const uint8_t* mask, int B, int T, int C, int H, int W,
                       int ph, int pw, float* out, cwm_stream_t stream);

/* Fused variants of cwm_patch_gather / cwm_unpatchify_scatter / cwm_vmae_forward that read the virtual video
 * (`src`) instead of a tensor: the 1.2 MB/sample `x_shift` is never written to or read from HBM.  Same semantics
 * as the tensor versions applied to cwm_cf_build_videos(src) viewed as [S, C, T, H, W]. */
int cwm_patch_gather_cf(const cwm_cf_source* src, int S, int C, int T, int H, int W, int pt, int ph, int pw,
                        const int32_t* perm, int Ntot, int rows_per_sample, const float* mean, const float* stdv,
                        uint16_t* out, cwm_stream_t stream);
int cwm_unpatchify_scatter_cf(const float* y, const cwm_cf_source* src, const int32_t* inv_perm, int S, int T, int C,
                              int H, int W, int pt, int ph, int pw, int Nvis, float* out, cwm_stream_t stream);

/* ==== SURVEY.md section 8(f) rank 2: flow-derived statistics of a counterfactual sweep ======================
 * `flows` is a logical [B, 2, H, W, S] fp32 tensor with arbitrary element strides fs[5] -- the reference hands the
 * permuted view '(b s) c h w -> b c h w s' of the flow network's output (cwm/models/segmentation.py:130-140).
 * fp32 like the reference; results match it to rounding (sums over pixels / samples are re-ordered). */

/* Per-sample statistics, stats fp32 [B, S, 5]:
 *   [0] patch_flow_mag: bilinear-downsampled (align_corners=False) magnitude averaged over the ACTIVE patches of
 *       frame 2 (`FlowSampleFilter.compute_flow_magnitude`, cwm/models/sampling.py:163-203); active = `active_patches`
 *       bytes, logical [B, 2*n_h*n_w, S] with element strides as[3], zero = active; NULL -> 0
 *   [1] flow_area: fraction of pixels whose magnitude exceeds magnitude_threshold (sampling.py:214-224)
 *   [2] num_corners: image corners whose magnitude exceeds magnitude_threshold (sampling.py:226-247)
 *   [3], [4] min / max magnitude over the image (`compute_flow_samples_magnitude`, segmentation.py:250-255) */
size_t cwm_flow_stats_workspace_bytes(int B, int H, int W, int S); /* scratch for the two reductions below */
int cwm_flow_sample_stats(const float* flows, const int64_t fs[5], int B, int H, int W, int S, const uint8_t* active,
                          const int64_t as[3], int n_h, int n_w, float magnitude_threshold, float* stats,
                          void* workspace, size_t workspace_bytes, cwm_stream_t stream);
/* filter_mask[b, s] = OR over the enabled methods (bit 0 patch_magnitude: stats[0] < magnitude_threshold; bit 1
 * flow_area: stats[1] > area_threshold; bit 2 num_corners: stats[2] >= corners_threshold) -- sampling.py:266-279. */
int cwm_flow_filter_mask(const float* stats, int B, int S, int methods, float magnitude_threshold, float area_threshold,
                         float corners_threshold, uint8_t* filter_mask, cwm_stream_t stream);
/* `flow_samples[filter_mask] = 0` in place on the strided view (sampling.py:281-284). */
int cwm_flow_zero_filtered(float* flows, const int64_t fs[5], int B, int H, int W, int S, const uint8_t* filter_mask,
                           cwm_stream_t stream);
/* sums[b, y, x] (+)= sum_s g_s(|flow[b, :, y, x, s]|): the numerator of `flow_mags.mean(-1)` (segmentation.py:257-267);
 * filter_mask (optional) treats filtered samples as zero flow; normalize_per_sample applies
 * (m - min_s) / max(max_s - min_s, eps) with stats from cwm_flow_sample_stats; accumulate != 0 adds to sums
 * (chunks of a sweep; partial sums of the ranks are combined with one all-reduce). */
int cwm_flow_magnitude_sum(const float* flows, const int64_t fs[5], int B, int H, int W, int S,
                           const uint8_t* filter_mask, const float* stats, int normalize_per_sample, float eps,
                           int accumulate, float* sums, void* workspace, size_t workspace_bytes, cwm_stream_t stream);
/* `FlowGenerator.compute_flow_corrs` with its default options (segmentation.py:478-547): per image the covariance
 * (use_covariance != 0, `torch.cov`) or correlation (`torch.corrcoef`, clamped to [-1, 1]) between the image locations
 * of the downsampled flow-magnitude samples, NaN -> 0.   out fp32 [B, N, N], N = (H/downsample)*(W/downsample). */
size_t cwm_flow_corrs_workspace_bytes(int B, int H, int W, int S, int downsample);
int cwm_flow_corrs(const float* flows, const int64_t fs[5], int B, int H, int W, int S, int downsample,
                   int use_covariance, float* out, void* workspace, size_t workspace_bytes, cwm_stream_t stream);
/* motion_map[b] = sums[b] / count, then (normalize != 0) minus its minimum and divided by its maximum clamped at eps
 * (`compute_mean_motion_map`, segmentation.py:268-276).  sums / motion_map fp32 [B, H, W]. */
int cwm_motion_map_finalize(const float* sums, int B, int H, int W, float count, int normalize, float eps,
                            float* motion_map, cwm_stream_t stream);

/* ==== SURVEY.md section 8(f) rank 3, first slice: the RAFT-specific stages of the flow network ================
 * (the convolutions of RAFT stay a caller-supplied torch module; these replace what is NOT a convolution.)
 * All tensors fp32, contiguous, NCHW like the reference.
 *
 * `CorrBlock.__init__` (cwm/models/raft/corr.py:12-28, :53-60): levels[0] = fmap1^T fmap2 / sqrt(D) as
 * [B*H*W, H, W]; levels[l] = avg_pool2d(levels[l-1], 2, stride=2) as [B*H*W, H>>l, W>>l].  `levels` is a HOST array
 * of num_levels DEVICE pointers (the reference's `corr_pyramid` list); fmap1 / fmap2 are [B, D, H, W].  Every level
 * must be at least 2x2 (the reference divides by size-1, cwm/models/raft/utils.py:64-65). */
int cwm_raft_corr_pyramid(const float* fmap1, const float* fmap2, int B, int D, int H, int W, int num_levels,
                          float* const* levels, cwm_stream_t stream);

/* The same pyramid with level 0 (the all-pairs volume) on the tensor cores: every operand is split into tf32(x) and
 * tf32(x - tf32(x)) and hi*hi + hi*lo + lo*hi is accumulated in fp32 by tcgen05.mma.kind::tf32 -- fp32 accuracy (the
 * dropped lo*lo term is 2^-22 of a product), corr.py:53-60.  D % 32 == 0; workspace: cwm_raft_corr_tc_workspace_bytes. */
size_t cwm_raft_corr_tc_workspace_bytes(int B, int D, int H, int W);
int cwm_raft_corr_volume_tc(const float* fmap1, const float* fmap2, int B, int D, int H, int W, float* out, void* workspace,
                            size_t workspace_bytes, cwm_stream_t stream);
int cwm_raft_corr_pyramid_tc(const float* fmap1, const float* fmap2, int B, int D, int H, int W, int num_levels,
                             float* const* levels, void* workspace, size_t workspace_bytes, cwm_stream_t stream);

/* The pyramid from f16 pixel-major feature rows (what the fused feature encoder writes: rows [n * H*W, D], D % 64 == 0):
 * f16 x f16 products are exact in the fp32 accumulator, so the mixed-precision path needs no fp32 cast, no NCHW transpose and
 * no hi / lo split -- one tcgen05.mma.kind::f16 per 16 channels.  n1 = 1: fmap1 is one image shared by all B samples (a
 * counterfactual sweep's frame 0), else n1 = B.  levels[l]: fp32 [B*H*W, H>>l, W>>l] as for cwm_raft_corr_pyramid. */
int cwm_raft_corr_volume_rows_f16(const uint16_t* rows1, int n1, const uint16_t* rows2, int B, int D, int H, int W, float* out,
                                  cwm_stream_t stream);
int cwm_raft_corr_pyramid_rows_f16(const uint16_t* rows1, int n1, const uint16_t* rows2, int B, int D, int H, int W,
                                   int num_levels, float* const* levels, cwm_stream_t stream);
/* The same with every level stored in f16 (levels[l]: f16 [B*H*W, H>>l, W>>l]; H*W % 8 == 0 for the vector stores): half the
 * bytes, so the level 0 of a 64-sample call (78 MB) stays in L2 across the 24 lookups.  Read by
 * cwm_raft_corr_lookup_f16_pyr16 (the separable f16 lookup: radius 4, <= 4 levels).  Mixed-precision path only: the lookup
 * output is f16 either way, the pyramid's own rounding (2^-11 relative) is of the same size. */
int cwm_raft_corr_volume_rows_f16_out16(const uint16_t* rows1, int n1, const uint16_t* rows2, int B, int D, int H, int W,
                                        uint16_t* out16, cwm_stream_t stream);
int cwm_raft_corr_pyramid_rows_f16_pyr16(const uint16_t* rows1, int n1, const uint16_t* rows2, int B, int D, int H, int W,
                                         int num_levels, uint16_t* const* levels, cwm_stream_t stream);
int cwm_raft_corr_lookup_f16_pyr16(const uint16_t* const* levels, int num_levels, int radius, const float* coords, int B, int H,
                                   int W, uint16_t* out16, int ld16, cwm_stream_t stream);
/* `CorrBlock.__call__` (corr.py:30-51) + `bilinear_sampler` (utils.py:60-80, grid_sample align_corners=True, zero
 * padding): coords [B, 2, H, W] (channel 0 = x, 1 = y, in level-0 pixels) -> out [B, num_levels*(2r+1)^2, H, W];
 * channel l*(2r+1)^2 + a*(2r+1) + b samples level l at (x/2^l + a - r, y/2^l + b - r). */
int cwm_raft_corr_lookup(const float* const* levels, int num_levels, int radius, const float* coords, int B, int H,
                         int W, float* out, cwm_stream_t stream);
/* `RAFT.upsample_flow` (cwm/models/raft/raft_model.py:175-186): flow [N, C, H, W], mask [N, 576, H, W] (9 x 8 x 8
 * logits per coarse pixel) -> out [N, C, 8H, 8W], the softmax-weighted combination of the 3x3 neighbours of 8*flow. */
int cwm_raft_upsample_flow(const float* flow, const float* mask, int N, int C, int H, int W, float* out,
                           cwm_stream_t stream);

/* ---- mixed-precision recurrent block of RAFT-large (cwm/models/raft/update.py:33-60, :79-98, :115-139) --------
 * The convolutions stay library calls WITHOUT bias on f16 pixel-major rows ([M = B*H*W, channels], "channels-last");
 * these entry points are everything between them.  Row pointers / leading dimensions must be 16-byte aligned. */
/* As cwm_raft_corr_lookup, but the result is written as f16 rows out16[M, ld16] (channels >= L*(2r+1)^2 zeroed):
 * the input layout of the first motion-encoder convolution (update.py:93). */
int cwm_raft_corr_lookup_f16(const float* const* levels, int num_levels, int radius, const float* coords, int B, int H,
                             int W, uint16_t* out16, int ld16, cwm_stream_t stream);
/* d1[m, c] (and d2[m, c] when d2 != NULL) = act(x[m, c] + bias[c]), c < C (C % 8 == 0; act 0 = none, 1 = relu); when
 * tail != NULL the last tail_cols columns are copied from tail[m, 0..tail_cols) instead -- `cat([out, flow])`,
 * update.py:98.  Replaces bias add + relu + torch.cat after a convolution. */
int cwm_raft_bias_act_f16(const uint16_t* x, int ldx, const float* bias, int act, int C, long long M, uint16_t* d1, int ld1,
                          uint16_t* d2, int ld2, const uint16_t* tail, int ldt, int tail_cols, cwm_stream_t stream);
/* zr[M, 2C] raw output of the stacked z|r convolution, bias[2C]: z_out[M, C] = sigmoid(z), rh[m, c] = sigmoid(r) * h
 * (update.py:46-47, :53-54); h / rh are slots (leading dimensions ldh / ldrh) of the GRU input rows. */
int cwm_raft_gru_gate_f16(const uint16_t* zr, const float* bias, const uint16_t* h, int ldh, int C, long long M,
                          uint16_t* z_out, uint16_t* rh, int ldrh, cwm_stream_t stream);
/* h <- (1 - z) * h + z * tanh(q + bias) in place (update.py:48-49, :55-56); h_dense (optional) gets a packed copy. */
int cwm_raft_gru_update_f16(const uint16_t* q, const float* bias, const uint16_t* z, uint16_t* h, int ldh, int C,
                            long long M, uint16_t* h_dense, cwm_stream_t stream);
/* coords1[B, 2, H, W] (fp32, in place) += delta[m, 0..1] + bias (raft_model.py:254); flow16[m, 0..7] = {coords1 - grid, 0..}
 * = the next iteration's flow input (raft_model.py:249). */
int cwm_raft_flow_update(const uint16_t* delta, int ldd, const float* bias, float* coords1, int B, int H, int W,
                         uint16_t* flow16, cwm_stream_t stream);

/* ---- convolutions of RAFT's recurrent block as implicit GEMMs on the tcgen05 kernel -------------------
 * Replaces the nn.Conv2d calls of cwm/models/raft/update.py:16-60 (SepConvGRU), :79-97 (BasicMotionEncoder), :6-14
 * (FlowHead) and :121-137 (the mask head) on f16 pixel-major rows.
 *
 * out[s, y, x, 0:Cout] = act(bias + sum over (ky, kx, c) of x[s, y + ky - pad_h, x + kx - pad_w, c] * w[n, ky, kx, c]),
 * stride 1, zero padding, kh = 2 pad_h + 1, kw = 2 pad_w + 1, image width <= 32.  `x` / `out` are NHWC f16 with pixel rows
 * ldx / ldo elements apart (so a convolution can read / write a column slice of a wider row buffer).  `w_packed` is
 * [Cout, kh, kw, cin_pad] f16 with cin_pad = Cin rounded up to 64 and zeros in the padding: cwm_conv2d_weight_k() columns.
 * The A tile of each (tap, 64-channel slab) k-step is one 4-D TMA box at the tap's offset; the padding is the TMA unit's
 * out-of-bounds zero fill, nothing is im2col-ed.  bias may be NULL; relu != 0 applies max(., 0). */
int cwm_conv2d_weight_k(int Cin, int kh, int kw);
int cwm_conv2d_f16(const uint16_t* x, int ldx, int S, int H, int W, int Cin, const uint16_t* w_packed, int Cout, int kh,
                   int kw, int pad_h, int pad_w, const float* bias, int relu, uint16_t* out, int ldo, cwm_stream_t stream);

/* The same with a stride (1 or 2) and no limit on the image width: the convolutions of RAFT's two encoders
 * (cwm/models/raft/extractor.py:6-56 ResidualBlock, :118-190 BasicEncoder: 3x3 / 1 and 3x3 / 2 convolutions and the 1x1 / 2
 * shortcuts on 112, 56 and 28 pixel maps).  H / W are the INPUT size; the output map is ((H - 1) / stride + 1) x ((W - 1) /
 * stride + 1) (nn.Conv2d with padding = k // 2) and `out` holds its pixel rows.  Maps wider than 32 pixels are tiled in both
 * directions (a 128-row tile = 8 x 16, 16 x 8 or 4 x 32 pixels); for stride 2 the A boxes are loaded through a tensor map
 * with traversal stride 2, so nothing is gathered or copied on the way either. */
int cwm_conv2d_strided_f16(const uint16_t* x, int ldx, int S, int H, int W, int Cin, const uint16_t* w_packed, int Cout, int kh,
                           int kw, int pad_h, int pad_w, int stride, const float* bias, int relu, uint16_t* out, int ldo,
                           cwm_stream_t stream);

/* cwm_conv2d_f16 writing its output rows to TWO row buffers (out2 may be NULL), with the last two output columns
 * optionally replaced by the 2 f16 at tail[row * ld_tail] (tail may be NULL): the motion encoder's last convolution
 * (126 features, update.py:96-98) feeds both GRU input buffers ([h | inp | motion, flow] and [r*h | inp | motion, flow],
 * :52-58) and closes their rows with the current flow -- no copy / concatenation kernel.  Maps of <= 32 pixels when a
 * tail is given. */
int cwm_conv2d_dual_f16(const uint16_t* x, int ldx, int S, int H, int W, int Cin, const uint16_t* w_packed, int Cout, int kh,
                        int kw, int pad_h, int pad_w, const float* bias, int relu, const uint16_t* tail, int ld_tail,
                        uint16_t* out, int ldo, uint16_t* out2, int ldo2, cwm_stream_t stream);

/* The two convolutions of one ConvGRU half-step with the gate arithmetic in their epilogues (update.py:43-60):
 *   gate:    [z | r] = sigmoid(conv(x, w_zr) + bias_zr), 2C output channels; z -> z_out [., C], r * h -> rh_out [., C]
 *            (the first C columns of the q convolution's input rows);
 *   update:  h <- (1 - z) * h + z * tanh(conv(x, w_q) + bias_q), in place in h's slot and, when h_dense != NULL, also to a
 *            dense [., C] copy.  h / z / outputs are f16 pixel rows with the given row strides; C % 64 == 0. */
int cwm_conv2d_gru_gate_f16(const uint16_t* x, int ldx, int S, int H, int W, int Cin, const uint16_t* w_zr, int C, int kh,
                            int kw, int pad_h, int pad_w, const float* bias_zr, const uint16_t* h, int ldh, uint16_t* z_out,
                            int ldz, uint16_t* rh_out, int ldrh, cwm_stream_t stream);
int cwm_conv2d_gru_update_f16(const uint16_t* x, int ldx, int S, int H, int W, int Cin, const uint16_t* w_q, int C, int kh,
                              int kw, int pad_h, int pad_w, const float* bias_q, const uint16_t* z, int ldz, uint16_t* h,
                              int ldh, uint16_t* h_dense, cwm_stream_t stream);

/* cwm_raft_flow_update with the flow head's last convolution (3x3, 256 -> 2; update.py:13-14) finished inside: `taps`
 * [B*H*W, ldt] holds per pixel the 18 per-tap products of a 1x1 GEMM (column (ky*3 + kx)*2 + co = w[co, :, ky, kx] . x[pixel]),
 * delta = bias + the 3x3 stencil sum of the neighbours' taps (zero outside the image); coords1 += delta; the new flow goes to
 * flow16 [., 8] and, as 2 f16, to dst1 / dst2 (the flow columns of the GRU input rows; either may be NULL). */
int cwm_raft_flow_update_taps(const uint16_t* taps, int ldt, const float* bias, float* coords1, int B, int H, int W,
                              uint16_t* flow16, uint16_t* dst1, int ld1, uint16_t* dst2, int ld2, cwm_stream_t stream);

/* im2col of the 2-channel flow rows for the k x k convolution of BasicMotionEncoder.convf1 (update.py:85): out[m, 2 tap + c],
 * taps in (ky, kx) order, zero outside the image and in the columns >= 2 k^2. */
int cwm_raft_im2col_flow(const uint16_t* flow16, int ldf, int B, int H, int W, int k, uint16_t* out, int ldo,
                         cwm_stream_t stream);

/* Instance normalisation of NHWC f16 maps fused with what follows it in RAFT's feature encoder (extractor.py:118-190,
 * nn.InstanceNorm2d: no affine, biased variance): out = relu_outer?( add? + relu_inner?( (x - mean[s,c]) * rstd[s,c] ) ).
 * x / add / out: [S, HW, C] f16, C % 8 == 0; fp32 statistics over the HW pixels of every (sample, channel). */
size_t cwm_instnorm_workspace_bytes(int S, int C);
int cwm_instnorm_f16(const uint16_t* x, int S, int HW, int C, float eps, int relu_inner, const uint16_t* add, int relu_outer,
                     uint16_t* out, void* workspace, size_t workspace_bytes, cwm_stream_t stream);

/* im2col of a few-channel NCHW fp32 image for a strided k x k convolution -- the encoders' 7x7 / 2 stem on the 3-channel frame
 * (extractor.py:132, :172): out[(s, oy, ox), (ky * k + kx) * Cin + c] = scale * img[s, c, stride oy + ky - pad, stride ox + kx
 * - pad] + shift, zero outside the image and in the columns [k k Cin, ldo); the stem is then one cwm_gemm_f16 with K = ldo.
 * scale / shift carry RAFT's input normalisation 2 (x / 255) - 1 (raft_model.py:205-206); sample_stride (elements, 0 = Cin*H*W)
 * lets `img` be one frame of a [S, T, Cin, H, W] movie, so the frames are neither copied nor normalised by separate kernels. */
int cwm_im2col_nchw_f16(const float* img, long long sample_stride, int S, int Cin, int H, int W, int k, int stride, int pad,
                        float scale, float shift, uint16_t* out, int ldo, cwm_stream_t stream);

/* out = relu?(a + b) over n f16 elements (n % 8 == 0): the residual joins of the context encoder, whose batch norms are
 * folded into the convolution weights at inference (extractor.py:46-56). */
int cwm_add_act_f16(const uint16_t* a, const uint16_t* b, long long n, int relu, uint16_t* out, cwm_stream_t stream);

/* ---- masks on device with a counter-based RNG (csrc/masks.cu) --------------------------------------
 * Opt-in stand-ins for the reference's host-side mask generation (cwm/models/masking.py:347-401 MaskingGenerator.
 * sample_mask_per_frame, :478-545 RotatedTableUniformMaskingGenerator; sampling.py:63-90 EnergySamplingMaskingGenerator;
 * utils.py:152-213 sample_from_energy; masking.py:100-132 RectangularizeMasks).  Draws come from Philox4x32-10 keyed by
 * `seed` with the counter (draw, GLOBAL sample index, stream, sub-stream): a sample's mask does not depend on the batch
 * split or the number of GPUs.  Masks are uint8 rows [frames * h * w], 1 = masked. */

/* Host-side reference point of the generator (no device work): out = Philox4x32-10(counter, key). */
int cwm_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);

/* rows masks [rows, (visible_frames + mask_frames) * h * w]: the first `visible_frames` frames fully visible, in every
 * masked frame `n_visible_cells` of the (h/clump) x (w/clump) clump cells visible, uniformly without replacement.
 * Row r is global sample row0 + r. */
int cwm_mask_uniform(uint64_t seed, int row0, int rows, int visible_frames, int mask_frames, int h, int w, int clump,
                     int n_visible_cells, uint8_t* masks, cwm_stream_t stream);

/* probs [B, n] (fp32 weights of the n clump cells of each image) -> inclusive integer cumulative table [B, n] of
 * floor(relu(p - min p + eps) / max * 2^24) (utils.py:160-163 with normalize=True, quantised). */
int cwm_mask_energy_table(const float* probs, int B, int n, float eps, uint64_t* table, cwm_stream_t stream);

/* masks [B * S, (visible_frames + 1) * h * w]: for image b and sample s (global index sample0 + s), `points` clump cells
 * drawn WITH replacement from table[b] are visible in the last frame (duplicates collapse, as in the reference). */
int cwm_mask_energy_sample(const uint64_t* table, int B, int h, int w, int clump, uint64_t seed, int sample0, int S,
                           int points, int visible_frames, uint8_t* masks, cwm_stream_t stream);

/* RectangularizeMasks('min') in place on masks [rows, N]: every row keeps `target_masked` masked tokens (< 0: the minimum
 * over the rows given); the revealed tokens of row r are the masked ones with the smallest Philox keys of global row
 * row0 + r.  Pass the global minimum as `target_masked` when the rows are a shard of a larger batch. */
size_t cwm_mask_rectangularize_workspace_bytes(int rows);
int cwm_mask_rectangularize(uint8_t* masks, int rows, int N, int row0, uint64_t seed, int target_masked, void* workspace,
                            size_t workspace_bytes, cwm_stream_t stream);

/* Number of kernel launches this thread enqueued through the library since the last cwm_vmae_forward began or
 * cwm_launch_count_reset() was called (for bench accounting). */
int cwm_last_forward_launches(void);
int cwm_launch_count_reset(void);
/* Monotonic count of every kernel launch this thread ever enqueued through the library (never reset): the difference
 * of two reads brackets a region, e.g. bench.py's timed steps. */
long long cwm_total_launches(void);

/* ---- tuning hooks -------------------------------------------------------------------------------------------------
 * Process-wide switches used by tools/kernel_bench.py and the A/B tests; they select between result-identical kernel
 * variants (or, for the polynomial share, variants inside the same tolerance) and are not part of the reference-facing
 * boundary. */
void cwm_debug_attention_poly(int eighths);        /* share (in 1/8) of the softmax exponentials evaluated on the FMA pipe */
void cwm_debug_attention_war_safe(int on);         /* conservative score-buffer reuse (debug) */
void cwm_debug_attention_stale_max(int on);        /* 1 (default): exponentials first, tile maximum checked afterwards */
void cwm_debug_attention_skip_idle(int on);        /* 1 (default): warps whose query rows all lie beyond the sequence skip the tile */
void cwm_debug_attention_persistent(int mode);     /* 1 = persistent CTAs (default), 3 = one work item per CTA; 2 / 4 = the same with watchdog waits */
void cwm_debug_attention_persist_map(int m);       /* work-item map of the persistent kernel: -1 auto, 0 ranges, 1 strided */
void cwm_debug_attn_mma_wide(int on);              /* small-attention kernel: 8 warps per K/V tile */
void cwm_debug_attn_mma_split(int keys);           /* small-attention kernel: key-axis split */
int cwm_debug_gemm_cta2(int enable);               /* CTA-pair GEMM kernels on (default) / off; returns CWM_OK */

/* ---- per-kernel timing (bench.py's roofline numbers) -------------------------------------------------
 * Between cwm_profile_begin() and cwm_profile_end() every launch made through this library is bracketed by two
 * CUDA events on its own stream.  cwm_profile_end() synchronises those events and aggregates per kernel class:
 * number of launches, total device milliseconds, algorithmic FLOPs and algorithmic bytes. */
typedef struct cwm_profile_entry {
  char name[32];
  int32_t launches;
  double ms;
  double flops;
  double bytes;
} cwm_profile_entry;
int cwm_profile_begin(void);
int cwm_profile_end(cwm_profile_entry* out, int max_entries, int* n_entries);

#ifdef __cplusplus
}
#endif
#endif /* CWM_B200_H_ */
