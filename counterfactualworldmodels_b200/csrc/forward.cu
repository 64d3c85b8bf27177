// forward.cu -- cwm_vmae_forward: the whole `PretrainVisionTransformer.forward(x, mask)` chain
// (cwm/models/VideoMAE/vmae.py:539-560) as a fixed sequence of launches on one stream.  No host
// synchronisation, no allocation: the caller provides the workspace.
#include "common.cuh"

namespace cwm {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Plan {
  int Ntot, Nvis, Nmask, out_rows, Kp;
  long long Me, Md, Mo;
  size_t off_xe, off_xd, off_a16, off_qkv, off_attn, off_h16, off_stats, total;
};
constexpr int kMaxLnParts = 16;  // partial-statistics planes of a LayerNorm-producer GEMM (2 per N tile, N <= 1024)

static int make_plan(const cwm_vmae_model* m, int B, int Nvis, Plan* p) {
  if (!m) return fail(CWM_ERR_INVALID, "cwm_vmae: null model");
  if (m->pt <= 0 || m->ph <= 0 || m->pw <= 0 || m->num_frames % m->pt || m->img_h % m->ph || m->img_w % m->pw)
    return fail(CWM_ERR_INVALID, "Input image size(%d,%d) must be divisible by patch size (%d,%d)", m->img_h,
                m->img_w, m->ph, m->pw);
  p->Ntot = (m->num_frames / m->pt) * (m->img_h / m->ph) * (m->img_w / m->pw);
  if (B < 0 || Nvis < 0 || Nvis > p->Ntot) return fail(CWM_ERR_INVALID, "cwm_vmae: bad B=%d / Nvis=%d (Ntot=%d)", B, Nvis, p->Ntot);
  if (m->enc_dim % 128 || m->dec_dim % 128 || m->enc_dim > 1024 || m->dec_dim > 1024)
    return fail(CWM_ERR_UNSUPPORTED, "cwm_vmae: embed dims (%d, %d) must be multiples of 128 and <= 1024", m->enc_dim, m->dec_dim);
  if (m->enc_dim != m->enc_heads * 64 || m->dec_dim != m->dec_heads * 64)
    return fail(CWM_ERR_UNSUPPORTED, "cwm_vmae: only head_dim 64 is implemented (enc %d/%d, dec %d/%d)", m->enc_dim,
                m->enc_heads, m->dec_dim, m->dec_heads);
  p->Nvis = Nvis;
  p->Nmask = p->Ntot - Nvis;
  p->out_rows = p->Nmask > 0 ? p->Nmask : p->Ntot;  // return_token_num == 0 -> all tokens (vmae.py:250-253)
  p->Kp = m->in_chans * m->pt * m->ph * m->pw;
  if (p->Kp % 16) return fail(CWM_ERR_UNSUPPORTED, "cwm_vmae: patch volume %d must be a multiple of 16", p->Kp);
  p->Me = static_cast<long long>(B) * Nvis;
  p->Md = static_cast<long long>(B) * p->Ntot;
  p->Mo = static_cast<long long>(B) * p->out_rows;
  auto mx = [](size_t a, size_t b) { return a > b ? a : b; };
  size_t off = 0;
  p->off_xe = off;  off = align_up(off + p->Me * m->enc_dim * 4, 1024);
  p->off_xd = off;  off = align_up(off + p->Md * m->dec_dim * 4, 1024);
  size_t a16 = mx(mx(p->Me * m->enc_dim, p->Md * m->dec_dim), mx(p->Me * p->Kp, p->Mo * m->dec_dim));
  p->off_a16 = off; off = align_up(off + a16 * 2, 1024);
  p->off_qkv = off; off = align_up(off + mx(p->Me * 3 * m->enc_dim, p->Md * 3 * m->dec_dim) * 2, 1024);
  p->off_attn = off; off = align_up(off + mx(p->Me * m->enc_dim, p->Md * m->dec_dim) * 2, 1024);
  p->off_h16 = off; off = align_up(off + mx(p->Me * m->enc_hidden, p->Md * m->dec_hidden) * 2, 1024);
  p->off_stats = off; off = align_up(off + mx(p->Me, p->Md) * kMaxLnParts * 8, 1024);
  p->total = off + 1024;
  return CWM_OK;
}

static cwm_gemm_epilogue epi_f16(const float* bias, float scale, int scale_cols, void* out, int ldo) {
  cwm_gemm_epilogue e = {};
  e.mode = CWM_EPI_F16; e.bias = bias; e.scale = scale; e.scale_cols = scale_cols; e.out = out; e.ldo = ldo;
  return e;
}
static cwm_gemm_epilogue epi_gelu(const float* bias, void* out, int ldo) {
  cwm_gemm_epilogue e = {};
  e.mode = CWM_EPI_GELU_F16; e.bias = bias; e.out = out; e.ldo = ldo;
  return e;
}
static cwm_gemm_epilogue epi_res(const float* bias, float* x, int ld) {
  cwm_gemm_epilogue e = {};
  e.mode = CWM_EPI_RES_F32; e.bias = bias; e.res = x; e.ldr = ld; e.out = x; e.ldo = ld;
  return e;
}

#define CWM_TRY(call)        \
  do {                       \
    int _rc = (call);        \
    if (_rc != CWM_OK) return _rc; \
  } while (0)

// x += Attn(LN1(x)); x += Mlp(LN2(x))   (cwm/models/VideoMAE/utils.py:146-153, gamma_* = None)
// inner attention width A = heads * head_dim; `attn` may alias `a16` (the LN output is dead once qkv is computed).
//
// Fused-LayerNorm path (folded weights present, a statistics buffer, attn != a16): no LayerNorm kernel at all.
// `a16` then holds f16(x) written by the epilogue of the GEMM that last updated x (or by the row-statistics kernel
// for the first block of a stream) together with per-row partial statistics in `stats`; the qkv / fc1 epilogues
// normalise.  *x16_valid says whether a16 / stats describe the current x on entry; emit_for_next asks fc2 to produce
// them for the following block.
static int run_block(const cwm_block_weights& w, float* x, int Bn, int N, int C, int heads, int head_dim, int hidden,
                     float eps, float qk_scale, uint16_t* a16, uint16_t* qkv, uint16_t* attn, uint16_t* h16,
                     cwm_stream_t st, float* stats = nullptr, bool* x16_valid = nullptr, bool emit_for_next = false) {
  const int M = Bn * N;
  const int A = heads * head_dim;
  const bool fused = w.w_qkv_ln != nullptr && w.w_fc1_ln != nullptr && stats != nullptr && x16_valid != nullptr &&
                     attn != a16 && C % 32 == 0 && cwm_gemm_ln_parts(C) <= kMaxLnParts;
  if (fused) {
    const int parts_c = cwm_gemm_ln_parts(C);
    int parts = parts_c;
    if (!*x16_valid) {
      CWM_TRY(cwm_rowstats_f16(x, M, C, a16, stats, st));
      parts = 1;
    }
    auto with_ln = [&](cwm_gemm_epilogue e, const float* colsum, int n_parts) {
      e.ln_stats_in = stats; e.ln_parts = n_parts; e.ln_colsum = colsum; e.ln_width = C; e.ln_eps = eps;
      return e;
    };
    auto with_emit = [&](cwm_gemm_epilogue e) {
      e.ln_x16 = a16; e.ln_ldx16 = C; e.ln_stats_out = stats;
      return e;
    };
    cwm_gemm_epilogue e = with_ln(epi_f16(w.c_qkv, qk_scale, A, qkv, 3 * A), w.s_qkv, parts);
    CWM_TRY(cwm_gemm_f16(a16, w.w_qkv_ln, M, 3 * A, C, &e, st));
    if (head_dim == 64) {
      CWM_TRY(cwm_attention_f16(qkv, Bn, N, heads, 64, attn, st));
    } else {
      CWM_TRY(cwm_attention_generic_f16(qkv, qkv + A, qkv + 2 * A, 3 * A, 3 * A, 3 * A, head_dim, head_dim, head_dim, Bn,
                                        N, N, heads, head_dim, attn, A, nullptr, 0, st));
    }
    e = with_emit(epi_res(w.b_proj, x, C));
    CWM_TRY(cwm_gemm_f16(attn, w.w_proj, M, C, A, &e, st));
    e = with_ln(epi_gelu(w.c_fc1, h16, hidden), w.s_fc1, parts_c);
    CWM_TRY(cwm_gemm_f16(a16, w.w_fc1_ln, M, hidden, C, &e, st));
    e = epi_res(w.b_fc2, x, C);
    if (emit_for_next) e = with_emit(e);
    CWM_TRY(cwm_gemm_f16(h16, w.w_fc2, M, C, hidden, &e, st));
    *x16_valid = emit_for_next;
    return CWM_OK;
  }
  if (x16_valid != nullptr) *x16_valid = false;
  CWM_TRY(cwm_layernorm_f16(x, M, C, w.ln1_g, w.ln1_b, eps, 0, 0, 0, a16, st));
  cwm_gemm_epilogue e = epi_f16(w.b_qkv, qk_scale, A, qkv, 3 * A);  // (xW + [q_bias,0,v_bias]); q *= scale
  CWM_TRY(cwm_gemm_f16(a16, w.w_qkv, M, 3 * A, C, &e, st));
  if (head_dim == 64) {
    CWM_TRY(cwm_attention_f16(qkv, Bn, N, heads, 64, attn, st));
  } else {
    CWM_TRY(cwm_attention_generic_f16(qkv, qkv + A, qkv + 2 * A, 3 * A, 3 * A, 3 * A, head_dim, head_dim, head_dim, Bn,
                                      N, N, heads, head_dim, attn, A, nullptr, 0, st));
  }
  e = epi_res(w.b_proj, x, C);
  CWM_TRY(cwm_gemm_f16(attn, w.w_proj, M, C, A, &e, st));
  CWM_TRY(cwm_layernorm_f16(x, M, C, w.ln2_g, w.ln2_b, eps, 0, 0, 0, a16, st));
  e = epi_gelu(w.b_fc1, h16, hidden);
  CWM_TRY(cwm_gemm_f16(a16, w.w_fc1, M, hidden, C, &e, st));
  e = epi_res(w.b_fc2, x, C);
  CWM_TRY(cwm_gemm_f16(h16, w.w_fc2, M, C, hidden, &e, st));
  return CWM_OK;
}

struct BlockPlan {
  size_t off_a16, off_qkv, off_h16, off_attn, off_stats, total;
};
static void make_block_plan(long long M, int C, int A, int hidden, BlockPlan* p) {
  size_t off = 0;
  p->off_a16 = off; off = align_up(off + M * (C > A ? C : A) * 2, 1024);
  p->off_qkv = off; off = align_up(off + M * 3 * A * 2, 1024);
  p->off_h16 = off; off = align_up(off + M * hidden * 2, 1024);
  // separate attention output + statistics planes: lets a block with LayerNorm-folded weights run without its second
  // LayerNorm kernel (the first one becomes the row-statistics pass: a single block call cannot know who wrote x last)
  p->off_attn = off; off = align_up(off + M * A * 2, 1024);
  p->off_stats = off; off = align_up(off + M * kMaxLnParts * 8, 1024);
  p->total = off + 1024;
}

struct CrossPlan {
  size_t off_a16, off_qkv, off_y, off_h16, off_a16s, off_qkvs, off_ys, off_h16s, off_attn_ws, attn_ws_bytes, total;
};
static void make_cross_plan(int B, int N, int M, int C, int Cs, int heads, int head_dim, int hidden, int hidden_s,
                            CrossPlan* p) {
  const long long Mx = static_cast<long long>(B) * N, Ms = static_cast<long long>(B) * M;
  const int D = heads * head_dim;
  size_t off = 0;
  p->off_a16 = off;  off = align_up(off + Mx * C * 2, 1024);
  p->off_qkv = off;  off = align_up(off + Mx * 3 * D * 2, 1024);
  p->off_y = off;    off = align_up(off + Mx * D * 2, 1024);
  p->off_h16 = off;  off = align_up(off + Mx * hidden * 2, 1024);
  p->off_a16s = off; off = align_up(off + Ms * Cs * 2, 1024);
  p->off_qkvs = off; off = align_up(off + Ms * 3 * D * 2, 1024);
  p->off_ys = off;   off = align_up(off + Ms * D * 2, 1024);
  p->off_h16s = off; off = align_up(off + Ms * hidden_s * 2, 1024);
  const size_t a = cwm_attention_generic_workspace_bytes(B, N, M, heads, head_dim);
  const size_t b = cwm_attention_generic_workspace_bytes(B, M, N, heads, head_dim);
  p->attn_ws_bytes = a > b ? a : b;
  p->off_attn_ws = off; off = align_up(off + p->attn_ws_bytes, 1024);
  p->total = off + 1024;
}

}  // namespace cwm

using namespace cwm;

extern "C" size_t cwm_vmae_workspace_bytes(const cwm_vmae_model* model, int B, int Nvis) {
  Plan p;
  if (make_plan(model, B, Nvis, &p) != CWM_OK) return 0;
  return p.total;
}

// The input is either a strided tensor (x, xs) or a counterfactual descriptor (cf): only the gather differs.
static int vmae_forward_impl(const cwm_vmae_model* m, const float* x, const int64_t* xs, const cwm_cf_source* cf, int B,
                             const float* norm_mean, const float* norm_std, const int32_t* perm, int Nvis,
                             float* y, void* workspace, size_t workspace_bytes, cwm_stream_t st) {
  reset_launches();
  Plan p;
  CWM_TRY(make_plan(m, B, Nvis, &p));
  CWM_REQUIRE(((x && xs) || cf) && perm && y && workspace, "cwm_vmae_forward: null pointer");
  if (workspace_bytes < p.total)
    return fail(CWM_ERR_WORKSPACE, "cwm_vmae_forward: workspace %zu bytes < required %zu", workspace_bytes, p.total);
  if (B == 0) return CWM_OK;
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  float* xe = reinterpret_cast<float*>(ws + p.off_xe);
  float* xd = reinterpret_cast<float*>(ws + p.off_xd);
  uint16_t* a16 = reinterpret_cast<uint16_t*>(ws + p.off_a16);
  uint16_t* qkv = reinterpret_cast<uint16_t*>(ws + p.off_qkv);
  uint16_t* attn = reinterpret_cast<uint16_t*>(ws + p.off_attn);
  uint16_t* h16 = reinterpret_cast<uint16_t*>(ws + p.off_h16);
  float* stats = reinterpret_cast<float*>(ws + p.off_stats);
  const int Ce = m->enc_dim, Cd = m->dec_dim;

  if (Nvis > 0) {
    // a1-a4: normalise + gather visible patches, embed, add positional embedding of the gathered tokens
    if (cf) {
      CWM_TRY(cwm_patch_gather_cf(cf, B, m->in_chans, m->num_frames, m->img_h, m->img_w, m->pt, m->ph, m->pw, perm,
                                  p.Ntot, Nvis, norm_mean, norm_std, a16, st));
    } else {
      CWM_TRY(cwm_patch_gather(x, xs, B, m->in_chans, m->num_frames, m->img_h, m->img_w, m->pt, m->ph, m->pw, perm,
                               p.Ntot, Nvis, norm_mean, norm_std, a16, st));
    }
    cwm_gemm_epilogue e = {};
    e.mode = CWM_EPI_RES_F32; e.bias = m->b_patch; e.res = m->pos_enc; e.ldr = Ce; e.res_gather = perm;
    e.gather_stride = p.Ntot; e.grp_rows = Nvis; e.grp_out_stride = Nvis; e.out = xe; e.ldo = Ce;
    CWM_TRY(cwm_gemm_f16(a16, m->w_patch, static_cast<int>(p.Me), Ce, p.Kp, &e, st));
    // a5-a7: encoder blocks
    bool x16_valid = false;
    for (int l = 0; l < m->enc_depth; ++l)
      CWM_TRY(run_block(m->enc_blocks[l], xe, B, Nvis, Ce, m->enc_heads, 64, m->enc_hidden, m->ln_eps, m->enc_qk_scale,
                        a16, qkv, attn, h16, st, stats, &x16_valid, l + 1 < m->enc_depth));
    // a8-a10: final norm, encoder_to_decoder (no bias) written at the visible rows of the decoder sequence with
    // the positional embedding of each visible token added
    CWM_TRY(cwm_layernorm_f16(xe, static_cast<int>(p.Me), Ce, m->enc_norm_g, m->enc_norm_b, m->ln_eps, 0, 0, 0, a16, st));
    e = {};
    e.mode = CWM_EPI_RES_F32; e.bias = nullptr; e.res = m->pos_dec; e.ldr = Cd; e.res_gather = perm;
    e.gather_stride = p.Ntot; e.grp_rows = Nvis; e.grp_out_stride = p.Ntot; e.out = xd; e.ldo = Cd;
    CWM_TRY(cwm_gemm_f16(a16, m->w_e2d, static_cast<int>(p.Me), Cd, Ce, &e, st));
  }
  CWM_TRY(cwm_fill_mask_tokens(m->mask_token, m->pos_dec, perm, B, p.Ntot, Nvis, Cd, xd, st));
  // a11: decoder blocks over all Ntot tokens, then head(norm(last Nmask tokens))
  bool xd16_valid = false;
  for (int l = 0; l < m->dec_depth; ++l)
    CWM_TRY(run_block(m->dec_blocks[l], xd, B, p.Ntot, Cd, m->dec_heads, 64, m->dec_hidden, m->ln_eps, m->dec_qk_scale, a16,
                      qkv, attn, h16, st, stats, &xd16_valid, l + 1 < m->dec_depth));
  if (p.Nmask > 0) {
    CWM_TRY(cwm_layernorm_f16(xd, static_cast<int>(p.Mo), Cd, m->dec_norm_g, m->dec_norm_b, m->ln_eps, p.Nmask, p.Ntot,
                              Nvis, a16, st));
  } else {
    CWM_TRY(cwm_layernorm_f16(xd, static_cast<int>(p.Mo), Cd, m->dec_norm_g, m->dec_norm_b, m->ln_eps, 0, 0, 0, a16, st));
  }
  cwm_gemm_epilogue e = {};
  e.mode = CWM_EPI_F32; e.bias = m->b_head; e.out = y; e.ldo = m->out_dim;
  CWM_TRY(cwm_gemm_f16(a16, m->w_head, static_cast<int>(p.Mo), m->out_dim, Cd, &e, st));
  return CWM_OK;
}

extern "C" int cwm_vmae_forward(const cwm_vmae_model* m, const float* x, const int64_t xs[5], int B,
                                const float* norm_mean, const float* norm_std, const int32_t* perm, int Nvis,
                                float* y, void* workspace, size_t workspace_bytes, cwm_stream_t st) {
  CWM_REQUIRE(x && xs, "cwm_vmae_forward: null pointer");
  return vmae_forward_impl(m, x, xs, nullptr, B, norm_mean, norm_std, perm, Nvis, y, workspace, workspace_bytes, st);
}

extern "C" int cwm_vmae_forward_cf(const cwm_vmae_model* m, const cwm_cf_source* src, int S, const float* norm_mean,
                                   const float* norm_std, const int32_t* perm, int Nvis, float* y, void* workspace,
                                   size_t workspace_bytes, cwm_stream_t st) {
  CWM_REQUIRE(src, "cwm_vmae_forward_cf: null pointer");
  return vmae_forward_impl(m, nullptr, nullptr, src, S, norm_mean, norm_std, perm, Nvis, y, workspace, workspace_bytes,
                           st);
}

// ---- block-level entry points used by the conjoined (IMU-conditioned) models ------------------------------------
static uint8_t* align_ws(void* workspace) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
}

extern "C" size_t cwm_block_workspace_bytes(int B, int N, int C, int heads, int head_dim, int hidden) {
  if (B < 0 || N < 0 || C <= 0 || heads <= 0 || head_dim <= 0 || hidden <= 0) return 0;
  BlockPlan p;
  make_block_plan(static_cast<long long>(B) * N, C, heads * head_dim, hidden, &p);
  return p.total;
}

extern "C" int cwm_block_forward(const cwm_block_weights* w, float* x, int B, int N, int C, int heads, int head_dim,
                                 int hidden, float ln_eps, float qk_scale, void* workspace, size_t workspace_bytes,
                                 cwm_stream_t st) {
  CWM_REQUIRE(w && x && workspace, "cwm_block_forward: null pointer");
  CWM_REQUIRE(B >= 0 && N > 0 && C > 0 && heads > 0 && head_dim > 0 && hidden > 0, "cwm_block_forward: bad shape");
  BlockPlan p;
  make_block_plan(static_cast<long long>(B) * N, C, heads * head_dim, hidden, &p);
  if (workspace_bytes < p.total)
    return fail(CWM_ERR_WORKSPACE, "cwm_block_forward: workspace %zu bytes < required %zu", workspace_bytes, p.total);
  if (B == 0) return CWM_OK;
  uint8_t* ws = align_ws(workspace);
  uint16_t* a16 = reinterpret_cast<uint16_t*>(ws + p.off_a16);
  bool x16_valid = false;
  return run_block(*w, x, B, N, C, heads, head_dim, hidden, ln_eps, qk_scale, a16,
                   reinterpret_cast<uint16_t*>(ws + p.off_qkv), reinterpret_cast<uint16_t*>(ws + p.off_attn),
                   reinterpret_cast<uint16_t*>(ws + p.off_h16), st, reinterpret_cast<float*>(ws + p.off_stats), &x16_valid,
                   false);
}

extern "C" size_t cwm_cross_block_workspace_bytes(int B, int N, int M, int C, int Cs, int heads, int head_dim,
                                                  int hidden, int hidden_s) {
  if (B < 0 || N <= 0 || M <= 0 || C <= 0 || Cs <= 0 || heads <= 0 || head_dim <= 0) return 0;
  CrossPlan p;
  make_cross_plan(B, N, M, C, Cs, heads, head_dim, hidden, hidden_s, &p);
  return p.total;
}

extern "C" int cwm_cross_block_forward(const cwm_cross_block_weights* w, float* x, float* src, int B, int N, int M,
                                       int C, int Cs, int heads, int head_dim, int hidden, int hidden_s, float eps,
                                       float scale, void* workspace, size_t workspace_bytes, cwm_stream_t st) {
  CWM_REQUIRE(w && x && src && workspace, "cwm_cross_block_forward: null pointer");
  CWM_REQUIRE(B >= 0 && N > 0 && M > 0 && C > 0 && Cs > 0 && heads > 0 && head_dim > 0 && hidden > 0 && hidden_s > 0,
              "cwm_cross_block_forward: bad shape");
  CrossPlan p;
  make_cross_plan(B, N, M, C, Cs, heads, head_dim, hidden, hidden_s, &p);
  if (workspace_bytes < p.total)
    return fail(CWM_ERR_WORKSPACE, "cwm_cross_block_forward: workspace %zu bytes < required %zu", workspace_bytes, p.total);
  if (B == 0) return CWM_OK;
  uint8_t* ws = align_ws(workspace);
  uint16_t* a16 = reinterpret_cast<uint16_t*>(ws + p.off_a16);
  uint16_t* qkv = reinterpret_cast<uint16_t*>(ws + p.off_qkv);
  uint16_t* y = reinterpret_cast<uint16_t*>(ws + p.off_y);
  uint16_t* h16 = reinterpret_cast<uint16_t*>(ws + p.off_h16);
  uint16_t* a16s = reinterpret_cast<uint16_t*>(ws + p.off_a16s);
  uint16_t* qkvs = reinterpret_cast<uint16_t*>(ws + p.off_qkvs);
  uint16_t* ys = reinterpret_cast<uint16_t*>(ws + p.off_ys);
  uint16_t* h16s = reinterpret_cast<uint16_t*>(ws + p.off_h16s);
  void* attn_ws = ws + p.off_attn_ws;
  const int Mx = B * N, Ms = B * M, D = heads * head_dim, hd = head_dim;

  // qk | v of both streams from the cross norms (transformer.py:539-546, :333-336); the softmax scale multiplies
  // both halves of the main stream's qk (transformer.py:359, :363), qk_src stays unscaled
  CWM_TRY(cwm_layernorm_f16(x, Mx, C, w->ln1_g, w->ln1_b, eps, 0, 0, 0, a16, st));
  cwm_gemm_epilogue e = epi_f16(nullptr, scale, 2 * D, qkv, 3 * D);
  CWM_TRY(cwm_gemm_f16(a16, w->w_qkv, Mx, 3 * D, C, &e, st));
  CWM_TRY(cwm_layernorm_f16(src, Ms, Cs, w->ln1s_g, w->ln1s_b, eps, 0, 0, 0, a16s, st));
  e = epi_f16(nullptr, 1.0f, 0, qkvs, 3 * D);
  CWM_TRY(cwm_gemm_f16(a16s, w->w_qkv_s, Ms, 3 * D, Cs, &e, st));
  // head h of qk occupies columns [2*hd*h, 2*hd*(h+1)): first hd = the "trg" similarity, last hd = the "src" one
  // attn     = softmax(qk[..., :hd] qk_src[..., :hd]^T);   y     = attn @ v_src      (transformer.py:358-361, :370)
  CWM_TRY(cwm_attention_generic_f16(qkv, qkvs, qkvs + 2 * D, 3 * D, 3 * D, 3 * D, 2 * hd, 2 * hd, hd, B, N, M, heads, hd, y,
                                    D, attn_ws, p.attn_ws_bytes, st));
  // attn_src = softmax(qk_src[..., hd:] qk[..., hd:]^T);   y_src = attn_src @ v      (transformer.py:362-365, :371)
  CWM_TRY(cwm_attention_generic_f16(qkvs + hd, qkv + hd, qkv + 2 * D, 3 * D, 3 * D, 3 * D, 2 * hd, 2 * hd, hd, B, M, N, heads,
                                    hd, ys, D, attn_ws, p.attn_ws_bytes, st));
  // x += projection(y); src += projection_src(y_src)   (transformer.py:374-375, :569-575 with gamma_1 = 0)
  e = epi_res(w->b_proj, x, C);
  CWM_TRY(cwm_gemm_f16(y, w->w_proj, Mx, C, D, &e, st));
  e = epi_res(w->b_proj_s, src, Cs);
  CWM_TRY(cwm_gemm_f16(ys, w->w_proj_s, Ms, Cs, D, &e, st));
  // per-stream MLP (transformer.py:578-580)
  CWM_TRY(cwm_layernorm_f16(x, Mx, C, w->ln2_g, w->ln2_b, eps, 0, 0, 0, a16, st));
  e = epi_gelu(w->b_fc1, h16, hidden);
  CWM_TRY(cwm_gemm_f16(a16, w->w_fc1, Mx, hidden, C, &e, st));
  e = epi_res(w->b_fc2, x, C);
  CWM_TRY(cwm_gemm_f16(h16, w->w_fc2, Mx, C, hidden, &e, st));
  CWM_TRY(cwm_layernorm_f16(src, Ms, Cs, w->ln2s_g, w->ln2s_b, eps, 0, 0, 0, a16s, st));
  e = epi_gelu(w->b_fc1_s, h16s, hidden_s);
  CWM_TRY(cwm_gemm_f16(a16s, w->w_fc1_s, Ms, hidden_s, Cs, &e, st));
  e = epi_res(w->b_fc2_s, src, Cs);
  CWM_TRY(cwm_gemm_f16(h16s, w->w_fc2_s, Ms, Cs, hidden_s, &e, st));
  return CWM_OK;
}
