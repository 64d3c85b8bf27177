// counterfactual.cu -- SURVEY.md section 8(f) rank 1: batched motion-counterfactual construction.
//
// The reference builds the S counterfactual prompts of a sweep one by one in Python
// (cwm/models/segmentation.py:321-338: pad, centre-crop, patchify, blend, unpatchify per sample).  Here the masks
// of all samples come from one launch, and the videos are either materialised by one HBM-write-bound kernel or --
// on the fused path -- never materialised: the patch gather and the final unpatchify read a *virtual* video
// described by (source image, pixel shift, shifted perturbation mask).
//
// Pixels are bit-exact with the reference: the blend `x_shift * (1 - m) + x * m` (perturbation.py:278-282) and
// `(1 - m) * x0 + m * x` (perturbation.py:146) are evaluated literally with one rounding per multiply and add.
#include "common.cuh"
#include "pixelsrc.cuh"

namespace cwm {

// ---------------------------------------------------------------------------------------------
// masks: one thread per token.  Algorithmic bytes per token: 2 read (+1 for the shifted source) + 1..2 written.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cf_shift_masks_kernel(const uint8_t* __restrict__ passive, const uint8_t* __restrict__ active,
                      const int32_t* __restrict__ mask_shift, int T, int n_h, int n_w, int frame, long long total,
                      uint8_t* __restrict__ shifted_active, uint8_t* __restrict__ mask_out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n_hw = n_h * n_w;
  const int N = T * n_hw;
  const int tok = static_cast<int>(i % N);
  const long long s = i / N;
  const int t = tok / n_hw;
  const int rem = tok - t * n_hw;
  const bool p = passive[i] != 0, a = active[i] != 0;
  // mask = minimum(masks, active); mask[~active] = 1   (segmentation.py:327, perturbation.py:106)
  const bool m0 = p || !a;
  bool pert = a;  // perturbation_mask = logical_not(perturbation_points) = active (perturbation.py:107)
  if (t == frame) {
    const int yy = rem / n_w, xx = rem - yy * n_w;
    const int ys = yy - mask_shift[2 * s], xs = xx - mask_shift[2 * s + 1];
    // CenterCrop(F.pad(m, mask_padding, value=1)) (perturbation.py:268-269): the source patch, 1 outside
    pert = (ys >= 0 && ys < n_h && xs >= 0 && xs < n_w) ? (active[s * N + frame * n_hw + ys * n_w + xs] != 0) : true;
    shifted_active[s * n_hw + rem] = pert ? 1 : 0;
  }
  mask_out[i] = (m0 && pert) ? 1 : 0;  // minimum(mask, mask_perturbed) (perturbation.py:109-110)
}

// ---------------------------------------------------------------------------------------------
// the virtual counterfactual video
// ---------------------------------------------------------------------------------------------
// CfSrc (the virtual video) lives in pixelsrc.cuh.
__device__ __forceinline__ float4 cf_load4(const CfSrc& s, long long i, int t, int c, int y, int x0) {
  return s.load4(i, t, c, y, x0);
}

// materialise: one thread per 4 output pixels, 16-byte coalesced stores.
__global__ void __launch_bounds__(256) cf_build_videos_kernel(CfSrc s, int T, int C, long long total, float4* out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int W4 = s.W >> 2;
  const int x4 = static_cast<int>(i % W4);
  long long r = i / W4;
  const int y = static_cast<int>(r % s.H);
  r /= s.H;
  const int c = static_cast<int>(r % C);
  r /= C;
  const int t = static_cast<int>(r % T);
  const long long smp = r / T;
  out[i] = cf_load4(s, smp, t, c, y, x4 << 2);
}

// MakeStatic: out[b,t] = (1 - m) * x[b,0] + m * x[b,t]
__global__ void __launch_bounds__(256)
cf_make_static_kernel(const float* __restrict__ x, int64_t sb, int64_t st, int64_t sc, int64_t sh, int64_t sw,
                      const uint8_t* __restrict__ mask, int T, int C, int H, int W, int ph, int pw, int vec_ok,
                      long long total, float4* out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int W4 = W >> 2;
  const int x4 = static_cast<int>(i % W4);
  long long r = i / W4;
  const int y = static_cast<int>(r % H);
  r /= H;
  const int c = static_cast<int>(r % C);
  r /= C;
  const int t = static_cast<int>(r % T);
  const long long b = r / T;
  const int n_h = H / ph, n_w = W / pw;
  const int x0 = x4 << 2;
  const float m = mask[(b * T + t) * (n_h * n_w) + (y / ph) * n_w + x0 / pw] ? 1.f : 0.f;
  const float* p0 = x + b * sb + c * sc + y * sh + x0 * sw;
  const float* pt_ = p0 + t * st;
  float4 a, d;
  if (vec_ok) {
    a = __ldg(reinterpret_cast<const float4*>(p0));
    d = __ldg(reinterpret_cast<const float4*>(pt_));
  } else {
    a.x = __ldg(p0); a.y = __ldg(p0 + sw); a.z = __ldg(p0 + 2 * sw); a.w = __ldg(p0 + 3 * sw);
    d.x = __ldg(pt_); d.y = __ldg(pt_ + sw); d.z = __ldg(pt_ + 2 * sw); d.w = __ldg(pt_ + 3 * sw);
  }
  const float om = __fsub_rn(1.f, m);
  float4 o;
  o.x = __fadd_rn(__fmul_rn(om, a.x), __fmul_rn(m, d.x));
  o.y = __fadd_rn(__fmul_rn(om, a.y), __fmul_rn(m, d.y));
  o.z = __fadd_rn(__fmul_rn(om, a.z), __fmul_rn(m, d.z));
  o.w = __fadd_rn(__fmul_rn(om, a.w), __fmul_rn(m, d.w));
  out[i] = o;
}

// fused patch gather (same thread mapping and output as patch_gather_kernel, elementwise.cu)
struct CfGatherParams {
  CfSrc s;
  int C, pt, K4;
  const int32_t* perm;
  int Ntot, rows_per_sample, n_tokens;
  float mean[8], stdv[8];
  int normalize;
  __half* out;
  long long total;
};

__global__ void __launch_bounds__(256) cf_patch_gather_kernel(CfGatherParams p) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  const int k4 = static_cast<int>(i % p.K4);
  const long long m = i / p.K4;
  const int j = static_cast<int>(m % p.rows_per_sample);
  const long long smp = m / p.rows_per_sample;
  const int tok = p.perm[smp * p.Ntot + j];
  uint2* dst = reinterpret_cast<uint2*>(p.out + m * (static_cast<long long>(p.K4) * 4) + k4 * 4);
  if (tok >= p.n_tokens) {
    *dst = make_uint2(0u, 0u);
    return;
  }
  const int n_hw = p.s.n_h * p.s.n_w;
  const int tt = tok / n_hw;
  const int rem = tok - tt * n_hw;
  const int hh = rem / p.s.n_w;
  const int ww = rem - hh * p.s.n_w;
  const int pw4 = p.s.pw >> 2;
  const int kw = (k4 % pw4) << 2;
  int r = k4 / pw4;
  const int kh = r % p.s.ph;
  r /= p.s.ph;
  const int kt = r % p.pt;
  const int c = r / p.pt;
  float4 v = cf_load4(p.s, smp, tt * p.pt + kt, c, hh * p.s.ph + kh, ww * p.s.pw + kw);
  if (p.normalize) {
    const float mu = p.mean[c], sd = p.stdv[c];
    v.x = __fdiv_rn(v.x - mu, sd);
    v.y = __fdiv_rn(v.y - mu, sd);
    v.z = __fdiv_rn(v.z - mu, sd);
    v.w = __fdiv_rn(v.w - mu, sd);
  }
  uint2 o;
  o.x = pack_half2(v.x, v.y);
  o.y = pack_half2(v.z, v.w);
  *dst = o;
}

// fused scatter + unpatchify (same mapping as unpatchify_scatter_kernel, elementwise.cu)
struct CfUnpatchParams {
  CfSrc s;
  const float* y;
  const int32_t* inv_perm;
  int T, C, pt, Ntot, Nvis, D;
  long long total;
  float4* out;
};

__global__ void __launch_bounds__(256) cf_unpatchify_scatter_kernel(CfUnpatchParams p) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  const int W4 = p.s.W >> 2;
  const int x4 = static_cast<int>(i % W4);
  long long r = i / W4;
  const int yy = static_cast<int>(r % p.s.H);
  r /= p.s.H;
  const int c = static_cast<int>(r % p.C);
  r /= p.C;
  const int t = static_cast<int>(r % p.T);
  const long long b = r / p.T;
  const int xx = x4 << 2;
  const int tt = t / p.pt, kt = t - tt * p.pt;
  const int hh = yy / p.s.ph, kh = yy - hh * p.s.ph;
  const int ww = xx / p.s.pw, kw = xx - ww * p.s.pw;
  const int tok = (tt * p.s.n_h + hh) * p.s.n_w + ww;
  const int pos = p.inv_perm[b * p.Ntot + tok];
  float4 v;
  if (pos < p.Nvis) {
    v = cf_load4(p.s, b, t, c, yy, xx);
  } else {
    const int Nmask = p.Ntot - p.Nvis;
    const float* src = p.y + (b * Nmask + (pos - p.Nvis)) * p.D + ((kt * p.s.ph + kh) * p.s.pw + kw) * p.C + c;
    v.x = __ldg(src);
    v.y = __ldg(src + p.C);
    v.z = __ldg(src + 2 * p.C);
    v.w = __ldg(src + 3 * p.C);
  }
  p.out[i] = v;
}

static int make_src(const cwm_cf_source* src, int T, int H, int W, int ph, int pw, CfSrc* s, const char* who) {
  if (!src || !src->x || !src->shift_px || !src->shifted_active)
    return fail(CWM_ERR_INVALID, "%s: null pointer in cwm_cf_source", who);
  if (ph <= 0 || pw <= 0 || H % ph || W % pw)
    return fail(CWM_ERR_INVALID, "%s: Input image size(%d,%d) must be divisible by patch size (%d,%d)", who, H, W, ph, pw);
  if (pw % 4 || W % 4) return fail(CWM_ERR_UNSUPPORTED, "%s: patch width %d must be a multiple of 4", who, pw);
  if (src->frame < 0 || src->frame >= T || src->static_frame >= T)
    return fail(CWM_ERR_INVALID, "%s: frame %d / static_frame %d out of range (T = %d)", who, src->frame, src->static_frame, T);
  s->x = src->x;
  s->sb = src->xs[0]; s->st = src->xs[1]; s->sc = src->xs[2]; s->sh = src->xs[3]; s->sw = src->xs[4];
  s->sample_image = src->sample_image;
  s->shift_px = src->shift_px;
  s->shifted_active = src->shifted_active;
  s->frame = src->frame;
  s->static_frame = src->static_frame;
  s->H = H; s->W = W; s->ph = ph; s->pw = pw; s->n_h = H / ph; s->n_w = W / pw;
  s->vec_ok = (s->sw == 1) && (reinterpret_cast<uintptr_t>(s->x) % 16 == 0) && (s->sb % 4 == 0) && (s->sc % 4 == 0) &&
              (s->st % 4 == 0) && (s->sh % 4 == 0);
  return CWM_OK;
}

}  // namespace cwm

using namespace cwm;

extern "C" int cwm_cf_shift_masks(const uint8_t* passive, const uint8_t* active, const int32_t* mask_shift, int S,
                                  int T, int n_h, int n_w, int frame, uint8_t* shifted_active, uint8_t* mask_out,
                                  cwm_stream_t stream) {
  CWM_REQUIRE(passive && active && mask_shift && shifted_active && mask_out, "cwm_cf_shift_masks: null pointer");
  CWM_REQUIRE(S >= 0 && T > 0 && n_h > 0 && n_w > 0, "cwm_cf_shift_masks: bad shape S=%d T=%d h=%d w=%d", S, T, n_h, n_w);
  CWM_REQUIRE(frame >= 0 && frame < T, "cwm_cf_shift_masks: frame %d out of range (T = %d)", frame, T);
  const long long total = static_cast<long long>(S) * T * n_h * n_w;
  if (total == 0) return CWM_OK;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "cf_shift_masks", 0.0, static_cast<double>(total) * 3.5);
  cf_shift_masks_kernel<<<static_cast<unsigned>(blocks), threads, 0, st>>>(passive, active, mask_shift, T, n_h, n_w, frame,
                                                                          total, shifted_active, mask_out);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_cf_build_videos(const cwm_cf_source* src, int S, int T, int C, int H, int W, int ph, int pw,
                                   float* out, cwm_stream_t stream) {
  CfSrc s;
  int rc = make_src(src, T, H, W, ph, pw, &s, "cwm_cf_build_videos");
  if (rc != CWM_OK) return rc;
  CWM_REQUIRE(out && S >= 0 && C > 0, "cwm_cf_build_videos: bad arguments");
  CWM_REQUIRE(reinterpret_cast<uintptr_t>(out) % 16 == 0, "cwm_cf_build_videos: out must be 16-byte aligned");
  const long long total = static_cast<long long>(S) * T * C * H * (W / 4);
  if (total == 0) return CWM_OK;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "cf_build_videos", 0.0, static_cast<double>(total) * 16.0);
  if (C == 3 && S <= 65535 && W / 4 <= 256) {
    UnpatchGeom g;
    g.y = nullptr; g.inv_perm = nullptr; g.T = T; g.H = H; g.W = W; g.pt = 1; g.ph = ph; g.pw = pw; g.n_h = s.n_h;
    g.n_w = s.n_w; g.Ntot = T * s.n_h * s.n_w; g.Nvis = g.Ntot; g.D = ph * pw * C; g.per_sample = T * H * (W / 4);
    g.out = out;
    launch_unpatchify2(s, g, S, st);
    CWM_LAUNCH_CHECK();
    return CWM_OK;
  }
  cf_build_videos_kernel<<<static_cast<unsigned>(blocks), threads, 0, st>>>(s, T, C, total, reinterpret_cast<float4*>(out));
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_cf_make_static(const float* x, const int64_t xs[5], const uint8_t* mask, int B, int T, int C, int H,
                                  int W, int ph, int pw, float* out, cwm_stream_t stream) {
  CWM_REQUIRE(x && xs && mask && out, "cwm_cf_make_static: null pointer");
  CWM_REQUIRE(ph > 0 && pw > 0 && H % ph == 0 && W % pw == 0,
              "cwm_cf_make_static: Input image size(%d,%d) must be divisible by patch size (%d,%d)", H, W, ph, pw);
  CWM_REQUIRE(pw % 4 == 0 && W % 4 == 0, "cwm_cf_make_static: patch width %d must be a multiple of 4", pw);
  CWM_REQUIRE(T > 1, "cwm_cf_make_static: needs T > 1 (perturbation.py:126-127)");
  const long long total = static_cast<long long>(B) * T * C * H * (W / 4);
  if (total == 0) return CWM_OK;
  const int vec_ok = (xs[4] == 1) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (xs[0] % 4 == 0) && (xs[1] % 4 == 0) &&
                     (xs[2] % 4 == 0) && (xs[3] % 4 == 0);
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "cf_make_static", 0.0, static_cast<double>(total) * 16.0 * 2.5);
  // x logical [B, T, C, H, W]
  cf_make_static_kernel<<<static_cast<unsigned>(blocks), threads, 0, st>>>(x, xs[0], xs[1], xs[2], xs[3], xs[4], mask, T, C,
                                                                          H, W, ph, pw, vec_ok, total,
                                                                          reinterpret_cast<float4*>(out));
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_patch_gather_cf(const cwm_cf_source* src, int S, int C, int T, int H, int W, int pt, int ph, int pw,
                                   const int32_t* perm, int Ntot, int rows_per_sample, const float* mean,
                                   const float* stdv, uint16_t* out, cwm_stream_t stream) {
  CfGatherParams p;
  int rc = make_src(src, T, H, W, ph, pw, &p.s, "cwm_patch_gather_cf");
  if (rc != CWM_OK) return rc;
  CWM_REQUIRE(perm && out, "cwm_patch_gather_cf: null pointer");
  CWM_REQUIRE(pt == 1, "cwm_patch_gather_cf: motion counterfactuals need a temporal patch size of 1 (got %d)", pt);
  CWM_REQUIRE(C <= 8, "cwm_patch_gather_cf: at most 8 input channels (got %d)", C);
  CWM_REQUIRE((mean == nullptr) == (stdv == nullptr), "cwm_patch_gather_cf: mean/std must both be set or both NULL");
  if (S == 0 || rows_per_sample == 0) return CWM_OK;
  const int K = C * pt * ph * pw;
  p.C = C; p.pt = pt; p.K4 = K / 4;
  p.perm = perm; p.Ntot = Ntot; p.rows_per_sample = rows_per_sample;
  p.n_tokens = (T / pt) * p.s.n_h * p.s.n_w;
  p.normalize = mean != nullptr;
  for (int c = 0; c < 8; ++c) { p.mean[c] = 0.f; p.stdv[c] = 1.f; }
  if (p.normalize)
    for (int c = 0; c < C; ++c) { p.mean[c] = mean[c]; p.stdv[c] = stdv[c]; }
  p.out = reinterpret_cast<__half*>(out);
  p.total = static_cast<long long>(S) * rows_per_sample * p.K4;
  const int threads = 256;
  const long long blocks = (p.total + threads - 1) / threads;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "patch_gather", 0.0, static_cast<double>(S) * rows_per_sample * K * 6.0);
  if (p.K4 <= 256 && S <= 65535 && ph < 256 && pw < 256) {
    GatherGeom g;
    g.C = C; g.pt = pt; g.ph = ph; g.pw = pw; g.n_h = p.s.n_h; g.n_w = p.s.n_w; g.K4 = p.K4; g.tpb = 256 / p.K4;
    g.perm = perm; g.Ntot = Ntot; g.rows_per_sample = rows_per_sample; g.n_tokens = p.n_tokens;
    for (int c = 0; c < 8; ++c) { g.mean[c] = p.mean[c]; g.stdv[c] = p.stdv[c]; }
    g.normalize = p.normalize; g.out = p.out;
    launch_patch_gather2(p.s, g, S, st);
    CWM_LAUNCH_CHECK();
    return CWM_OK;
  }
  cf_patch_gather_kernel<<<static_cast<unsigned>(blocks), threads, 0, st>>>(p);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_unpatchify_scatter_cf(const float* y, const cwm_cf_source* src, const int32_t* inv_perm, int S, int T,
                                         int C, int H, int W, int pt, int ph, int pw, int Nvis, float* out,
                                         cwm_stream_t stream) {
  CfUnpatchParams p;
  int rc = make_src(src, T, H, W, ph, pw, &p.s, "cwm_unpatchify_scatter_cf");
  if (rc != CWM_OK) return rc;
  CWM_REQUIRE(inv_perm && out, "cwm_unpatchify_scatter_cf: null pointer");
  CWM_REQUIRE(pt == 1, "cwm_unpatchify_scatter_cf: motion counterfactuals need a temporal patch size of 1 (got %d)", pt);
  p.y = y; p.inv_perm = inv_perm;
  p.T = T; p.C = C; p.pt = pt;
  p.Ntot = (T / pt) * p.s.n_h * p.s.n_w;
  p.Nvis = Nvis;
  p.D = pt * ph * pw * C;
  CWM_REQUIRE(Nvis >= 0 && Nvis <= p.Ntot, "cwm_unpatchify_scatter_cf: Nvis=%d out of range", Nvis);
  CWM_REQUIRE(y != nullptr || Nvis == p.Ntot, "cwm_unpatchify_scatter_cf: y is NULL but there are masked tokens");
  p.total = static_cast<long long>(S) * T * C * H * (W / 4);
  p.out = reinterpret_cast<float4*>(out);
  if (p.total == 0) return CWM_OK;
  const int threads = 256;
  const long long blocks = (p.total + threads - 1) / threads;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "unpatchify_scatter", 0.0, static_cast<double>(p.total) * 4 * 8.0);
  if (C == 3 && S <= 65535 && W / 4 <= 256 && (y == nullptr || (reinterpret_cast<uintptr_t>(y) % 16 == 0 && p.D % 4 == 0)) &&
      reinterpret_cast<uintptr_t>(out) % 16 == 0) {
    UnpatchGeom g;
    g.y = y; g.inv_perm = inv_perm; g.T = T; g.H = H; g.W = W; g.pt = pt; g.ph = ph; g.pw = pw; g.n_h = p.s.n_h;
    g.n_w = p.s.n_w; g.Ntot = p.Ntot; g.Nvis = Nvis; g.D = p.D; g.per_sample = T * H * (W / 4); g.out = out;
    launch_unpatchify2(p.s, g, S, st);
    CWM_LAUNCH_CHECK();
    return CWM_OK;
  }
  cf_unpatchify_scatter_kernel<<<static_cast<unsigned>(blocks), threads, 0, st>>>(p);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}
