// pixelsrc.cuh -- pixel sources and the two pixel<->token kernels shared by the tensor path (elementwise.cu) and the
// fused counterfactual path (counterfactual.cu).
//
// A *source* answers "4 consecutive pixels of v[b, t, c, y, x0..x0+3]" (x0 % 4 == 0, inside one patch):
//   TensorSrc  a strided fp32 tensor (the reference's transposed view, cwm/models/prediction.py:304-312)
//   CfSrc      the virtual motion-counterfactual video (cwm/models/perturbation.py:245-289), never materialised
//
// Kernel shape (v2): grid.y = sample, grid.x = a 32-bit index inside the sample, so there is no 64-bit division on
// the address path (v1 spent most of its time there: 1.4 TB/s gather, 2.3 TB/s scatter).  The gather decodes the
// patch-volume index through a shared-memory table, the scatter handles all channels of its 4 pixels in one thread
// so that the prediction row y[b, rank, (kt, kh, kw, c)] is read with contiguous 16-byte loads.
#pragma once
#include "common.cuh"

namespace cwm {

struct TensorSrc {
  const float* x;
  int64_t sb, st, sc, sh, sw;
  int vec_ok;
  __device__ __forceinline__ float4 load4(long long b, int t, int c, int y, int x0) const {
    const float* src = x + b * sb + t * st + c * sc + y * sh + x0 * sw;
    float4 v;
    if (vec_ok) {
      v = __ldg(reinterpret_cast<const float4*>(src));
    } else {
      v.x = __ldg(src); v.y = __ldg(src + sw); v.z = __ldg(src + 2 * sw); v.w = __ldg(src + 3 * sw);
    }
    return v;
  }
};

// Per-(sample, frame, row, 4-pixel column) context: everything that does not depend on the channel is computed once.
struct TensorCtx {
  const float* base;
};
__device__ __forceinline__ TensorCtx src_begin(const TensorSrc& s, long long b, int t, int y, int x0) {
  TensorCtx c;
  c.base = s.x + b * s.sb + t * s.st + y * s.sh + x0 * s.sw;
  return c;
}
__device__ __forceinline__ float4 src_load4(const TensorSrc& s, const TensorCtx& k, int c) {
  const float* src = k.base + c * s.sc;
  float4 v;
  if (s.vec_ok) {
    v = __ldg(reinterpret_cast<const float4*>(src));
  } else {
    v.x = __ldg(src); v.y = __ldg(src + s.sw); v.z = __ldg(src + 2 * s.sw); v.w = __ldg(src + 3 * s.sw);
  }
  return v;
}

struct CfSrc {
  const float* x;
  int64_t sb, st, sc, sh, sw;  // logical [B_img, T, C, H, W]
  const int32_t* sample_image;
  const int32_t* shift_px;
  const uint8_t* shifted_active;
  int frame, static_frame;
  int H, W, ph, pw, n_h, n_w;
  int vec_ok;
  __device__ __forceinline__ float4 load4(long long i, int t, int c, int y, int x0) const {
    const int b = sample_image ? sample_image[i] : 0;
    const int ts = static_frame >= 0 ? static_frame : t;
    const float* img = x + b * sb + ts * st + c * sc;
    const float* src = img + y * sh + x0 * sw;
    float4 o;
    if (vec_ok) {
      o = __ldg(reinterpret_cast<const float4*>(src));
    } else {
      o.x = __ldg(src); o.y = __ldg(src + sw); o.z = __ldg(src + 2 * sw); o.w = __ldg(src + 3 * sw);
    }
    if (t != frame) return o;
    const float m = shifted_active[i * (n_h * n_w) + (y / ph) * n_w + x0 / pw] ? 1.f : 0.f;
    // Fast path for the (vast majority of) patches that keep the original frame: x_shift * 0 + x * 1 is x bit for bit
    // whenever x != 0 and x_shift is finite (x_shift * 0 = +-0); only a zero pixel needs the sign of the shifted one.
    // (Non-finite pixels are outside the contract: the reference would turn them into NaN through 0 * inf.)
    if (m != 0.f && o.x != 0.f && o.y != 0.f && o.z != 0.f && o.w != 0.f) return o;
    const int sy = shift_px[2 * i], sx = shift_px[2 * i + 1];
    const int ys = y - sy, xs = x0 - sx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // F.pad(..., value=0) (perturbation.py:258)
    if (ys >= 0 && ys < H) {
      const float* sp = img + ys * sh;
      if (vec_ok && (sx & 3) == 0) {
        if (xs >= 0 && xs < W) v = __ldg(reinterpret_cast<const float4*>(sp + xs));
      } else {
        if (xs >= 0 && xs < W) v.x = __ldg(sp + xs * sw);
        if (xs + 1 >= 0 && xs + 1 < W) v.y = __ldg(sp + (xs + 1) * sw);
        if (xs + 2 >= 0 && xs + 2 < W) v.z = __ldg(sp + (xs + 2) * sw);
        if (xs + 3 >= 0 && xs + 3 < W) v.w = __ldg(sp + (xs + 3) * sw);
      }
    }
    // x_shift * (1 - m) + x * m, literally (perturbation.py:278-282): one rounding per multiply and add
    const float om = __fsub_rn(1.f, m);
    float4 r;
    r.x = __fadd_rn(__fmul_rn(v.x, om), __fmul_rn(o.x, m));
    r.y = __fadd_rn(__fmul_rn(v.y, om), __fmul_rn(o.y, m));
    r.z = __fadd_rn(__fmul_rn(v.z, om), __fmul_rn(o.z, m));
    r.w = __fadd_rn(__fmul_rn(v.w, om), __fmul_rn(o.w, m));
    return r;
  }
};

struct CfCtx {
  const float* o;   // original pixel row position (channel 0)
  const float* v;   // shifted source position (channel 0); only read when blend != 0
  float m;
  int blend;        // 0: the pixel is the original one; 1: evaluate the blend
  int vmask;        // bit e set: shifted pixel e is inside the image
  int vvec;         // the 4 shifted pixels can be read with one aligned 16-byte load
};
__device__ __forceinline__ CfCtx src_begin(const CfSrc& s, long long i, int t, int y, int x0) {
  CfCtx k;
  const int b = s.sample_image ? s.sample_image[i] : 0;
  const int ts = s.static_frame >= 0 ? s.static_frame : t;
  const float* img = s.x + b * s.sb + ts * s.st;
  k.o = img + y * s.sh + x0 * s.sw;
  k.v = k.o;
  k.m = 1.f;
  k.blend = 0;
  k.vmask = 0;
  k.vvec = 0;
  if (t != s.frame) return k;
  k.blend = 1;
  k.m = s.shifted_active[i * (s.n_h * s.n_w) + (y / s.ph) * s.n_w + x0 / s.pw] ? 1.f : 0.f;
  const int sy = s.shift_px[2 * i], sx = s.shift_px[2 * i + 1];
  const int ys = y - sy, xs = x0 - sx;
  if (ys >= 0 && ys < s.H) {
    k.v = img + ys * s.sh + xs * s.sw;
#pragma unroll
    for (int e = 0; e < 4; ++e) k.vmask |= (xs + e >= 0 && xs + e < s.W) ? (1 << e) : 0;
    k.vvec = s.vec_ok && (sx & 3) == 0 && k.vmask == 15;
  }
  return k;
}
__device__ __forceinline__ float4 src_load4(const CfSrc& s, const CfCtx& k, int c) {
  const float* src = k.o + c * s.sc;
  float4 o;
  if (s.vec_ok) {
    o = __ldg(reinterpret_cast<const float4*>(src));
  } else {
    o.x = __ldg(src); o.y = __ldg(src + s.sw); o.z = __ldg(src + 2 * s.sw); o.w = __ldg(src + 3 * s.sw);
  }
  if (!k.blend) return o;
  // fast path (see CfSrc::load4): x_shift * 0 + x * 1 == x bit for bit unless x is a (signed) zero
  if (k.m != 0.f && o.x != 0.f && o.y != 0.f && o.z != 0.f && o.w != 0.f) return o;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // F.pad(..., value=0) (perturbation.py:258)
  const float* sp = k.v + c * s.sc;
  if (k.vvec) {
    v = __ldg(reinterpret_cast<const float4*>(sp));
  } else {
    if (k.vmask & 1) v.x = __ldg(sp);
    if (k.vmask & 2) v.y = __ldg(sp + s.sw);
    if (k.vmask & 4) v.z = __ldg(sp + 2 * s.sw);
    if (k.vmask & 8) v.w = __ldg(sp + 3 * s.sw);
  }
  const float om = __fsub_rn(1.f, k.m);
  float4 r;
  r.x = __fadd_rn(__fmul_rn(v.x, om), __fmul_rn(o.x, k.m));
  r.y = __fadd_rn(__fmul_rn(v.y, om), __fmul_rn(o.y, k.m));
  r.z = __fadd_rn(__fmul_rn(v.z, om), __fmul_rn(o.z, k.m));
  r.w = __fadd_rn(__fmul_rn(v.w, om), __fmul_rn(o.w, k.m));
  return r;
}

// ---------------------------------------------------------------------------------------------
// patch gather v2: A[m, (c,kt,kh,kw)] = norm(v[b, tt*pt+kt, c, hh*ph+kh, ww*pw+kw]) for the visible token m.
// Block = 256 threads = tpb tokens x K4 vectors (K4 = K/4 <= 256); the K4 decode lives in shared memory.
// ---------------------------------------------------------------------------------------------
struct GatherGeom {
  int C, pt, ph, pw, n_h, n_w, K4, tpb;
  const int32_t* perm;
  int Ntot, rows_per_sample, n_tokens;
  float mean[8], stdv[8];
  int normalize;
  __half* out;
};

template <class Src, int U>  // U consecutive 4-pixel vectors per thread (K4 % U == 0): U independent 16-byte loads in flight
__global__ void __launch_bounds__(256) patch_gather2_kernel(Src s, GatherGeom p) {
  __shared__ uint32_t dec[256];
  __shared__ float s_mean[8], s_stdv[8];
  if (threadIdx.x >= 248) {  // (dynamic indexing of kernel-parameter arrays would go through local memory)
    const int c = threadIdx.x - 248;
    float mu = 0.f, sd = 1.f;
    switch (c) {
      case 0: mu = p.mean[0]; sd = p.stdv[0]; break;
      case 1: mu = p.mean[1]; sd = p.stdv[1]; break;
      case 2: mu = p.mean[2]; sd = p.stdv[2]; break;
      case 3: mu = p.mean[3]; sd = p.stdv[3]; break;
      case 4: mu = p.mean[4]; sd = p.stdv[4]; break;
      case 5: mu = p.mean[5]; sd = p.stdv[5]; break;
      case 6: mu = p.mean[6]; sd = p.stdv[6]; break;
      default: mu = p.mean[7]; sd = p.stdv[7]; break;
    }
    s_mean[c] = mu;
    s_stdv[c] = sd;
  }
  if (threadIdx.x < p.K4) {
    const int k4 = threadIdx.x;
    const int pw4 = p.pw >> 2;
    const int kw = (k4 % pw4) << 2;
    int r = k4 / pw4;
    const int kh = r % p.ph;
    r /= p.ph;
    const int kt = r % p.pt;
    const int c = r / p.pt;
    dec[k4] = (c << 24) | (kt << 16) | (kh << 8) | kw;
  }
  __syncthreads();
  const int tpt = p.K4 / U;  // threads per token
  const int tl = threadIdx.x / tpt;
  const int k4 = (threadIdx.x - tl * tpt) * U;
  const int j = blockIdx.x * p.tpb + tl;
  if (tl >= p.tpb || j >= p.rows_per_sample) return;
  const long long b = blockIdx.y;
  const int tok = p.perm[b * p.Ntot + j];
  uint2* dst = reinterpret_cast<uint2*>(p.out + ((b * p.rows_per_sample + j) * p.K4 + k4) * 4);
  if (tok >= p.n_tokens) {  // padding position of a PaddedVisionTransformer (conjoined_vmae.py:130-133)
#pragma unroll
    for (int u = 0; u < U; ++u) dst[u] = make_uint2(0u, 0u);
    return;
  }
  const int n_hw = p.n_h * p.n_w;
  const int tt = tok / n_hw;
  const int rem = tok - tt * n_hw;
  const int hh = rem / p.n_w;
  const int ww = rem - hh * p.n_w;
  float4 v[U];
  int ch[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t d = dec[k4 + u];
    const int c = d >> 24, kt = (d >> 16) & 255, kh = (d >> 8) & 255, kw = d & 255;
    ch[u] = c;
    v[u] = s.load4(b, tt * p.pt + kt, c, hh * p.ph + kh, ww * p.pw + kw);
  }
  uint2 o[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (p.normalize) {
      // same operation order as imagenet_normalize: (x - mean) / std, IEEE division
      const float mu = s_mean[ch[u]], sd = s_stdv[ch[u]];
      v[u].x = __fdiv_rn(v[u].x - mu, sd);
      v[u].y = __fdiv_rn(v[u].y - mu, sd);
      v[u].z = __fdiv_rn(v[u].z - mu, sd);
      v[u].w = __fdiv_rn(v[u].w - mu, sd);
    }
    o[u].x = pack_half2(v[u].x, v[u].y);
    o[u].y = pack_half2(v[u].z, v[u].w);
  }
  if (U % 2 == 0) {  // k4 % U == 0 -> the U * 8 bytes are 16-byte aligned
#pragma unroll
    for (int u = 0; u < U; u += 2)
      reinterpret_cast<uint4*>(dst)[u >> 1] = make_uint4(o[u].x, o[u].y, o[u + 1].x, o[u + 1].y);
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u) dst[u] = o[u];
  }
}

// host-side launch shared by cwm_patch_gather and cwm_patch_gather_cf
template <class Src>
static inline void launch_patch_gather2(const Src& src, GatherGeom g, int B, cudaStream_t st) {
  if (g.K4 % 4 == 0) {
    g.tpb = 256 / (g.K4 / 4);
    dim3 grid((g.rows_per_sample + g.tpb - 1) / g.tpb, B);
    patch_gather2_kernel<Src, 4><<<grid, 256, 0, st>>>(src, g);
  } else {
    g.tpb = 256 / g.K4;
    dim3 grid((g.rows_per_sample + g.tpb - 1) / g.tpb, B);
    patch_gather2_kernel<Src, 1><<<grid, 256, 0, st>>>(src, g);
  }
}

// ---------------------------------------------------------------------------------------------
// scatter + unpatchify v2: one thread per 4 consecutive pixels of a row and ALL C channels.
//   visible patch -> C 16-byte loads from the source; masked patch -> C contiguous 16-byte loads of
//   y[b, rank, ((kt*ph+kh)*pw + kw)*C ...] transposed in registers; C 16-byte stores.
// y == nullptr with Nvis == Ntot materialises the source (cwm_cf_build_videos).
// ---------------------------------------------------------------------------------------------
struct UnpatchGeom {
  const float* y;
  const int32_t* inv_perm;
  int T, H, W, pt, ph, pw, n_h, n_w, Ntot, Nvis, D;
  int per_sample;  // T * H * W / 4
  int ph_shift, pw_shift;  // log2 of the patch sides when both are powers of two, else -1
  float* out;
};

template <class Src, int C>
__global__ void __launch_bounds__(256) unpatchify2_kernel(Src s, UnpatchGeom p) {
  // blockDim = (W/4, rows per block): no division to find the pixel; grid = (row blocks, T, samples)
  const int x4 = threadIdx.x;
  const int yy = blockIdx.x * blockDim.y + threadIdx.y;
  if (yy >= p.H) return;
  const int t = blockIdx.y;
  const long long b = blockIdx.z;
  const int xx = x4 << 2;
  const int tt = t / p.pt, kt = t - tt * p.pt;
  int hh, ww;
  if (p.ph_shift >= 0) { hh = yy >> p.ph_shift; ww = xx >> p.pw_shift; } else { hh = yy / p.ph; ww = xx / p.pw; }
  const int kh = yy - hh * p.ph, kw = xx - ww * p.pw;
  const int tok = (tt * p.n_h + hh) * p.n_w + ww;
  const int pos = p.inv_perm ? p.inv_perm[b * p.Ntot + tok] : 0;
  float4 v[C];
  if (pos < p.Nvis) {
    const auto ctx = src_begin(s, b, t, yy, xx);
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = src_load4(s, ctx, c);
  } else {
    const int Nmask = p.Ntot - p.Nvis;
    const float4* src = reinterpret_cast<const float4*>(p.y + (b * Nmask + (pos - p.Nvis)) * p.D +
                                                        ((kt * p.ph + kh) * p.pw + kw) * C);
    float f[4 * C];
#pragma unroll
    for (int q = 0; q < C; ++q) {
      const float4 g = __ldg(src + q);
      f[4 * q] = g.x; f[4 * q + 1] = g.y; f[4 * q + 2] = g.z; f[4 * q + 3] = g.w;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = make_float4(f[c], f[C + c], f[2 * C + c], f[3 * C + c]);
  }
  const long long plane = static_cast<long long>(p.H) * p.W;
  float* o = p.out + ((b * p.T + t) * C) * plane + static_cast<long long>(yy) * p.W + xx;
#pragma unroll
  for (int c = 0; c < C; ++c) *reinterpret_cast<float4*>(o + c * plane) = v[c];
}

static inline int log2_exact(int v) {
  for (int k = 0; k < 31; ++k)
    if ((1 << k) == v) return k;
  return -1;
}

// host-side launch shared by cwm_unpatchify_scatter, cwm_unpatchify_scatter_cf and cwm_cf_build_videos.
// Requires W / 4 <= 256 and B <= 65535 (checked by the callers).
template <class Src>
static inline void launch_unpatchify2(const Src& src, UnpatchGeom g, int B, cudaStream_t st) {
  g.ph_shift = log2_exact(g.ph);
  g.pw_shift = log2_exact(g.pw);
  if (g.ph_shift < 0 || g.pw_shift < 0) g.ph_shift = g.pw_shift = -1;
  const int W4 = g.W / 4;
  const int rows = 256 / W4 > 0 ? 256 / W4 : 1;
  dim3 block(W4, rows);
  dim3 grid((g.H + rows - 1) / rows, g.T, B);
  unpatchify2_kernel<Src, 3><<<grid, block, 0, st>>>(src, g);
}

}  // namespace cwm
