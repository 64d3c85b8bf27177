// elementwise.cu -- the HBM-bound kernels of the path: patch gather (a1+a2 input side), LayerNorm (a5/a8/a11),
// mask-token fill (a10) and scatter+unpatchify (a12).  All are coalesced, 128-bit vectorised where alignment
// allows, and sized so the grid is many waves over 148 SMs.
#include "common.cuh"
#include "pixelsrc.cuh"

namespace cwm {

// ---------------------------------------------------------------------------------------------
// patch gather: A[m, (c,kt,kh,kw)] = norm(x[b, c, tt*pt+kt, hh*ph+kh, ww*pw+kw]) for visible token m.
// One thread per 4 consecutive kw (pw % 4 == 0): one 16-byte load, one 8-byte store; consecutive threads
// write consecutive addresses of A.
// Algorithmic bytes per visible token: K*4 read + K*2 written.
// ---------------------------------------------------------------------------------------------
struct GatherParams {
  const float* x;
  int64_t sb, sc, st, sh, sw;
  int C, pt, ph, pw, n_h, n_w;
  int K4;  // K / 4
  const int32_t* perm;
  int Ntot, rows_per_sample;
  int n_tokens;  // real tokens (T/pt * n_h * n_w); ids >= n_tokens are padding positions -> zero rows
  float mean[8], stdv[8];
  int normalize;
  __half* out;
  long long total;  // B * rows_per_sample * K4
  int vec_ok;       // sw == 1 and 16-byte aligned rows
};

__global__ void __launch_bounds__(256) patch_gather_kernel(GatherParams p) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  const int k4 = static_cast<int>(i % p.K4);
  const long long m = i / p.K4;
  const int j = static_cast<int>(m % p.rows_per_sample);
  const int b = static_cast<int>(m / p.rows_per_sample);
  const int tok = p.perm[static_cast<long long>(b) * p.Ntot + j];
  if (tok >= p.n_tokens) {  // padding position of a PaddedVisionTransformer (conjoined_vmae.py:130-133)
    *reinterpret_cast<uint2*>(p.out + m * (static_cast<long long>(p.K4) * 4) + k4 * 4) = make_uint2(0u, 0u);
    return;
  }
  const int n_hw = p.n_h * p.n_w;
  const int tt = tok / n_hw;
  const int rem = tok - tt * n_hw;
  const int hh = rem / p.n_w;
  const int ww = rem - hh * p.n_w;
  // k = ((c*pt + kt)*ph + kh)*pw + kw, kw = 4*(k4 % (pw/4))
  const int pw4 = p.pw >> 2;
  const int kw = (k4 % pw4) << 2;
  int r = k4 / pw4;
  const int kh = r % p.ph;
  r /= p.ph;
  const int kt = r % p.pt;
  const int c = r / p.pt;
  const float* src = p.x + b * p.sb + c * p.sc + (tt * p.pt + kt) * p.st + (hh * p.ph + kh) * p.sh +
                     (ww * p.pw + kw) * p.sw;
  float4 v;
  if (p.vec_ok) {
    v = __ldg(reinterpret_cast<const float4*>(src));
  } else {
    v.x = __ldg(src);
    v.y = __ldg(src + p.sw);
    v.z = __ldg(src + 2 * p.sw);
    v.w = __ldg(src + 3 * p.sw);
  }
  if (p.normalize) {
    // same operation order as imagenet_normalize: (x - mean) / std, IEEE division
    const float mu = p.mean[c], sd = p.stdv[c];
    v.x = __fdiv_rn(v.x - mu, sd);
    v.y = __fdiv_rn(v.y - mu, sd);
    v.z = __fdiv_rn(v.z - mu, sd);
    v.w = __fdiv_rn(v.w - mu, sd);
  }
  __half2 lo = __floats2half2_rn(v.x, v.y);
  __half2 hi = __floats2half2_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&lo);
  o.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p.out + m * (static_cast<long long>(p.K4) * 4) + k4 * 4) = o;
}


// scalar variant for patch widths that are not a multiple of 4 (the IMU "video" [B, 6, 400, 1, 1] with a
// (16, 1, 1) tubelet, conjoined_vmae.py:1013-1038): one thread per element; K4 holds K here.
__global__ void __launch_bounds__(256) patch_gather_scalar_kernel(GatherParams p) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  const int K = p.K4;
  const int k = static_cast<int>(i % K);
  const long long m = i / K;
  const int j = static_cast<int>(m % p.rows_per_sample);
  const int b = static_cast<int>(m / p.rows_per_sample);
  const int tok = p.perm[static_cast<long long>(b) * p.Ntot + j];
  float v = 0.f;
  if (tok < p.n_tokens) {
    const int n_hw = p.n_h * p.n_w;
    const int tt = tok / n_hw;
    const int rem = tok - tt * n_hw;
    const int hh = rem / p.n_w;
    const int ww = rem - hh * p.n_w;
    const int kw = k % p.pw;
    int r = k / p.pw;
    const int kh = r % p.ph;
    r /= p.ph;
    const int kt = r % p.pt;
    const int c = r / p.pt;
    v = __ldg(p.x + b * p.sb + c * p.sc + (tt * p.pt + kt) * p.st + (hh * p.ph + kh) * p.sh + (ww * p.pw + kw) * p.sw);
    if (p.normalize) v = __fdiv_rn(v - p.mean[c], p.stdv[c]);
  }
  p.out[m * K + k] = __float2half_rn(v);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (C/128 float4 per lane), two-pass statistics in
// fp32 (mean, then centred sum of squares) like ATen's CPU/CUDA LayerNorm; output f16.
// Algorithmic bytes per row: 4*C read + 2*C written.
// ---------------------------------------------------------------------------------------------
template <int VEC>  // VEC = C / 128 float4 per lane
__global__ void __launch_bounds__(256)
layernorm_f16_kernel(const float* __restrict__ x, int M, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, int grp_rows, int grp_stride, int grp_offset,
                     __half* __restrict__ out) {
  constexpr int C = VEC * 128;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  long long in_row = warp;
  if (grp_rows > 0) in_row = static_cast<long long>(warp / grp_rows) * grp_stride + grp_offset + warp % grp_rows;
  const float4* xr = reinterpret_cast<const float4*>(x + in_row * C);
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.0f / C) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* orow = reinterpret_cast<uint2*>(out + static_cast<long long>(warp) * C);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i);
    const float4 bb = __ldg(b4 + lane + 32 * i);
    const float y0 = (v[i].x - mean) * rstd * g.x + bb.x;
    const float y1 = (v[i].y - mean) * rstd * g.y + bb.y;
    const float y2 = (v[i].z - mean) * rstd * g.z + bb.z;
    const float y3 = (v[i].w - mean) * rstd * g.w + bb.w;
    uint2 o;
    o.x = pack_half2(y0, y1);
    o.y = pack_half2(y2, y3);
    orow[lane + 32 * i] = o;
  }
}


// First LayerNorm of a stream when the LayerNorm is fused into the consumer GEMM (gemm.cu): f16 copy of the rows and
// per-row (sum, sum of squares) in the same format the residual-GEMM epilogues emit (one partial plane).
__global__ void __launch_bounds__(256)
rowstats_f16_kernel(const float* __restrict__ x, int M, int C4, __half* __restrict__ x16, float2* __restrict__ stats) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x) + static_cast<long long>(warp) * C4;
  uint2* orow = reinterpret_cast<uint2*>(x16) + static_cast<long long>(warp) * C4;
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < C4; i += 32) {
    const float4 v = xr[i];
    s1 += (v.x + v.y) + (v.z + v.w);
    s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    uint2 o;
    o.x = pack_half2(v.x, v.y);
    o.y = pack_half2(v.z, v.w);
    orow[i] = o;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) stats[warp] = make_float2(s1, s2);
}

// Same algorithm for widths that are not a multiple of 128 (context-stream widths 64 / 192, conjoined_vmae.py:
// 1198-1204): C % 4 == 0, C <= 1024; lanes beyond the row end hold zeros and are excluded from the statistics.
__global__ void __launch_bounds__(256)
layernorm_f16_generic_kernel(const float* __restrict__ x, int M, int C, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float eps, int grp_rows, int grp_stride, int grp_offset,
                             __half* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  long long in_row = warp;
  if (grp_rows > 0) in_row = static_cast<long long>(warp / grp_rows) * grp_stride + grp_offset + warp % grp_rows;
  const int C4 = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + in_row * C);
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c4 = lane + 32 * i;
    v[i] = (c4 < C4) ? xr[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (lane + 32 * i < C4) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / static_cast<float>(C) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* orow = reinterpret_cast<uint2*>(out + static_cast<long long>(warp) * C);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c4 = lane + 32 * i;
    if (c4 < C4) {
      const float4 g = __ldg(g4 + c4);
      const float4 bb = __ldg(b4 + c4);
      uint2 o;
      o.x = pack_half2((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
      o.y = pack_half2((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
      orow[c4] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// padding rows: x[b, j, :] = value (or 0) where perm[b, perm_offset + j] >= first_pad_token
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fill_pad_rows_kernel(float4* __restrict__ x, int rows, int C4, const int32_t* __restrict__ perm, int perm_stride,
                     int perm_offset, int first_pad_token, const float4* __restrict__ value, long long total) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = static_cast<int>(i % C4);
  const long long r = i / C4;
  const int j = static_cast<int>(r % rows);
  const long long b = r / rows;
  if (perm[b * perm_stride + perm_offset + j] < first_pad_token) return;
  x[i] = value != nullptr ? __ldg(value + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------
// mask-token rows of the decoder input: x_full[b, Nvis + j, :] = mask_token + pos[perm[b, Nvis + j], :]
// Algorithmic bytes per masked token: 4*C written (pos table and mask token are L2 resident).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fill_mask_tokens_kernel(const float4* __restrict__ mask_token, const float4* __restrict__ pos,
                        const int32_t* __restrict__ perm, int Ntot, int Nvis, int C4, long long total,
                        float4* __restrict__ x_full) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = static_cast<int>(i % C4);
  const long long r = i / C4;
  const int Nmask = Ntot - Nvis;
  const int j = static_cast<int>(r % Nmask);
  const long long b = r / Nmask;
  const int tok = perm[b * Ntot + Nvis + j];
  const float4 mt = __ldg(mask_token + c4);
  const float4 pe = __ldg(pos + static_cast<long long>(tok) * C4 + c4);
  float4 o;
  o.x = mt.x + pe.x;
  o.y = mt.y + pe.y;
  o.z = mt.z + pe.z;
  o.w = mt.w + pe.w;
  x_full[(b * Ntot + Nvis + j) * C4 + c4] = o;
}

// v2: grid.y = sample, 32-bit index inside the sample (no 64-bit division on the address path); U float4 per thread
template <int U>
__global__ void __launch_bounds__(256)
fill_mask_tokens2_kernel(const float4* __restrict__ mask_token, const float4* __restrict__ pos,
                         const int32_t* __restrict__ perm, int Ntot, int Nvis, int C4, int per_sample,
                         float4* __restrict__ x_full) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * U;  // C4 % U == 0: the U vectors belong to one token
  if (i >= per_sample) return;
  const long long b = blockIdx.y;
  const int j = i / C4;
  const int c4 = i - j * C4;
  const int tok = perm[b * Ntot + Nvis + j];
  const float4* pe = pos + static_cast<long long>(tok) * C4 + c4;
  float4 a[U], m[U];
#pragma unroll
  for (int u = 0; u < U; ++u) { a[u] = __ldg(pe + u); m[u] = __ldg(mask_token + c4 + u); }
  float4* dst = x_full + (b * Ntot + Nvis + j) * C4 + c4;
#pragma unroll
  for (int u = 0; u < U; ++u) dst[u] = make_float4(m[u].x + a[u].x, m[u].y + a[u].y, m[u].z + a[u].z, m[u].w + a[u].w);
}

// ---------------------------------------------------------------------------------------------
// scatter + unpatchify: one thread per 4 consecutive output pixels of a row (pw % 4 == 0 so the 4 pixels
// belong to one patch).  Visible patches are copied from the raw input (bit-exact), masked patches come from
// the prediction y[b, rank, ((kt*ph + kh)*pw + kw)*C + c]   (patches.py:72-74 layout "(pt ph pw) c").
// Algorithmic bytes per sample: T*C*H*W*4 written + the same amount read (x_raw or y).
// ---------------------------------------------------------------------------------------------
struct UnpatchParams {
  const float* y;
  const float* x;
  int64_t sb, sc, st, sh, sw;
  const int32_t* inv_perm;
  int T, C, H, W, pt, ph, pw, n_h, n_w, Ntot, Nvis, D;
  long long total;  // B*T*C*H*W/4
  float* out;
  int vec_ok;
};

__global__ void __launch_bounds__(256) unpatchify_scatter_kernel(UnpatchParams p) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  const int W4 = p.W >> 2;
  const int x4 = static_cast<int>(i % W4);
  long long r = i / W4;
  const int yy = static_cast<int>(r % p.H);
  r /= p.H;
  const int c = static_cast<int>(r % p.C);
  r /= p.C;
  const int t = static_cast<int>(r % p.T);
  const long long b = r / p.T;
  const int xx = x4 << 2;
  const int tt = t / p.pt, kt = t - tt * p.pt;
  const int hh = yy / p.ph, kh = yy - hh * p.ph;
  const int ww = xx / p.pw, kw = xx - ww * p.pw;
  const int tok = (tt * p.n_h + hh) * p.n_w + ww;
  const int pos = p.inv_perm[b * p.Ntot + tok];
  float4 v;
  if (pos < p.Nvis) {
    const float* src = p.x + b * p.sb + t * p.st + c * p.sc + yy * p.sh + xx * p.sw;
    if (p.vec_ok) {
      v = __ldg(reinterpret_cast<const float4*>(src));
    } else {
      v.x = __ldg(src);
      v.y = __ldg(src + p.sw);
      v.z = __ldg(src + 2 * p.sw);
      v.w = __ldg(src + 3 * p.sw);
    }
  } else {
    const int Nmask = p.Ntot - p.Nvis;
    const float* src = p.y + (b * Nmask + (pos - p.Nvis)) * p.D + ((kt * p.ph + kh) * p.pw + kw) * p.C + c;
    v.x = __ldg(src);
    v.y = __ldg(src + p.C);
    v.z = __ldg(src + 2 * p.C);
    v.w = __ldg(src + 3 * p.C);
  }
  reinterpret_cast<float4*>(p.out)[i] = v;
}

}  // namespace cwm

using namespace cwm;

extern "C" int cwm_patch_gather(const float* x, const int64_t xs[5], int B, int C, int T, int H, int W, int pt,
                                int ph, int pw, const int32_t* perm, int Ntot, int rows_per_sample,
                                const float* mean, const float* stdv, uint16_t* out, cwm_stream_t stream) {
  CWM_REQUIRE(x && xs && perm && out, "cwm_patch_gather: null pointer");
  CWM_REQUIRE(pt > 0 && ph > 0 && pw > 0 && T % pt == 0 && H % ph == 0 && W % pw == 0,
              "Input image size(%d,%d) must be divisible by patch size (%d,%d)", H, W, ph, pw);
  CWM_REQUIRE(C <= 8, "cwm_patch_gather: at most 8 input channels (got %d)", C);
  CWM_REQUIRE((mean == nullptr) == (stdv == nullptr), "cwm_patch_gather: mean/std must both be set or both NULL");
  if (B == 0 || rows_per_sample == 0) return CWM_OK;
  GatherParams p;
  p.x = x;
  p.sb = xs[0]; p.sc = xs[1]; p.st = xs[2]; p.sh = xs[3]; p.sw = xs[4];
  p.C = C; p.pt = pt; p.ph = ph; p.pw = pw; p.n_h = H / ph; p.n_w = W / pw;
  const int K = C * pt * ph * pw;
  const bool vec4 = (pw % 4 == 0);
  p.K4 = vec4 ? K / 4 : K;
  p.perm = perm; p.Ntot = Ntot; p.rows_per_sample = rows_per_sample;
  p.n_tokens = (T / pt) * p.n_h * p.n_w;
  p.normalize = mean != nullptr;
  for (int c = 0; c < 8; ++c) { p.mean[c] = 0.f; p.stdv[c] = 1.f; }
  if (p.normalize) {
    // mean/std are HOST pointers (3 floats): they are constants of the caller, not tensors
    for (int c = 0; c < C; ++c) { p.mean[c] = mean[c]; p.stdv[c] = stdv[c]; }
  }
  p.out = reinterpret_cast<__half*>(out);
  p.total = static_cast<long long>(B) * rows_per_sample * p.K4;
  p.vec_ok = (p.sw == 1) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (p.sb % 4 == 0) && (p.sc % 4 == 0) &&
             (p.st % 4 == 0) && (p.sh % 4 == 0);
  const int threads = 256;
  const long long blocks = (p.total + threads - 1) / threads;
  ProfileScope prof(static_cast<cudaStream_t>(stream), "patch_gather", 0.0,
                    static_cast<double>(B) * rows_per_sample * K * 6.0);
  if (vec4 && p.K4 <= 256 && B <= 65535 && pt < 256 && ph < 256 && pw < 256) {
    // v2: grid.y = sample, 32-bit index math, shared-memory decode of the patch-volume index (pixelsrc.cuh)
    TensorSrc src;
    src.x = x; src.sb = p.sb; src.st = p.st; src.sc = p.sc; src.sh = p.sh; src.sw = p.sw; src.vec_ok = p.vec_ok;
    GatherGeom g;
    g.C = C; g.pt = pt; g.ph = ph; g.pw = pw; g.n_h = p.n_h; g.n_w = p.n_w; g.K4 = p.K4; g.tpb = 256 / p.K4;
    g.perm = perm; g.Ntot = Ntot; g.rows_per_sample = rows_per_sample; g.n_tokens = p.n_tokens;
    for (int c = 0; c < 8; ++c) { g.mean[c] = p.mean[c]; g.stdv[c] = p.stdv[c]; }
    g.normalize = p.normalize; g.out = p.out;
    launch_patch_gather2(src, g, B, static_cast<cudaStream_t>(stream));
    CWM_LAUNCH_CHECK();
    return CWM_OK;
  }
  if (vec4)
    patch_gather_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else
    patch_gather_scalar_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_layernorm_f16(const float* x, int M, int C, const float* gamma, const float* beta, float eps,
                                 int grp_rows, int grp_stride, int grp_offset, uint16_t* out,
                                 cwm_stream_t stream) {
  CWM_REQUIRE(x && gamma && beta && out, "cwm_layernorm_f16: null pointer");
  CWM_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024, "cwm_layernorm_f16: C=%d must be a multiple of 4 in [4,1024]", C);
  if (M == 0) return CWM_OK;
  const int threads = 256;
  const int rows_per_cta = threads / 32;
  const unsigned blocks = (M + rows_per_cta - 1) / rows_per_cta;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __half* o = reinterpret_cast<__half*>(out);
  ProfileScope prof(s, "layernorm_f16", 0.0, static_cast<double>(M) * C * 6.0);
  if (C % 128 != 0) {
    layernorm_f16_generic_kernel<<<blocks, threads, 0, s>>>(x, M, C, gamma, beta, eps, grp_rows, grp_stride, grp_offset, o);
    CWM_LAUNCH_CHECK();
    return CWM_OK;
  }
#define LN_CASE(V)                                                                                        \
  case V:                                                                                                 \
    layernorm_f16_kernel<V><<<blocks, threads, 0, s>>>(x, M, gamma, beta, eps, grp_rows, grp_stride,      \
                                                       grp_offset, o);                                    \
    break;
  switch (C / 128) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
    default:
      return fail(CWM_ERR_UNSUPPORTED, "cwm_layernorm_f16: unsupported C=%d", C);
  }
#undef LN_CASE
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_rowstats_f16(const float* x, int M, int C, uint16_t* x16, float* stats, cwm_stream_t stream) {
  CWM_REQUIRE(x && x16 && stats, "cwm_rowstats_f16: null pointer");
  CWM_REQUIRE(C % 4 == 0 && C >= 4, "cwm_rowstats_f16: C=%d must be a multiple of 4", C);
  if (M == 0) return CWM_OK;
  const int threads = 256;
  const unsigned blocks = (M + 7) / 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProfileScope prof(s, "rowstats_f16", 0.0, static_cast<double>(M) * C * 6.0);
  rowstats_f16_kernel<<<blocks, threads, 0, s>>>(x, M, C / 4, reinterpret_cast<__half*>(x16),
                                                 reinterpret_cast<float2*>(stats));
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_fill_mask_tokens(const float* mask_token, const float* pos, const int32_t* perm, int B, int Ntot,
                                    int Nvis, int C, float* x_full, cwm_stream_t stream) {
  CWM_REQUIRE(mask_token && pos && perm && x_full, "cwm_fill_mask_tokens: null pointer");
  CWM_REQUIRE(C % 4 == 0, "cwm_fill_mask_tokens: C=%d must be a multiple of 4", C);
  CWM_REQUIRE(Nvis >= 0 && Nvis <= Ntot, "cwm_fill_mask_tokens: Nvis=%d out of range", Nvis);
  const long long total = static_cast<long long>(B) * (Ntot - Nvis) * (C / 4);
  if (total == 0) return CWM_OK;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  ProfileScope prof(static_cast<cudaStream_t>(stream), "fill_mask_tokens", 0.0,
                    static_cast<double>(B) * (Ntot - Nvis) * C * 4.0);
  const long long per_sample = static_cast<long long>(Ntot - Nvis) * (C / 4);
  if (B <= 65535 && per_sample < (1ll << 30)) {
    if (false && (C / 4) % 4 == 0) {  // measured: 4 vectors per thread breaks store coalescing (27 -> 39 us)
      dim3 grid(static_cast<unsigned>((per_sample / 4 + threads - 1) / threads), B);
      fill_mask_tokens2_kernel<4><<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(
          reinterpret_cast<const float4*>(mask_token), reinterpret_cast<const float4*>(pos), perm, Ntot, Nvis, C / 4,
          static_cast<int>(per_sample), reinterpret_cast<float4*>(x_full));
    } else {
      dim3 grid(static_cast<unsigned>((per_sample + threads - 1) / threads), B);
      fill_mask_tokens2_kernel<1><<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(
          reinterpret_cast<const float4*>(mask_token), reinterpret_cast<const float4*>(pos), perm, Ntot, Nvis, C / 4,
          static_cast<int>(per_sample), reinterpret_cast<float4*>(x_full));
    }
    CWM_LAUNCH_CHECK();
    return CWM_OK;
  }
  fill_mask_tokens_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(mask_token), reinterpret_cast<const float4*>(pos), perm, Ntot, Nvis, C / 4,
      total, reinterpret_cast<float4*>(x_full));
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_unpatchify_scatter(const float* y, const float* x_raw, const int64_t xs[5],
                                      const int32_t* inv_perm, int B, int T, int C, int H, int W, int pt, int ph,
                                      int pw, int Nvis, float* out, cwm_stream_t stream) {
  CWM_REQUIRE(x_raw && xs && inv_perm && out, "cwm_unpatchify_scatter: null pointer");
  CWM_REQUIRE(pt > 0 && ph > 0 && pw > 0 && T % pt == 0 && H % ph == 0 && W % pw == 0,
              "cwm_unpatchify_scatter: video (%d,%d,%d) not divisible by patch (%d,%d,%d)", T, H, W, pt, ph, pw);
  CWM_REQUIRE(pw % 4 == 0 && W % 4 == 0, "cwm_unpatchify_scatter: patch width %d must be a multiple of 4", pw);
  UnpatchParams p;
  p.y = y; p.x = x_raw;
  // x_raw logical [B, T, C, H, W]
  p.sb = xs[0]; p.st = xs[1]; p.sc = xs[2]; p.sh = xs[3]; p.sw = xs[4];
  p.inv_perm = inv_perm;
  p.T = T; p.C = C; p.H = H; p.W = W; p.pt = pt; p.ph = ph; p.pw = pw;
  p.n_h = H / ph; p.n_w = W / pw;
  p.Ntot = (T / pt) * p.n_h * p.n_w;
  p.Nvis = Nvis;
  p.D = pt * ph * pw * C;
  CWM_REQUIRE(Nvis >= 0 && Nvis <= p.Ntot, "cwm_unpatchify_scatter: Nvis=%d out of range", Nvis);
  CWM_REQUIRE(y != nullptr || Nvis == p.Ntot, "cwm_unpatchify_scatter: y is NULL but there are masked tokens");
  p.total = static_cast<long long>(B) * T * C * H * (W / 4);
  p.out = out;
  p.vec_ok = (p.sw == 1) && (reinterpret_cast<uintptr_t>(x_raw) % 16 == 0) && (p.sb % 4 == 0) &&
             (p.sc % 4 == 0) && (p.st % 4 == 0) && (p.sh % 4 == 0);
  if (p.total == 0) return CWM_OK;
  const int threads = 256;
  const long long blocks = (p.total + threads - 1) / threads;
  ProfileScope prof(static_cast<cudaStream_t>(stream), "unpatchify_scatter", 0.0,
                    static_cast<double>(p.total) * 4 * 8.0);
  if (C == 3 && B <= 65535 && W / 4 <= 256 && (y == nullptr || (reinterpret_cast<uintptr_t>(y) % 16 == 0 && p.D % 4 == 0)) &&
      reinterpret_cast<uintptr_t>(out) % 16 == 0) {
    // v2: one thread per 4 pixels x all channels, contiguous 16-byte reads of the prediction rows (pixelsrc.cuh)
    TensorSrc src;
    src.x = x_raw; src.sb = p.sb; src.st = p.st; src.sc = p.sc; src.sh = p.sh; src.sw = p.sw; src.vec_ok = p.vec_ok;
    UnpatchGeom g;
    g.y = y; g.inv_perm = inv_perm; g.T = T; g.H = H; g.W = W; g.pt = pt; g.ph = ph; g.pw = pw; g.n_h = p.n_h;
    g.n_w = p.n_w; g.Ntot = p.Ntot; g.Nvis = Nvis; g.D = p.D; g.per_sample = T * H * (W / 4); g.out = out;
    launch_unpatchify2(src, g, B, static_cast<cudaStream_t>(stream));
    CWM_LAUNCH_CHECK();
    return CWM_OK;
  }
  unpatchify_scatter_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_fill_pad_rows(float* x, int B, int rows, int C, const int32_t* perm, int perm_stride,
                                 int perm_offset, int first_pad_token, const float* value, cwm_stream_t stream) {
  CWM_REQUIRE(x && perm, "cwm_fill_pad_rows: null pointer");
  CWM_REQUIRE(C % 4 == 0 && rows >= 0 && B >= 0 && perm_offset >= 0 && perm_offset + rows <= perm_stride,
              "cwm_fill_pad_rows: bad shape rows=%d C=%d perm_offset=%d perm_stride=%d", rows, C, perm_offset, perm_stride);
  const long long total = static_cast<long long>(B) * rows * (C / 4);
  if (total == 0) return CWM_OK;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  ProfileScope prof(static_cast<cudaStream_t>(stream), "fill_pad_rows", 0.0, static_cast<double>(B) * rows * 4.0);
  fill_pad_rows_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(x), rows, C / 4, perm, perm_stride, perm_offset, first_pad_token,
      reinterpret_cast<const float4*>(value), total);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}
