// flowstats.cu -- SURVEY.md section 8(f) rank 2: flow-derived statistics of a counterfactual sweep on device.
//
//   FlowSampleFilter.compute_flow_magnitude / filter_by_* / forward      cwm/models/sampling.py:163-286
//   FlowGenerator.compute_flow_samples_magnitude / compute_mean_motion_map  cwm/models/segmentation.py:250-276
//
// The flow samples are a logical [B, 2, H, W, S] fp32 tensor with arbitrary element strides: the reference hands a
// permuted view of the flow network's [(b s), 2, H, W] output (segmentation.py:130-140), which is the layout these
// kernels are fastest on (x contiguous); nothing is ever re-laid-out.  All kernels are HBM-bound streaming
// reductions: every flow sample is read once per statistic (8*H*W bytes), outputs are O(B*S) or O(B*H*W).
// Floating point: fp32 like the reference; sums over samples / pixels are tree- or atomics-ordered, so results match
// the reference to rounding (tests: 1e-5 relative), thresholds are compared exactly as the reference does.
#include "common.cuh"

namespace cwm {

struct FlowView {
  const float* p;
  int64_t sb, sc, sh, sw, ss;
  int H, W, S;
  __device__ __forceinline__ float mag(int b, int y, int x, int s) const {
    const float* q = p + b * sb + y * sh + x * sw + s * ss;
    const float u = __ldg(q), v = __ldg(q + sc);
    // flow_samples.norm(dim=1, p=2) / square().sum().sqrt(): sqrt(u*u + v*v) with separate roundings
    return __fsqrt_rn(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)));
  }
};

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;  // valid in warp 0
}
__device__ __forceinline__ float block_max(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = -INFINITY;
  if (warp == 0) {
    r = (lane < (blockDim.x >> 5)) ? sh[lane] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
  }
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------------
// per-sample statistics: kStatChunks CTAs per (b, s), each streaming a slice of the image (16-byte loads when the
// rows are contiguous), combined with order-independent atomics -- a pixel COUNT (exact in fp32 up to 2^24) and
// min / max on the bit patterns of the non-negative magnitudes -- so the result is deterministic.  Chunk 0 also
// evaluates the few active patches and the corners.  acc[b, s, :] = {patch sum, patch count, pixels above the
// threshold, corners, min bits, max bits} (zero / +inf-initialised by the host wrapper); flow_stats_finalize_kernel
// turns it into
//   stats[b, s, 0] = patch_flow_mag   (sampling.py:189-200: bilinear-downsampled magnitude averaged over the active
//                                       patches of frame 2; align_corners=False source index like ATen)
//   stats[b, s, 1] = flow_area        (:222-223: fraction of pixels with magnitude > threshold)
//   stats[b, s, 2] = num_corners      (:232-247)
//   stats[b, s, 3] = min magnitude, stats[b, s, 4] = max magnitude (per-sample normalisation, segmentation.py:252-254)
// ---------------------------------------------------------------------------------------------
constexpr int kStatChunks = 8;

__global__ void __launch_bounds__(256)
flow_sample_stats_kernel(FlowView f, const uint8_t* __restrict__ active, int64_t ab, int64_t an, int64_t as_, int n_h,
                         int n_w, float thr, int vec_ok, float* __restrict__ acc) {
  __shared__ float sh[8];
  const int s = blockIdx.x % f.S, b = blockIdx.x / f.S;
  const int chunk = blockIdx.y;
  const int HW = f.H * f.W;
  float cnt = 0.f, mn = INFINITY, mx = 0.f;
  if (vec_ok) {  // sw == 1, W % 4 == 0, 16-byte aligned rows
    const int W4 = f.W >> 2, n4 = HW >> 2;
    const float* base = f.p + b * f.sb + s * f.ss;
    for (int i = chunk * blockDim.x + threadIdx.x; i < n4; i += blockDim.x * kStatChunks) {
      const int y = i / W4, x4 = i - y * W4;
      const float4 u = __ldg(reinterpret_cast<const float4*>(base + y * f.sh) + x4);
      const float4 v = __ldg(reinterpret_cast<const float4*>(base + f.sc + y * f.sh) + x4);
      const float m0 = __fsqrt_rn(__fadd_rn(__fmul_rn(u.x, u.x), __fmul_rn(v.x, v.x)));
      const float m1 = __fsqrt_rn(__fadd_rn(__fmul_rn(u.y, u.y), __fmul_rn(v.y, v.y)));
      const float m2 = __fsqrt_rn(__fadd_rn(__fmul_rn(u.z, u.z), __fmul_rn(v.z, v.z)));
      const float m3 = __fsqrt_rn(__fadd_rn(__fmul_rn(u.w, u.w), __fmul_rn(v.w, v.w)));
      cnt += ((m0 > thr) ? 1.f : 0.f) + ((m1 > thr) ? 1.f : 0.f) + ((m2 > thr) ? 1.f : 0.f) + ((m3 > thr) ? 1.f : 0.f);
      mn = fminf(fminf(mn, m0), fminf(m1, fminf(m2, m3)));
      mx = fmaxf(fmaxf(mx, m0), fmaxf(m1, fmaxf(m2, m3)));
    }
  } else {
    for (int i = chunk * blockDim.x + threadIdx.x; i < HW; i += blockDim.x * kStatChunks) {
      const int y = i / f.W, x = i - y * f.W;
      const float m = f.mag(b, y, x, s);
      cnt += (m > thr) ? 1.f : 0.f;
      mn = fminf(mn, m);
      mx = fmaxf(mx, m);
    }
  }
  float psum = 0.f, pcnt = 0.f;
  if (chunk == 0 && active != nullptr) {
    const float sy = static_cast<float>(f.H) / n_h, sx = static_cast<float>(f.W) / n_w;
    const int n = n_h * n_w;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      // active_second = 1 - active_patches[:, h*w:, :]  (sampling.py:190)
      if (active[b * ab + (n + i) * an + s * as_] != 0) continue;
      const int py = i / n_w, px = i - py * n_w;
      // F.interpolate(mode='bilinear', align_corners=False): src = scale * (dst + 0.5) - 0.5, clamped at 0
      float fy = fmaxf(sy * (py + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * (px + 0.5f) - 0.5f, 0.f);
      const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
      const int y1 = min(y0 + 1, f.H - 1), x1 = min(x0 + 1, f.W - 1);
      const float ly = fy - y0, lx = fx - x0;
      const float v = (1.f - ly) * ((1.f - lx) * f.mag(b, y0, x0, s) + lx * f.mag(b, y0, x1, s)) +
                      ly * ((1.f - lx) * f.mag(b, y1, x0, s) + lx * f.mag(b, y1, x1, s));
      psum += v;
      pcnt += 1.f;
    }
  }
  const float t_cnt = block_sum(cnt, sh);
  const float t_mn = -block_max(-mn, sh);
  const float t_mx = block_max(mx, sh);
  float* o = acc + (static_cast<long long>(b) * f.S + s) * 6;
  if (chunk == 0) {  // block-uniform branch: the patch sums are written by exactly one CTA (deterministic order)
    const float t_psum = block_sum(psum, sh);
    const float t_pcnt = block_sum(pcnt, sh);
    if (threadIdx.x == 0) {
      o[0] = t_psum;
      o[1] = t_pcnt;
      o[3] = ((f.mag(b, 0, 0, s) > thr) ? 1.f : 0.f) + ((f.mag(b, 0, f.W - 1, s) > thr) ? 1.f : 0.f) +
             ((f.mag(b, f.H - 1, 0, s) > thr) ? 1.f : 0.f) + ((f.mag(b, f.H - 1, f.W - 1, s) > thr) ? 1.f : 0.f);
    }
  }
  if (threadIdx.x == 0) {
    atomicAdd(o + 2, t_cnt);  // integer-valued: exact and order-independent
    atomicMin(reinterpret_cast<int*>(o + 4), __float_as_int(t_mn));  // magnitudes are >= 0: bit order == value order
    atomicMax(reinterpret_cast<int*>(o + 5), __float_as_int(t_mx));
  }
}

__global__ void flow_stats_finalize_kernel(const float* __restrict__ acc, int n, float hw, float* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* a = acc + static_cast<long long>(i) * 6;
  float* o = stats + static_cast<long long>(i) * 5;
  o[0] = a[0] / (a[1] + 1e-12f);
  o[1] = a[2] / hw;
  o[2] = a[3];
  o[3] = a[4];
  o[4] = a[5];
}

__global__ void flow_stats_init_kernel(float* __restrict__ acc, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* a = acc + static_cast<long long>(i) * 6;
  a[0] = a[1] = a[2] = a[3] = 0.f;
  a[4] = INFINITY;
  a[5] = 0.f;
}

// filter mask [B, S] from the statistics (sampling.py:202-247, :266-279); methods bit 0 = patch_magnitude,
// bit 1 = flow_area, bit 2 = num_corners
__global__ void flow_filter_mask_kernel(const float* __restrict__ stats, int n, int methods, float mag_thr, float area_thr,
                                        float corners_thr, uint8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* s = stats + static_cast<long long>(i) * 5;
  bool f = false;
  if (methods & 1) f = f || (s[0] < mag_thr);
  if (methods & 2) f = f || (s[1] > area_thr);
  if (methods & 4) f = f || (s[2] >= corners_thr);
  out[i] = f ? 1 : 0;
}

// flow_samples[filter_mask] = 0  (sampling.py:282-284), in place on the strided view; one thread per element
__global__ void __launch_bounds__(256)
flow_zero_filtered_kernel(float* p, int64_t sb, int64_t sc, int64_t sh_, int64_t sw, int64_t ss, int H, int W, int S,
                          const uint8_t* __restrict__ filt, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % W);
  long long r = i / W;
  const int y = static_cast<int>(r % H);
  r /= H;
  const int c = static_cast<int>(r % 2);
  r /= 2;
  const int s = static_cast<int>(r % S);
  const long long b = r / S;
  if (filt[b * S + s]) p[b * sb + c * sc + y * sh_ + x * sw + s * ss] = 0.f;
}

// sum over samples of the (optionally per-sample normalised, optionally filtered) magnitude.  One thread per pixel
// and sample slice (grid.y slices keep > 100k threads in flight for any S); the slice sums go to `partial`
// [slices, B*H*W] and flow_partial_reduce_kernel adds them in slice order (deterministic).
//   g_s(m) = normalize_per_sample ? (m - min_s) / max(max_s - min_s, eps) : m
__global__ void __launch_bounds__(256)
flow_magnitude_sum_kernel(FlowView f, const uint8_t* __restrict__ filt, const float* __restrict__ stats,
                          int normalize_per_sample, float eps, int per_slice, long long total, float* __restrict__ partial) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % f.W);
  long long r = i / f.W;
  const int y = static_cast<int>(r % f.H);
  const int b = static_cast<int>(r / f.H);
  const int s0 = blockIdx.y * per_slice, s1 = min(f.S, s0 + per_slice);
  float acc = 0.f;
#pragma unroll 4
  for (int s = s0; s < s1; ++s) {
    const bool dropped = filt != nullptr && filt[b * f.S + s];
    float m = dropped ? 0.f : f.mag(b, y, x, s);
    if (normalize_per_sample) {
      // a filtered sample is all zeros: min = max = 0 -> (0 - 0) / clamp(0, eps) = 0
      const float* st = stats + (static_cast<long long>(b) * f.S + s) * 5;
      const float mn = dropped ? 0.f : st[3], mx = dropped ? 0.f : st[4];
      m = __fdiv_rn(m - mn, fmaxf(mx - mn, eps));
    }
    acc += m;
  }
  partial[blockIdx.y * total + i] = acc;
}

__global__ void __launch_bounds__(256)
flow_partial_reduce_kernel(const float* __restrict__ partial, int slices, long long total, int accumulate,
                           float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float acc = accumulate ? out[i] : 0.f;
  for (int k = 0; k < slices; ++k) acc += partial[k * total + i];
  out[i] = acc;
}

// motion_map = sums / count (`flow_mags.mean(-1)`); then (optional) minus its min over (H, W), divided by max.clamp(eps)
// (segmentation.py:268-275).  One CTA per image.
__global__ void __launch_bounds__(1024)
motion_map_finalize_kernel(const float* __restrict__ sums, int HW, float count, int normalize, float eps,
                           float* __restrict__ out) {
  __shared__ float sh[32];
  __shared__ float s_mn, s_mx;
  const float* in = sums + static_cast<long long>(blockIdx.x) * HW;
  float* o = out + static_cast<long long>(blockIdx.x) * HW;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float v = __fdiv_rn(in[i], count);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  const float t_mn = -block_max(-mn, sh);
  const float t_mx = block_max(mx, sh);
  if (threadIdx.x == 0) { s_mn = t_mn; s_mx = t_mx; }
  __syncthreads();
  const float lo = normalize ? s_mn : 0.f;
  const float den = normalize ? fmaxf(s_mx - s_mn, eps) : 1.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float v = __fdiv_rn(in[i], count);
    o[i] = normalize ? __fdiv_rn(v - lo, den) : v;
  }
}

// ---------------------------------------------------------------------------------------------
// motion covariance / correlation between image locations over the samples of a sweep
// (`FlowGenerator.compute_flow_corrs`, segmentation.py:478-547, default options):
//   X[n, s]  = sqrt(mean_c(avg_pool_ds(flow)[c, n, s]^2))          (ChannelMSE against zeros, utils.py:510-513)
//   C        = (X - mean_s X)(X - mean_s X)^T / (S - 1)             (torch.cov), or
//   R[i, j]  = clamp(C[i, j] / sqrt(C[i, i]) / sqrt(C[j, j]), -1, 1) (torch.corrcoef);   NaN -> 0
// Kernel 1 writes the centred features Xc [B, N, S] and the row standard deviations; kernel 2 is an fp32
// outer-product GEMM with K = S (tiny), i.e. bound by the N^2 * 4 bytes it writes.
// ---------------------------------------------------------------------------------------------
// one thread per (location, sample)
__global__ void __launch_bounds__(256)
flow_corr_features_kernel(FlowView f, int ds, int n_h, int n_w, int K, long long total, float* __restrict__ xc) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // location fastest: neighbouring threads read neighbouring ds-pixel groups of one flow sample (coalesced reads of
  // the 8*H*W bytes per sample); the 4-byte writes into xc[b, n, s] are strided but 1/(2*ds*ds) of the traffic
  const int N = n_h * n_w;
  const int n = static_cast<int>(i % N);
  long long r = i / N;
  const int s = static_cast<int>(r % K);
  const int b = static_cast<int>(r / K);
  const int py = n / n_w, px = n - py * n_w;
  const float* q = f.p + b * f.sb + s * f.ss;
  float su = 0.f, sv = 0.f;
  for (int dy = 0; dy < ds; ++dy)
    for (int dx = 0; dx < ds; ++dx) {
      const float* e = q + (py * ds + dy) * f.sh + (px * ds + dx) * f.sw;
      su += __ldg(e);
      sv += __ldg(e + f.sc);
    }
  const float area = static_cast<float>(ds * ds);
  const float u = __fdiv_rn(su, area), v = __fdiv_rn(sv, area);
  xc[(static_cast<long long>(b) * N + n) * K + s] = __fsqrt_rn(__fdiv_rn(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)), 2.f));
}

// one warp per location: centre the K samples and record the standard deviation
__global__ void __launch_bounds__(256)
flow_corr_center_kernel(float* __restrict__ xc, int K, long long rows, float* __restrict__ stdev) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* x = xc + row * K;
  float sum = 0.f;
  for (int s = lane; s < K; s += 32) sum += x[s];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = __fdiv_rn(sum, static_cast<float>(K));
  float ss = 0.f;
  for (int s = lane; s < K; s += 32) {
    const float c = x[s] - mean;
    x[s] = c;
    ss += c * c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) stdev[row] = __fsqrt_rn(__fdiv_rn(ss, static_cast<float>(K - 1)));
}

// C tile 64 x 64 per CTA, 4 x 4 outputs per thread, K (= samples) streamed through shared memory in chunks of 32
__global__ void __launch_bounds__(256)
flow_cov_kernel(const float* __restrict__ xc, const float* __restrict__ stdev, int N, int K, int use_covariance,
                float* __restrict__ out) {
  __shared__ float a_s[32][65], b_s[32][65];
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const float* xb = xc + static_cast<long long>(b) * N * K;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int e = threadIdx.x; e < 64 * 32; e += 256) {
      const int r = e >> 5, k = e & 31;
      a_s[k][r] = (i0 + r < N && k0 + k < K) ? xb[static_cast<long long>(i0 + r) * K + k0 + k] : 0.f;
      b_s[k][r] = (j0 + r < N && k0 + k < K) ? xb[static_cast<long long>(j0 + r) * K + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { av[u] = a_s[k][ty * 4 + u]; bv[u] = b_s[k][tx * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
    }
    __syncthreads();
  }
  const float norm = static_cast<float>(K - 1);
  const float* sd = stdev + static_cast<long long>(b) * N;
  float* ob = out + static_cast<long long>(b) * N * N;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ty * 4 + u;
    if (i >= N) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int j = j0 + tx * 4 + v;
      if (j >= N) continue;
      float c = __fdiv_rn(acc[u][v], norm);
      if (!use_covariance) {
        c = __fdiv_rn(__fdiv_rn(c, sd[i]), sd[j]);
        c = fminf(fmaxf(c, -1.f), 1.f);  // fminf/fmaxf drop a NaN operand, so test it first
        if (!(acc[u][v] == acc[u][v]) || !(sd[i] > 0.f) || !(sd[j] > 0.f)) c = __int_as_float(0x7fc00000);
      }
      if (c != c) c = 0.f;  // flow_corrs_b[isnan] = 0 (segmentation.py:541)
      ob[static_cast<long long>(i) * N + j] = c;
    }
  }
}

static int make_view(const float* flows, const int64_t fs[5], int B, int H, int W, int S, FlowView* v, const char* who) {
  if (!flows || !fs) return fail(CWM_ERR_INVALID, "%s: null pointer", who);
  if (B < 0 || H <= 0 || W <= 0 || S < 0) return fail(CWM_ERR_INVALID, "%s: bad shape B=%d H=%d W=%d S=%d", who, B, H, W, S);
  v->p = flows;
  v->sb = fs[0]; v->sc = fs[1]; v->sh = fs[2]; v->sw = fs[3]; v->ss = fs[4];
  v->H = H; v->W = W; v->S = S;
  return CWM_OK;
}

}  // namespace cwm

using namespace cwm;

extern "C" size_t cwm_flow_stats_workspace_bytes(int B, int H, int W, int S) {
  if (B < 0 || H <= 0 || W <= 0 || S < 0) return 0;
  const size_t acc = static_cast<size_t>(B) * S * 6 * sizeof(float);
  const size_t partial = static_cast<size_t>(32) * B * H * W * sizeof(float);
  return (acc > partial ? acc : partial) + 256;
}

extern "C" int cwm_flow_sample_stats(const float* flows, const int64_t fs[5], int B, int H, int W, int S,
                                     const uint8_t* active, const int64_t as[3], int n_h, int n_w,
                                     float magnitude_threshold, float* stats, void* workspace, size_t workspace_bytes,
                                     cwm_stream_t stream) {
  FlowView v;
  int rc = make_view(flows, fs, B, H, W, S, &v, "cwm_flow_sample_stats");
  if (rc != CWM_OK) return rc;
  CWM_REQUIRE(stats && workspace, "cwm_flow_sample_stats: null pointer");
  CWM_REQUIRE(active == nullptr || (as && n_h > 0 && n_w > 0), "cwm_flow_sample_stats: active patches need strides and a grid");
  if (workspace_bytes < static_cast<size_t>(B) * S * 6 * sizeof(float))
    return fail(CWM_ERR_WORKSPACE, "cwm_flow_sample_stats: workspace %zu bytes too small", workspace_bytes);
  const int n = B * S;
  if (n == 0) return CWM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* acc = reinterpret_cast<float*>(workspace);
  const int vec_ok = (fs[3] == 1) && (W % 4 == 0) && (reinterpret_cast<uintptr_t>(flows) % 16 == 0) && (fs[0] % 4 == 0) &&
                     (fs[1] % 4 == 0) && (fs[2] % 4 == 0) && (fs[4] % 4 == 0);
  flow_stats_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(acc, n);
  CWM_LAUNCH_CHECK();
  {
    ProfileScope prof(st, "flow_sample_stats", 0.0, static_cast<double>(B) * S * H * W * 8.0);
    flow_sample_stats_kernel<<<dim3(n, kStatChunks), 256, 0, st>>>(v, active, active ? as[0] : 0, active ? as[1] : 0,
                                                                   active ? as[2] : 0, n_h, n_w, magnitude_threshold,
                                                                   vec_ok, acc);
    CWM_LAUNCH_CHECK();
  }
  flow_stats_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(acc, n, static_cast<float>(H) * W, stats);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_flow_filter_mask(const float* stats, int B, int S, int methods, float magnitude_threshold,
                                    float area_threshold, float corners_threshold, uint8_t* filter_mask,
                                    cwm_stream_t stream) {
  CWM_REQUIRE(stats && filter_mask, "cwm_flow_filter_mask: null pointer");
  const int n = B * S;
  if (n == 0) return CWM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "flow_filter_mask", 0.0, n * 21.0);
  flow_filter_mask_kernel<<<(n + 255) / 256, 256, 0, st>>>(stats, n, methods, magnitude_threshold, area_threshold,
                                                          corners_threshold, filter_mask);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_flow_zero_filtered(float* flows, const int64_t fs[5], int B, int H, int W, int S,
                                      const uint8_t* filter_mask, cwm_stream_t stream) {
  FlowView v;
  int rc = make_view(flows, fs, B, H, W, S, &v, "cwm_flow_zero_filtered");
  if (rc != CWM_OK) return rc;
  CWM_REQUIRE(filter_mask, "cwm_flow_zero_filtered: null pointer");
  const long long total = static_cast<long long>(B) * S * 2 * H * W;
  if (total == 0) return CWM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "flow_zero_filtered", 0.0, static_cast<double>(total) * 4.0);
  flow_zero_filtered_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(flows, fs[0], fs[1], fs[2], fs[3],
                                                                                        fs[4], H, W, S, filter_mask, total);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_flow_magnitude_sum(const float* flows, const int64_t fs[5], int B, int H, int W, int S,
                                      const uint8_t* filter_mask, const float* stats, int normalize_per_sample,
                                      float eps, int accumulate, float* sums, void* workspace, size_t workspace_bytes,
                                      cwm_stream_t stream) {
  FlowView v;
  int rc = make_view(flows, fs, B, H, W, S, &v, "cwm_flow_magnitude_sum");
  if (rc != CWM_OK) return rc;
  CWM_REQUIRE(sums && workspace, "cwm_flow_magnitude_sum: null pointer");
  CWM_REQUIRE(!normalize_per_sample || stats, "cwm_flow_magnitude_sum: per-sample normalisation needs the statistics");
  const long long total = static_cast<long long>(B) * H * W;
  if (total == 0) return CWM_OK;
  // enough sample slices for ~300k threads in flight, at most 32, at least 4 samples per slice
  int slices = static_cast<int>((300000 + total - 1) / total);
  if (slices > 32) slices = 32;
  if (slices > (S + 3) / 4) slices = (S + 3) / 4;
  if (slices < 1) slices = 1;
  const int per_slice = S > 0 ? (S + slices - 1) / slices : 1;
  slices = S > 0 ? (S + per_slice - 1) / per_slice : 1;
  if (workspace_bytes < static_cast<size_t>(slices) * total * sizeof(float))
    return fail(CWM_ERR_WORKSPACE, "cwm_flow_magnitude_sum: workspace %zu bytes too small", workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = reinterpret_cast<float*>(workspace);
  {
    ProfileScope prof(st, "flow_magnitude_sum", 0.0, static_cast<double>(total) * (8.0 * S + 4.0));
    flow_magnitude_sum_kernel<<<dim3(static_cast<unsigned>((total + 255) / 256), slices), 256, 0, st>>>(
        v, filter_mask, stats, normalize_per_sample, eps, per_slice, total, partial);
    CWM_LAUNCH_CHECK();
  }
  flow_partial_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(partial, slices, total, accumulate,
                                                                                         sums);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_motion_map_finalize(const float* sums, int B, int H, int W, float count, int normalize, float eps,
                                       float* motion_map, cwm_stream_t stream) {
  CWM_REQUIRE(sums && motion_map, "cwm_motion_map_finalize: null pointer");
  CWM_REQUIRE(B >= 0 && H > 0 && W > 0 && count > 0.f, "cwm_motion_map_finalize: bad shape");
  if (B == 0) return CWM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "motion_map_finalize", 0.0, static_cast<double>(B) * H * W * 12.0);
  motion_map_finalize_kernel<<<B, 1024, 0, st>>>(sums, H * W, count, normalize, eps, motion_map);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" size_t cwm_flow_corrs_workspace_bytes(int B, int H, int W, int S, int downsample) {
  if (B < 0 || H <= 0 || W <= 0 || S < 0 || downsample <= 0) return 0;
  const size_t N = static_cast<size_t>(H / downsample) * (W / downsample);
  return (static_cast<size_t>(B) * N * (S + 1)) * sizeof(float) + 256;
}

extern "C" int cwm_flow_corrs(const float* flows, const int64_t fs[5], int B, int H, int W, int S, int downsample,
                              int use_covariance, float* out, void* workspace, size_t workspace_bytes,
                              cwm_stream_t stream) {
  FlowView v;
  int rc = make_view(flows, fs, B, H, W, S, &v, "cwm_flow_corrs");
  if (rc != CWM_OK) return rc;
  CWM_REQUIRE(out && workspace, "cwm_flow_corrs: null pointer");
  CWM_REQUIRE(downsample > 0 && H % downsample == 0 && W % downsample == 0,
              "cwm_flow_corrs: image (%d,%d) not divisible by downsample %d", H, W, downsample);
  CWM_REQUIRE(S >= 1, "cwm_flow_corrs: needs at least one sample (the caller substitutes a zero sample, segmentation.py:494-497)");
  const int n_h = H / downsample, n_w = W / downsample, N = n_h * n_w;
  if (workspace_bytes < cwm_flow_corrs_workspace_bytes(B, H, W, S, downsample))
    return fail(CWM_ERR_WORKSPACE, "cwm_flow_corrs: workspace %zu bytes too small", workspace_bytes);
  if (B == 0) return CWM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* xc = reinterpret_cast<float*>(workspace);
  float* sd = xc + static_cast<size_t>(B) * N * S;
  {
    ProfileScope prof(st, "flow_corr_features", 0.0, static_cast<double>(B) * S * H * W * 8.0);
    const long long total = static_cast<long long>(B) * N * S;
    flow_corr_features_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(v, downsample, n_h, n_w, S, total,
                                                                                          xc);
    CWM_LAUNCH_CHECK();
    const long long rows = static_cast<long long>(B) * N;
    flow_corr_center_kernel<<<static_cast<unsigned>((rows * 32 + 255) / 256), 256, 0, st>>>(xc, S, rows, sd);
    CWM_LAUNCH_CHECK();
  }
  {
    ProfileScope prof(st, "flow_cov", 2.0 * B * static_cast<double>(N) * N * S, static_cast<double>(B) * N * N * 4.0);
    flow_cov_kernel<<<dim3((N + 63) / 64, (N + 63) / 64, B), 256, 0, st>>>(xc, sd, N, S, use_covariance, out);
    CWM_LAUNCH_CHECK();
  }
  return CWM_OK;
}
