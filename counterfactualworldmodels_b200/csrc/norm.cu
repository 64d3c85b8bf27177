// norm.cu -- instance normalisation of NHWC f16 feature maps for RAFT's feature encoder (cwm/models/raft/extractor.py:
// 118-190, `norm_fn='instance'`: nn.InstanceNorm2d, no affine, eps 1e-5, biased variance), fused with what follows it:
//   y = relu?( (x - mean[s, c]) * rstd[s, c] )        and, at the end of a residual block,   out = relu( shortcut + y ).
// In eager PyTorch every norm is a statistics kernel + a transform kernel on NCHW tensors, plus cuDNN's NCHW <-> NHWC
// transforms around the neighbouring convolutions, a bias add and a relu: 4.5 of the 20.4 ms of a 64-sample flow call.
// Here the activations stay NHWC f16 (what the tensor-core convolutions consume), statistics are fp32:
//   instnorm_stats_kernel   grid (slabs, S): per (sample, pixel slab) partial sum / sum of squares of every channel
//   instnorm_apply_kernel   grid (pixel blocks, S): sums the slab partials in a fixed order (deterministic), then streams
// Both are HBM bound: 2 bytes read per element for the statistics, 2 (+2) read + 2 written for the transform.
#include "common.cuh"

namespace cwm {

constexpr int kNormThreads = 256;

__device__ __forceinline__ void h8_to_f32(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__global__ void __launch_bounds__(kNormThreads)
instnorm_stats_kernel(const __half* __restrict__ x, int HW, int C, int nslab, float2* __restrict__ partial) {
  extern __shared__ float2 red[];   // [pixel lanes][C]
  grid_dep_sync();
  const int s = blockIdx.y, slab = blockIdx.x;
  const int groups = C >> 3;                       // 8-channel groups per pixel
  const int lanes = kNormThreads / groups;         // pixels in flight
  const int cg = threadIdx.x % groups, pl = threadIdx.x / groups;
  const int per = (HW + nslab - 1) / nslab;
  const int p0 = slab * per, p1 = min(HW, p0 + per);
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  if (pl < lanes) {
    const __half* xs = x + (static_cast<size_t>(s) * HW) * C + cg * 8;
#pragma unroll 4   // four 16-byte loads in flight per thread (ncu of the rolled loop: 27 % of the warps resident, latency bound)
    for (int p = p0 + pl; p < p1; p += lanes) {
      float f[8];
      h8_to_f32(__ldg(reinterpret_cast<const uint4*>(xs + static_cast<size_t>(p) * C)), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s1[i] += f[i];
        s2[i] = fmaf(f[i], f[i], s2[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[pl * C + cg * 8 + i] = make_float2(s1[i], s2[i]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kNormThreads) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) {
      const float2 v = red[l * C + c];
      a += v.x;
      b += v.y;
    }
    partial[(static_cast<size_t>(s) * nslab + slab) * C + c] = make_float2(a, b);
  }
}

__global__ void __launch_bounds__(kNormThreads)
instnorm_apply_kernel(const __half* __restrict__ x, const float2* __restrict__ partial, int nslab, int HW, int C, float eps,
                      int relu_inner, const __half* __restrict__ add, int relu_outer, int pixels_per_block,
                      __half* __restrict__ out) {
  extern __shared__ float2 stat[];   // [C]: (mean, rstd)
  grid_dep_sync();
  const int s = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += kNormThreads) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < nslab; ++l) {
      const float2 v = partial[(static_cast<size_t>(s) * nslab + l) * C + c];
      a += v.x;
      b += v.y;
    }
    const float mean = a / static_cast<float>(HW);
    const float var = fmaxf(b / static_cast<float>(HW) - mean * mean, 0.f);
    stat[c] = make_float2(mean, rsqrtf(var + eps));
  }
  __syncthreads();
  const int groups = C >> 3;
  const int p0 = blockIdx.x * pixels_per_block, p1 = min(HW, p0 + pixels_per_block);
  const size_t base = static_cast<size_t>(s) * HW * C;
#pragma unroll 2
  for (int u = p0 * groups + threadIdx.x; u < p1 * groups; u += kNormThreads) {
    const int p = u / groups, cg = u - p * groups;
    const size_t off = base + static_cast<size_t>(p) * C + cg * 8;
    float f[8], g[8];
    h8_to_f32(__ldg(reinterpret_cast<const uint4*>(x + off)), f);
    if (add != nullptr) h8_to_f32(__ldg(reinterpret_cast<const uint4*>(add + off)), g);
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const float2 st0 = stat[cg * 8 + i], st1 = stat[cg * 8 + i + 1];
      float y0 = (f[i] - st0.x) * st0.y, y1 = (f[i + 1] - st1.x) * st1.y;
      if (relu_inner) {
        y0 = fmaxf(y0, 0.f);
        y1 = fmaxf(y1, 0.f);
      }
      if (add != nullptr) {
        y0 += g[i];
        y1 += g[i + 1];
      }
      if (relu_outer) {
        y0 = fmaxf(y0, 0.f);
        y1 = fmaxf(y1, 0.f);
      }
      o[i >> 1] = pack_half2(y0, y1);
    }
    *reinterpret_cast<uint4*>(out + off) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// im2col of a few-channel NCHW fp32 image for a strided k x k convolution (the encoders' 7x7 / 2 stem on 3 channels,
// extractor.py:132): row m = (s, oy, ox) of `out` holds  scale * img[s, c, stride*oy + ky - pad, stride*ox + kx - pad] + shift
// at column (ky * k + kx) * Cin + c, zeros outside the image (the padding of the already-normalised image) and in the
// columns [k*k*Cin, ldo).  The convolution is then ONE plain GEMM with K = ldo on the tcgen05 kernel.  One thread = one
// 16-byte store (8 columns).
constexpr int kIm2colTile = 32;   // output pixels of one image row per CTA
// CIN / K / STRIDE / LDO > 0: compile-time shapes (the index arithmetic is all divisions by these; with run-time divisors
// the kernel is instruction-bound: 370 us instead of ~100 us for 65 frames); 0 = take the run-time arguments.
template <int CIN, int K, int STRIDE, int LDO>
__global__ void __launch_bounds__(256)
im2col_nchw_kernel(const float* __restrict__ img, long long sample_stride, int Cin_rt, int H, int W, int Ho, int Wo, int k_rt,
                   int stride_rt, int pad, float scale, float shift, __half* __restrict__ out, int ldo_rt) {
  const int Cin = CIN > 0 ? CIN : Cin_rt, k = K > 0 ? K : k_rt, stride = STRIDE > 0 ? STRIDE : stride_rt;
  const int ldo = LDO > 0 ? LDO : ldo_rt;
  // the input patch of this CTA's 32 output pixels: [Cin][k][WT] floats, WT = 31 * stride + k, staged with coalesced
  // loads (normalised, zero outside the image); the im2col rows are then assembled from shared memory
  extern __shared__ float patch[];
  const int WT = (kIm2colTile - 1) * stride + k;
  const int kk = k * k * Cin;
  int* col_off = reinterpret_cast<int*>(patch + Cin * k * WT);   // [ldo]: patch offset of im2col column (ky, kx, c); -1 = pad
  const int ox0 = blockIdx.x * kIm2colTile, oy = blockIdx.y;
  const long long s = blockIdx.z;
  const float* im = img + s * sample_stride;
  const int gx0 = ox0 * stride - pad, gy0 = oy * stride - pad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rc = warp; rc < Cin * k; rc += 8) {          // one (channel, patch row) per warp pass: no per-element divisions
    const int c = rc / k, r = rc - c * k;
    const int gy = gy0 + r;
    const float* src = im + (static_cast<long long>(c) * H + gy) * W;
    for (int x = lane; x < WT; x += 32) {
      const int gx = gx0 + x;
      float v = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = fmaf(__ldg(src + gx), scale, shift);
      patch[rc * WT + x] = v;
    }
  }
  for (int col = threadIdx.x; col < ldo; col += 256) {
    int off = -1;
    if (col < kk) {
      const int tap = col / Cin, c = col - tap * Cin;
      const int ky = tap / k;
      off = (c * k + ky) * WT + (tap - ky * k);
    }
    col_off[col] = off;
  }
  __syncthreads();
  const int units = ldo >> 3;
  for (int task = threadIdx.x; task < kIm2colTile * units; task += 256) {
    const int px = task / units, u = task - px * units;
    if (ox0 + px >= Wo) continue;
    const int col0 = u * 8;
    const float* pp = patch + px * stride;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int off = col_off[col0 + i];
      v[i] = off >= 0 ? pp[off] : 0.f;
    }
    const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]), h2 = __floats2half2_rn(v[4], v[5]),
                  h3 = __floats2half2_rn(v[6], v[7]);
    const long long m = (s * Ho + oy) * Wo + ox0 + px;
    *reinterpret_cast<uint4*>(out + m * ldo + col0) =
        make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                   *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
  }
}

// out = relu?(a + b) on f16 rows (the residual joins of the batch-norm context encoder, whose norms are folded into the
// convolution weights: extractor.py:46-56)
__global__ void __launch_bounds__(256)
add_act_f16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, long long n16, int relu, uint4* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n16) return;
  float fa[8], fb[8];
  h8_to_f32(__ldg(a + i), fa);
  h8_to_f32(__ldg(b + i), fb);
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x0 = fa[2 * j] + fb[2 * j], x1 = fa[2 * j + 1] + fb[2 * j + 1];
    if (relu) {
      x0 = fmaxf(x0, 0.f);
      x1 = fmaxf(x1, 0.f);
    }
    const __half2 h = __floats2half2_rn(x0, x1);
    o[j] = *reinterpret_cast<const uint32_t*>(&h);
  }
  out[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

}  // namespace cwm

using namespace cwm;

extern "C" size_t cwm_instnorm_workspace_bytes(int S, int C) { return static_cast<size_t>(S) * 16 * C * sizeof(float2); }

// out[s, p, c] = post((x[s, p, c] - mean[s, c]) / sqrt(var[s, c] + eps)), statistics over the HW pixels of sample s;
// post: optional relu, then optional `+ add[s, p, c]`, then optional relu.  x / add / out: NHWC f16 [S, HW, C], C % 8 == 0.
extern "C" int cwm_instnorm_f16(const uint16_t* x, int S, int HW, int C, float eps, int relu_inner, const uint16_t* add,
                                int relu_outer, uint16_t* out, void* workspace, size_t workspace_bytes, cwm_stream_t stream) {
  CWM_REQUIRE(S >= 0 && HW >= 1 && C >= 8 && C % 8 == 0 && C <= 2048, "cwm_instnorm_f16: bad shape S=%d HW=%d C=%d (C %% 8 == 0)", S, HW, C);
  if (S == 0) return CWM_OK;
  CWM_REQUIRE(x && out && workspace, "cwm_instnorm_f16: null pointer");
  CWM_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(add)) & 15) == 0,
              "cwm_instnorm_f16: tensors must be 16-byte aligned");
  CWM_REQUIRE(workspace_bytes >= cwm_instnorm_workspace_bytes(S, C), "cwm_instnorm_f16: workspace too small");
  CWM_REQUIRE(S <= 65535, "cwm_instnorm_f16: batch %d > 65535", S);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // enough (sample, slab) CTAs to fill the machine, at most 16 slabs (the workspace holds 16)
  int nslab = (8 * num_sms() + S - 1) / S;
  nslab = nslab < 1 ? 1 : (nslab > 16 ? 16 : nslab);
  if (nslab > HW) nslab = HW;
  float2* partial = static_cast<float2*>(workspace);
  const int groups = C / 8, lanes = kNormThreads / groups > 0 ? kNormThreads / groups : 1;
  CWM_REQUIRE(groups <= kNormThreads, "cwm_instnorm_f16: C too large for one CTA row");
  {
    ProfileScope prof(st, "instnorm_stats", 0.0, static_cast<double>(S) * HW * C * 2.0);
    const size_t smem = static_cast<size_t>(lanes) * C * sizeof(float2);
    static size_t configured = 0;
    if (smem > configured && smem > 48 * 1024) {
      CWM_CUDA_CHECK(cudaFuncSetAttribute(instnorm_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      configured = smem;
    }
    CWM_CUDA_CHECK(launch_pdl(instnorm_stats_kernel, dim3(nslab, S), dim3(kNormThreads), smem, st, reinterpret_cast<const __half*>(x),
                              HW, C, nslab, partial));
    CWM_LAUNCH_CHECK();
  }
  {
    ProfileScope prof(st, "instnorm_apply", 0.0, static_cast<double>(S) * HW * C * (add ? 6.0 : 4.0));
    int blocks = (8 * num_sms() + S - 1) / S;
    if (blocks < 1) blocks = 1;
    int ppb = (HW + blocks - 1) / blocks;
    if (ppb < 8) ppb = 8;
    blocks = (HW + ppb - 1) / ppb;
    CWM_CUDA_CHECK(launch_pdl(instnorm_apply_kernel, dim3(blocks, S), dim3(kNormThreads), C * sizeof(float2), st,
                              reinterpret_cast<const __half*>(x), static_cast<const float2*>(partial), nslab, HW, C, eps, relu_inner,
                              reinterpret_cast<const __half*>(add), relu_outer, ppb, reinterpret_cast<__half*>(out)));
    CWM_LAUNCH_CHECK();
  }
  return CWM_OK;
}

extern "C" int cwm_im2col_nchw_f16(const float* img, long long sample_stride, int S, int Cin, int H, int W, int k, int stride,
                                   int pad, float scale, float shift, uint16_t* out, int ldo, cwm_stream_t stream) {
  if (sample_stride == 0) sample_stride = static_cast<long long>(Cin) * H * W;
  CWM_REQUIRE(sample_stride >= static_cast<long long>(Cin) * H * W, "cwm_im2col_nchw_f16: sample stride %lld < Cin*H*W", sample_stride);
  CWM_REQUIRE(S >= 0 && Cin >= 1 && H >= 1 && W >= 1 && k >= 1 && stride >= 1 && pad >= 0 && ldo % 8 == 0 && ldo >= k * k * Cin,
              "cwm_im2col_nchw_f16: bad shape S=%d Cin=%d H=%d W=%d k=%d stride=%d pad=%d ldo=%d", S, Cin, H, W, k, stride, pad, ldo);
  if (S == 0) return CWM_OK;
  CWM_REQUIRE(img && out && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "cwm_im2col_nchw_f16: null or misaligned pointer");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  CWM_REQUIRE(Ho >= 1 && Wo >= 1, "cwm_im2col_nchw_f16: empty output");
  const long long M = static_cast<long long>(S) * Ho * Wo;
  CWM_REQUIRE(S <= 65535 && Ho <= 65535, "cwm_im2col_nchw_f16: S=%d / Ho=%d exceed the grid limits", S, Ho);
  const size_t smem = (static_cast<size_t>(Cin) * k * ((cwm::kIm2colTile - 1) * stride + k) + ldo) * sizeof(float);
  CWM_REQUIRE(smem <= 48 * 1024, "cwm_im2col_nchw_f16: Cin=%d k=%d stride=%d needs %zu bytes of shared memory", Cin, k, stride, smem);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "im2col_nchw", 0.0, static_cast<double>(M) * ldo * 2.0 + static_cast<double>(S) * Cin * H * W * 4.0);
  const dim3 grid((Wo + cwm::kIm2colTile - 1) / cwm::kIm2colTile, Ho, S);
  if (Cin == 3 && k == 7 && stride == 2 && ldo == 152)   // the encoders' stem (extractor.py:132)
    cwm::im2col_nchw_kernel<3, 7, 2, 152><<<grid, 256, smem, st>>>(img, sample_stride, Cin, H, W, Ho, Wo, k, stride, pad, scale, shift,
                                                                   reinterpret_cast<__half*>(out), ldo);
  else
    cwm::im2col_nchw_kernel<0, 0, 0, 0><<<grid, 256, smem, st>>>(img, sample_stride, Cin, H, W, Ho, Wo, k, stride, pad, scale, shift,
                                                                 reinterpret_cast<__half*>(out), ldo);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_add_act_f16(const uint16_t* a, const uint16_t* b, long long n, int relu, uint16_t* out, cwm_stream_t stream) {
  CWM_REQUIRE(n >= 0 && n % 8 == 0, "cwm_add_act_f16: element count %lld must be a multiple of 8", n);
  if (n == 0) return CWM_OK;
  CWM_REQUIRE(a && b && out, "cwm_add_act_f16: null pointer");
  CWM_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              "cwm_add_act_f16: tensors must be 16-byte aligned");
  const long long n16 = n / 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "add_act_f16", 0.0, static_cast<double>(n) * 6.0);
  cwm::add_act_f16_kernel<<<static_cast<unsigned>((n16 + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), n16, relu, reinterpret_cast<uint4*>(out));
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}
