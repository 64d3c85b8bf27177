// raftcorr.cu -- SURVEY.md section 8(f) rank 3, first slice: the RAFT-specific (non-convolution) stages of the flow
// network that runs right after the VMAE path in every counterfactual (cwm/models/segmentation.py:431).
//
//   CorrBlock.corr / __init__   all-pairs correlation volume + average-pooled pyramid   cwm/models/raft/corr.py:12-28, 53-60
//   CorrBlock.__call__          (2r+1)^2 bilinear window lookup at every level          cwm/models/raft/corr.py:30-51
//   bilinear_sampler            pixel coords -> grid_sample(align_corners=True)         cwm/models/raft/utils.py:60-80
//   RAFT.upsample_flow          convex 8x upsampling of the 1/8-resolution flow         cwm/models/raft/raft_model.py:175-186
//
// Everything is fp32 like the reference (RAFT calls `.float()` on the feature maps before the correlation,
// raft_model.py:224-225).  The lookup repeats the reference's coordinate arithmetic operation by operation
// (normalise to [-1, 1], un-normalise inside grid_sample, floor, corner weights, zero padding), so the only
// differences left are summation order in the volume (fp32 FMA chain over the feature channels instead of a BLAS
// order) and FMA contraction: tests hold 2e-5 relative to the volume's scale.
//
// Data layout in HBM: level l of the pyramid is fp32 [B*H*W, H>>l, W>>l] (the reference's `corr_pyramid[l]` without
// its singleton channel); 3.25 MB per sample at 28x28 / 4 levels.  The lookup output is [B, L*(2r+1)^2, H, W].
#include <cstdlib>

#include "common.cuh"

namespace cwm {

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// all-pairs correlation: C[b, i, j] = sum_c f1[b, c, i] * f2[b, c, j] / sqrt(D).   fmaps are [B, D, HW] (NCHW), so
// both operands are contiguous along the output index -> coalesced 16-byte loads with no transposition.
// 64 x 64 outputs per CTA, 4 x 4 per thread, the channel axis streamed through shared memory 16 at a time with the
// next slab prefetched into registers.  fp32 FMA pipe bound (2*HW^2*D FLOP against HW^2*4 bytes written).
// ---------------------------------------------------------------------------------------------
template <bool kVec>
__global__ void __launch_bounds__(256)
raft_corr_volume_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int D, int HW, float div,
                        float* __restrict__ out) {
  __shared__ __align__(16) float a_s[16][64], b_s[16][64];
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const float* f1b = f1 + static_cast<size_t>(b) * D * HW;
  const float* f2b = f2 + static_cast<size_t>(b) * D * HW;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // loader mapping: thread -> (channel lk of the slab, 4 consecutive positions lr)
  const int lk = threadIdx.x >> 4, lr = (threadIdx.x & 15) * 4;
  float4 ra, rb;
  auto fetch = [&](int k0) {
    const int c = k0 + lk;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = ra;
    if (c < D) {
      const float* pa = f1b + static_cast<size_t>(c) * HW + i0 + lr;
      const float* pb = f2b + static_cast<size_t>(c) * HW + j0 + lr;
      if (kVec) {
        if (i0 + lr < HW) ra = __ldg(reinterpret_cast<const float4*>(pa));
        if (j0 + lr < HW) rb = __ldg(reinterpret_cast<const float4*>(pb));
      } else {
        float* va = reinterpret_cast<float*>(&ra);
        float* vb = reinterpret_cast<float*>(&rb);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (i0 + lr + u < HW) va[u] = __ldg(pa + u);
          if (j0 + lr + u < HW) vb[u] = __ldg(pb + u);
        }
      }
    }
  };
  float acc[4][4] = {};
  fetch(0);
  for (int k0 = 0; k0 < D; k0 += 16) {
    *reinterpret_cast<float4*>(&a_s[lk][lr]) = ra;
    *reinterpret_cast<float4*>(&b_s[lk][lr]) = rb;
    __syncthreads();
    if (k0 + 16 < D) fetch(k0 + 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&a_s[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&b_s[k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], bb[v], acc[u][v]);
    }
    __syncthreads();
  }
  float* ob = out + static_cast<size_t>(b) * HW * HW;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ty * 4 + u;
    if (i >= HW) continue;
    const int j = j0 + tx * 4;
    // corr / torch.sqrt(torch.tensor(dim).float())  (corr.py:60): div = sqrtf(D)
    float o[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) o[v] = __fdiv_rn(acc[u][v], div);
    float* dst = ob + static_cast<size_t>(i) * HW + j;
    if (kVec) {
      if (j < HW) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (j + v < HW) dst[v] = o[v];
    }
  }
}

// F.avg_pool2d(corr, 2, stride=2) (corr.py:26-27): row-major sum of the 2 x 2 window divided by 4, odd trailing
// row / column dropped.  One thread per output element; HBM-bound (reads 16 B, writes 4 B per element).
__device__ __forceinline__ float pyr_load(const float* p) { return __ldg(p); }
__device__ __forceinline__ float pyr_load(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ void pyr_store(float* p, float v) { *p = v; }
__device__ __forceinline__ void pyr_store(__half* p, float v) { *p = __float2half_rn(v); }

// T = float: the reference's fp32 pyramid; T = __half: the f16 pyramid of the mixed-precision path (half the bytes: the
// 64-sample level 0 then fits the L2 across the 24 lookups)
template <typename T>
__global__ void __launch_bounds__(256)
raft_corr_pool_kernel(const T* __restrict__ in, T* __restrict__ out, long long total, int hi, int wi, int ho, int wo) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int x = static_cast<int>(e % wo);
  const long long t = e / wo;
  const int y = static_cast<int>(t % ho);
  const long long p = t / ho;
  const T* src = in + (p * hi + 2 * y) * wi + 2 * x;
  const float s = ((pyr_load(src) + pyr_load(src + 1)) + pyr_load(src + wi)) + pyr_load(src + wi + 1);
  pyr_store(out + e, __fdiv_rn(s, 4.f));
}

// ---------------------------------------------------------------------------------------------
// pyramid lookup.  One CTA = 32 consecutive query pixels (flat over b, y, x) x all channels of the output, so the
// [B, L*(2r+1)^2, H, W] result is written as full 128-byte rows.  Work item = (pixel, level), one warp each:
//   * 2r+1 lanes restate the reference's x arithmetic, 2r+1 lanes the y arithmetic (it is separable: the sample
//     position of tap (a, b) is (cx + a - r, cy + b - r), corr.py:38-44 -- note the reference's meshgrid puts the
//     FIRST window axis on x);
//   * the warp stages the (2r+4)^2 source window that every tap of the item can touch in shared memory (zero-filled
//     outside the map = grid_sample's zero padding), 144 loads instead of 4*81; the window of the NEXT item is
//     already in flight (registers) while the taps of the current one are combined;
//   * every lane then combines 4 corners per tap in ATen's order (nw, ne, sw, se): 4 LDS + 4 FMUL + 4 FFMA.
// The radius is a template parameter for RAFT's two values (4: large, 3: small) so the window arithmetic is
// constant-folded; R = -1 is the generic fallback.  The kernel is issue-bound, not HBM-bound (ncu: DESIGN.md 8).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxLevels = 8;
constexpr int kMaxMapSide = 4096;  // beyond |coord| ~ 2^21 fp32 rounding exceeds a pixel; maps are far smaller
struct CorrLevels {
  const float* p[kMaxLevels];
  int h[kMaxLevels], w[kMaxLevels];
};

// One window offset along one axis -> index of the low corner RELATIVE to the staged window (w0 = its origin) and the
// two corner weights.  A corner pair that falls outside the staged window can only be a sample far outside the map
// (see kMaxMapSide), i.e. all zeros in the reference: it gets weight 0 at a harmless index.
__device__ __forceinline__ void raft_axis(float c, int tap_off, int size, int w0, int ws, int* rel, float* w_lo,
                                          float* w_hi) {
  // centroid_lvl + delta_lvl (corr.py:42-44)
  const float pos = __fadd_rn(c, static_cast<float>(tap_off));
  // bilinear_sampler: 2*x/(W-1) - 1 (utils.py:64-65)
  const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, pos), static_cast<float>(size - 1)), 1.f);
  // grid_sampler_unnormalize, align_corners=True: ((g + 1) / 2) * (size - 1)
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), static_cast<float>(size - 1));
  int r = 0;
  float lo = 0.f, hi = 0.f;
  if (fabsf(ix) < 1.0e8f) {  // false for NaN / inf / absurd coordinates: every corner out of bounds
    const float fl = floorf(ix);
    const int i0 = static_cast<int>(fl) - w0;
    if (i0 >= 0 && i0 <= ws - 2) {
      r = i0;
      lo = __fsub_rn(__fadd_rn(fl, 1.f), ix);  // (ix_se - ix): weight of the low corner
      hi = __fsub_rn(ix, fl);                  // (ix - ix_nw): weight of the high corner
    }
  }
  *rel = r;
  *w_lo = lo;
  *w_hi = hi;
}

template <int R>
__global__ void __launch_bounds__(256)
raft_corr_lookup_kernel(const __grid_constant__ CorrLevels lv, int L, int r_rt, const float* __restrict__ coords,
                        long long P, int HW, float* __restrict__ out, __half* __restrict__ out16, int ld16) {
  extern __shared__ __align__(16) float smem[];
  const int r = (R >= 0) ? R : r_rt;
  const int n1 = 2 * r + 1, n2 = n1 * n1, nch = L * n2;
  const int WS = 2 * r + 4, WW = WS * WS;
  constexpr int kWinRounds = (R >= 0) ? ((2 * R + 4) * (2 * R + 4) + 31) / 32 : 11;  // r <= 7: 18*18 / 32
  constexpr int kTapRounds = (R >= 0) ? ((2 * R + 1) * (2 * R + 1) + 31) / 32 : 8;   // r <= 7: 15*15 / 32
  float* tile = smem;  // [nch][33]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per warp: the staged window (WW floats, padded to 16 bytes), then 2 x 16 axis entries {rel index, w_lo, w_hi, -}
  const int per_warp = ((WW + 3) & ~3) + 128;
  float* win = smem + ((nch * 33 + 3) & ~3) + warp * per_warp;
  float4* ax = reinterpret_cast<float4*>(win + ((WW + 3) & ~3));
  const long long p0 = static_cast<long long>(blockIdx.x) * 32;

  // lane-constant index arithmetic, done once: the pixel this lane owns in the write phase (and whose centre it
  // broadcasts), the window elements it stages and the taps it combines
  const long long my_p = p0 + lane;
  const bool my_valid = my_p < P;
  const long long my_b = my_valid ? my_p / HW : 0;
  const int my_hw = static_cast<int>(my_p - my_b * HW);
  float my_cx = 0.f, my_cy = 0.f;
  if (my_valid) {
    my_cx = __ldg(coords + (my_b * 2) * HW + my_hw);
    my_cy = __ldg(coords + (my_b * 2 + 1) * HW + my_hw);
  }
  int win_yx[kWinRounds];  // (wy << 16) | wx; elements past the window get a row far outside any map
#pragma unroll
  for (int q = 0; q < kWinRounds; ++q) {
    const int e = lane + 32 * q;
    const int wy = e / WS;
    win_yx[q] = (e < WW) ? ((wy << 16) | (e - wy * WS)) : (0x4000 << 16);
  }
  int tap_ab[kTapRounds];  // (a << 8) | b, a: x offset index, b: y offset index; -1 past the last tap
#pragma unroll
  for (int t = 0; t < kTapRounds; ++t) {
    const int k = lane + 32 * t;
    const int a = k / n1;
    tap_ab[t] = (k < n2) ? ((a << 8) | (k - a * n1)) : -1;
  }

  // state of the item whose window is in flight: pixel warp + 8*f_q, level f_lvl
  float pre[kWinRounds];
  float f_cx = 0.f, f_cy = 0.f;
  int f_wx0 = 0, f_wy0 = 0, f_Hl = 2, f_Wl = 2, f_q = 0, f_lvl = 0;
  bool f_live = false;
  auto prefetch = [&]() {
    f_live = false;
    if (f_q >= 4) return;
    const int pix = warp + 8 * f_q;
    if (p0 + pix >= P) return;  // warp-uniform
    f_live = true;
    const float scale = 1.f / static_cast<float>(1 << f_lvl);  // coords / 2**i (exact)
    f_cx = __shfl_sync(0xffffffffu, my_cx, pix) * scale;
    f_cy = __shfl_sync(0xffffffffu, my_cy, pix) * scale;
    f_Hl = lv.h[f_lvl];
    f_Wl = lv.w[f_lvl];
    const float* src = lv.p[f_lvl] + (p0 + pix) * (static_cast<long long>(f_Hl) * f_Wl);
    // window origin; clamped so absurd coordinates cannot overflow the int conversion
    f_wx0 = static_cast<int>(floorf(fminf(fmaxf(f_cx, -1.0e6f), 1.0e6f))) - r - 1;
    f_wy0 = static_cast<int>(floorf(fminf(fmaxf(f_cy, -1.0e6f), 1.0e6f))) - r - 1;
#pragma unroll
    for (int q = 0; q < kWinRounds; ++q) {
      const int gy = f_wy0 + (win_yx[q] >> 16), gx = f_wx0 + (win_yx[q] & 0xffff);
      const bool in = static_cast<unsigned>(gy) < static_cast<unsigned>(f_Hl) &&
                      static_cast<unsigned>(gx) < static_cast<unsigned>(f_Wl);
      pre[q] = in ? __ldg(src + gy * f_Wl + gx) : 0.f;
    }
  };

  prefetch();
  for (int it = 0; it < 4 * L; ++it) {
    const bool cur = f_live;
    const int tile_off = f_lvl * n2 * 33 + warp + 8 * f_q;
    if (cur) {
#pragma unroll
      for (int q = 0; q < kWinRounds; ++q)
        if (lane + 32 * q < WW) win[lane + 32 * q] = pre[q];
      // lanes 0..2r: the x arithmetic of the 2r+1 offsets, lanes 16..16+2r: the y arithmetic (one pass, no divergence)
      const int sel = lane >> 4, t = lane & 15;
      if (t < n1) {
        int rel;
        float lo, hi;
        raft_axis(sel ? f_cy : f_cx, t - r, sel ? f_Hl : f_Wl, sel ? f_wy0 : f_wx0, WS, &rel, &lo, &hi);
        ax[lane] = make_float4(__int_as_float(sel ? rel * WS : rel), lo, hi, 0.f);
      }
    }
    __syncwarp();
    if (++f_lvl == L) {  // all levels of one pixel in a row, then the warp's next pixel
      f_lvl = 0;
      ++f_q;
    }
    prefetch();  // global loads of the next window overlap the taps below
    if (cur) {
#pragma unroll
      for (int t = 0; t < kTapRounds; ++t) {
        if (tap_ab[t] >= 0) {
          const float4 X = ax[tap_ab[t] >> 8], Y = ax[16 + (tap_ab[t] & 0xff)];
          const float* w = win + __float_as_int(Y.x) + __float_as_int(X.x);
          float acc = 0.f;
          acc = fmaf(w[0], __fmul_rn(X.y, Y.y), acc);       // nw
          acc = fmaf(w[1], __fmul_rn(X.z, Y.y), acc);       // ne
          acc = fmaf(w[WS], __fmul_rn(X.y, Y.z), acc);      // sw
          acc = fmaf(w[WS + 1], __fmul_rn(X.z, Y.z), acc);  // se
          tile[tile_off + (lane + 32 * t) * 33] = acc;
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (out16 != nullptr) {
    // pixel-major f16 rows [P, ld16] (channels past nch are zero): the layout the f16 recurrent block consumes
    for (int q = 0; q < 4; ++q) {
      const int pix = warp + 8 * q;
      if (p0 + pix >= P) break;
      __half2* dst = reinterpret_cast<__half2*>(out16 + (p0 + pix) * ld16);
      for (int c2 = lane; 2 * c2 < ld16; c2 += 32) {
        const int ch = 2 * c2;
        const float v0 = ch < nch ? tile[ch * 33 + pix] : 0.f;
        const float v1 = ch + 1 < nch ? tile[(ch + 1) * 33 + pix] : 0.f;
        dst[c2] = __floats2half2_rn(v0, v1);
      }
    }
  } else if (my_valid) {
    float* dst = out + my_b * static_cast<long long>(nch) * HW + my_hw;
    for (int ch = warp; ch < nch; ch += 8) dst[static_cast<long long>(ch) * HW] = tile[ch * 33 + lane];
  }
}

// ---------------------------------------------------------------------------------------------
// convex upsampling (raft_model.py:175-186): out[n, c, 8y+i, 8x+j] = sum_k softmax_k(mask[n, k*64 + i*8 + j, y, x]) *
// 8*flow[n, c, y + k/3 - 1, x + k%3 - 1] (zero padded).  One CTA per (n, y, 32-column slab, half of the 8 sub-rows):
// its 288 mask rows are streamed into shared memory with 16-byte cp.async (no register staging, all in flight at
// once; pitch 36 keeps the j-fastest reads conflict-free), every thread produces outputs with j fastest so each warp
// writes one 128-byte row segment.  HBM-bound: reads 576*4 B, writes 64*C*4 B per 1/8-resolution pixel.
// ---------------------------------------------------------------------------------------------
constexpr int kUpPitch = 36;
constexpr int kUpRows = 9 * 32;  // 9 neighbours x 4 sub-rows x 8 sub-columns
__global__ void __launch_bounds__(256)
raft_upsample_kernel(const float* __restrict__ flow, const float* __restrict__ mask, int C, int H, int W, int vec_ok,
                     float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* m_s = smem;                       // [288][36]: row k*32 + il*8 + j
  float* f_s = smem + kUpRows * kUpPitch;  // [C][3][34]
  const int n = blockIdx.z, y = blockIdx.y, x0 = (blockIdx.x >> 1) * 32, ih = blockIdx.x & 1;
  const int XW = min(32, W - x0);
  const size_t HWs = static_cast<size_t>(H) * W;
  const float* mrow = mask + (static_cast<size_t>(n) * 576) * HWs + static_cast<size_t>(y) * W + x0;
  if (vec_ok) {  // W % 4 == 0 and 16-byte aligned base: XW % 4 == 0 too
    const int cpr = XW >> 2, total = kUpRows * cpr;
    for (int c = threadIdx.x; c < total; c += 256) {
      const int row = c / cpr, col = c - row * cpr;
      const int k = row >> 5, rest = row & 31;
      cp_async_16(m_s + row * kUpPitch + col * 4, mrow + (k * 64 + ih * 32 + rest) * HWs + col * 4);
    }
    cp_async_commit();
  } else {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < XW)
      for (int row = warp; row < kUpRows; row += 8)
        m_s[row * kUpPitch + lane] = __ldg(mrow + ((row >> 5) * 64 + ih * 32 + (row & 31)) * HWs + lane);
  }
  for (int e = threadIdx.x; e < C * 3 * 34; e += 256) {
    const int xx = e % 34, t = e / 34, dy = t % 3, c = t / 3;
    const int gy = y + dy - 1, gx = x0 + xx - 1;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W)
      v = __fmul_rn(8.f, __ldg(flow + (static_cast<size_t>(n) * C + c) * HWs + static_cast<size_t>(gy) * W + gx));
    f_s[e] = v;
  }
  if (vec_ok) cp_async_wait_all();
  __syncthreads();
  const int per_row = 8 * XW;
  for (int t = threadIdx.x; t < 4 * per_row; t += 256) {
    const int il = t / per_row, rem = t - il * per_row;
    const int x = rem >> 3, j = rem & 7;
    float m[9];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      m[k] = m_s[(k * 32 + il * 8 + j) * kUpPitch + x];
      mx = fmaxf(mx, m[k]);
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      m[k] = expf(m[k] - mx);
      sum += m[k];
    }
    const float inv = __fdiv_rn(1.f, sum);
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] *= inv;
    const int i = ih * 4 + il;
    for (int c = 0; c < C; ++c) {
      const float* fc = f_s + c * 3 * 34;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) acc = fmaf(m[k], fc[(k / 3) * 34 + x + (k % 3)], acc);
      out[((static_cast<size_t>(n) * C + c) * (8 * H) + 8 * y + i) * (8 * static_cast<size_t>(W)) + 8 * (x0 + x) + j] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The elementwise half of RAFT-large's recurrent block (cwm/models/raft/update.py:33-60, :79-98, :115-139) for the
// mixed-precision path.  The convolutions stay cuDNN calls WITHOUT bias on f16 pixel-major (channels-last) rows; these
// kernels apply bias + activation to the raw convolution output and write it straight into its slot of the next
// convolution's input (the reference's torch.cat / bias add / relu / sigmoid / tanh / gate arithmetic: ~55 eager
// launches per iteration become 11).  All of them are HBM-bound streams of 16-byte vectors; math in fp32.
// ---------------------------------------------------------------------------------------------
union H8 {
  uint4 u;
  __half2 h2[4];
};
__device__ __forceinline__ void h8_to_f(const uint4& u, float* f) {
  H8 v;
  v.u = u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(v.h2[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 f_to_h8(const float* f) {
  H8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.h2[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v.u;
}
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// d1[m, c] (and d2[m, c]) = act(x[m, c] + bias[c]) for c < C; the last tail_cols columns are taken from tail[m, :]
// instead (BasicMotionEncoder's `cat([out, flow])`, update.py:98).  act: 0 none, 1 relu.
__global__ void __launch_bounds__(256)
raft_bias_act_kernel(const __half* __restrict__ x, int ldx, const float* __restrict__ bias, int act, int C, long long M,
                     __half* __restrict__ d1, int ld1, __half* __restrict__ d2, int ld2, const __half* __restrict__ tail,
                     int ldt, int tail_cols) {
  const int cv = C >> 3;
  const long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= M * cv) return;
  const long long row = v / cv;
  const int c8 = static_cast<int>(v - row * cv) << 3;
  float f[8];
  h8_to_f(*reinterpret_cast<const uint4*>(x + row * ldx + c8), f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[i] += bias ? __ldg(bias + c8 + i) : 0.f;
    if (act == 1) f[i] = fmaxf(f[i], 0.f);
    if (tail != nullptr && c8 + i >= C - tail_cols) f[i] = __half2float(tail[row * ldt + (c8 + i - (C - tail_cols))]);
  }
  const uint4 o = f_to_h8(f);
  *reinterpret_cast<uint4*>(d1 + row * ld1 + c8) = o;
  if (d2 != nullptr) *reinterpret_cast<uint4*>(d2 + row * ld2 + c8) = o;
}

// z = sigmoid(zr[:, :C] + b), r = sigmoid(zr[:, C:] + b) (update.py:46-47 / :53-54); z -> z_out, r * h -> rh
__global__ void __launch_bounds__(256)
raft_gru_gate_kernel(const __half* __restrict__ zr, const float* __restrict__ bias, const __half* __restrict__ h, int ldh,
                     int C, long long M, __half* __restrict__ z_out, __half* __restrict__ rh, int ldrh) {
  const int cv = C >> 3;
  const long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= M * cv) return;
  const long long row = v / cv;
  const int c8 = static_cast<int>(v - row * cv) << 3;
  float fz[8], fr[8], fh[8];
  h8_to_f(*reinterpret_cast<const uint4*>(zr + row * (2 * C) + c8), fz);
  h8_to_f(*reinterpret_cast<const uint4*>(zr + row * (2 * C) + C + c8), fr);
  h8_to_f(*reinterpret_cast<const uint4*>(h + row * ldh + c8), fh);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    fz[i] = sigmoidf_(fz[i] + __ldg(bias + c8 + i));
    fr[i] = sigmoidf_(fr[i] + __ldg(bias + C + c8 + i)) * fh[i];
  }
  *reinterpret_cast<uint4*>(z_out + row * C + c8) = f_to_h8(fz);
  *reinterpret_cast<uint4*>(rh + row * ldrh + c8) = f_to_h8(fr);
}

// h <- (1 - z) * h + z * tanh(q + b) (update.py:48-49 / :55-56), in place in its slot and optionally to a dense copy
__global__ void __launch_bounds__(256)
raft_gru_update_kernel(const __half* __restrict__ q, const float* __restrict__ bias, const __half* __restrict__ z,
                       __half* __restrict__ h, int ldh, int C, long long M, __half* __restrict__ h_dense) {
  const int cv = C >> 3;
  const long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= M * cv) return;
  const long long row = v / cv;
  const int c8 = static_cast<int>(v - row * cv) << 3;
  float fq[8], fz[8], fh[8];
  h8_to_f(*reinterpret_cast<const uint4*>(q + row * C + c8), fq);
  h8_to_f(*reinterpret_cast<const uint4*>(z + row * C + c8), fz);
  h8_to_f(*reinterpret_cast<const uint4*>(h + row * ldh + c8), fh);
#pragma unroll
  for (int i = 0; i < 8; ++i) fh[i] = (1.f - fz[i]) * fh[i] + fz[i] * tanhf(fq[i] + __ldg(bias + c8 + i));
  const uint4 o = f_to_h8(fh);
  *reinterpret_cast<uint4*>(h + row * ldh + c8) = o;
  if (h_dense != nullptr) *reinterpret_cast<uint4*>(h_dense + row * C + c8) = o;
}

// coords1 += delta + b (raft_model.py:254), flow = coords1 - coords0 as the next iteration's f16 input row
// [fx, fy, 0 x 6].  delta: raw flow-head output rows [M, ldd] (columns 0, 1), coords1: fp32 [B, 2, H, W].
__global__ void __launch_bounds__(256)
raft_flow_update_kernel(const __half* __restrict__ delta, int ldd, const float* __restrict__ bias, float* __restrict__ coords1,
                        int HW, int W, long long M, __half* __restrict__ flow16) {
  const long long m = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long b = m / HW;
  const int hw = static_cast<int>(m - b * HW);
  float* cx = coords1 + (b * 2) * HW + hw;
  float* cy = cx + HW;
  const float nx = *cx + (__half2float(delta[m * ldd]) + __ldg(bias));
  const float ny = *cy + (__half2float(delta[m * ldd + 1]) + __ldg(bias + 1));
  *cx = nx;
  *cy = ny;
  float f[8] = {nx - static_cast<float>(hw % W), ny - static_cast<float>(hw / W), 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  *reinterpret_cast<uint4*>(flow16 + m * 8) = f_to_h8(f);
}

// The same update with the flow head's last convolution (3x3, 256 -> 2; update.py:13-14) finished here: `taps` [M, ldt]
// holds, per pixel q, the 18 per-tap products  w[co, :, ky, kx] . fh1[q, :]  at column (ky * 3 + kx) * 2 + co (one 1x1
// GEMM with N = 18 instead of a 3x3 implicit GEMM whose 64-wide N tile would be 97 % padding), and
//   delta[p, co] = bias[co] + sum over (ky, kx) of taps[p + (ky - 1, kx - 1)][(ky * 3 + kx) * 2 + co]   (zero outside the image).
// The new flow is also written (2 f16) into up to two further row buffers: the flow slots of the GRU input rows.
__global__ void __launch_bounds__(256)
raft_flow_update_taps_kernel(const __half* __restrict__ taps, int ldt, const float* __restrict__ bias, float* __restrict__ coords1,
                             int H, int W, long long M, __half* __restrict__ flow16, __half* __restrict__ dst1, int ld1,
                             __half* __restrict__ dst2, int ld2) {
  grid_dep_sync();
  const long long m = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int HW = H * W;
  const long long b = m / HW;
  const int hw = static_cast<int>(m - b * HW);
  const int y = hw / W, x = hw - y * W;
  float dx = __ldg(bias), dy = __ldg(bias + 1);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y + ky - 1;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = x + kx - 1;
      if (xx < 0 || xx >= W) continue;
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(taps + (b * HW + yy * W + xx) * ldt + (ky * 3 + kx) * 2));
      dx += t.x;
      dy += t.y;
    }
  }
  float* cx = coords1 + (b * 2) * HW + hw;
  float* cy = cx + HW;
  const float nx = *cx + dx, ny = *cy + dy;
  *cx = nx;
  *cy = ny;
  float f[8] = {nx - static_cast<float>(x), ny - static_cast<float>(y), 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const uint4 o = f_to_h8(f);
  *reinterpret_cast<uint4*>(flow16 + m * 8) = o;
  if (dst1 != nullptr) *reinterpret_cast<uint32_t*>(dst1 + m * ld1) = o.x;
  if (dst2 != nullptr) *reinterpret_cast<uint32_t*>(dst2 + m * ld2) = o.x;
}

// ---- f16 lookup for the mixed-precision recurrent block (cwm_raft_corr_lookup_f16), radius 4, <= 4 levels ----
// All (2r+1)^2 taps of a level are the centre plus INTEGER offsets, so they share one pair of bilinear fractions
// (fx, fy): the 9x9 output window is the separable blend of a 10x10 source window,
//   T[i][j] = W[i][j] + fx (W[i][j+1] - W[i][j]),   out[i][j] = T[i][j] + fy (T[i+1][j] - T[i][j]),
// ~10 instructions per output instead of the ~120 of raft_corr_lookup_kernel, which restates the reference's normalise /
// unnormalise coordinate arithmetic operation by operation for the fp32 parity path (issue-bound: 1240 warp instructions per
// pixel).  The fractions here come from floor() directly; they differ from the reference's round trip by a few ulp, far
// below the f16 rounding of the output.  Zero outside the map (grid_sample's zero padding); a non-finite centre gives zeros.
// CTA = kFastPix pixels; phase 1: the 10x10 windows of every (pixel, level) are staged with coalesced loads (lanes run
// along window rows); phase 2: one thread per (pixel, level, window row) blends its 9 outputs from shared memory and puts
// them at channel l*81 + j*9 + i (x offset major, corr.py:37-47); phase 3: the rows leave as 16-byte segments.
constexpr int kFastPix = 16;
constexpr int kFastR = 4, kFastN1 = 9, kFastWin = 10, kFastPitch = 11;   // window rows padded to 11 floats: conflict-free
constexpr int kFastWarps = 9;
constexpr int kFastFlight = 4;   // windows a warp has in flight in phase 1
constexpr int kFastThreads = kFastWarps * 32;                              // 288 = half of the 16 x 4 x 9 blend tasks
constexpr int kFastItems = kFastPix * 4;                                   // (pixel, level) windows per CTA

struct FastItem {
  const void* src;    // the (pixel, level) map (fp32 or f16); nullptr: nothing to load (level >= L, pixel >= P)
  int gx0, gy0;       // map coordinates of window element (0, 0)
  int Hl, Wl;
  float fx, fy;
};

template <typename T>
__global__ void __launch_bounds__(kFastThreads)
raft_corr_lookup_fast_kernel(const __grid_constant__ CorrLevels lv, int L, const float* __restrict__ coords, long long P,
                             int HW, __half* __restrict__ out16, int ld16) {
  __shared__ float win[kFastItems * kFastWin * kFastPitch];     // [pix][lvl][10][11]
  __shared__ __align__(16) __half rows[kFastPix * 328];         // [pix][ld <= 328]
  __shared__ FastItem items[kFastItems];
  grid_dep_sync();   // (launched with programmatic stream serialization)
  const long long p0 = static_cast<long long>(blockIdx.x) * kFastPix;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < kFastItems) {
    const int pix = tid >> 2, lvl = tid & 3;
    const long long p = p0 + pix;
    FastItem it;
    it.src = nullptr;
    it.gx0 = it.gy0 = 0;
    it.Hl = it.Wl = 1;
    it.fx = it.fy = 0.f;
    if (p < P && lvl < L) {
      const long long b = p / HW;
      const int hw = static_cast<int>(p - b * HW);
      float cx = __ldg(coords + (b * 2) * HW + hw);
      float cy = __ldg(coords + (b * 2 + 1) * HW + hw);
      // non-finite or absurd centres: every tap is out of bounds
      if (!(fabsf(cx) < 1.0e6f) || !(fabsf(cy) < 1.0e6f)) cx = cy = -1.0e6f;
      const float sc = 1.f / static_cast<float>(1 << lvl);      // coords / 2**lvl (exact)
      const float x = cx * sc, y = cy * sc;
      const float flx = floorf(x), fly = floorf(y);
      it.Hl = lv.h[lvl];
      it.Wl = lv.w[lvl];
      it.src = reinterpret_cast<const T*>(lv.p[lvl]) + p * (static_cast<long long>(it.Hl) * it.Wl);
      it.gx0 = static_cast<int>(flx) - kFastR;
      it.gy0 = static_cast<int>(fly) - kFastR;
      it.fx = x - flx;
      it.fy = y - fly;
    }
    items[tid] = it;
  }
  // pad columns [L*81, ld16)
  for (int e = tid; e < kFastPix * (ld16 - L * 81); e += kFastThreads) {
    const int pix = e / (ld16 - L * 81);
    rows[pix * 328 + L * 81 + (e - pix * (ld16 - L * 81))] = __float2half_rn(0.f);
  }
  // lane-constant window positions of the four load rounds (elements lane, lane + 32, ... of the 10 x 10 window)
  int wyx[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e = lane + 32 * q;
    const int wy = e / kFastWin;
    wyx[q] = (e < kFastWin * kFastWin) ? ((wy << 8) | (e - wy * kFastWin)) : -1;
  }
  __syncthreads();
  // ---- phase 1: a warp stages whole windows (coalesced along the window rows), four windows in flight (the kernel is latency bound: 16 loads per lane outstanding) ----
  for (int i0 = warp; i0 < kFastItems; i0 += kFastFlight * kFastWarps) {
    float v[kFastFlight][4];
#pragma unroll
    for (int h = 0; h < kFastFlight; ++h) {
      const int item = i0 + h * kFastWarps;
#pragma unroll
      for (int q = 0; q < 4; ++q) v[h][q] = 0.f;
      if (item < kFastItems) {
        const FastItem it = items[item];
        if (it.src != nullptr) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int gy = it.gy0 + (wyx[q] >> 8), gx = it.gx0 + (wyx[q] & 0xff);
            if (wyx[q] >= 0 && static_cast<unsigned>(gx) < static_cast<unsigned>(it.Wl) &&
                static_cast<unsigned>(gy) < static_cast<unsigned>(it.Hl))
              v[h][q] = pyr_load(static_cast<const T*>(it.src) + gy * it.Wl + gx);
          }
        }
      }
    }
#pragma unroll
    for (int h = 0; h < kFastFlight; ++h) {
      const int item = i0 + h * kFastWarps;
      if (item < kFastItems) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (wyx[q] >= 0) win[item * (kFastWin * kFastPitch) + (wyx[q] >> 8) * kFastPitch + (wyx[q] & 0xff)] = v[h][q];
      }
    }
  }
  __syncthreads();
  // ---- phase 2: separable blend, one task = (pixel, level, window row): 576 tasks = 2 rounds of 288 threads ----
#pragma unroll
  for (int round = 0; round < (kFastItems * kFastN1) / kFastThreads; ++round) {
    const int task = tid + round * kFastThreads;
    const int item = task / kFastN1;
    const int i = task - item * kFastN1;
    const int pix = item >> 2, lvl = item & 3;
    if (lvl < L) {
      const float fx = items[item].fx, fy = items[item].fy;
      const float* w0 = win + item * (kFastWin * kFastPitch) + i * kFastPitch;
      float t0[kFastN1], t1[kFastN1];
      float a = w0[0], b = w0[kFastPitch];
#pragma unroll
      for (int j = 0; j < kFastN1; ++j) {
        const float a1 = w0[j + 1], b1 = w0[kFastPitch + j + 1];
        t0[j] = fmaf(fx, a1 - a, a);
        t1[j] = fmaf(fx, b1 - b, b);
        a = a1;
        b = b1;
      }
      __half* dst = rows + pix * 328 + lvl * (kFastN1 * kFastN1) + i;
#pragma unroll
      for (int j = 0; j < kFastN1; ++j) dst[j * kFastN1] = __float2half_rn(fmaf(fy, t1[j] - t0[j], t0[j]));
    }
  }
  __syncthreads();
  // ---- phase 3: rows out ----
  const int units = ld16 >> 3;   // ld16 % 8 == 0
  for (int e = tid; e < kFastPix * units; e += kFastThreads) {
    const int pix = e / units, u = e - pix * units;
    if (p0 + pix < P)
      *reinterpret_cast<uint4*>(out16 + (p0 + pix) * ld16 + u * 8) = *reinterpret_cast<const uint4*>(rows + pix * 328 + u * 8);
  }
}

static int level_dims(int H, int W, int L, int* hs, int* ws, const char* who) {
  CWM_REQUIRE(L >= 1 && L <= kMaxLevels, "%s: num_levels %d not in [1, %d]", who, L, kMaxLevels);
  hs[0] = H;
  ws[0] = W;
  for (int l = 1; l < L; ++l) {
    hs[l] = hs[l - 1] / 2;
    ws[l] = ws[l - 1] / 2;
  }
  // the reference divides by (size - 1) of every level (utils.py:64-65): a 1-pixel level would produce NaN there
  CWM_REQUIRE(hs[L - 1] >= 2 && ws[L - 1] >= 2, "%s: level %d of a (%d,%d) map is smaller than 2x2", who, L - 1, H, W);
  return CWM_OK;
}

}  // namespace cwm

using namespace cwm;

extern "C" int cwm_raft_corr_pyramid(const float* fmap1, const float* fmap2, int B, int D, int H, int W,
                                     int num_levels, float* const* levels, cwm_stream_t stream) {
  CWM_REQUIRE(B >= 0 && D >= 1 && H >= 1 && W >= 1, "cwm_raft_corr_pyramid: bad shape B=%d D=%d H=%d W=%d", B, D, H, W);
  int hs[kMaxLevels], ws[kMaxLevels];
  int rc = level_dims(H, W, num_levels, hs, ws, "cwm_raft_corr_pyramid");
  if (rc != CWM_OK) return rc;
  CWM_REQUIRE(levels, "cwm_raft_corr_pyramid: null level table");
  if (B == 0) return CWM_OK;
  CWM_REQUIRE(fmap1 && fmap2, "cwm_raft_corr_pyramid: null feature map");
  for (int l = 0; l < num_levels; ++l) CWM_REQUIRE(levels[l], "cwm_raft_corr_pyramid: null level %d", l);
  const int HW = H * W;
  CWM_REQUIRE(B <= 65535, "cwm_raft_corr_pyramid: batch %d > 65535 (chunk the sweep)", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float div = sqrtf(static_cast<float>(D));
  {
    ProfileScope prof(st, "raft_corr_volume", 2.0 * B * static_cast<double>(HW) * HW * D,
                      static_cast<double>(B) * HW * (static_cast<double>(HW) + 2.0 * D) * 4.0);
    const dim3 grid((HW + 63) / 64, (HW + 63) / 64, B);
    const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(fmap1) | reinterpret_cast<uintptr_t>(fmap2) |
                                        reinterpret_cast<uintptr_t>(levels[0])) % 16 == 0);
    if (vec)
      raft_corr_volume_kernel<true><<<grid, 256, 0, st>>>(fmap1, fmap2, D, HW, div, levels[0]);
    else
      raft_corr_volume_kernel<false><<<grid, 256, 0, st>>>(fmap1, fmap2, D, HW, div, levels[0]);
    CWM_LAUNCH_CHECK();
  }
  for (int l = 1; l < num_levels; ++l) {
    const long long total = static_cast<long long>(B) * HW * hs[l] * ws[l];
    ProfileScope prof(st, "raft_corr_pool", 0.0, static_cast<double>(total) * 20.0);
    raft_corr_pool_kernel<float><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(levels[l - 1], levels[l], total,
                                                                                      hs[l - 1], ws[l - 1], hs[l], ws[l]);
    CWM_LAUNCH_CHECK();
  }
  return CWM_OK;
}

// The same pyramid with level 0 on the tensor cores (raftcorr_tc.cu: 3xTF32 split, fp32 accuracy); D % 32 == 0.
extern "C" int cwm_raft_corr_volume_tc(const float* fmap1, const float* fmap2, int B, int D, int H, int W, float* out,
                                       void* workspace, size_t workspace_bytes, cwm_stream_t stream);

extern "C" int cwm_raft_corr_pyramid_tc(const float* fmap1, const float* fmap2, int B, int D, int H, int W, int num_levels,
                                        float* const* levels, void* workspace, size_t workspace_bytes, cwm_stream_t stream) {
  CWM_REQUIRE(B >= 0 && D >= 1 && H >= 1 && W >= 1, "cwm_raft_corr_pyramid_tc: bad shape B=%d D=%d H=%d W=%d", B, D, H, W);
  int hs[kMaxLevels], ws[kMaxLevels];
  int rc = level_dims(H, W, num_levels, hs, ws, "cwm_raft_corr_pyramid_tc");
  if (rc != CWM_OK) return rc;
  CWM_REQUIRE(levels, "cwm_raft_corr_pyramid_tc: null level table");
  if (B == 0) return CWM_OK;
  for (int l = 0; l < num_levels; ++l) CWM_REQUIRE(levels[l], "cwm_raft_corr_pyramid_tc: null level %d", l);
  rc = cwm_raft_corr_volume_tc(fmap1, fmap2, B, D, H, W, levels[0], workspace, workspace_bytes, stream);
  if (rc != CWM_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int HW = H * W;
  for (int l = 1; l < num_levels; ++l) {
    const long long total = static_cast<long long>(B) * HW * hs[l] * ws[l];
    ProfileScope prof(st, "raft_corr_pool", 0.0, static_cast<double>(total) * 20.0);
    raft_corr_pool_kernel<float><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(levels[l - 1], levels[l], total,
                                                                                      hs[l - 1], ws[l - 1], hs[l], ws[l]);
    CWM_LAUNCH_CHECK();
  }
  return CWM_OK;
}

extern "C" int cwm_raft_corr_volume_rows_f16(const uint16_t* rows1, int n1, const uint16_t* rows2, int B, int D, int H, int W,
                                             float* out, cwm_stream_t stream);
extern "C" int cwm_raft_corr_volume_rows_f16_out16(const uint16_t* rows1, int n1, const uint16_t* rows2, int B, int D, int H,
                                                   int W, uint16_t* out16, cwm_stream_t stream);

// the pyramid in f16 (every level): half the bytes of the fp32 one -- the 64-sample level 0 (78 MB) then stays in the
// 126 MB L2 across the 24 lookups of a flow call; read by cwm_raft_corr_lookup_f16_pyr16
extern "C" int cwm_raft_corr_pyramid_rows_f16_pyr16(const uint16_t* rows1, int n1, const uint16_t* rows2, int B, int D, int H,
                                                    int W, int num_levels, uint16_t* const* levels, cwm_stream_t stream) {
  CWM_REQUIRE(B >= 0 && D >= 1 && H >= 1 && W >= 1, "cwm_raft_corr_pyramid_rows_f16_pyr16: bad shape B=%d D=%d H=%d W=%d", B, D, H, W);
  int hs[kMaxLevels], ws[kMaxLevels];
  int rc = level_dims(H, W, num_levels, hs, ws, "cwm_raft_corr_pyramid_rows_f16_pyr16");
  if (rc != CWM_OK) return rc;
  CWM_REQUIRE(levels, "cwm_raft_corr_pyramid_rows_f16_pyr16: null level table");
  if (B == 0) return CWM_OK;
  for (int l = 0; l < num_levels; ++l) CWM_REQUIRE(levels[l], "cwm_raft_corr_pyramid_rows_f16_pyr16: null level %d", l);
  rc = cwm_raft_corr_volume_rows_f16_out16(rows1, n1, rows2, B, D, H, W, levels[0], stream);
  if (rc != CWM_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int HW = H * W;
  for (int l = 1; l < num_levels; ++l) {
    const long long total = static_cast<long long>(B) * HW * hs[l] * ws[l];
    ProfileScope prof(st, "raft_corr_pool", 0.0, static_cast<double>(total) * 10.0);
    raft_corr_pool_kernel<__half><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const __half*>(levels[l - 1]), reinterpret_cast<__half*>(levels[l]), total, hs[l - 1], ws[l - 1], hs[l], ws[l]);
    CWM_LAUNCH_CHECK();
  }
  return CWM_OK;
}

extern "C" int cwm_raft_corr_pyramid_rows_f16(const uint16_t* rows1, int n1, const uint16_t* rows2, int B, int D, int H, int W,
                                              int num_levels, float* const* levels, cwm_stream_t stream) {
  CWM_REQUIRE(B >= 0 && D >= 1 && H >= 1 && W >= 1, "cwm_raft_corr_pyramid_rows_f16: bad shape B=%d D=%d H=%d W=%d", B, D, H, W);
  int hs[kMaxLevels], ws[kMaxLevels];
  int rc = level_dims(H, W, num_levels, hs, ws, "cwm_raft_corr_pyramid_rows_f16");
  if (rc != CWM_OK) return rc;
  CWM_REQUIRE(levels, "cwm_raft_corr_pyramid_rows_f16: null level table");
  if (B == 0) return CWM_OK;
  for (int l = 0; l < num_levels; ++l) CWM_REQUIRE(levels[l], "cwm_raft_corr_pyramid_rows_f16: null level %d", l);
  rc = cwm_raft_corr_volume_rows_f16(rows1, n1, rows2, B, D, H, W, levels[0], stream);
  if (rc != CWM_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int HW = H * W;
  for (int l = 1; l < num_levels; ++l) {
    const long long total = static_cast<long long>(B) * HW * hs[l] * ws[l];
    ProfileScope prof(st, "raft_corr_pool", 0.0, static_cast<double>(total) * 20.0);
    raft_corr_pool_kernel<float><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(levels[l - 1], levels[l], total,
                                                                                      hs[l - 1], ws[l - 1], hs[l], ws[l]);
    CWM_LAUNCH_CHECK();
  }
  return CWM_OK;
}

static int corr_lookup_impl(const float* const* levels, int num_levels, int radius, const float* coords, int B, int H, int W,
                            float* out, __half* out16, int ld16, cwm_stream_t stream, bool src_f16 = false) {
  CWM_REQUIRE(B >= 0 && H >= 1 && W >= 1, "cwm_raft_corr_lookup: bad shape B=%d H=%d W=%d", B, H, W);
  CWM_REQUIRE(radius >= 0 && radius <= 7, "cwm_raft_corr_lookup: radius %d not in [0, 7]", radius);
  CorrLevels lv;
  int rc = level_dims(H, W, num_levels, lv.h, lv.w, "cwm_raft_corr_lookup");
  if (rc != CWM_OK) return rc;
  CWM_REQUIRE(levels, "cwm_raft_corr_lookup: null level table");
  if (B == 0) return CWM_OK;
  CWM_REQUIRE(coords && (out || out16), "cwm_raft_corr_lookup: null pointer");
  for (int l = 0; l < num_levels; ++l) {
    CWM_REQUIRE(levels[l], "cwm_raft_corr_lookup: null level %d", l);
    lv.p[l] = levels[l];
  }
  const int n1 = 2 * radius + 1, nch = num_levels * n1 * n1, WS = 2 * radius + 4;
  const size_t smem = (static_cast<size_t>((nch * 33 + 3) & ~3) + 8 * (((WS * WS + 3) & ~3) + 128)) * sizeof(float);
  CWM_REQUIRE(smem <= 200 * 1024, "cwm_raft_corr_lookup: %d levels x radius %d needs %zu bytes of shared memory", num_levels,
              radius, smem);
  CWM_REQUIRE(H <= kMaxMapSide && W <= kMaxMapSide, "cwm_raft_corr_lookup: map (%d,%d) larger than %d", H, W, kMaxMapSide);
  cudaStream_t st_fast = static_cast<cudaStream_t>(stream);
  static int fast_env = -1;
  if (fast_env < 0) {
    const char* e = getenv("CWM_RAFT_LOOKUP");
    fast_env = (e != nullptr && (e[0] == 'e' || e[0] == '0')) ? 0 : 1;   // CWM_RAFT_LOOKUP=exact: the reference-arithmetic kernel
  }
  if (out16 != nullptr && out == nullptr && fast_env && radius == kFastR && num_levels <= 4 && ld16 % 8 == 0 && ld16 <= 328 &&
      ld16 >= num_levels * 81 && reinterpret_cast<uintptr_t>(out16) % 16 == 0) {
    const long long Pf = static_cast<long long>(B) * H * W;
    double pyr_f = 0.0;
    for (int l = 0; l < num_levels; ++l) pyr_f += static_cast<double>(min(kFastWin, lv.w[l])) * min(kFastWin, lv.h[l]);
    ProfileScope prof(st_fast, "raft_corr_lookup", 0.0, static_cast<double>(Pf) * (0.5 * ld16 + pyr_f + 2.0) * 4.0);
    if (src_f16)
      CWM_CUDA_CHECK(launch_pdl(raft_corr_lookup_fast_kernel<__half>, dim3(static_cast<unsigned>((Pf + kFastPix - 1) / kFastPix)),
                                dim3(kFastThreads), 0, st_fast, lv, num_levels, coords, Pf, H * W, out16, ld16));
    else
      CWM_CUDA_CHECK(launch_pdl(raft_corr_lookup_fast_kernel<float>, dim3(static_cast<unsigned>((Pf + kFastPix - 1) / kFastPix)),
                                dim3(kFastThreads), 0, st_fast, lv, num_levels, coords, Pf, H * W, out16, ld16));
    CWM_LAUNCH_CHECK();
    return CWM_OK;
  }
  CWM_REQUIRE(!src_f16, "cwm_raft_corr_lookup_f16_pyr16: the f16 pyramid is read by the fast lookup only (radius 4, <= 4 levels, "
                        "row length a multiple of 8 and <= 328)");
  auto kernel = radius == 4 ? raft_corr_lookup_kernel<4> : radius == 3 ? raft_corr_lookup_kernel<3> : raft_corr_lookup_kernel<-1>;
  static size_t configured[3] = {0, 0, 0};  // grow-only opt-in for > 48 KB of dynamic shared memory, per instantiation
  size_t& conf = configured[radius == 4 ? 0 : radius == 3 ? 1 : 2];
  if (smem > conf) {
    CWM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    conf = smem;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long P = static_cast<long long>(B) * H * W;
  double pyr = 0.0;
  for (int l = 0; l < num_levels; ++l) {
    const int ws = min(WS, lv.w[l]), hs = min(WS, lv.h[l]);
    pyr += static_cast<double>(ws) * hs;
  }
  if (out16 != nullptr)
    CWM_REQUIRE(ld16 >= nch && ld16 % 2 == 0 && reinterpret_cast<uintptr_t>(out16) % 4 == 0,
                "cwm_raft_corr_lookup_f16: row length %d must be even and >= %d channels", ld16, nch);
  ProfileScope prof(st, "raft_corr_lookup", 0.0,
                    static_cast<double>(P) * ((out16 ? 0.5 * ld16 : static_cast<double>(nch)) + pyr + 2.0) * 4.0);
  kernel<<<static_cast<unsigned>((P + 31) / 32), 256, smem, st>>>(lv, num_levels, radius, coords, P, H * W, out, out16, ld16);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_raft_corr_lookup(const float* const* levels, int num_levels, int radius, const float* coords, int B,
                                    int H, int W, float* out, cwm_stream_t stream) {
  return corr_lookup_impl(levels, num_levels, radius, coords, B, H, W, out, nullptr, 0, stream);
}

extern "C" int cwm_raft_corr_lookup_f16(const float* const* levels, int num_levels, int radius, const float* coords, int B,
                                        int H, int W, uint16_t* out16, int ld16, cwm_stream_t stream) {
  return corr_lookup_impl(levels, num_levels, radius, coords, B, H, W, nullptr, reinterpret_cast<__half*>(out16), ld16, stream);
}

extern "C" int cwm_raft_corr_lookup_f16_pyr16(const uint16_t* const* levels, int num_levels, int radius, const float* coords,
                                              int B, int H, int W, uint16_t* out16, int ld16, cwm_stream_t stream) {
  return corr_lookup_impl(reinterpret_cast<const float* const*>(levels), num_levels, radius, coords, B, H, W, nullptr,
                          reinterpret_cast<__half*>(out16), ld16, stream, true);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static unsigned blocks_for(long long n) { return static_cast<unsigned>((n + 255) / 256); }

extern "C" int cwm_raft_bias_act_f16(const uint16_t* x, int ldx, const float* bias, int act, int C, long long M, uint16_t* d1,
                                     int ld1, uint16_t* d2, int ld2, const uint16_t* tail, int ldt, int tail_cols,
                                     cwm_stream_t stream) {
  CWM_REQUIRE(M >= 0 && C >= 8 && C % 8 == 0 && (act == 0 || act == 1), "cwm_raft_bias_act_f16: bad C=%d / act=%d", C, act);
  if (M == 0) return CWM_OK;
  CWM_REQUIRE(x && d1 && ldx >= C && ld1 >= C && ldx % 8 == 0 && ld1 % 8 == 0 && aligned16(x) && aligned16(d1),
              "cwm_raft_bias_act_f16: rows must be 16-byte aligned and at least C wide");
  CWM_REQUIRE(d2 == nullptr || (ld2 >= C && ld2 % 8 == 0 && aligned16(d2)), "cwm_raft_bias_act_f16: bad second destination");
  CWM_REQUIRE(tail == nullptr || (tail_cols >= 1 && tail_cols <= C && ldt >= tail_cols), "cwm_raft_bias_act_f16: bad tail");
  CWM_REQUIRE(M * (C / 8) < (1LL << 39), "cwm_raft_bias_act_f16: too many rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "raft_bias_act", 0.0, static_cast<double>(M) * C * (d2 ? 6.0 : 4.0));
  raft_bias_act_kernel<<<blocks_for(M * (C / 8)), 256, 0, st>>>(
      reinterpret_cast<const __half*>(x), ldx, bias, act, C, M, reinterpret_cast<__half*>(d1), ld1,
      reinterpret_cast<__half*>(d2), ld2, reinterpret_cast<const __half*>(tail), ldt, tail ? tail_cols : 0);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_raft_gru_gate_f16(const uint16_t* zr, const float* bias, const uint16_t* h, int ldh, int C, long long M,
                                     uint16_t* z_out, uint16_t* rh, int ldrh, cwm_stream_t stream) {
  CWM_REQUIRE(M >= 0 && C >= 8 && C % 8 == 0, "cwm_raft_gru_gate_f16: bad C=%d", C);
  if (M == 0) return CWM_OK;
  CWM_REQUIRE(zr && bias && h && z_out && rh && ldh >= C && ldrh >= C && ldh % 8 == 0 && ldrh % 8 == 0 && aligned16(zr) &&
                  aligned16(h) && aligned16(z_out) && aligned16(rh),
              "cwm_raft_gru_gate_f16: null pointer or rows not 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "raft_gru_gate", 0.0, static_cast<double>(M) * C * 10.0);
  raft_gru_gate_kernel<<<blocks_for(M * (C / 8)), 256, 0, st>>>(reinterpret_cast<const __half*>(zr), bias,
                                                               reinterpret_cast<const __half*>(h), ldh, C, M,
                                                               reinterpret_cast<__half*>(z_out),
                                                               reinterpret_cast<__half*>(rh), ldrh);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_raft_gru_update_f16(const uint16_t* q, const float* bias, const uint16_t* z, uint16_t* h, int ldh, int C,
                                       long long M, uint16_t* h_dense, cwm_stream_t stream) {
  CWM_REQUIRE(M >= 0 && C >= 8 && C % 8 == 0, "cwm_raft_gru_update_f16: bad C=%d", C);
  if (M == 0) return CWM_OK;
  CWM_REQUIRE(q && bias && z && h && ldh >= C && ldh % 8 == 0 && aligned16(q) && aligned16(z) && aligned16(h) &&
                  (h_dense == nullptr || aligned16(h_dense)),
              "cwm_raft_gru_update_f16: null pointer or rows not 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "raft_gru_update", 0.0, static_cast<double>(M) * C * (h_dense ? 10.0 : 8.0));
  raft_gru_update_kernel<<<blocks_for(M * (C / 8)), 256, 0, st>>>(reinterpret_cast<const __half*>(q), bias,
                                                                 reinterpret_cast<const __half*>(z),
                                                                 reinterpret_cast<__half*>(h), ldh, C, M,
                                                                 reinterpret_cast<__half*>(h_dense));
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_raft_flow_update(const uint16_t* delta, int ldd, const float* bias, float* coords1, int B, int H, int W,
                                    uint16_t* flow16, cwm_stream_t stream) {
  CWM_REQUIRE(B >= 0 && H >= 1 && W >= 1 && ldd >= 2, "cwm_raft_flow_update: bad shape B=%d H=%d W=%d ld=%d", B, H, W, ldd);
  if (B == 0) return CWM_OK;
  CWM_REQUIRE(delta && bias && coords1 && flow16 && aligned16(flow16), "cwm_raft_flow_update: null or misaligned pointer");
  const long long M = static_cast<long long>(B) * H * W;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "raft_flow_update", 0.0, static_cast<double>(M) * 36.0);
  raft_flow_update_kernel<<<blocks_for(M), 256, 0, st>>>(reinterpret_cast<const __half*>(delta), ldd, bias, coords1, H * W, W, M,
                                                        reinterpret_cast<__half*>(flow16));
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_raft_flow_update_taps(const uint16_t* taps, int ldt, const float* bias, float* coords1, int B, int H, int W,
                                         uint16_t* flow16, uint16_t* dst1, int ld1, uint16_t* dst2, int ld2,
                                         cwm_stream_t stream) {
  CWM_REQUIRE(B >= 0 && H >= 1 && W >= 1 && ldt >= 18 && ldt % 2 == 0 && ld1 % 2 == 0 && ld2 % 2 == 0,
              "cwm_raft_flow_update_taps: bad shape B=%d H=%d W=%d ldt=%d ld1=%d ld2=%d", B, H, W, ldt, ld1, ld2);
  if (B == 0) return CWM_OK;
  CWM_REQUIRE(taps && bias && coords1 && flow16 && aligned16(flow16) && (reinterpret_cast<uintptr_t>(taps) & 3) == 0 &&
                  (reinterpret_cast<uintptr_t>(dst1) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst2) & 3) == 0,
              "cwm_raft_flow_update_taps: null or misaligned pointer");
  const long long M = static_cast<long long>(B) * H * W;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "raft_flow_update_taps", 0.0, static_cast<double>(M) * (36.0 + 36.0 + 8.0));
  CWM_CUDA_CHECK(launch_pdl(raft_flow_update_taps_kernel, dim3(blocks_for(M)), dim3(256), 0, st,
                            reinterpret_cast<const __half*>(taps), ldt, bias, coords1, H, W, M, reinterpret_cast<__half*>(flow16),
                            reinterpret_cast<__half*>(dst1), ld1, reinterpret_cast<__half*>(dst2), ld2));
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

// im2col of the 2-channel flow field for BasicMotionEncoder.convf1 (7x7, padding 3; update.py:85): row m of `out` holds
// the 49 taps x 2 channels of pixel m's neighbourhood in (ky, kx, c) order (zero outside the image), padded to `ldo`
// columns -- a K = 98 convolution becomes one plain K = 128 GEMM instead of 49 nearly empty 64-channel k-steps.
namespace cwm {
// K > 0: compile-time kernel size (the index arithmetic is divisions by k and by the map size; with run-time divisors the
// kernel was instruction bound at 15 us for 12.8 MB)
template <int K>
__global__ void __launch_bounds__(256)
raft_im2col_flow_kernel(const __half* __restrict__ flow16, int ldf, int H, int W, long long M, int k_rt,
                        __half* __restrict__ out, int ldo) {
  grid_dep_sync();
  const int k = K > 0 ? K : k_rt;
  const int units = ldo / 8;   // one thread writes four taps x two channels = one 16-byte store
  // grid.y = image row (sample, y): no division by the map size; threadIdx / blockIdx.x walk (x, unit)
  const long long row = blockIdx.y;                 // (sample * H + y)
  const int y = static_cast<int>(row % H);
  const int xu = blockIdx.x * blockDim.x + threadIdx.x;
  if (xu >= W * units) return;
  const int x = xu / units, u = xu - x * units;
  const int tap0 = u * 4;
  const long long m = row * W + x;
  const __half* img = flow16 + (row - y) * W * ldf;      // first pixel of this sample
  const int py = y - k / 2, px = x - k / 2;
  uint32_t v[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int tap = tap0 + t;
    v[t] = 0u;
    if (tap < k * k) {
      const int ty = tap / k;
      const int yy = py + ty, xx = px + tap - ty * k;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v[t] = *reinterpret_cast<const uint32_t*>(img + (static_cast<long long>(yy) * W + xx) * ldf);
    }
  }
  *reinterpret_cast<uint4*>(out + m * ldo + 2 * tap0) = make_uint4(v[0], v[1], v[2], v[3]);
}
}  // namespace cwm

extern "C" int cwm_raft_im2col_flow(const uint16_t* flow16, int ldf, int B, int H, int W, int k, uint16_t* out, int ldo,
                                    cwm_stream_t stream) {
  CWM_REQUIRE(B >= 0 && H >= 1 && W >= 1 && k >= 1 && (k & 1) && ldf >= 2 && ldf % 2 == 0 && ldo % 8 == 0 && ldo >= 2 * k * k,
              "cwm_raft_im2col_flow: bad shape B=%d H=%d W=%d k=%d ldf=%d ldo=%d", B, H, W, k, ldf, ldo);
  if (B == 0) return CWM_OK;
  CWM_REQUIRE(flow16 && out && aligned16(out), "cwm_raft_im2col_flow: null or misaligned pointer");
  const long long M = static_cast<long long>(B) * H * W;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "raft_im2col_flow", 0.0, static_cast<double>(M) * (ldo * 2.0 + 4.0 * k * k));
  CWM_REQUIRE(static_cast<long long>(B) * H <= 65535, "cwm_raft_im2col_flow: B * H = %lld exceeds the grid limit", static_cast<long long>(B) * H);
  const dim3 grid((W * (ldo / 8) + 255) / 256, B * H);
  if (k == 7)
    CWM_CUDA_CHECK(cwm::launch_pdl(cwm::raft_im2col_flow_kernel<7>, grid, dim3(256), 0, st, reinterpret_cast<const __half*>(flow16),
                                   ldf, H, W, M, k, reinterpret_cast<__half*>(out), ldo));
  else
    CWM_CUDA_CHECK(cwm::launch_pdl(cwm::raft_im2col_flow_kernel<0>, grid, dim3(256), 0, st, reinterpret_cast<const __half*>(flow16),
                                   ldf, H, W, M, k, reinterpret_cast<__half*>(out), ldo));
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_raft_upsample_flow(const float* flow, const float* mask, int N, int C, int H, int W, float* out,
                                      cwm_stream_t stream) {
  CWM_REQUIRE(N >= 0 && C >= 1 && C <= 16 && H >= 1 && W >= 1, "cwm_raft_upsample_flow: bad shape N=%d C=%d H=%d W=%d", N, C,
              H, W);
  if (N == 0) return CWM_OK;
  CWM_REQUIRE(flow && mask && out, "cwm_raft_upsample_flow: null pointer");
  CWM_REQUIRE(N <= 65535 && H <= 65535, "cwm_raft_upsample_flow: N=%d / H=%d exceed the grid limits", N, H);
  const size_t smem = (kUpRows * kUpPitch + static_cast<size_t>(C) * 3 * 34) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    CWM_CUDA_CHECK(cudaFuncSetAttribute(raft_upsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    configured = smem;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vec_ok = (W % 4 == 0) && (reinterpret_cast<uintptr_t>(mask) % 16 == 0);
  ProfileScope prof(st, "raft_upsample", 0.0, static_cast<double>(N) * H * W * (576.0 + C + 64.0 * C) * 4.0);
  raft_upsample_kernel<<<dim3(2 * ((W + 31) / 32), H, N), 256, smem, st>>>(flow, mask, C, H, W, vec_ok, out);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}
