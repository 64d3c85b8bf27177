// masks.cu -- SURVEY 8(f) rank 4: mask generation, energy sampling of patches and mask rectangularisation ON DEVICE with a
// counter-based RNG, so that the masks of a sweep depend only on (seed, global sample index) -- not on the batch split,
// the number of GPUs or the order in which samples are generated.
//
// What it stands in for (cwm/models, opt-in: the reference draws from `torch.randperm` / `Categorical.sample` on the
// global generator, whose streams cannot be reproduced by a parallel generator -- `masking.py` of this package remains
// the bit-exact mirror of those):
//   * MaskingGenerator.sample_mask_per_frame (masking.py:347-376): k of the (h/c)*(w/c) clump cells of every masked
//     frame are visible, uniformly without replacement; RotatedTableUniformMaskingGenerator (:478-545) prepends the fully
//     visible frames.
//   * EnergySamplingMaskingGenerator.sample_mask_per_frame (sampling.py:63-90) + sample_from_energy /
//     sample_image_inds_from_probs (utils.py:152-213): P cells per sample drawn WITH replacement from the categorical
//     distribution relu(p - min p + eps) / sum.
//   * RectangularizeMasks('min') (masking.py:100-132): rows with more masked tokens than the minimum get that many
//     randomly chosen masked tokens revealed.
//
// RNG: Philox4x32-10 (Salmon et al., SC'11; Random123), key = the 64-bit seed, counter = (draw index, global sample
// index, stream id, sub-stream).  All selections are integer: uniform subsets are the k smallest of the 64-bit keys
// (random word << 32 | cell), categorical draws invert an INTEGER cumulative table (probabilities quantised to 24 bits
// of the largest one), so the numpy restatement in oracle/device_masks_oracle.py is bit-exact.
#include "common.cuh"

namespace cwm {

enum { kStreamUniform = 0, kStreamEnergy = 1, kStreamRect = 2 };

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t out[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(M0) * c0;
    const uint64_t p1 = static_cast<uint64_t>(M1) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0, c1 = n1, c2 = n2, c3 = n3;
    k0 += W0, k1 += W1;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

// word `i` of the stream (sample, stream id, sub): draws come in blocks of four
__device__ __forceinline__ uint32_t philox_word(uint64_t seed, uint32_t sample, uint32_t stream, uint32_t sub, uint32_t i) {
  uint32_t o[4];
  philox4x32_10(i >> 2, sample, stream, sub, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), o);
  return o[i & 3];
}

// ascending bitonic sort of n_pad (power of two) 64-bit keys in shared memory, whole CTA
__device__ void bitonic_sort(uint64_t* keys, int n_pad) {
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const uint64_t a = keys[i], b = keys[p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            keys[i] = b;
            keys[p] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

constexpr int kMaskThreads = 256;

// One CTA per (sample, masked frame).  mask row layout: [visible_frames + mask_frames, h, w] bytes, 1 = masked.
__global__ void __launch_bounds__(kMaskThreads)
mask_uniform_kernel(uint64_t seed, int row0, int visible_frames, int mask_frames, int h, int w, int cf, int n_visible,
                    int n_pad, uint8_t* __restrict__ out) {
  extern __shared__ uint64_t keys[];
  const int row = blockIdx.x, f = blockIdx.y;
  const int gh = h / cf, gw = w / cf, n = gh * gw;
  const uint32_t sample = static_cast<uint32_t>(row0 + row);
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x)
    keys[i] = i < n ? (static_cast<uint64_t>(philox_word(seed, sample, kStreamUniform, f, i)) << 32) | static_cast<uint32_t>(i)
                    : ~0ull;
  uint8_t* frame = out + (static_cast<size_t>(row) * (visible_frames + mask_frames) + visible_frames + f) * h * w;
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) frame[i] = 1;
  if (f == 0) {
    uint8_t* vis = out + static_cast<size_t>(row) * (visible_frames + mask_frames) * h * w;
    for (int i = threadIdx.x; i < visible_frames * h * w; i += blockDim.x) vis[i] = 0;
  }
  __syncthreads();
  bitonic_sort(keys, n_pad);
  for (int i = threadIdx.x; i < n_visible * cf * cf; i += blockDim.x) {
    const int cell = static_cast<int>(keys[i / (cf * cf)] & 0xffffffffu);
    const int sub = i % (cf * cf);
    frame[((cell / gw) * cf + sub / cf) * w + (cell % gw) * cf + sub % cf] = 0;
  }
}

// One CTA per image: probabilities [n] (any non-negative-after-shift weights) -> integer cumulative table [n] (uint64,
// inclusive): q_i = floor(relu(p_i - min p + eps) / max * 2^24) -- cells more than 2^24 times less likely than the likeliest
// one are never drawn (the reference's eps = 1e-16 puts them at ~1e-17).
__global__ void __launch_bounds__(kMaskThreads)
mask_energy_table_kernel(const float* __restrict__ probs, int n, float eps, unsigned long long* __restrict__ table) {
  __shared__ float red[kMaskThreads];
  __shared__ unsigned long long wsum[kMaskThreads];
  const float* p = probs + static_cast<size_t>(blockIdx.x) * n;
  unsigned long long* t = table + static_cast<size_t>(blockIdx.x) * n;
  float mn = INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mn = fminf(mn, p[i]);
  red[threadIdx.x] = mn;
  __syncthreads();
  for (int s = kMaskThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fminf(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  mn = red[0];
  __syncthreads();
  float mx = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, fmaxf(__fadd_rn(__fsub_rn(p[i], mn), eps), 0.f));
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int s = kMaskThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  mx = red[0];
  // contiguous chunk per thread, then an exclusive scan of the chunk sums (integers: any order gives the same table)
  const int chunk = (n + kMaskThreads - 1) / kMaskThreads;
  const int i0 = min(static_cast<int>(threadIdx.x) * chunk, n), i1 = min(i0 + chunk, n);
  auto weight = [&](int i) -> unsigned long long {
    const float v = fmaxf(__fadd_rn(__fsub_rn(p[i], mn), eps), 0.f);
    if (!(mx > 0.f)) return 1ull;  // flat (or all-zero) energy: uniform
    return static_cast<unsigned long long>(__fmul_rn(__fdiv_rn(v, mx), 16777216.0f));
  };
  unsigned long long s = 0;
  for (int i = i0; i < i1; ++i) s += weight(i);
  wsum[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int k = 0; k < kMaskThreads; ++k) {
      const unsigned long long v = wsum[k];
      wsum[k] = run;
      run += v;
    }
  }
  __syncthreads();
  unsigned long long run = wsum[threadIdx.x];
  for (int i = i0; i < i1; ++i) {
    run += weight(i);
    t[i] = run;
  }
}

// One CTA per (image b, sample s): P draws with replacement from image b's table -> mask row (b * S + s).
__global__ void __launch_bounds__(64)
mask_energy_sample_kernel(const unsigned long long* __restrict__ table, int n, uint64_t seed, int sample0, int S, int P,
                          int visible_frames, int h, int w, int cf, uint8_t* __restrict__ out) {
  const int b = blockIdx.y, s = blockIdx.x;
  const unsigned long long* t = table + static_cast<size_t>(b) * n;
  uint8_t* row = out + (static_cast<size_t>(b) * S + s) * (visible_frames + 1) * h * w;
  uint8_t* frame = row + static_cast<size_t>(visible_frames) * h * w;
  for (int i = threadIdx.x; i < visible_frames * h * w; i += blockDim.x) row[i] = 0;
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) frame[i] = 1;
  __syncthreads();
  const int gw = w / cf;
  const unsigned long long total = t[n - 1];
  const uint32_t sample = static_cast<uint32_t>(sample0 + s);
  for (int pnt = threadIdx.x; pnt < P; pnt += blockDim.x) {
    uint32_t o[4];
    philox4x32_10(pnt, sample, kStreamEnergy, b, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), o);
    const unsigned long long r = (static_cast<unsigned long long>(o[0]) << 32) | o[1];
    const unsigned long long target = __umul64hi(r, total);  // uniform in [0, total)
    int lo = 0, hi = n - 1;                                   // first cell whose inclusive sum exceeds target
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (t[mid] > target) hi = mid; else lo = mid + 1;
    }
    const int cy = (lo / gw) * cf, cx = (lo % gw) * cf;
    for (int dy = 0; dy < cf; ++dy)
      for (int dx = 0; dx < cf; ++dx) frame[(cy + dy) * w + cx + dx] = 0;   // duplicates collapse, like the reference
  }
}

__global__ void mask_count_kernel(const uint8_t* __restrict__ masks, int N, int* __restrict__ counts, int* __restrict__ min_count) {
  __shared__ int red[kMaskThreads];
  const uint8_t* row = masks + static_cast<size_t>(blockIdx.x) * N;
  int c = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) c += row[i] != 0;
  red[threadIdx.x] = c;
  __syncthreads();
  for (int s = kMaskThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts[blockIdx.x] = red[0];
    atomicMin(min_count, red[0]);
  }
}

// One CTA per row: reveal (masked_count - target) masked tokens, chosen as the smallest Philox keys among the masked ones.
__global__ void __launch_bounds__(kMaskThreads)
mask_rectangularize_kernel(uint8_t* __restrict__ masks, int N, int n_pad, const int* __restrict__ counts,
                           const int* __restrict__ target_ptr, int target_fixed, uint64_t seed, int row0) {
  extern __shared__ uint64_t keys[];
  const int row = blockIdx.x;
  const int target = target_ptr ? *target_ptr : target_fixed;
  const int excess = counts[row] - target;
  if (excess <= 0) return;   // rows at (or, with a caller-given target, below) the target stay as they are
  uint8_t* m = masks + static_cast<size_t>(row) * N;
  const uint32_t sample = static_cast<uint32_t>(row0 + row);
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x)
    keys[i] = (i < N && m[i]) ? (static_cast<uint64_t>(philox_word(seed, sample, kStreamRect, 0, i)) << 32) | static_cast<uint32_t>(i)
                              : ~0ull;
  __syncthreads();
  bitonic_sort(keys, n_pad);
  for (int i = threadIdx.x; i < excess; i += blockDim.x) m[keys[i] & 0xffffffffu] = 0;
}

static int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace cwm

using namespace cwm;

extern "C" int cwm_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
  CWM_REQUIRE(counter && key && out, "cwm_philox4x32_10: null pointer");
  philox4x32_10(counter[0], counter[1], counter[2], counter[3], key[0], key[1], out);
  return CWM_OK;
}

extern "C" int cwm_mask_uniform(uint64_t seed, int row0, int rows, int visible_frames, int mask_frames, int h, int w,
                                int clump, int n_visible_cells, uint8_t* masks, cwm_stream_t stream) {
  CWM_REQUIRE(rows >= 0 && visible_frames >= 0 && mask_frames >= 1 && h > 0 && w > 0 && clump >= 1,
              "cwm_mask_uniform: bad shape (rows=%d frames=%d+%d grid=%dx%d clump=%d)", rows, visible_frames, mask_frames, h, w, clump);
  CWM_REQUIRE(h % clump == 0 && w % clump == 0, "cwm_mask_uniform: the %dx%d patch grid is not a multiple of the clump size %d", h, w, clump);
  const int n = (h / clump) * (w / clump);
  CWM_REQUIRE(n_visible_cells >= 0 && n_visible_cells <= n, "cwm_mask_uniform: %d visible cells of %d", n_visible_cells, n);
  CWM_REQUIRE(n <= 8192, "cwm_mask_uniform: at most 8192 clump cells per frame (got %d)", n);
  if (rows == 0) return CWM_OK;
  CWM_REQUIRE(masks, "cwm_mask_uniform: null pointer");
  const int n_pad = next_pow2(n);
  static bool attr = false;
  if (!attr) {
    CWM_CUDA_CHECK(cudaFuncSetAttribute(mask_uniform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
    attr = true;
  }
  ProfileScope prof(static_cast<cudaStream_t>(stream), "mask_uniform", 0.0,
                    static_cast<double>(rows) * (visible_frames + mask_frames) * h * w);
  mask_uniform_kernel<<<dim3(rows, mask_frames), kMaskThreads, n_pad * 8, static_cast<cudaStream_t>(stream)>>>(
      seed, row0, visible_frames, mask_frames, h, w, clump, n_visible_cells, n_pad, masks);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_mask_energy_table(const float* probs, int B, int n, float eps, uint64_t* table, cwm_stream_t stream) {
  CWM_REQUIRE(B >= 0 && n > 0, "cwm_mask_energy_table: bad shape B=%d n=%d", B, n);
  if (B == 0) return CWM_OK;
  CWM_REQUIRE(probs && table, "cwm_mask_energy_table: null pointer");
  ProfileScope prof(static_cast<cudaStream_t>(stream), "mask_energy_table", 0.0, static_cast<double>(B) * n * 12.0);
  mask_energy_table_kernel<<<B, kMaskThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      probs, n, eps, reinterpret_cast<unsigned long long*>(table));
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_mask_energy_sample(const uint64_t* table, int B, int h, int w, int clump, uint64_t seed, int sample0,
                                      int S, int points, int visible_frames, uint8_t* masks, cwm_stream_t stream) {
  CWM_REQUIRE(B >= 0 && S >= 0 && points >= 0 && visible_frames >= 0 && h > 0 && w > 0 && clump >= 1,
              "cwm_mask_energy_sample: bad shape");
  CWM_REQUIRE(h % clump == 0 && w % clump == 0, "cwm_mask_energy_sample: the %dx%d patch grid is not a multiple of the clump size %d", h, w, clump);
  if (B == 0 || S == 0) return CWM_OK;
  CWM_REQUIRE(table && masks, "cwm_mask_energy_sample: null pointer");
  CWM_REQUIRE(B <= 65535, "cwm_mask_energy_sample: at most 65535 images per call");
  ProfileScope prof(static_cast<cudaStream_t>(stream), "mask_energy_sample", 0.0,
                    static_cast<double>(B) * S * (visible_frames + 1) * h * w);
  mask_energy_sample_kernel<<<dim3(S, B), 64, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const unsigned long long*>(table), (h / clump) * (w / clump), seed, sample0, S, points,
      visible_frames, h, w, clump, masks);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" size_t cwm_mask_rectangularize_workspace_bytes(int rows) { return static_cast<size_t>(rows + 1) * sizeof(int); }

extern "C" int cwm_mask_rectangularize(uint8_t* masks, int rows, int N, int row0, uint64_t seed, int target_masked,
                                       void* workspace, size_t workspace_bytes, cwm_stream_t stream) {
  CWM_REQUIRE(rows >= 0 && N > 0, "cwm_mask_rectangularize: bad shape rows=%d N=%d", rows, N);
  CWM_REQUIRE(N <= 16384, "cwm_mask_rectangularize: at most 16384 tokens per row (got %d)", N);
  if (rows == 0) return CWM_OK;
  CWM_REQUIRE(masks && workspace, "cwm_mask_rectangularize: null pointer");
  CWM_REQUIRE(workspace_bytes >= cwm_mask_rectangularize_workspace_bytes(rows), "cwm_mask_rectangularize: workspace too small");
  int* counts = static_cast<int*>(workspace);
  int* min_count = counts + rows;
  const int n_pad = next_pow2(N);
  static bool attr = false;
  if (!attr) {
    CWM_CUDA_CHECK(cudaFuncSetAttribute(mask_rectangularize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    attr = true;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProfileScope prof(s, "mask_rectangularize", 0.0, static_cast<double>(rows) * N * 2.0);
  CWM_CUDA_CHECK(cudaMemsetAsync(min_count, 0x7f, sizeof(int), s));
  mask_count_kernel<<<rows, kMaskThreads, 0, s>>>(masks, N, counts, min_count);
  CWM_LAUNCH_CHECK();
  mask_rectangularize_kernel<<<rows, kMaskThreads, n_pad * 8, s>>>(masks, N, n_pad, counts,
                                                                  target_masked < 0 ? min_count : nullptr, target_masked, seed, row0);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}
