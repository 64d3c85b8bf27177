// raftcorr_tc.cu -- RAFT's all-pairs correlation volume (cwm/models/raft/corr.py:53-60:
// corr[s, i, j] = <fmap1[s, :, i], fmap2[s, :, j]> / sqrt(D)) on the tensor cores with fp32 accuracy.
//
// A plain TF32 GEMM would lose 13 mantissa bits of every operand, and the reference computes this product in fp32
// (raft_model.py:224-225 casts the feature maps to float before CorrBlock).  So every operand is split once into
// hi = tf32(x) and lo = tf32(x - hi) and the product is accumulated as  hi*hi + hi*lo + lo*hi  (the dropped lo*lo term is
// 2^-22 of a product) on `tcgen05.mma.kind::tf32` with fp32 accumulators in tensor memory -- the "3xTF32" scheme.
//   corr_split_kernel   NCHW fp32 [S, D, HW] -> pixel-major hi / lo [S, HW, D] (K-major operands; the transpose rides on
//                       the split pass, which has to touch every element anyway)
//   corr_tf32_kernel    persistent (one CTA per SM) over 128 x 128 output tiles: warp 0 TMA producer (4 operand tiles per
//                       32-channel k-slab, 3-stage mbarrier ring), warp 1 MMA issuer (12 MMAs of K = 8 per slab, two
//                       accumulator stages in tensor memory), warps 2-5 epilogue (tcgen05.ld -> scale -> 128-bit stores)
//                       overlapping the next tile's MMAs
// Roofline: 3 x 2 * HW^2 * D FLOP per sample on the TF32 pipe (half the f16 rate) against HW^2 * 4 bytes written: the
// 784 x 784 x 256 volumes of a 224-px frame pair are tensor bound by a factor ~3, so the goal is simply to leave the fp32
// SIMT pipe (33.6 TFLOP/s, round 1).
#include "common.cuh"

namespace cwm {

constexpr int kCtThreads = 192;
constexpr int kCtStages = 3;
constexpr int kCtTileBytes = 128 * 32 * 4;          // 128 rows x 32 fp32 = 16 KB, 128B-swizzled rows
constexpr int kCtStageBytes = 4 * kCtTileBytes;     // A_hi, A_lo, B_hi, B_lo
constexpr int kCtSmemBytes = 1024 + kCtStages * kCtStageBytes + 256;

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// [S, D, HW] -> hi, lo [S, HW, D]; 32 x 32 tiles through shared memory (coalesced on both sides)
__global__ void __launch_bounds__(256)
corr_split_kernel(const float* __restrict__ x, int D, int HW, float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int s = blockIdx.z, d0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows per pass
  const float* xs = x + static_cast<size_t>(s) * D * HW;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int d = d0 + r, p = p0 + tx;
    tile[r][tx] = (d < D && p < HW) ? xs[static_cast<size_t>(d) * HW + p] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, d = d0 + tx;
    if (p < HW && d < D) {
      const float v = tile[tx][r];
      const float h = to_tf32(v);
      const size_t o = (static_cast<size_t>(s) * HW + p) * D + d;
      hi[o] = h;
      lo[o] = to_tf32(v - h);
    }
  }
}

// kind::tf32 instruction descriptor: fp32 accumulate, TF32 operands (format code 2), both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kCtThreads, 1)
corr_tf32_kernel(const __grid_constant__ CUtensorMap tma_ahi, const __grid_constant__ CUtensorMap tma_alo,
                 const __grid_constant__ CUtensorMap tma_bhi, const __grid_constant__ CUtensorMap tma_blo, int HW, int D,
                 int num_samples, float div, float inv_div_exact, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCtStages * kCtStageBytes);
  uint64_t* full_bar = bars;                 // kCtStages
  uint64_t* empty_bar = bars + kCtStages;    // kCtStages
  uint64_t* tfull_bar = bars + 2 * kCtStages;  // 2 accumulator stages
  uint64_t* tempty_bar = tfull_bar + 2;        // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int num_kb = D / 32;
  // persistent: CTA c walks tiles c, c + grid, ...; tile -> (sample, m block, n block), n fastest so that concurrently
  // running CTAs share the A tiles and the sample's B tiles in L2
  const int tiles_1d = (HW + 127) / 128;
  const int tiles_per_sample = tiles_1d * tiles_1d;
  const int num_tiles = tiles_per_sample * num_samples;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_ahi);
    tma_prefetch_desc(&tma_alo);
    tma_prefetch_desc(&tma_bhi);
    tma_prefetch_desc(&tma_blo);
    for (int i = 0; i < kCtStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);   // one elected lane per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int s = tile / tiles_per_sample;
      const int r = tile - s * tiles_per_sample;
      const int mb = r / tiles_1d, nb = r - mb * tiles_1d;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* st = smem + stage * kCtStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], kCtStageBytes);
          tma_load_3d(st, &tma_ahi, &full_bar[stage], kb * 32, mb * 128, s);
          tma_load_3d(st + kCtTileBytes, &tma_alo, &full_bar[stage], kb * 32, mb * 128, s);
          tma_load_3d(st + 2 * kCtTileBytes, &tma_bhi, &full_bar[stage], kb * 32, nb * 128, s);
          tma_load_3d(st + 3 * kCtTileBytes, &tma_blo, &full_bar[stage], kb * 32, nb * 128, s);
        }
        __syncwarp();
        if (++stage == kCtStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_tf32(128, 128);
    const uint64_t desc0 = umma_desc_kmajor_sw128(smem_u32(smem));
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * 128;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t ahi = desc0 + static_cast<uint64_t>((stage * kCtStageBytes) >> 4);
        const uint64_t alo = ahi + (kCtTileBytes >> 4), bhi = ahi + 2 * (kCtTileBytes >> 4), blo = ahi + 3 * (kCtTileBytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 8 fp32 = 32 bytes per MMA: the descriptor start advances by 2 x 16 B
            umma_tf32(tmem_d, ahi + 2 * k, bhi + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_tf32(tmem_d, ahi + 2 * k, blo + 2 * k, idesc, 1u);
            umma_tf32(tmem_d, alo + 2 * k, bhi + 2 * k, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++stage == kCtStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else {
    // ===================== epilogue: one thread per output row; overlaps the next tile's MMAs =====================
    const int quad = warp & 3;
    const bool vec = (HW % 4 == 0);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int s = tile / tiles_per_sample;
      const int r = tile - s * tiles_per_sample;
      const int mb = r / tiles_1d, nb = r - mb * tiles_1d;
      const int row = mb * 128 + quad * 32 + lane;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      float* orow = out + (static_cast<size_t>(s) * HW + row) * HW + nb * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t acc[32];
        tmem_ld_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * 128 + c * 32, acc);
        tmem_ld_wait();
        if (c == 3) {   // all of this warp's reads of the accumulator stage are done
          tc_fence_before();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        float v[32];
        if (inv_div_exact != 0.f) {   // sqrt(D) is a power of two: multiplying by its reciprocal IS the division
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) * inv_div_exact;
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __fdiv_rn(__uint_as_float(acc[i]), div);
        }
        if (row < HW) {
          const int n0 = nb * 128 + c * 32;
          if (vec && n0 + 32 <= HW) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(orow + c * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + i < HW) orow[c * 32 + i] = v[i];
          }
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---- the same volume from f16 pixel-major rows (the fused feature encoder's output) ----
// fmap rows [S, HW, D] f16 are already the K-major operands, and f16 x f16 products are exact in the fp32 accumulator, so
// the mixed-precision path needs neither the fp32 cast, nor the NCHW transpose, nor the hi / lo split: one
// tcgen05.mma.kind::f16 per 16 channels (a third of the MMAs of the 3xTF32 scheme, half the operand bytes).  Same tile walk,
// accumulator double buffering and epilogue as corr_tf32_kernel.  a_shared: fmap1 is ONE image shared by all samples (a
// counterfactual sweep's frame 0).
constexpr int kCfStages = 4;
constexpr int kCfTileBytes = 128 * 64 * 2;          // 128 rows x 64 f16 = 16 KB, 128B-swizzled rows
constexpr int kCfStageBytes = 2 * kCfTileBytes;     // A, B
constexpr int kCfSmemBytes = 1024 + kCfStages * kCfStageBytes + 256;

template <typename TO>   // float: the fp32 volume; __half: the f16 pyramid of the mixed-precision path
__global__ void __launch_bounds__(kCtThreads, 1)
corr_f16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int HW, int D,
                int num_samples, int a_shared, float div, float inv_div_exact, TO* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCfStages * kCfStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kCfStages;
  uint64_t* tfull_bar = bars + 2 * kCfStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int num_kb = D / 64;
  const int tiles_1d = (HW + 127) / 128;
  const int tiles_per_sample = tiles_1d * tiles_1d;
  const int num_tiles = tiles_per_sample * num_samples;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < kCfStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int s = tile / tiles_per_sample;
      const int r = tile - s * tiles_per_sample;
      const int mb = r / tiles_1d, nb = r - mb * tiles_1d;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* st = smem + stage * kCfStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], kCfStageBytes);
          tma_load_4d(st, &tma_a, &full_bar[stage], kb * 64, mb * 128, 0, a_shared ? 0 : s);
          tma_load_4d(st + kCfTileBytes, &tma_b, &full_bar[stage], kb * 64, nb * 128, 0, s);
        }
        __syncwarp();
        if (++stage == kCfStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_f16(128, 128, 0, 0);
    const uint64_t desc0 = umma_desc_kmajor_sw128(smem_u32(smem));
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * 128;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t a = desc0 + static_cast<uint64_t>((stage * kCfStageBytes) >> 4);
        const uint64_t b = a + (kCfTileBytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_d, a + 2 * k, b + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++stage == kCfStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else {
    const int quad = warp & 3;
    const bool vec = (HW % 8 == 0);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int s = tile / tiles_per_sample;
      const int r = tile - s * tiles_per_sample;
      const int mb = r / tiles_1d, nb = r - mb * tiles_1d;
      const int row = mb * 128 + quad * 32 + lane;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      TO* orow = out + (static_cast<size_t>(s) * HW + row) * HW + nb * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t acc[32];
        tmem_ld_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * 128 + c * 32, acc);
        tmem_ld_wait();
        if (c == 3) {
          tc_fence_before();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        float v[32];
        if (inv_div_exact != 0.f) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) * inv_div_exact;
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __fdiv_rn(__uint_as_float(acc[i]), div);
        }
        if (row < HW) {
          const int n0 = nb * 128 + c * 32;
          if (vec && n0 + 32 <= HW) {
            if constexpr (sizeof(TO) == 4) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(orow + c * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 8)
                *reinterpret_cast<uint4*>(orow + c * 32 + i) = make_uint4(pack_half2(v[i], v[i + 1]), pack_half2(v[i + 2], v[i + 3]),
                                                                          pack_half2(v[i + 4], v[i + 5]), pack_half2(v[i + 6], v[i + 7]));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + i < HW) {
                if constexpr (sizeof(TO) == 4) orow[c * 32 + i] = v[i]; else orow[c * 32 + i] = __float2half_rn(v[i]);
              }
          }
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace cwm

using namespace cwm;

extern "C" size_t cwm_raft_corr_tc_workspace_bytes(int B, int D, int H, int W) {
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  return 4ull * static_cast<size_t>(B) * H * W * D * sizeof(float);
}

// Level 0 of the correlation pyramid on the tensor cores: out [B, HW, HW] = fmap1^T fmap2 / sqrt(D).  D % 32 == 0.
extern "C" int cwm_raft_corr_volume_tc(const float* fmap1, const float* fmap2, int B, int D, int H, int W, float* out,
                                       void* workspace, size_t workspace_bytes, cwm_stream_t stream) {
  CWM_REQUIRE(B >= 0 && D >= 32 && D % 32 == 0 && H >= 1 && W >= 1, "cwm_raft_corr_volume_tc: bad shape B=%d D=%d H=%d W=%d (D %% 32 == 0)", B, D, H, W);
  if (B == 0) return CWM_OK;
  CWM_REQUIRE(fmap1 && fmap2 && out && workspace, "cwm_raft_corr_volume_tc: null pointer");
  CWM_REQUIRE(workspace_bytes >= cwm_raft_corr_tc_workspace_bytes(B, D, H, W) && reinterpret_cast<uintptr_t>(workspace) % 16 == 0,
              "cwm_raft_corr_volume_tc: workspace too small or misaligned");
  CWM_REQUIRE(B <= 65535, "cwm_raft_corr_volume_tc: batch %d > 65535 (chunk the sweep)", B);
  const int HW = H * W;
  const size_t n = static_cast<size_t>(B) * HW * D;
  float* ws = static_cast<float*>(workspace);
  float *ahi = ws, *alo = ws + n, *bhi = ws + 2 * n, *blo = ws + 3 * n;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ProfileScope prof(st, "raft_corr_split", 0.0, static_cast<double>(n) * 2.0 * 12.0);
    const dim3 grid((HW + 31) / 32, D / 32, B);
    corr_split_kernel<<<grid, 256, 0, st>>>(fmap1, D, HW, ahi, alo);
    CWM_LAUNCH_CHECK();
    corr_split_kernel<<<grid, 256, 0, st>>>(fmap2, D, HW, bhi, blo);
    CWM_LAUNCH_CHECK();
  }
  CUtensorMap ta, tb, tc, td;
  int rc = make_tmap_3d_f32(&ta, ahi, B, HW, D, 128, 32);
  if (rc) return rc;
  if ((rc = make_tmap_3d_f32(&tb, alo, B, HW, D, 128, 32))) return rc;
  if ((rc = make_tmap_3d_f32(&tc, bhi, B, HW, D, 128, 32))) return rc;
  if ((rc = make_tmap_3d_f32(&td, blo, B, HW, D, 128, 32))) return rc;
  static bool attr = false;
  if (!attr) {
    CWM_CUDA_CHECK(cudaFuncSetAttribute(corr_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtSmemBytes));
    attr = true;
  }
  ProfileScope prof(st, "raft_corr_volume_tf32x3", 3.0 * 2.0 * B * static_cast<double>(HW) * HW * D,
                    static_cast<double>(B) * HW * (static_cast<double>(HW) * 4.0 + 4.0 * D * 4.0));
  const long long tiles = static_cast<long long>((HW + 127) / 128) * ((HW + 127) / 128) * B;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  const float div = sqrtf(static_cast<float>(D));
  int e = 0;
  const float m = frexpf(div, &e);                  // div = m * 2^e; m == 0.5 <=> a power of two
  const float inv_exact = (m == 0.5f) ? 1.0f / div : 0.f;
  corr_tf32_kernel<<<grid, kCtThreads, kCtSmemBytes, st>>>(ta, tb, tc, td, HW, D, B, div, inv_exact, out);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

// Level 0 from f16 pixel-major feature rows: rows1 [n1 * HW, D] (n1 = 1: one image shared by all B samples, or n1 = B),
// rows2 [B * HW, D]; out [B, HW, HW] (fp32, or f16 for the f16 pyramid) = rows1 rows2^T / sqrt(D).  D % 64 == 0.
template <typename TO>
static int corr_volume_rows_impl(const uint16_t* rows1, int n1, const uint16_t* rows2, int B, int D, int H, int W, TO* out,
                                 cwm_stream_t stream) {
  CWM_REQUIRE(B >= 0 && D >= 64 && D % 64 == 0 && H >= 1 && W >= 1 && (n1 == 1 || n1 == B),
              "cwm_raft_corr_volume_rows_f16: bad shape B=%d n1=%d D=%d H=%d W=%d (D %% 64 == 0, n1 in {1, B})", B, n1, D, H, W);
  if (B == 0) return CWM_OK;
  CWM_REQUIRE(rows1 && rows2 && out && reinterpret_cast<uintptr_t>(out) % 16 == 0, "cwm_raft_corr_volume_rows_f16: null / misaligned pointer");
  CWM_REQUIRE(B <= 65535, "cwm_raft_corr_volume_rows_f16: batch %d > 65535 (chunk the sweep)", B);
  const int HW = H * W;
  CUtensorMap ta, tb;
  int rc = make_tmap_nhwc(&ta, rows1, n1, 1, HW, D, D, 1, 128, 64);
  if (rc) return rc;
  if ((rc = make_tmap_nhwc(&tb, rows2, B, 1, HW, D, D, 1, 128, 64))) return rc;
  static bool attr = false;
  if (!attr) {
    CWM_CUDA_CHECK(cudaFuncSetAttribute(corr_f16_kernel<TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCfSmemBytes));
    attr = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfileScope prof(st, "raft_corr_volume_f16", 2.0 * B * static_cast<double>(HW) * HW * D,
                    static_cast<double>(B) * HW * (static_cast<double>(HW) * sizeof(TO) + 2.0 * D * 2.0));
  const long long tiles = static_cast<long long>((HW + 127) / 128) * ((HW + 127) / 128) * B;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  const float div = sqrtf(static_cast<float>(D));
  int e = 0;
  const float m = frexpf(div, &e);
  const float inv_exact = (m == 0.5f) ? 1.0f / div : 0.f;
  CWM_CUDA_CHECK(launch_pdl(corr_f16_kernel<TO>, dim3(grid), dim3(kCtThreads), kCfSmemBytes, st, ta, tb, HW, D, B,
                            n1 == 1 && B > 1 ? 1 : 0, div, inv_exact, out));
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

extern "C" int cwm_raft_corr_volume_rows_f16(const uint16_t* rows1, int n1, const uint16_t* rows2, int B, int D, int H, int W,
                                             float* out, cwm_stream_t stream) {
  return corr_volume_rows_impl<float>(rows1, n1, rows2, B, D, H, W, out, stream);
}

extern "C" int cwm_raft_corr_volume_rows_f16_out16(const uint16_t* rows1, int n1, const uint16_t* rows2, int B, int D, int H,
                                                   int W, uint16_t* out16, cwm_stream_t stream) {
  return corr_volume_rows_impl<__half>(rows1, n1, rows2, B, D, H, W, reinterpret_cast<__half*>(out16), stream);
}
