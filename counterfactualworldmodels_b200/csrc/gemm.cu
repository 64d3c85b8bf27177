// gemm.cu -- Y[M,N] = epilogue(A[M,K] . W[N,K]^T) on tcgen05 tensor cores (sm_100a).
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer: 128x64 A tile + BNx64 W tile per stage, 128B-swizzled, mbarrier pipelined
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//               fp32 accumulators in tensor memory, two accumulator stages so the epilogue of tile i overlaps
//               the main loop of tile i+1
//   warp 2      TMEM allocator
//   warps 4-11  epilogue: tcgen05.ld (32 lanes x 32/64 columns) -> registers -> per-warp padded smem transpose
//               -> fully coalesced 128-byte global stores (and residual loads).  Two warps per TMEM lane
//               quadrant split the column chunks.
// Fused epilogues (cwm_epilogue): +bias, q-scale, GELU(erf), +residual (optionally row-gathered = positional
// embedding of visible tokens), output row remap (writes x_vis straight into the decoder sequence).
//
// Roofline: tensor bound for K >= 512; the fp32-residual epilogues of the narrow decoder (K = 384/512) are HBM/L2
// bound (8 bytes of residual traffic per output element) -- see DESIGN.md.
#include "common.cuh"

namespace cwm {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 f16 = 128 bytes = one swizzle row
constexpr int kGemmThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kStagingWords = 32 * 33;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 192 ? 4 : 6);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int kSmemBytes =
      1024 + kStages * kStageBytes + kEpiWarps * kStagingWords * 4 + 256;
};

struct EpiDev {
  int mode;
  const float* bias;
  float scale;
  int scale_cols;
  const float* res;
  int ldr;
  const int32_t* res_gather;
  int gather_stride;
  int grp_rows;
  int grp_out_stride;
  void* out;
  int ldo;
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w, int M, int N,
                int K, EpiDev ep) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint32_t* staging = reinterpret_cast<uint32_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + kEpiWarps * kStagingWords * 4);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
  uint64_t* tempty_bar = bars + 2 * Cfg::kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = (M + BM - 1) / BM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / tiles_n;
        const int n_blk = tile - m_blk * tiles_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_a + stage * Cfg::kABytes, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tma_w, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = umma_desc_kmajor_sw128(a_addr + k * 32);
            const uint64_t bdesc = umma_desc_kmajor_sw128(b_addr + k * 32);
            umma_ss(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when the MMAs above have read it
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;       // 0..7
    const int quad = ew & 3;       // TMEM lane quadrant == warp % 4
    const int half = ew >> 2;      // which column chunks this warp takes
    uint32_t* stg = staging + ew * kStagingWords;
    const bool f16_out = (ep.mode == CWM_EPI_F16 || ep.mode == CWM_EPI_GELU_F16);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / tiles_n;
      const int n_blk = tile - m_blk * tiles_n;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      const int row0 = m_blk * BM + quad * 32;
      if (f16_out) {
        // 64-column chunks; math in row-owner layout, packed to half2, transposed through smem
        constexpr int kChunks = BN / 64;
        for (int c = half; c < kChunks; c += 2) {
          const int n0 = n_blk * BN + c * 64;
          uint32_t acc0[32], acc1[32];
          tmem_ld_x32(tmem_acc + c * 64, acc0);
          tmem_ld_x32(tmem_acc + c * 64 + 32, acc1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float v[4];
            v[0] = __uint_as_float(acc0[j]);
            v[1] = __uint_as_float(acc0[j + 1]);
            v[2] = __uint_as_float(acc1[j]);
            v[3] = __uint_as_float(acc1[j + 1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int n = n0 + (q >> 1) * 32 + j + (q & 1);
              if (ep.bias != nullptr && n < N) v[q] += __ldg(ep.bias + n);
              if (ep.mode == CWM_EPI_GELU_F16) {
                v[q] = gelu_erf(v[q]);
              } else if (n < ep.scale_cols) {
                v[q] *= ep.scale;
              }
            }
            stg[lane * 33 + (j >> 1)] = pack_half2(v[0], v[1]);
            stg[lane * 33 + 16 + (j >> 1)] = pack_half2(v[2], v[3]);
          }
          __syncwarp();
          const int n = n0 + 2 * lane;
          if (n < N) {
            __half* outp = reinterpret_cast<__half*>(ep.out);
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
              const int m = row0 + r;
              if (m >= M) break;
              long long orow = m;
              if (ep.grp_rows > 0) orow = static_cast<long long>(m / ep.grp_rows) * ep.grp_out_stride + m % ep.grp_rows;
              *reinterpret_cast<uint32_t*>(outp + orow * ep.ldo + n) = stg[r * 33 + lane];
            }
          }
          __syncwarp();
        }
      } else {
        // 32-column chunks; raw accumulators transposed through smem, math in column-owner layout
        constexpr int kChunks = BN / 32;
        for (int c = half; c < kChunks; c += 2) {
          const int n0 = n_blk * BN + c * 32;
          uint32_t acc[32];
          tmem_ld_x32(tmem_acc + c * 32, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = acc[j];
          __syncwarp();
          const int n = n0 + lane;
          if (n < N) {
            const float bias = ep.bias != nullptr ? __ldg(ep.bias + n) : 0.f;
            float* outp = reinterpret_cast<float*>(ep.out);
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
              const int m = row0 + r;
              if (m >= M) break;
              long long orow = m;
              int g = 0, j = m;
              if (ep.grp_rows > 0) {
                g = m / ep.grp_rows;
                j = m - g * ep.grp_rows;
                orow = static_cast<long long>(g) * ep.grp_out_stride + j;
              }
              float v = __uint_as_float(stg[r * 33 + lane]) + bias;
              if (ep.mode == CWM_EPI_RES_F32) {
                long long rrow = orow;
                if (ep.res_gather != nullptr)
                  rrow = ep.res_gather[static_cast<long long>(g) * ep.gather_stride + j];
                v += ep.res[rrow * ep.ldr + n];
              }
              outp[orow * ep.ldo + n] = v;
            }
          }
          __syncwarp();
        }
      }
      // all tcgen05.ld of this warp have completed (tmem_ld_wait) -> release the accumulator stage
      tc_fence_before();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, int M, int N, int K, const EpiDev& ep,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CWM_CUDA_CHECK(cudaFuncSetAttribute(gemm_f16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_f16_kernel<BN><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tw, M, N, K, ep);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

int pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  if (N <= 192) return 192;
  return 256;
}

}  // namespace cwm

using namespace cwm;

extern "C" int cwm_gemm_f16(const uint16_t* A, const uint16_t* W, int M, int N, int K, const cwm_gemm_epilogue* e,
                            cwm_stream_t stream) {
  CWM_REQUIRE(A && W && e && e->out, "cwm_gemm_f16: null pointer");
  CWM_REQUIRE(M >= 0 && N > 0 && K >= 16 && K % 16 == 0, "cwm_gemm_f16: bad shape M=%d N=%d K=%d (K %% 16 == 0)", M, N, K);
  CWM_REQUIRE(e->mode >= CWM_EPI_F16 && e->mode <= CWM_EPI_F32, "cwm_gemm_f16: bad epilogue mode %d", e->mode);
  CWM_REQUIRE(e->mode != CWM_EPI_RES_F32 || e->res != nullptr, "cwm_gemm_f16: residual epilogue without residual");
  CWM_REQUIRE(e->res_gather == nullptr || e->grp_rows > 0, "cwm_gemm_f16: res_gather needs grp_rows > 0");
  const bool f16_out = (e->mode == CWM_EPI_F16 || e->mode == CWM_EPI_GELU_F16);
  CWM_REQUIRE(!f16_out || (N % 2 == 0 && e->ldo % 2 == 0 && reinterpret_cast<uintptr_t>(e->out) % 4 == 0),
              "cwm_gemm_f16: f16 output needs even N/ldo and 4-byte aligned out");
  if (M == 0) return CWM_OK;
  EpiDev ep;
  ep.mode = e->mode; ep.bias = e->bias; ep.scale = e->scale; ep.scale_cols = (e->mode == CWM_EPI_F16) ? e->scale_cols : 0;
  ep.res = e->res; ep.ldr = e->ldr; ep.res_gather = e->res_gather; ep.gather_stride = e->gather_stride;
  ep.grp_rows = e->grp_rows; ep.grp_out_stride = e->grp_out_stride; ep.out = e->out; ep.ldo = e->ldo;
  const int bn = pick_bn(N);
  CUtensorMap ta, tw;
  int rc = make_tmap_2d_f16(&ta, A, M, K, K, BM, BK);
  if (rc) return rc;
  rc = make_tmap_2d_f16(&tw, W, N, K, K, bn, BK);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const char* kNames[4] = {"gemm_f16_out", "gemm_gelu_f16_out", "gemm_residual_f32", "gemm_f32_out"};
  const double out_bytes = static_cast<double>(M) * N * (f16_out ? 2.0 : (e->mode == CWM_EPI_RES_F32 ? 8.0 : 4.0));
  ProfileScope prof(s, kNames[e->mode], 2.0 * M * N * K,
                    (static_cast<double>(M) * K + static_cast<double>(N) * K) * 2.0 + out_bytes);
  switch (bn) {
    case 64: return launch_gemm<64>(ta, tw, M, N, K, ep, s);
    case 128: return launch_gemm<128>(ta, tw, M, N, K, ep, s);
    case 192: return launch_gemm<192>(ta, tw, M, N, K, ep, s);
    default: return launch_gemm<256>(ta, tw, M, N, K, ep, s);
  }
}
