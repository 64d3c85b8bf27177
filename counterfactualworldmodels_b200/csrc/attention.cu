// attention.cu -- softmax(q k^T) v, head_dim 64, unmasked, f16 operands, fp32 softmax statistics and fp32
// accumulation in tensor memory (a6: cwm/models/VideoMAE/utils.py:108-113).
//
// Flash-style, warp-specialised, TMA-pipelined (sm_100a).  One CTA = 256 query rows (two 128-row tiles) of one
// (sample, head); K/V are streamed once per CTA in 128-row tiles.
//   warp 0       TMA producer (Q once, then K_j / V_j through a kKVStages-deep mbarrier ring)
//   warp 1       MMA issuer: S_t = Q_t K_j^T (SS, M=128 N=128 K=64) and O_t += P_t V_j (TS: A = P from tensor
//                memory, B = V from smem MN-major, M=128 N=64 K=128), interleaved over the two query tiles so the
//                tensor core works on one tile while the other tile is in softmax
//   warps 4-7    softmax warpgroup for query tile 0, warps 8-11 for tile 1: one thread per query row (no
//                shuffles): tcgen05.ld S row -> running max with lazy rescale (O is only rescaled when the max
//                grows by more than 2^8) -> exp2 -> fp32 row sum -> f16 P written back over S in tensor memory
// Tensor memory map (512 columns): three 128-column score buffers B0..B2 used round-robin -- the scores of
// (query tile t, kv tile j) live in B[(2j + t) % 3] and P (f16 pairs) overwrites their first 64 columns -- so that
// S_0(j+1) is computed while the softmax of S_0(j) is still running; O0 [384,448), O1 [448,512).
// Ordering facts relied on: tcgen05.mma from one thread execute in issue order (a buffer's previous P has been
// consumed by its PV before the next QK overwrites it), and a tcgen05.commit arrives only after ALL earlier MMAs
// completed (pv_done[t] after PV_t(j) => O_t may be rescaled / read by the softmax warpgroup).
// The last kv tile is issued with N = round_up(valid columns, 32), so ragged sequence lengths (788, 3140 ...)
// do not pay for a full 128-column tile.
#include "common.cuh"

namespace cwm {

constexpr int kAttnThreads = 384;
constexpr int kKVStages = 4;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB
constexpr int kAttnSmemBytes = 1024 + (2 + 2 * kKVStages) * kTileBytes + 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // in log2 units

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed dual-fp32 math (Blackwell FFMA2 / FADD2): one issue slot for two lanes of work
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

template <int REGS>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
}
template <int REGS>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_f16_kernel(const __grid_constant__ CUtensorMap tma_qkv, int N, int H, __half* __restrict__ out,
                     float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                               // 2 tiles
  uint8_t* smem_k = smem + 2 * kTileBytes;              // kKVStages tiles
  uint8_t* smem_v = smem_k + kKVStages * kTileBytes;    // kKVStages tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + kKVStages * kTileBytes);
  uint64_t* q_full = bars;                    // 1
  uint64_t* k_full = bars + 1;                // kKVStages
  uint64_t* k_empty = k_full + kKVStages;
  uint64_t* v_full = k_empty + kKVStages;
  uint64_t* v_empty = v_full + kKVStages;
  uint64_t* s_full = v_empty + kKVStages;     // 2
  uint64_t* p_full = s_full + 2;              // 2
  uint64_t* pv_done = p_full + 2;             // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int C = H * 64;
  const int q0 = blockIdx.x * 256;               // first query row of this CTA within the sample
  const bool t1_valid = (q0 + 128) < N;          // does query tile 1 contain any valid row?
  const int n_kv = (N + 127) / 128;
  const int row_base = b * N;                    // row of token 0 of this sample in the [B*N, 3C] matrix
  const int cols_last = ((N - (n_kv - 1) * 128) + 31) & ~31;  // MMA columns of the last kv tile (32..128)

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tma_qkv);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);  // one elected lane per softmax warp
      mbar_init(&pv_done[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    reg_dec<56>();
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, (t1_valid ? 2 : 1) * kTileBytes);
        tma_load_2d(smem_q, &tma_qkv, q_full, h * 64, row_base + q0);
        if (t1_valid) tma_load_2d(smem_q + kTileBytes, &tma_qkv, q_full, h * 64, row_base + q0 + 128);
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_kv; ++j) {
          mbar_wait(&k_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&k_full[stage], kTileBytes);
          tma_load_2d(smem_k + stage * kTileBytes, &tma_qkv, &k_full[stage], C + h * 64, row_base + j * 128);
          mbar_wait(&v_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&v_full[stage], kTileBytes);
          tma_load_2d(smem_v + stage * kTileBytes, &tma_qkv, &v_full[stage], 2 * C + h * 64, row_base + j * 128);
          if (++stage == kKVStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, 0, 1);  // B = V is MN-major
        const uint32_t q_addr0 = smem_u32(smem_q);
        const uint32_t q_addr1 = smem_u32(smem_q + kTileBytes);
        const uint32_t tm_o0 = tmem_base + 384, tm_o1 = tmem_base + 448;

        auto issue_qk = [&](uint32_t q_addr, uint32_t k_addr, uint32_t tm_s, int ncols) {
          const uint32_t idesc_qk = umma_idesc_f16(128, ncols, 0, 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tm_s, umma_desc_kmajor_sw128(q_addr + k * 32), umma_desc_kmajor_sw128(k_addr + k * 32),
                    idesc_qk, k != 0 ? 1u : 0u);
        };
        auto issue_pv = [&](uint32_t tm_p, uint32_t v_addr, uint32_t tm_o, bool accumulate, int ksteps) {
          for (int k = 0; k < ksteps; ++k)  // 16 kv rows per MMA: 8 TMEM columns of packed f16 pairs, 2048 B of V
            umma_ts(tm_o, tm_p + k * 8, umma_desc_mnmajor_sw128(v_addr + k * 2048, 0), idesc_pv,
                    (accumulate || k != 0) ? 1u : 0u);
        };
        auto sbuf = [&](int j, int t) { return tmem_base + static_cast<uint32_t>(((2 * j + t) % 3) * 128); };

        mbar_wait(q_full, 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        {
          const int nc0 = (n_kv == 1) ? cols_last : 128;
          issue_qk(q_addr0, smem_u32(smem_k), sbuf(0, 0), nc0);
          umma_commit(&s_full[0]);
          if (t1_valid) {
            issue_qk(q_addr1, smem_u32(smem_k), sbuf(0, 1), nc0);
            umma_commit(&s_full[1]);
          }
          umma_commit(&k_empty[0]);
        }

        int stage = 0;       // stage of K_j / V_j
        uint32_t phase = 0;
        for (int j = 0; j < n_kv; ++j) {
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == kKVStages) {
            nstage = 0;
            nphase ^= 1;
          }
          const bool has_next = (j + 1) < n_kv;
          const int nc_next = (j + 2 == n_kv) ? cols_last : 128;
          const int ksteps = ((j + 1 == n_kv) ? cols_last : 128) >> 4;
          const uint32_t v_addr = smem_u32(smem_v + stage * kTileBytes);
          const uint32_t kn_addr = smem_u32(smem_k + nstage * kTileBytes);
          // ---- S_0(j+1) first: it lands in the buffer freed by PV_1(j-1), so tile 0 never waits for its scores
          if (has_next) {
            mbar_wait(&k_full[nstage], nphase);
            tc_fence_after();
            issue_qk(q_addr0, kn_addr, sbuf(j + 1, 0), nc_next);
            umma_commit(&s_full[0]);
          }
          // ---- O_0 += P_0(j) V_j
          mbar_wait(&p_full[0], j & 1);
          mbar_wait(&v_full[stage], phase);
          tc_fence_after();
          issue_pv(sbuf(j, 0), v_addr, tm_o0, j > 0, ksteps);
          umma_commit(&pv_done[0]);
          if (t1_valid) {
            // ---- S_1(j+1) reuses the buffer P_0(j) just vacated (in-order execution after PV_0(j))
            if (has_next) {
              issue_qk(q_addr1, kn_addr, sbuf(j + 1, 1), nc_next);
              umma_commit(&s_full[1]);
            }
            mbar_wait(&p_full[1], j & 1);
            tc_fence_after();
            issue_pv(sbuf(j, 1), v_addr, tm_o1, j > 0, ksteps);
            umma_commit(&pv_done[1]);
          }
          umma_commit(&v_empty[stage]);
          if (has_next) umma_commit(&k_empty[nstage]);
          stage = nstage;
          phase = nphase;
        }
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    reg_inc<208>();
    const int t = (warp - 4) >> 2;          // query tile
    const int quad = warp & 3;              // TMEM lane quadrant
    const int row_in_tile = quad * 32 + lane;
    const int q_row = q0 + t * 128 + row_in_tile;  // row within the sample
    if (t == 0 || t1_valid) {
      const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
      const uint32_t tm_o = tmem_base + lane_off + 384 + t * 64;
      float m_ref = 0.f;   // running reference max (raw score units)
      float l_sum = 0.f;   // running sum of exp2((s - m_ref) * scale_log2)
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t tm_s = tmem_base + lane_off + static_cast<uint32_t>(((2 * j + t) % 3) * 128);
        const int ncols = (j + 1 == n_kv) ? cols_last : 128;  // columns the MMA produced for this tile
        const int kv_valid = N - j * 128;                     // columns >= kv_valid are padding
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        uint32_t s[128];
        // all four chunks are always read (columns >= ncols hold stale data and are masked below); only the
        // exp / P-store work is skipped for chunks the last, ragged MMA did not produce
        tmem_ld_x32(tm_s + 0, s);
        tmem_ld_x32(tm_s + 32, s + 32);
        tmem_ld_x32(tm_s + 64, s + 64);
        tmem_ld_x32(tm_s + 96, s + 96);
        tmem_ld_wait();
        if (kv_valid < 128) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= kv_valid) s[i] = 0xff800000u;  // -inf
        }
        float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]);
        float mx2 = __uint_as_float(s[2]), mx3 = __uint_as_float(s[3]);
#pragma unroll
        for (int i = 4; i < 128; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(s[i]));
          mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
        }
        const float row_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (j == 0) {
          m_ref = row_max;
        } else {
          const bool need = (row_max - m_ref) * scale_log2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {
            const float f = need ? ex2((m_ref - row_max) * scale_log2) : 1.0f;
            if (need) m_ref = row_max;
            l_sum *= f;
            // O_t is quiescent once PV_t(j-1) has completed: PV_t(j) is not issued before this warpgroup
            // signals p_full[t].  (At step j the barrier has completed phase j-1 at most, so the parity is
            // unambiguous even though this wait is only executed when a rescale is needed.)
            mbar_wait(&pv_done[t], (j - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t o[32];
              tmem_ld_x32(tm_o + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
              tmem_st_x32(tm_o + c * 32, o);
            }
            tmem_st_wait();
          }
        }
        const float neg_m = -m_ref * scale_log2;
        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 32 < ncols) {
            uint32_t p[16];  // 32 kv elements as packed f16 pairs; P aliases S columns that are already in registers
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const int e = c * 32 + i;
              const float p0 = ex2(fmaf(__uint_as_float(s[e]), scale_log2, neg_m));
              const float p1 = ex2(fmaf(__uint_as_float(s[e + 1]), scale_log2, neg_m));
              const float p2 = ex2(fmaf(__uint_as_float(s[e + 2]), scale_log2, neg_m));
              const float p3 = ex2(fmaf(__uint_as_float(s[e + 3]), scale_log2, neg_m));
              sum0 += p0;
              sum1 += p1;
              sum2 += p2;
              sum3 += p3;
              p[(i >> 1)] = pack_half2(p0, p1);
              p[(i >> 1) + 1] = pack_half2(p2, p3);
            }
            tmem_st_x16(tm_s + c * 16, p);
          }
        }
        tmem_st_wait();
        l_sum += (sum0 + sum1) + (sum2 + sum3);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      // ---- finalise: O / l -> f16 -> global
      mbar_wait(&pv_done[t], (n_kv - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.0f / l_sum;
      __half* orow = out + (static_cast<long long>(row_base) + q_row) * C + h * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld_x32(tm_o + c * 32, o);
        tmem_ld_wait();
        if (q_row < N) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 v;
            v.x = pack_half2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
            v.y = pack_half2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
            v.z = pack_half2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
            v.w = pack_half2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c * 32 + i) = v;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace cwm

using namespace cwm;

extern "C" int cwm_attention_f16(const uint16_t* qkv, int B, int N, int H, int head_dim, uint16_t* out,
                                 cwm_stream_t stream) {
  CWM_REQUIRE(qkv && out, "cwm_attention_f16: null pointer");
  CWM_REQUIRE(B >= 0 && N > 0 && H > 0, "cwm_attention_f16: bad shape B=%d N=%d H=%d", B, N, H);
  if (head_dim != 64) return fail(CWM_ERR_UNSUPPORTED, "cwm_attention_f16: head_dim %d (only 64 is implemented)", head_dim);
  CWM_REQUIRE(H <= 65535 && B <= 65535, "cwm_attention_f16: grid too large");
  if (B == 0) return CWM_OK;
  static bool attr_set = false;
  if (!attr_set) {
    CWM_CUDA_CHECK(cudaFuncSetAttribute(attention_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kAttnSmemBytes));
    attr_set = true;
  }
  const int C = H * 64;
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, qkv, CWM_TMAP_F16, static_cast<uint64_t>(B) * N, 3ull * C, 3ull * C, 128, 64);
  if (rc) return rc;
  ProfileScope prof(static_cast<cudaStream_t>(stream), "attention_f16", 4.0 * B * H * static_cast<double>(N) * N * 64,
                    static_cast<double>(B) * N * C * 2.0 * 4.0);
  dim3 grid((N + 255) / 256, H, B);
  attention_f16_kernel<<<grid, kAttnThreads, kAttnSmemBytes, static_cast<cudaStream_t>(stream)>>>(
      tm, N, H, reinterpret_cast<__half*>(out), kLog2e);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}
