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
constexpr int kDefaultPoly = 2;            // columns out of 8 whose exp2 runs on the FMA pipe (tools/kernel_bench.py sweep)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA / ALU pipes (no MUFU): Cody-Waite range reduction with the 1.5 * 2^23 magic constant and a
// degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, ~6x below the f16 rounding of P).
// head dim 64 needs one exp per 256 tensor FLOPs and the SM has only 16 MUFU lanes, so a fixed fraction of the
// elements of every row takes this path to balance the MUFU pipe against the FMA pipe.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;   // low mantissa bits of t = round(x) (two's complement)
  const float fi = t - 12582912.0f;
  const float f = x - fi;
  float p = fmaf(0.05517105013132095f, f, 0.24260960519313812f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999281764030457f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));  // p * 2^round(x)
}
// which of every 8 consecutive columns use ex2_poly, for kPoly = 0..4 offloaded columns out of 8
__device__ __forceinline__ constexpr bool use_poly(int kPoly, int i) {
  return kPoly == 1 ? (i % 8 == 7)
       : kPoly == 2 ? (i % 4 == 3)
       : kPoly == 3 ? (i % 8 == 2 || i % 8 == 5 || i % 8 == 7)
       : kPoly == 4 ? (i % 2 == 1)
                    : false;
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

__device__ __forceinline__ void fma2g(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "mov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
// two polynomial exp2 at once with packed FMA-pipe instructions (10 issue slots per pair instead of 18)
__device__ __forceinline__ void ex2_poly2(float& r0, float& r1, float x0, float x1) {
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  float t0, t1, g0, g1, f0, f1, p0, p1;
  fadd2(t0, t1, x0, x1, 12582912.0f, 12582912.0f);
  fadd2(g0, g1, t0, t1, -12582912.0f, -12582912.0f);        // round(x)
  fma2g(f0, f1, g0, g1, -1.0f, -1.0f, x0, x1);               // x - round(x)
  fma2g(p0, p1, f0, f1, 0.05517105013132095f, 0.05517105013132095f, 0.24260960519313812f, 0.24260960519313812f);
  fma2g(p0, p1, p0, p1, f0, f1, 0.6932609677314758f, 0.6932609677314758f);
  fma2g(p0, p1, p0, p1, f0, f1, 0.9999281764030457f, 0.9999281764030457f);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
// packed variant: which PAIRS (e, e+1), e even, take the polynomial path
__device__ __forceinline__ constexpr bool pair_poly(int kPoly, int e) {
  return kPoly == 1 ? (e % 16 == 14)
       : kPoly == 2 ? (e % 8 == 6)
       : kPoly == 3 ? (e % 16 == 2 || e % 8 == 6)
       : kPoly == 4 ? (e % 4 == 2)
                    : false;
}

template <int REGS>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
}
template <int REGS>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}

// kTrace: CTA (0,0,0) records clock64() at the pipeline events of its first 16 kv iterations
// (trace[role][iter][event], role 0 = MMA issuer, 1 / 2 = softmax warpgroup 0 / 1) -- tools/attn_trace.py.
#define CWM_TRACE(role, it, ev)                                                              \
  do {                                                                                       \
    if (kTrace && trace_on && (it) < 16 && (threadIdx.x & 31) == 0) trace[((role) * 16 + (it)) * 8 + (ev)] = clock64(); \
  } while (0)

template <bool kTrace, int kPoly, bool kPacked>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_f16_kernel(const __grid_constant__ CUtensorMap tma_qkv, int N, int H, __half* __restrict__ out,
                     float scale_log2, long long* __restrict__ trace) {
  const bool trace_on = kTrace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
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

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
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
    // The producer / MMA roles run warp-uniformly; only the TMA / tcgen05 instructions are issued by one elected
    // lane (see the note in gemm.cu: a role under `if (lane == 0)` costs ~100 cycles of issue per MMA).
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, (t1_valid ? 2 : 1) * kTileBytes);
        tma_load_2d(smem_q, &tma_qkv, q_full, h * 64, row_base + q0);
        if (t1_valid) tma_load_2d(smem_q + kTileBytes, &tma_qkv, q_full, h * 64, row_base + q0 + 128);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[stage], kTileBytes);
          tma_load_2d(smem_k + stage * kTileBytes, &tma_qkv, &k_full[stage], C + h * 64, row_base + j * 128);
        }
        __syncwarp();
        mbar_wait(&v_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[stage], kTileBytes);
          tma_load_2d(smem_v + stage * kTileBytes, &tma_qkv, &v_full[stage], 2 * C + h * 64, row_base + j * 128);
        }
        __syncwarp();
        if (++stage == kKVStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, 0, 1);  // B = V is MN-major
      const uint32_t tm_o0 = tmem_base + 384, tm_o1 = tmem_base + 448;
      // Descriptors are built once; a stage / k-step only adds to the 14-bit start-address field (>> 4 units).
      const uint64_t dq0 = umma_desc_kmajor_sw128(smem_u32(smem_q));
      const uint64_t dq1 = umma_desc_kmajor_sw128(smem_u32(smem_q + kTileBytes));
      const uint64_t dk_base = umma_desc_kmajor_sw128(smem_u32(smem_k));
      const uint64_t dv_base = umma_desc_mnmajor_sw128(smem_u32(smem_v), 0);
      const uint32_t idesc_full = umma_idesc_f16(128, 128, 0, 0);
      const uint32_t idesc_last = umma_idesc_f16(128, cols_last, 0, 0);
      auto issue_qk = [&](uint64_t dq, uint64_t dk, uint32_t tm_s, uint32_t idesc_qk, uint64_t* bar) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tm_s, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
          umma_commit(bar);
        }
        __syncwarp();
      };
      // 16 kv rows per MMA: 8 TMEM columns of packed f16 pairs, 2048 B of V
      auto issue_pv = [&](uint32_t tm_p, uint64_t dv, uint32_t tm_o, bool accumulate, int ksteps, uint64_t* bar) {
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k)
            umma_ts(tm_o, tm_p + k * 8, dv + 128 * k, idesc_pv, (accumulate || k != 0) ? 1u : 0u);
          umma_commit(bar);
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
      };
      auto sbuf = [&](int j, int t) { return tmem_base + static_cast<uint32_t>(((2 * j + t) % 3) * 128); };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      {
        const uint32_t id0 = (n_kv == 1) ? idesc_last : idesc_full;
        issue_qk(dq0, dk_base, sbuf(0, 0), id0, &s_full[0]);
        if (t1_valid) issue_qk(dq1, dk_base, sbuf(0, 1), id0, &s_full[1]);
        commit(&k_empty[0]);
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
        const uint32_t id_next = (j + 2 == n_kv) ? idesc_last : idesc_full;
        const int ksteps = ((j + 1 == n_kv) ? cols_last : 128) >> 4;
        const uint64_t dv = dv_base + static_cast<uint64_t>(stage * (kTileBytes >> 4));
        const uint64_t dkn = dk_base + static_cast<uint64_t>(nstage * (kTileBytes >> 4));
        // ---- S_0(j+1) first: it lands in the buffer freed by PV_1(j-1), so tile 0 never waits for its scores
        if (has_next) {
          mbar_wait(&k_full[nstage], nphase);
          tc_fence_after();
          CWM_TRACE(0, j, 0);
          issue_qk(dq0, dkn, sbuf(j + 1, 0), id_next, &s_full[0]);
          CWM_TRACE(0, j, 1);
        }
        // ---- O_0 += P_0(j) V_j
        mbar_wait(&p_full[0], j & 1);
        mbar_wait(&v_full[stage], phase);
        tc_fence_after();
        CWM_TRACE(0, j, 2);
        issue_pv(sbuf(j, 0), dv, tm_o0, j > 0, ksteps, &pv_done[0]);
        CWM_TRACE(0, j, 3);
        if (t1_valid) {
          // ---- S_1(j+1) reuses the buffer P_0(j) just vacated (in-order execution after PV_0(j))
          if (has_next) issue_qk(dq1, dkn, sbuf(j + 1, 1), id_next, &s_full[1]);
          CWM_TRACE(0, j, 4);
          mbar_wait(&p_full[1], j & 1);
          tc_fence_after();
          CWM_TRACE(0, j, 5);
          issue_pv(sbuf(j, 1), dv, tm_o1, j > 0, ksteps, &pv_done[1]);
          CWM_TRACE(0, j, 6);
        }
        commit(&v_empty[stage]);
        if (has_next) commit(&k_empty[nstage]);
        stage = nstage;
        phase = nphase;
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
        const bool tr = (quad == 0 && lane == 0);
        if (tr) CWM_TRACE(1 + t, j, 0);
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        if (tr) CWM_TRACE(1 + t, j, 1);
        uint32_t s[128];
        // all four chunks are always read (columns >= ncols hold stale data and are masked below); only the
        // exp / P-store work is skipped for chunks the last, ragged MMA did not produce
        tmem_ld_x32(tm_s + 0, s);
        tmem_ld_x32(tm_s + 32, s + 32);
        tmem_ld_x32(tm_s + 64, s + 64);
        tmem_ld_x32(tm_s + 96, s + 96);
        tmem_ld_wait();
        if (tr) CWM_TRACE(1 + t, j, 2);
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
        if (tr) CWM_TRACE(1 + t, j, 3);
        const float neg_m = -m_ref * scale_log2;
        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 32 < ncols) {
            uint32_t p[16];  // 32 kv elements as packed f16 pairs; P aliases S columns that are already in registers
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const int e = c * 32 + i;
              float x0, x1, x2, x3;
              if (kPacked) {
                ffma2(x0, x1, __uint_as_float(s[e]), __uint_as_float(s[e + 1]), scale_log2, neg_m);
                ffma2(x2, x3, __uint_as_float(s[e + 2]), __uint_as_float(s[e + 3]), scale_log2, neg_m);
              } else {
                x0 = fmaf(__uint_as_float(s[e]), scale_log2, neg_m);
                x1 = fmaf(__uint_as_float(s[e + 1]), scale_log2, neg_m);
                x2 = fmaf(__uint_as_float(s[e + 2]), scale_log2, neg_m);
                x3 = fmaf(__uint_as_float(s[e + 3]), scale_log2, neg_m);
              }
              float p0, p1, p2, p3;
              if (kPacked) {
                if (pair_poly(kPoly, e)) {
                  ex2_poly2(p0, p1, x0, x1);
                } else {
                  p0 = ex2(x0);
                  p1 = ex2(x1);
                }
                if (pair_poly(kPoly, e + 2)) {
                  ex2_poly2(p2, p3, x2, x3);
                } else {
                  p2 = ex2(x2);
                  p3 = ex2(x3);
                }
              } else {
                p0 = use_poly(kPoly, e) ? ex2_poly(x0) : ex2(x0);
                p1 = use_poly(kPoly, e + 1) ? ex2_poly(x1) : ex2(x1);
                p2 = use_poly(kPoly, e + 2) ? ex2_poly(x2) : ex2(x2);
                p3 = use_poly(kPoly, e + 3) ? ex2_poly(x3) : ex2(x3);
              }
              if (kPacked) {
                fadd2(sum0, sum1, sum0, sum1, p0, p1);
                fadd2(sum2, sum3, sum2, sum3, p2, p3);
              } else {
                sum0 += p0;
                sum1 += p1;
                sum2 += p2;
                sum3 += p3;
              }
              p[(i >> 1)] = pack_half2(p0, p1);
              p[(i >> 1) + 1] = pack_half2(p2, p3);
            }
            tmem_st_x16(tm_s + c * 16, p);
          }
        }
        if (tr) CWM_TRACE(1 + t, j, 4);
        tmem_st_wait();
        l_sum += (sum0 + sum1) + (sum2 + sum3);
        tc_fence_before();
        __syncwarp();
        if (tr) CWM_TRACE(1 + t, j, 5);
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

// Debug hook (not part of the public header): device buffer of 3*16*8 int64 that receives the pipeline trace of
// CTA (0,0,0) of every following cwm_attention_f16 call; NULL switches tracing off.
static long long* g_attn_trace = nullptr;
extern "C" void cwm_debug_attention_trace(long long* device_buffer) { g_attn_trace = device_buffer; }
// Debug hook: how many of every 8 score columns take the polynomial exp2 (0..4); default kDefaultPoly.
static int g_attn_poly = kDefaultPoly;
static int g_attn_packed = 1;  // FFMA2 / FADD2 variant is the default (+4 % alone, +13 % with kPoly = 2)
extern "C" void cwm_debug_attention_poly(int eighths) {
  g_attn_packed = eighths >= 10;  // 10 + e selects the FFMA2 / FADD2 variant
  if (eighths >= 10) eighths -= 10;
  g_attn_poly = eighths < 0 ? 0 : (eighths > 4 ? 4 : eighths);
}

extern "C" int cwm_attention_f16(const uint16_t* qkv, int B, int N, int H, int head_dim, uint16_t* out,
                                 cwm_stream_t stream) {
  CWM_REQUIRE(qkv && out, "cwm_attention_f16: null pointer");
  CWM_REQUIRE(B >= 0 && N > 0 && H > 0, "cwm_attention_f16: bad shape B=%d N=%d H=%d", B, N, H);
  if (head_dim != 64) {  // 32 / 96 / 128 / 192: the generic warp-level tensor-core kernel (attn_mma.cu)
    const int A = H * head_dim;
    return cwm_attention_generic_f16(qkv, qkv + A, qkv + 2 * A, 3 * A, 3 * A, 3 * A, head_dim, head_dim, head_dim, B, N, N,
                                     H, head_dim, out, A, nullptr, 0, stream);
  }
  CWM_REQUIRE(H <= 65535 && B <= 65535, "cwm_attention_f16: grid too large");
  if (B == 0) return CWM_OK;
  using KernelFn = void (*)(const CUtensorMap, int, int, __half*, float, long long*);
  static const KernelFn kernels[10] = {
      attention_f16_kernel<false, 0, false>, attention_f16_kernel<false, 1, false>, attention_f16_kernel<false, 2, false>,
      attention_f16_kernel<false, 3, false>, attention_f16_kernel<false, 4, false>, attention_f16_kernel<false, 0, true>,
      attention_f16_kernel<false, 1, true>,  attention_f16_kernel<false, 2, true>,  attention_f16_kernel<false, 3, true>,
      attention_f16_kernel<false, 4, true>};
  static bool attr_set = false;
  if (!attr_set) {
    for (int i = 0; i < 10; ++i)
      CWM_CUDA_CHECK(cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    CWM_CUDA_CHECK(cudaFuncSetAttribute(attention_f16_kernel<true, kDefaultPoly, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    attr_set = true;
  }
  const int C = H * 64;
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, qkv, CWM_TMAP_F16, static_cast<uint64_t>(B) * N, 3ull * C, 3ull * C, 128, 64);
  if (rc) return rc;
  ProfileScope prof(static_cast<cudaStream_t>(stream), "attention_f16", 4.0 * B * H * static_cast<double>(N) * N * 64,
                    static_cast<double>(B) * N * C * 2.0 * 4.0);
  dim3 grid((N + 255) / 256, H, B);
  if (g_attn_trace != nullptr)
    attention_f16_kernel<true, kDefaultPoly, true><<<grid, kAttnThreads, kAttnSmemBytes, static_cast<cudaStream_t>(stream)>>>(
        tm, N, H, reinterpret_cast<__half*>(out), kLog2e, g_attn_trace);
  else
    kernels[g_attn_poly + (g_attn_packed ? 5 : 0)]<<<grid, kAttnThreads, kAttnSmemBytes, static_cast<cudaStream_t>(stream)>>>(
        tm, N, H, reinterpret_cast<__half*>(out), kLog2e, nullptr);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}
