// attention.cu -- softmax(q k^T) v, head_dim 64, unmasked, f16 operands, fp32 softmax statistics and fp32
// accumulation in tensor memory (a6: cwm/models/VideoMAE/utils.py:108-113).
//
// Flash-style, warp-specialised, TMA-pipelined (sm_100a).  One work item = 256 query rows (two 128-row tiles) of one
// (sample, head); K/V are streamed once per item in 128-row tiles.
//   warp 0       TMA producer (Q, then K_j / V_j through a kKVStages-deep mbarrier ring)
//   warp 1       MMA issuer: S_t = Q_t K_j^T (SS, M=128 N=128 K=64) and O_t += P_t V_j (TS: A = P from tensor
//                memory, B = V from smem MN-major, M=128 N=64 K=128), interleaved over the two query tiles so the
//                tensor core works on one tile while the other tile is in softmax
//   warps 4-7    softmax warpgroup for query tile 0, warps 8-11 for tile 1: one thread per query row (no
//                shuffles): tcgen05.ld S row -> running max with lazy rescale (O is only rescaled when the max
//                grows by more than 2^8) -> exp2 (packed FFMA2 / FADD2 math, 2 of every 8 on the FMA pipe through a
//                degree-3 polynomial) -> fp32 row sum -> f16 P written back over S in tensor memory
// Tensor memory map (512 columns): three 128-column score buffers used round-robin -- the scores of (query tile t,
// kv step g) live in buffer (2g + t) % 3 and P (f16 pairs) overwrites their first 64 columns -- so that S_0(g+1) is
// computed while the softmax of S_0(g) is still running; O0 [384,448), O1 [448,512).
// The last kv tile is issued with N = round_up(valid columns, 32), so ragged sequence lengths (788, 3140 ...) do
// not pay for a full 128-column tile.
//
// Synchronisation discipline (what tools/debug/attn_stress.py checks).  The three actors are deliberately allowed
// to drift: the issuer runs up to two score tiles ahead of a softmax warpgroup, and the four warps of a warpgroup
// are not synchronised with each other.  Therefore
//   * every producer/consumer barrier whose consumer may lag by more than one phase exists TWICE and alternates by
//     step parity (s_full, p_full, pv_done): a waiter can then never see a barrier two phases ahead (which a parity
//     wait cannot distinguish from "not yet"), and a fast warp can never arrive twice in one phase;
//   * a score tile is written into a buffer only after the PV product that read P from that buffer has COMPLETED
//     (the issuer waits on pv_done), instead of relying on in-order execution inside the tensor pipe (war_safe).
// The first version of this kernel had a single barrier per hand-off; it produced wrong rows (max error > 1) for
// logits with a realistic spread at N = 788 / 1568 once more CTAs than SMs were in flight -- the parity tests, whose
// random-init logits never trigger the rescale path or much warp drift, did not see it.

#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace cwm {

constexpr int kAttnThreads = 384;
constexpr int kKVStages = 4;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // in log2 units
constexpr int kDefaultPoly = 2;            // columns out of 8 whose exp2 runs on the FMA pipe (tools/kernel_bench.py sweep)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA / ALU pipes (no MUFU), see ex2_poly2 below: Cody-Waite range reduction with the 1.5 * 2^23 magic
// constant and a degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, ~6x below the f16
// rounding of P).  head dim 64 needs one exp per 256 tensor FLOPs and the SM has only 16 MUFU lanes, so a fixed
// fraction of the elements of every row takes this path to balance the MUFU pipe against the FMA pipe.

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

// ---------------------------------------------------------------------------------------------------------
// The kernel.  Default launch: persistent -- one CTA per SM walks a range of (q block, head, sample) work items
// and software-pipelines ACROSS items -- the producer prefetches the next item's Q (double-buffered) and K/V through
// the same ring, the MMA warp issues S(next item, 0) while the softmax of the current item's last tile is still
// running, and the softmax warpgroups normalise / store O of item i while the tensor core already works on item
// i+1.  This hides the per-CTA launch, TMEM allocation, pipeline ramp-up and epilogue that dominate short sequences
// (N = 788: 7 kv tiles per item).  Items of one (sample, head) are consecutive, so K/V are re-read from L2, and a
// contiguous range keeps DRAM traffic at the algorithmic bytes.
// All mbarrier phases are driven by running counters: g = global kv step of this CTA (S-buffer rotation, K/V ring),
// g_t = steps that involved query tile t (tile 1 is absent in the last, <= 128-row q block of a sequence).
// O_t needs no extra barrier: the softmax warpgroup reads O_t of item i (epilogue) before it signals p_full[t] for
// step 0 of item i+1, and PV_t(i+1, 0) -- the first MMA that overwrites O_t -- is only issued after that p_full.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPersistSmemBytes = 1024 + (4 + 2 * kKVStages) * kTileBytes + 512;

struct ItemCoord {
  int h, row_base, q0;
  bool t1;
};
__device__ __forceinline__ ItemCoord decode_item(int item, int nq, int H, int N) {
  const int grp = item / nq;
  const int qb = item - grp * nq;
  ItemCoord c;
  c.h = grp % H;
  c.row_base = (grp / H) * N;
  c.q0 = qb * 256;
  c.t1 = (c.q0 + 128) < N;
  return c;
}

template <int kPoly, bool kDbg, bool kCompact>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_persist_kernel(const __grid_constant__ CUtensorMap tma_qkv, int N, int H, int B, __half* __restrict__ out,
                         float scale_log2, int strided, int war_safe, int stale_max) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                               // 2 buffers x 2 tiles
  uint8_t* smem_k = smem + 4 * kTileBytes;              // kKVStages tiles
  uint8_t* smem_v = smem_k + kKVStages * kTileBytes;    // kKVStages tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + kKVStages * kTileBytes);
  uint64_t* q_full = bars;                    // 2
  uint64_t* q_empty = bars + 2;               // 2
  uint64_t* k_full = bars + 4;                // kKVStages
  uint64_t* k_empty = k_full + kKVStages;
  uint64_t* v_full = k_empty + kKVStages;
  uint64_t* v_empty = v_full + kKVStages;
  uint64_t* s_full = v_empty + kKVStages;     // 2 tiles x 2 (alternating by step parity, see below)
  uint64_t* p_full = s_full + 4;              // 2 tiles x 2 (alternating by step parity)
  uint64_t* pv_done = p_full + 4;             // 2 tiles x 2 (alternating by step parity)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int C = H * 64;
  // kDbg: every mbarrier wait gives up after ~2^26 polls, reports who waited for what and traps (debug hook only)
  auto WAIT = [&](uint64_t* bar, uint32_t parity, int tag, int step) {
    if (!kDbg) {
      mbar_wait(bar, parity);
    } else {
      long long n = 0;
      while (!mbar_try_wait(bar, parity)) {
        if (++n > (1ll << 24)) {
          if (lane == 0)
            printf("attention_persist: block %d warp %d stuck tag %d step %d parity %u\n", blockIdx.x, warp, tag, step, parity);
          __trap();
        }
      }
    }
  };
  const int n_kv = (N + 127) / 128;
  const int cols_last = ((N - (n_kv - 1) * 128) + 31) & ~31;
  // ---- this CTA's contiguous, cost-balanced item range (a q block with both tiles costs 2, a half block 1) ----
  const int nq = (N + 255) / 256;
  const int cost_grp = 2 * nq - (((N - (nq - 1) * 256) <= 128) ? 1 : 0);
  const long long total_cost = static_cast<long long>(H) * B * cost_grp;
  auto first_item = [&](long long t) -> int {
    const long long g = t / cost_grp;
    const int r = static_cast<int>(t - g * cost_grp);
    return static_cast<int>(g * nq + ((r + 1) >> 1));
  };
  // Two item -> CTA maps.  Contiguous (short sequences): CTA c walks a cost-balanced contiguous range, so the K/V of a
  // (sample, head) are re-read from L2 by the same SM.  Strided (long sequences, where the K/V of 148 concurrent
  // (sample, head) groups would overflow L2): item = c + i * gridDim.x, so that the CTAs running at the same time
  // share a few (sample, head) groups, exactly like a plain grid launch.
  const int n_items = nq * H * B;
  int item_first, item_step, n_my;
  if (strided) {
    item_first = blockIdx.x;
    item_step = gridDim.x;
    n_my = (n_items > static_cast<int>(blockIdx.x)) ? (n_items - 1 - static_cast<int>(blockIdx.x)) / item_step + 1 : 0;
  } else {
    item_first = first_item(total_cost * blockIdx.x / gridDim.x);
    item_step = 1;
    n_my = first_item(total_cost * (blockIdx.x + 1) / gridDim.x) - item_first;
  }
  auto item_of = [&](int e) { return item_first + e * item_step; };

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tma_qkv);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[2 * i], 1);
      mbar_init(&s_full[2 * i + 1], 1);
      mbar_init(&p_full[2 * i], 4);  // one elected lane per softmax warp
      mbar_init(&p_full[2 * i + 1], 4);
      mbar_init(&pv_done[2 * i], 1);
      mbar_init(&pv_done[2 * i + 1], 1);
    }
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
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
  grid_dep_sync();   // programmatic dependent launch: the set-up above overlapped the previous kernel's drain

  if (warp < 4) {
    reg_dec<64>();
    if (warp == 0) {
      // ===================== TMA producer =====================
      int g = 0;  // global kv step
      for (int e = 0; e < n_my; ++e) {  // e = item ordinal
        const ItemCoord ic = decode_item(item_of(e), nq, H, N);
        const int qb = e & 1;
        WAIT(&q_empty[qb], ((e >> 1) & 1) ^ 1, 1, e);
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[qb], (ic.t1 ? 2 : 1) * kTileBytes);
          tma_load_2d(smem_q + (2 * qb) * kTileBytes, &tma_qkv, &q_full[qb], ic.h * 64, ic.row_base + ic.q0);
          if (ic.t1)
            tma_load_2d(smem_q + (2 * qb + 1) * kTileBytes, &tma_qkv, &q_full[qb], ic.h * 64, ic.row_base + ic.q0 + 128);
        }
        __syncwarp();
        for (int j = 0; j < n_kv; ++j, ++g) {
          const int stage = g % kKVStages;
          const uint32_t phase = (g / kKVStages) & 1;
          WAIT(&k_empty[stage], phase ^ 1, 2, g);
          if (elect_one()) {
            mbar_arrive_expect_tx(&k_full[stage], kTileBytes);
            tma_load_2d(smem_k + stage * kTileBytes, &tma_qkv, &k_full[stage], C + ic.h * 64, ic.row_base + j * 128);
          }
          __syncwarp();
          WAIT(&v_empty[stage], phase ^ 1, 3, g);
          if (elect_one()) {
            mbar_arrive_expect_tx(&v_full[stage], kTileBytes);
            tma_load_2d(smem_v + stage * kTileBytes, &tma_qkv, &v_full[stage], 2 * C + ic.h * 64, ic.row_base + j * 128);
          }
          __syncwarp();
        }
      }
    } else if (warp == 1 && n_my > 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, 0, 1);  // B = V is MN-major
      const uint32_t tm_o0 = tmem_base + 384, tm_o1 = tmem_base + 448;
      const uint64_t dq_base = umma_desc_kmajor_sw128(smem_u32(smem_q));
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
      auto sbuf = [&](int gs, int t) { return tmem_base + static_cast<uint32_t>(((2 * gs + t) % 3) * 128); };
      auto dq_of = [&](int e, int t) { return dq_base + static_cast<uint64_t>(((e & 1) * 2 + t) * (kTileBytes >> 4)); };

      // The issuer runs up to TWO score tiles ahead of a softmax warpgroup (S(g+1) is issued before it waits for
      // P(g), and S(g+2) right after): with a single barrier per tile the warpgroup could find the barrier two phases
      // ahead -- indistinguishable from "not yet" by parity -- whenever it is delayed (e.g. by an item's epilogue).
      // Scores of tile t therefore alternate between two barriers, s_full[2t + (n & 1)], n = running tile-t step.
      // p_full alternates the same way for a different reason: the four warps of a softmax warpgroup are not
      // synchronised with each other, and because S(g+1) is available before P(g) is consumed a fast warp could
      // arrive for step g+1 before a slow warp has arrived for step g -- on one barrier those two arrivals would
      // complete phase g without the slow warp.  With two barriers a warp can only arrive on the same barrier again
      // after S(g+2) exists, i.e. after the issuer has seen all four arrivals of step g.
      int n_s0 = 1, n_s1 = 0;
      // ---- prologue: scores of the first step ----
      {
        const ItemCoord ic = decode_item(item_of(0), nq, H, N);
        WAIT(&q_full[0], 0, 4, 0);
        WAIT(&k_full[0], 0, 5, 0);
        tc_fence_after();
        const uint32_t id0 = (n_kv == 1) ? idesc_last : idesc_full;
        issue_qk(dq_of(0, 0), dk_base, sbuf(0, 0), id0, &s_full[0]);
        if (ic.t1) {
          issue_qk(dq_of(0, 1), dk_base, sbuf(0, 1), id0, &s_full[2]);
          n_s1 = 1;
        }
        commit(&k_empty[0]);
        if (n_kv == 1) commit(&q_empty[0]);
      }
      int g = 0, g1 = 0;
      for (int e = 0; e < n_my; ++e) {
        const bool t1_cur = decode_item(item_of(e), nq, H, N).t1;
        for (int j = 0; j < n_kv; ++j, ++g) {
          const bool same_item = (j + 1) < n_kv;
          const bool has_next = same_item || (e + 1) < n_my;
          const int next_j = same_item ? j + 1 : 0;
          const int next_e = same_item ? e : e + 1;
          const bool t1_next = has_next && (same_item ? t1_cur : decode_item(item_of(e + 1), nq, H, N).t1);
          const int stage = g % kKVStages;
          const uint32_t phase = (g / kKVStages) & 1;
          const int nstage = (g + 1) % kKVStages;
          const uint32_t nphase = ((g + 1) / kKVStages) & 1;
          const uint32_t id_next = (next_j == n_kv - 1) ? idesc_last : idesc_full;
          const int ksteps = ((j + 1 == n_kv) ? cols_last : 128) >> 4;
          const uint64_t dv = dv_base + static_cast<uint64_t>(stage * (kTileBytes >> 4));
          const uint64_t dkn = dk_base + static_cast<uint64_t>(nstage * (kTileBytes >> 4));
          // ---- S_0(next) first: it lands in the buffer freed by PV_1 of the previous step
          if (has_next) {
            if (next_j == 0) WAIT(&q_full[next_e & 1], (next_e >> 1) & 1, 6, g);
            WAIT(&k_full[nstage], nphase, 7, g);
            // war_safe: S_0(g+1) overwrites the buffer P_1 of the previous tile-1 step was read from -- wait until
            // that PV has COMPLETED instead of relying on in-order execution of the MMA pipe
            if (war_safe && g1 > 0) WAIT(&pv_done[2 + ((g1 - 1) & 1)], ((g1 - 1) >> 1) & 1, 17, g);
            tc_fence_after();
            issue_qk(dq_of(next_e, 0), dkn, sbuf(g + 1, 0), id_next, &s_full[n_s0 & 1]);
            ++n_s0;
          }
          // ---- O_0 += P_0 V
          WAIT(&p_full[g & 1], (g >> 1) & 1, 8, g);
          WAIT(&v_full[stage], phase, 9, g);
          tc_fence_after();
          issue_pv(sbuf(g, 0), dv, tm_o0, j > 0, ksteps, &pv_done[g & 1]);
          // ---- S_1(next) reuses the buffer P_0 just vacated (in-order execution after PV_0)
          if (t1_next) {
            if (war_safe) {  // S_1(g+1) overwrites the buffer PV_0(g) reads P_0(g) from
              WAIT(&pv_done[g & 1], (g >> 1) & 1, 18, g);
              tc_fence_after();
            }
            issue_qk(dq_of(next_e, 1), dkn, sbuf(g + 1, 1), id_next, &s_full[2 + (n_s1 & 1)]);
            ++n_s1;
          }
          if (t1_cur) {
            WAIT(&p_full[2 + (g1 & 1)], (g1 >> 1) & 1, 10, g);
            tc_fence_after();
            issue_pv(sbuf(g, 1), dv, tm_o1, j > 0, ksteps, &pv_done[2 + (g1 & 1)]);
            ++g1;
          }
          commit(&v_empty[stage]);
          if (has_next) {
            commit(&k_empty[nstage]);
            if (next_j == n_kv - 1) commit(&q_empty[next_e & 1]);  // the last scores that read this Q buffer are issued
          }
        }
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    reg_inc<208>();
    const int t = (warp - 4) >> 2;          // query tile
    const int quad = warp & 3;              // TMEM lane quadrant
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tm_o = tmem_base + lane_off + 384 + t * 64;
    int g = 0;   // global kv step of the first tile of the current item
    int gt = 0;  // steps this tile took part in (barrier phases)
    for (int e = 0; e < n_my; ++e, g += n_kv) {
      const ItemCoord ic = decode_item(item_of(e), nq, H, N);
      if (t == 1 && !ic.t1) continue;
      const int q_row = ic.q0 + t * 128 + row_in_tile;  // row within the sample
      float m_ref = 0.f;   // running reference max (raw score units)
      float l_sum = 0.f;   // running sum of exp2((s - m_ref) * scale_log2)
      for (int j = 0; j < n_kv; ++j, ++gt) {
        const uint32_t tm_s = tmem_base + lane_off + static_cast<uint32_t>(((2 * (g + j) + t) % 3) * 128);
        const int ncols = (j + 1 == n_kv) ? cols_last : 128;  // columns the MMA produced for this tile
        const int kv_valid = N - j * 128;                     // columns >= kv_valid are padding
        WAIT(&s_full[2 * t + (gt & 1)], (gt >> 1) & 1, 11 + t, gt);
        tc_fence_after();
        if ((stale_max & 2) && ic.q0 + t * 128 + quad * 32 >= N) {
          // every query row of this warp lies beyond the sequence (the ragged last tile: N = 788 leaves 20 rows, N = 1568
          // leaves 32): its rows of P feed rows of O that are never stored, so the tile's exponentials are skipped and the
          // buffer is handed on as it is -- the XU / issue slots go to the warps that have rows
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[2 * t + (gt & 1)]);
          continue;
        }
        uint32_t s[128];
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
        // Row maximum of this tile.  Only the FIRST tile of an item needs it before the exponentials; afterwards the
        // exponentials use the running reference m_ref as it stands (the lazy rescale keeps it within 2^8 of the true
        // maximum in all but rare tiles), the tile maximum is reduced in the same instruction stream (ALU pipe, beside the
        // MUFU / FMA work), and only if it turns out to exceed m_ref by more than the threshold the tile is redone with
        // the new reference.  This removes the serialised "load everything -> max -> exp" phase during which a warp keeps
        // the XU pipe idle.
        auto tile_max = [&]() {
          float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]);
          float mx2 = __uint_as_float(s[2]), mx3 = __uint_as_float(s[3]);
#pragma unroll
          for (int i = 4; i < 128; i += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(s[i]));
            mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
            mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
          }
          return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        };
        // exp2((s - m_ref) * scale) for the whole tile -> f16 P over S in tensor memory; returns the fp32 row sum
        auto exp_tile = [&](float neg_m) {
          float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c * 32 < ncols) {
              uint32_t p[16];  // 32 kv elements as packed f16 pairs; P aliases S columns that are already in registers
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const int el = c * 32 + i;
                float x0, x1, x2, x3, p0, p1, p2, p3;
                ffma2(x0, x1, __uint_as_float(s[el]), __uint_as_float(s[el + 1]), scale_log2, neg_m);
                ffma2(x2, x3, __uint_as_float(s[el + 2]), __uint_as_float(s[el + 3]), scale_log2, neg_m);
                if (pair_poly(kPoly, el)) {
                  ex2_poly2(p0, p1, x0, x1);
                } else {
                  p0 = ex2(x0);
                  p1 = ex2(x1);
                }
                if (pair_poly(kPoly, el + 2)) {
                  ex2_poly2(p2, p3, x2, x3);
                } else {
                  p2 = ex2(x2);
                  p3 = ex2(x3);
                }
                fadd2(sum0, sum1, sum0, sum1, p0, p1);
                fadd2(sum2, sum3, sum2, sum3, p2, p3);
                p[(i >> 1)] = pack_half2(p0, p1);
                p[(i >> 1) + 1] = pack_half2(p2, p3);
              }
              tmem_st_x16(tm_s + c * 16, p);
            }
          }
          return (sum0 + sum1) + (sum2 + sum3);
        };
        // O_t *= f, l *= f for the rows whose reference maximum moved (all lanes take part: the TMEM accesses are warp-wide)
        auto rescale = [&](float row_max, bool need) {
          const float f = need ? ex2((m_ref - row_max) * scale_log2) : 1.0f;
          if (need) m_ref = row_max;
          l_sum *= f;
          // O_t is quiescent once PV_t of the previous step has completed (PV_t of this step is not issued before
          // this warpgroup signals p_full[t]).  pv_done alternates between two barriers by step parity: this warp
          // does not observe every completion (the wait is lazy), and a warp may be a full step ahead of the
          // slowest warp of its group, so on a single barrier "PV(gt-2) still pending" and "PV(gt-1) done" would
          // have the same parity.  On barrier (gt-1)&1 the previous completion is PV(gt-3), which is known to be
          // complete because S(gt-1) -- issued after it and already consumed by this warp -- has completed.
          WAIT(&pv_done[2 * t + ((gt - 1) & 1)], ((gt - 1) >> 1) & 1, 13 + t, gt);
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
        };
        float tile_sum = 0.f;
        if constexpr (kCompact) {
        // One call site for the exponentials (the kernel's largest block of code: four inlined copies made it 68 KB of
        // SASS).  Before the pass: the first tile of an item takes its maximum as the reference; the exact (non-stale)
        // mode checks every tile's maximum up front.  After the pass, in stale mode: the tile maximum is checked, and in the
        // rare case it exceeds the reference by more than the threshold the tile is redone (second pass) after the rescale.
        if (j == 0) {
          m_ref = tile_max();
        } else if (!(stale_max & 1)) {
          const float row_max = tile_max();
          const bool need = (row_max - m_ref) * scale_log2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) rescale(row_max, need);
        }
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          tile_sum = exp_tile(-m_ref * scale_log2);
          if (pass == 1 || j == 0 || !(stale_max & 1)) break;
          const float row_max = tile_max();
          const bool need = (row_max - m_ref) * scale_log2 > kRescaleThreshold;
          if (!__any_sync(0xffffffffu, need)) break;   // the common case
          tmem_st_wait();
          rescale(row_max, need);
        }
        } else {
        if (j == 0) {
          m_ref = tile_max();
          tile_sum = exp_tile(-m_ref * scale_log2);
        } else if (stale_max & 1) {
          tile_sum = exp_tile(-m_ref * scale_log2);
          const float row_max = tile_max();
          const bool need = (row_max - m_ref) * scale_log2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {   // rare: the tile's exponentials were taken against a stale reference
            tmem_st_wait();
            rescale(row_max, need);
            tile_sum = exp_tile(-m_ref * scale_log2);
          }
        } else {
          const float row_max = tile_max();
          const bool need = (row_max - m_ref) * scale_log2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) rescale(row_max, need);
          tile_sum = exp_tile(-m_ref * scale_log2);
        }
        }
        tmem_st_wait();
        l_sum += tile_sum;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * t + (gt & 1)]);
      }
      // ---- finalise this item: O / l -> f16 -> global (the tensor core is already busy with the next item)
      WAIT(&pv_done[2 * t + ((gt - 1) & 1)], ((gt - 1) >> 1) & 1, 15 + t, gt);
      tc_fence_after();
      const float inv_l = 1.0f / l_sum;
      __half* orow = out + (static_cast<long long>(ic.row_base) + q_row) * C + ic.h * 64;
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

// Debug hook: how many of every 8 score columns take the polynomial exp2 (0..4); default kDefaultPoly.
static int g_attn_poly = kDefaultPoly;
extern "C" void cwm_debug_attention_poly(int eighths) {
  if (eighths >= 10) eighths -= 10;  // (the scalar-math variant selected by values < 10 no longer exists)
  g_attn_poly = eighths < 0 ? 0 : (eighths > 4 ? 4 : eighths);
}

// Debug hook: 1 (default) = persistent CTAs (one per SM) that pipeline across work items; 3 = one work item per CTA
// (same kernel, grid = items); 2 / 4 = modes 1 / 3 with watchdog waits (report a stuck mbarrier wait and trap).
static int g_attn_mode = 1;
static int g_attn_persist_map = -1;  // -1 = automatic, 0 = contiguous ranges, 1 = strided
static int g_attn_war_safe = 1;
static int g_attn_stale_max = 3;  // bit 0: exponentials against the running reference, tile maximum checked afterwards;
                                  // bit 1: warps whose 32 query rows all lie beyond the sequence skip the tile's exponentials
extern "C" void cwm_debug_attention_stale_max(int on) { g_attn_stale_max = (g_attn_stale_max & 2) | (on ? 1 : 0); }
extern "C" void cwm_debug_attention_skip_idle(int on) { g_attn_stale_max = (g_attn_stale_max & 1) | (on ? 2 : 0); }
extern "C" void cwm_debug_attention_war_safe(int on) { g_attn_war_safe = on; }
extern "C" void cwm_debug_attention_persistent(int mode) { g_attn_mode = (mode <= 0) ? 3 : mode; }
extern "C" void cwm_debug_attention_persist_map(int m) { g_attn_persist_map = m; }

extern "C" int cwm_attention_f16(const uint16_t* qkv, int B, int N, int H, int head_dim, uint16_t* out,
                                 cwm_stream_t stream) {
  CWM_REQUIRE(qkv && out, "cwm_attention_f16: null pointer");
  CWM_REQUIRE(B >= 0 && N > 0 && H > 0, "cwm_attention_f16: bad shape B=%d N=%d H=%d", B, N, H);
  if (head_dim != 64) {  // 32 / 96 / 128 / 192: the generic warp-level tensor-core kernel (attn_mma.cu)
    const int A = H * head_dim;
    return cwm_attention_generic_f16(qkv, qkv + A, qkv + 2 * A, 3 * A, 3 * A, 3 * A, head_dim, head_dim, head_dim, B, N, N,
                                     H, head_dim, out, A, nullptr, 0, stream);
  }
  if (B == 0) return CWM_OK;
  const long long n_items = static_cast<long long>((N + 255) / 256) * H * B;
  CWM_REQUIRE(n_items < (1ll << 31), "cwm_attention_f16: too many work items");
  using KernelFn = void (*)(const CUtensorMap, int, int, int, __half*, float, int, int, int);
  // [0..4] = exponentials-on-the-FMA-pipe share kPoly, [5] = the debug build, [6] = the default kPoly with the softmax tile
  // as ONE inlined call site (2864 instead of 4256 SASS instructions; CWM_ATTN_COMPACT=1)
  static const KernelFn kernels[7] = {attention_persist_kernel<0, false, false>, attention_persist_kernel<1, false, false>,
                                      attention_persist_kernel<2, false, false>, attention_persist_kernel<3, false, false>,
                                      attention_persist_kernel<4, false, false>, attention_persist_kernel<kDefaultPoly, true, false>,
                                      attention_persist_kernel<kDefaultPoly, false, true>};
  static bool attr_set = false;
  static int compact_env = 0;
  if (!attr_set) {
    const char* v = getenv("CWM_ATTN_COMPACT");
    compact_env = (v != nullptr) ? atoi(v) : 0;
    for (int i = 0; i < 7; ++i)
      CWM_CUDA_CHECK(cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kPersistSmemBytes));
    attr_set = true;
  }
  const int C = H * 64;
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, qkv, CWM_TMAP_F16, static_cast<uint64_t>(B) * N, 3ull * C, 3ull * C, 128, 64);
  if (rc) return rc;
  ProfileScope prof(static_cast<cudaStream_t>(stream), "attention_f16", 4.0 * B * H * static_cast<double>(N) * N * 64,
                    static_cast<double>(B) * N * C * 2.0 * 4.0);
  const bool multi = (g_attn_mode == 1 || g_attn_mode == 2);
  const int grid = multi ? static_cast<int>(n_items < num_sms() ? n_items : num_sms()) : static_cast<int>(n_items);
  // multi-item maps: contiguous ranges while the K/V of the concurrently resident (sample, head) groups fit in
  // (part of) the 126 MB L2, strided otherwise; one item per CTA is the strided map with grid = items
  const double kv_resident = 2.0 * N * 64 * 2 * grid;
  const int strided = !multi ? 1 : (g_attn_persist_map >= 0) ? g_attn_persist_map : (kv_resident > 48e6 ? 1 : 0);
  CWM_CUDA_CHECK(launch_pdl(kernels[(g_attn_mode == 2 || g_attn_mode == 4) ? 5 : ((compact_env && g_attn_poly == kDefaultPoly) ? 6 : g_attn_poly)],
                            dim3(grid), dim3(kAttnThreads), kPersistSmemBytes, static_cast<cudaStream_t>(stream), tm, N, H, B,
                            reinterpret_cast<__half*>(out), kLog2e, strided, g_attn_war_safe, g_attn_stale_max));
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}
