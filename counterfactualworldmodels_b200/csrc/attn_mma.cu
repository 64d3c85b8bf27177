// attn_mma.cu -- generic flash attention softmax(q k^T) v for the SMALL attentions of the conjoined
// (IMU-conditioned) model: head dims 32 / 96 / 128 / 192 and problems where one side has only 1..64 tokens:
//   * context-stream self-attention (25 / 26 / 50 IMU tokens, head dim 32: cwm/models/VideoMAE/utils.py:87-121
//     instantiated by conjoined_vmae.py:1198-1216),
//   * bidirectional cross-attention between the N main tokens and the M <= 64 context tokens
//     (cwm/models/transformer.py:314-378): "trg" direction = N queries x M keys, "src" direction = M queries x N
//     keys (split over the key axis, partial results merged by attn_combine_kernel).
// These are < 1 % of the FLOPs of a forward and their operand tiles (M <= 64 rows) are too small for a 128-row
// tcgen05 tile, so they run on the warp-level tensor-core path (mma.sync m16n8k16, f16 operands, fp32 accumulate,
// fp32 softmax statistics) -- the main 64-dim attention over 788..6336 tokens is the tcgen05 kernel in attention.cu.
//
// One CTA = kWarps * 16 query rows of one (sample, head); K/V are streamed in 64-row tiles through shared memory
// (rows padded by 16 bytes: conflict-free ldmatrix).  q is expected pre-scaled (the qkv GEMM epilogue applies the
// softmax scale), like cwm_attention_f16.
#include <cstdlib>

#include "common.cuh"

namespace cwm {

constexpr float kLog2eMma = 1.4426950408889634f;

struct AttnMmaParams {
  const __half* q;
  const __half* k;
  const __half* v;
  long long ldq, ldk, ldv;  // row strides (elements)
  int q_hs, k_hs, v_hs;     // column stride between heads (elements)
  int Nq, Nk, H;
  __half* out;
  long long ldo;
  int n_splits, kv_per_split;  // kv_per_split % 64 == 0
  float* part_o;               // [B, H, n_splits, Nq, HD]  (n_splits > 1)
  float* part_ml;              // [B, H, n_splits, Nq, 2]   running max (raw score units), sum
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." CWM_MMA_SYNC_TYPES ".f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// rows [row0, row0 + n_rows) of a [*, HD] f16 matrix (row stride ld) -> smem tile with padded rows; rows >= limit
// are zero-filled.
template <int HD, int kThreads>
__device__ __forceinline__ void load_tile(__half* dst, const __half* src, long long ld, int row0, int n_rows,
                                          int limit) {
  constexpr int kVec = HD / 8;       // uint4 per row
  constexpr int kStride = HD + 8;    // padded row (elements)
  for (int i = threadIdx.x; i < n_rows * kVec; i += kThreads) {
    const int r = i / kVec, c = i - r * kVec;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + r < limit) val = __ldg(reinterpret_cast<const uint4*>(src + (row0 + r) * ld) + c);
    *reinterpret_cast<uint4*>(dst + r * kStride + c * 8) = val;
  }
}

// ---- cp.async helpers (16-byte copies; src_bytes = 0 zero-fills the destination) ----
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <int HD, int kThreads>
__device__ __forceinline__ void load_tile_async(__half* dst, const __half* src, long long ld, int row0, int n_rows, int limit) {
  constexpr int kVec = HD / 8;
  constexpr int kStride = HD + 8;
  for (int i = threadIdx.x; i < n_rows * kVec; i += kThreads) {
    const int r = i / kVec, c = i - r * kVec;
    const bool in = row0 + r < limit;
    cp_async16(smem_u32(dst + r * kStride + c * 8), reinterpret_cast<const uint4*>(src + (in ? row0 + r : 0) * ld) + c, in ? 16 : 0);
  }
}

// kPipe: K / V tiles are double-buffered with cp.async (the next tile is in flight while the current one is multiplied) --
// the split-key "src" direction of the cross-attention, where a CTA streams 512 keys for a handful of queries and the
// un-pipelined load -> sync -> compute loop left the memory system idle most of the time.
template <int HD, int kWarps, int BN, bool kPipe = false>  // BN = key tile (64, or 32 when there are at most 32 keys: half the MMA work)
__global__ void __launch_bounds__(kWarps * 32)
attn_mma_kernel(AttnMmaParams p) {
  constexpr int kThreads = kWarps * 32;
  constexpr int BM = kWarps * 16;
  constexpr int NT = BN / 8;   // 8-key score tiles per warp row slab
  constexpr int kStride = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_raw_mma[];
  __half* Qs = reinterpret_cast<__half*>(smem_raw_mma);
  __half* Ks = Qs + BM * kStride;
  __half* Vs = Ks + BN * kStride;           // kPipe: stage 1 = the same pair 2 * BN rows further

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int b = blockIdx.z / p.n_splits;
  const int split = blockIdx.z - b * p.n_splits;
  const int q0 = blockIdx.x * BM;
  const int kv_begin = split * p.kv_per_split;
  const int kv_end = min(p.Nk, kv_begin + p.kv_per_split);

  const __half* qb = p.q + static_cast<long long>(b) * p.Nq * p.ldq + h * p.q_hs;
  const __half* kb = p.k + static_cast<long long>(b) * p.Nk * p.ldk + h * p.k_hs;
  const __half* vb = p.v + static_cast<long long>(b) * p.Nk * p.ldv + h * p.v_hs;

  if constexpr (kPipe) {
    load_tile_async<HD, kThreads>(Qs, qb, p.ldq, q0, BM, p.Nq);
    load_tile_async<HD, kThreads>(Ks, kb, p.ldk, kv_begin, BN, kv_end);
    load_tile_async<HD, kThreads>(Vs, vb, p.ldv, kv_begin, BN, kv_end);
    cp_async_commit();
  } else {
    load_tile<HD, kThreads>(Qs, qb, p.ldq, q0, BM, p.Nq);
  }

  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};  // rows lane/4 and lane/4 + 8 of this warp's 16-row slab
  float l_run[2] = {0.f, 0.f};

  const uint32_t q_addr = smem_u32(Qs + (warp * 16 + (lane & 15)) * kStride + (lane >> 4) * 8);
  const uint32_t k_addr0 = smem_u32(Ks + ((lane & 7) + ((lane >> 4) << 3)) * kStride + ((lane >> 3) & 1) * 8);
  const uint32_t v_addr0 = smem_u32(Vs + ((lane & 7) + (((lane >> 3) & 1) << 3)) * kStride + (lane >> 4) * 8);
  constexpr uint32_t kStageBytes = 2 * BN * kStride * 2;

  int stage = 0;
  for (int kv0 = kv_begin; kv0 < kv_end; kv0 += BN) {
    uint32_t k_addr = k_addr0, v_addr = v_addr0;
    if constexpr (kPipe) {
      __syncthreads();  // the other stage (tile - 1) is fully consumed
      if (kv0 + BN < kv_end) {
        load_tile_async<HD, kThreads>(Ks + (stage ^ 1) * 2 * BN * kStride, kb, p.ldk, kv0 + BN, BN, kv_end);
        load_tile_async<HD, kThreads>(Vs + (stage ^ 1) * 2 * BN * kStride, vb, p.ldv, kv0 + BN, BN, kv_end);
      }
      cp_async_commit();
      cp_async_wait<1>();   // this tile (and Q) landed
      __syncthreads();
      k_addr += stage * kStageBytes;
      v_addr += stage * kStageBytes;
      stage ^= 1;
    } else {
      __syncthreads();  // previous tile fully consumed (and Q visible on the first pass)
      load_tile<HD, kThreads>(Ks, kb, p.ldk, kv0, BN, kv_end);
      load_tile<HD, kThreads>(Vs, vb, p.ldv, kv0, BN, kv_end);
      __syncthreads();
    }

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4(q_addr + kk * 32, a0, a1, a2, a3);
#pragma unroll
      for (int nt = 0; nt < NT / 2; ++nt) {  // two 8-key tiles per ldmatrix.x4
        uint32_t b0, b1, b2, b3;
        ldsm_x4(k_addr + (nt * 16 * kStride + kk * 16) * 2, b0, b1, b2, b3);
        mma16816(s[2 * nt], a0, a1, a2, a3, b0, b1);
        mma16816(s[2 * nt + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    // ---- mask the ragged tail, online softmax ----
    const int valid = kv_end - kv0;  // >= 1
    if (valid < BN) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = nt * 8 + (lane & 3) * 2;
        if (c >= valid) s[nt][0] = s[nt][2] = -INFINITY;
        if (c + 1 >= valid) s[nt][1] = s[nt][3] = -INFINITY;
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], neg_m[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_new = fmaxf(m_run[r], mx[r]);  // finite: every tile has at least one valid key
      corr[r] = ex2f((m_run[r] - m_new) * kLog2eMma);
      m_run[r] = m_new;
      neg_m[r] = -m_new * kLog2eMma;
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = ex2f(fmaf(s[nt][0], kLog2eMma, neg_m[0]));
      s[nt][1] = ex2f(fmaf(s[nt][1], kLog2eMma, neg_m[0]));
      s[nt][2] = ex2f(fmaf(s[nt][2], kLog2eMma, neg_m[1]));
      s[nt][3] = ex2f(fmaf(s[nt][3], kLog2eMma, neg_m[1]));
      sum[0] += s[nt][0] + s[nt][1];
      sum[1] += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + sum[r];  // per-thread partial; reduced at the end
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o[i][0] *= corr[0];
      o[i][1] *= corr[0];
      o[i][2] *= corr[1];
      o[i][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < BN / 16; ++kk) {  // 16 keys per step
      const uint32_t a0 = pack_half2(s[2 * kk][0], s[2 * kk][1]);
      const uint32_t a1 = pack_half2(s[2 * kk][2], s[2 * kk][3]);
      const uint32_t a2 = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      const uint32_t a3 = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dt = 0; dt < HD / 16; ++dt) {  // two 8-wide d tiles per ldmatrix.x4.trans
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(v_addr + (kk * 16 * kStride + dt * 16) * 2, b0, b1, b2, b3);
        mma16816(o[2 * dt], a0, a1, a2, a3, b0, b1);
        mma16816(o[2 * dt + 1], a0, a1, a2, a3, b2, b3);
      }
    }
  }

  // ---- finalise ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const int row_a = q0 + warp * 16 + (lane >> 2);
  const int col = (lane & 3) * 2;
  if (p.n_splits == 1) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = row_a + r * 8;
      if (row >= p.Nq) continue;
      const float inv = 1.0f / l_run[r];
      __half* orow = p.out + (static_cast<long long>(b) * p.Nq + row) * p.ldo + h * HD + col;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i)
        *reinterpret_cast<uint32_t*>(orow + i * 8) = pack_half2(o[i][2 * r] * inv, o[i][2 * r + 1] * inv);
    }
  } else {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = row_a + r * 8;
      if (row >= p.Nq) continue;
      const long long idx = ((static_cast<long long>(b) * p.H + h) * p.n_splits + split) * p.Nq + row;
      float* po = p.part_o + idx * HD + col;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) *reinterpret_cast<float2*>(po + i * 8) = make_float2(o[i][2 * r], o[i][2 * r + 1]);
      if ((lane & 3) == 0) *reinterpret_cast<float2*>(p.part_ml + idx * 2) = make_float2(m_run[r], l_run[r]);
    }
  }
}

// The "trg" direction of the bidirectional cross-attention (transformer.py:314-378): thousands of main-stream queries
// against the <= 64 context keys of one (sample, head).  The generic kernel above gives every 64-row query tile its own CTA,
// which re-stages K / V per tile and runs load -> sync -> compute -> store with nothing overlapped (measured 222-232 us =
// 1.3 TB/s for 3140 x 25 / 6336 x 50).  Here a CTA keeps K and V resident, walks the query tiles q0 = blockIdx.x, +gridDim.x,
// ... with the NEXT tile's Q in flight (cp.async, two buffers) while the current one is computed, and writes O through the
// consumed Q buffer as full 16-byte row segments.  One kv tile means a plain (single-pass) softmax.
template <int HD, int kWarps, int BN>
__global__ void __launch_bounds__(kWarps * 32)
attn_mma_trg_kernel(AttnMmaParams p) {
  constexpr int kThreads = kWarps * 32;
  constexpr int BM = kWarps * 16;
  constexpr int NT = BN / 8;
  constexpr int kStride = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_raw_mma[];
  __half* Qs = reinterpret_cast<__half*>(smem_raw_mma);   // 2 x [BM][kStride]
  __half* Ks = Qs + 2 * BM * kStride;
  __half* Vs = Ks + BN * kStride;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const __half* qb = p.q + static_cast<long long>(b) * p.Nq * p.ldq + h * p.q_hs;
  const __half* kb = p.k + static_cast<long long>(b) * p.Nk * p.ldk + h * p.k_hs;
  const __half* vb = p.v + static_cast<long long>(b) * p.Nk * p.ldv + h * p.v_hs;
  const int n_tiles = (p.Nq + BM - 1) / BM;

  load_tile_async<HD, kThreads>(Ks, kb, p.ldk, 0, BN, p.Nk);
  load_tile_async<HD, kThreads>(Vs, vb, p.ldv, 0, BN, p.Nk);
  int tile = blockIdx.x;
  if (tile < n_tiles) load_tile_async<HD, kThreads>(Qs, qb, p.ldq, tile * BM, BM, p.Nq);
  cp_async_commit();

  const uint32_t k_addr = smem_u32(Ks + ((lane & 7) + ((lane >> 4) << 3)) * kStride + ((lane >> 3) & 1) * 8);
  const uint32_t v_addr = smem_u32(Vs + ((lane & 7) + (((lane >> 3) & 1) << 3)) * kStride + (lane >> 4) * 8);
  const int valid = p.Nk;  // <= BN

  for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
    __half* Qc = Qs + (it & 1) * BM * kStride;
    __syncthreads();   // everyone is done with the other buffer (tile it - 1: its O staging has been written out)
    const int next = tile + gridDim.x;
    if (next < n_tiles) load_tile_async<HD, kThreads>(Qs + ((it + 1) & 1) * BM * kStride, qb, p.ldq, next * BM, BM, p.Nq);
    cp_async_commit();
    cp_async_wait<1>();   // this tile's Q (and, the first time, K / V) have landed for this thread
    __syncthreads();      // ... and for every other thread

    const uint32_t q_addr = smem_u32(Qc + (warp * 16 + (lane & 15)) * kStride + (lane >> 4) * 8);
    float s[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4(q_addr + kk * 32, a0, a1, a2, a3);
#pragma unroll
      for (int nt = 0; nt < NT / 2; ++nt) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(k_addr + (nt * 16 * kStride + kk * 16) * 2, b0, b1, b2, b3);
        mma16816(s[2 * nt], a0, a1, a2, a3, b0, b1);
        mma16816(s[2 * nt + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    if (valid < BN) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = nt * 8 + (lane & 3) * 2;
        if (c >= valid) s[nt][0] = s[nt][2] = -INFINITY;
        if (c + 1 >= valid) s[nt][1] = s[nt][3] = -INFINITY;
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mx[r] *= -kLog2eMma;
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = ex2f(fmaf(s[nt][0], kLog2eMma, mx[0]));
      s[nt][1] = ex2f(fmaf(s[nt][1], kLog2eMma, mx[0]));
      s[nt][2] = ex2f(fmaf(s[nt][2], kLog2eMma, mx[1]));
      s[nt][3] = ex2f(fmaf(s[nt][3], kLog2eMma, mx[1]));
      sum[0] += s[nt][0] + s[nt][1];
      sum[1] += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
      sum[r] = 1.0f / sum[r];
    }
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < BN / 16; ++kk) {
      const uint32_t a0 = pack_half2(s[2 * kk][0], s[2 * kk][1]);
      const uint32_t a1 = pack_half2(s[2 * kk][2], s[2 * kk][3]);
      const uint32_t a2 = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      const uint32_t a3 = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dt = 0; dt < HD / 16; ++dt) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(v_addr + (kk * 16 * kStride + dt * 16) * 2, b0, b1, b2, b3);
        mma16816(o[2 * dt], a0, a1, a2, a3, b0, b1);
        mma16816(o[2 * dt + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    // ---- O through this warp's own 16 rows of the consumed Q buffer, then 16-byte row segments to global ----
    __syncwarp();   // every lane's ldmatrix reads of these rows are done
    {
      __half* orow = Qc + (warp * 16 + (lane >> 2)) * kStride + (lane & 3) * 2;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        *reinterpret_cast<uint32_t*>(orow + i * 8) = pack_half2(o[i][0] * sum[0], o[i][1] * sum[0]);
        *reinterpret_cast<uint32_t*>(orow + 8 * kStride + i * 8) = pack_half2(o[i][2] * sum[1], o[i][3] * sum[1]);
      }
    }
    __syncwarp();
    constexpr int kVec = HD / 8;
    for (int i = lane; i < 16 * kVec; i += 32) {
      const int r = i / kVec, c = i - r * kVec;
      const int row = tile * BM + warp * 16 + r;
      if (row < p.Nq)
        *reinterpret_cast<uint4*>(p.out + (static_cast<long long>(b) * p.Nq + row) * p.ldo + h * HD + c * 8) =
            *reinterpret_cast<const uint4*>(Qc + (warp * 16 + r) * kStride + c * 8);
    }
  }
  cp_async_wait<0>();
}

// merge the split-KV partials: out = sum_s exp(m_s - m) O_s / sum_s exp(m_s - m) l_s.  One CTA per (row, h, b).
__global__ void attn_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml, int Nq,
                                    int H, int n_splits, int HD, __half* __restrict__ out, long long ldo) {
  const int row = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const long long base = (static_cast<long long>(b) * H + h) * n_splits;
  float m = -INFINITY;
  for (int s = 0; s < n_splits; ++s) m = fmaxf(m, part_ml[((base + s) * Nq + row) * 2]);
  float l = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const float* ml = part_ml + ((base + s) * Nq + row) * 2;
    l += ml[1] * ex2f((ml[0] - m) * kLog2eMma);
  }
  const float inv = 1.0f / l;
  for (int c = threadIdx.x; c < HD; c += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < n_splits; ++s) {
      const long long idx = (base + s) * Nq + row;
      acc += part_o[idx * HD + c] * ex2f((part_ml[idx * 2] - m) * kLog2eMma);
    }
    out[(static_cast<long long>(b) * Nq + row) * ldo + h * HD + c] = __float2half_rn(acc * inv);
  }
}

static int g_attn_mma_wide = 1;
static int g_attn_mma_pipe = 1;   // CWM_XATTN_PIPE=0: the un-pipelined split-key loop
static int g_attn_mma_split = 512;  // keys per CTA when few queries face a long key axis (multiple of 64)

static void pick_splits(int Nq, int Nk, int* n_splits, int* kv_per_split) {
  // few query rows against a long key axis (cross-attention "src" direction): split the keys for parallelism
  if (Nq <= 64 && Nk >= 1024) {
    *kv_per_split = g_attn_mma_split;
    *n_splits = (Nk + g_attn_mma_split - 1) / g_attn_mma_split;
  } else {
    *kv_per_split = ((Nk + 63) / 64) * 64;
    *n_splits = 1;
  }
}

template <int HD, int kWarps, int BN, bool kPipe = false>
static int launch_attn_mma_bn(const AttnMmaParams& p, int B, cudaStream_t s) {
  constexpr int smem = (kWarps * 16 + (kPipe ? 4 : 2) * BN) * (HD + 8) * 2;
  static bool attr_set = false;
  if (!attr_set) {
    CWM_CUDA_CHECK(cudaFuncSetAttribute(attn_mma_kernel<HD, kWarps, BN, kPipe>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((p.Nq + kWarps * 16 - 1) / (kWarps * 16), p.H, B * p.n_splits);
  attn_mma_kernel<HD, kWarps, BN, kPipe><<<grid, kWarps * 32, smem, s>>>(p);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

template <int HD, int kWarps>
static int launch_attn_mma(const AttnMmaParams& p, int B, cudaStream_t s) {
  // at most 32 keys in total (the 25 IMU tokens of the encoder-side cross attention): 32-row key tile
  if (p.Nk <= 32 && p.n_splits == 1) return launch_attn_mma_bn<HD, kWarps, 32>(p, B, s);
  // split keys (few queries against a long key axis): 32-key tiles, double-buffered
  if (p.n_splits > 1 && g_attn_mma_pipe) return launch_attn_mma_bn<HD, kWarps, 32, true>(p, B, s);
  return launch_attn_mma_bn<HD, kWarps, 64>(p, B, s);
}

template <int HD, int kWarps, int BN>
static int launch_attn_trg(const AttnMmaParams& p, int B, cudaStream_t s) {
  constexpr int smem = (2 * kWarps * 16 + 2 * BN) * (HD + 8) * 2;
  static bool attr_set = false;
  if (!attr_set) {
    CWM_CUDA_CHECK(cudaFuncSetAttribute(attn_mma_trg_kernel<HD, kWarps, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  // query tiles per (sample, head) are spread over gx CTAs so that the grid is ~4 CTAs per SM: each CTA then walks several
  // tiles with K / V staged once and the next Q tile in flight
  const int n_tiles = (p.Nq + kWarps * 16 - 1) / (kWarps * 16);
  int gx = (4 * num_sms() + p.H * B - 1) / (p.H * B);
  gx = gx < 1 ? 1 : (gx > n_tiles ? n_tiles : gx);
  dim3 grid(gx, p.H, B);
  attn_mma_trg_kernel<HD, kWarps, BN><<<grid, kWarps * 32, smem, s>>>(p);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}

}  // namespace cwm

using namespace cwm;

extern "C" size_t cwm_attention_generic_workspace_bytes(int B, int Nq, int Nk, int H, int head_dim) {
  int n_splits, kvp;
  pick_splits(Nq, Nk, &n_splits, &kvp);
  if (n_splits == 1) return 0;
  return static_cast<size_t>(B) * H * n_splits * Nq * (head_dim + 2) * sizeof(float) + 256;
}

extern "C" int cwm_attention_generic_f16(const uint16_t* q, const uint16_t* k, const uint16_t* v, int ldq, int ldk,
                                         int ldv, int q_head_stride, int k_head_stride, int v_head_stride, int B,
                                         int Nq, int Nk, int H, int head_dim, uint16_t* out, int ldo, void* workspace,
                                         size_t workspace_bytes, cwm_stream_t stream) {
  CWM_REQUIRE(q && k && v && out, "cwm_attention_generic_f16: null pointer");
  CWM_REQUIRE(B >= 0 && Nq > 0 && Nk > 0 && H > 0 && H <= 65535, "cwm_attention_generic_f16: bad shape B=%d Nq=%d Nk=%d H=%d", B, Nq, Nk, H);
  auto al = [](const void* ptr) { return reinterpret_cast<uintptr_t>(ptr) % 16 == 0; };
  CWM_REQUIRE(al(q) && al(k) && al(v) && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && q_head_stride % 8 == 0 &&
                  k_head_stride % 8 == 0 && v_head_stride % 8 == 0 && ldo % 2 == 0 &&
                  reinterpret_cast<uintptr_t>(out) % 4 == 0,
              "cwm_attention_generic_f16: operands must be 16-byte aligned (strides multiples of 8 elements)");
  if (B == 0) return CWM_OK;
  AttnMmaParams p;
  p.q = reinterpret_cast<const __half*>(q);
  p.k = reinterpret_cast<const __half*>(k);
  p.v = reinterpret_cast<const __half*>(v);
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv;
  p.q_hs = q_head_stride; p.k_hs = k_head_stride; p.v_hs = v_head_stride;
  p.Nq = Nq; p.Nk = Nk; p.H = H;
  p.out = reinterpret_cast<__half*>(out);
  p.ldo = ldo;
  pick_splits(Nq, Nk, &p.n_splits, &p.kv_per_split);
  p.part_o = nullptr;
  p.part_ml = nullptr;
  CWM_REQUIRE(static_cast<long long>(B) * p.n_splits <= 65535, "cwm_attention_generic_f16: grid too large");
  if (p.n_splits > 1) {
    const size_t need = cwm_attention_generic_workspace_bytes(B, Nq, Nk, H, head_dim);
    if (workspace == nullptr || workspace_bytes < need)
      return fail(CWM_ERR_WORKSPACE, "cwm_attention_generic_f16: workspace %zu bytes < required %zu", workspace_bytes, need);
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    p.part_o = reinterpret_cast<float*>(ws);
    p.part_ml = p.part_o + static_cast<size_t>(B) * H * p.n_splits * Nq * head_dim;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProfileScope prof(s, "attention_small_mma", 4.0 * B * H * static_cast<double>(Nq) * Nk * head_dim,
                    (static_cast<double>(B) * (Nq * 2.0 + Nk * 2.0) * H * head_dim) * 2.0);
  const bool narrow = Nq <= 32;
  // many query rows against a handful of keys (cross-attention "trg" direction): 8 warps share one K/V tile
  // (head dims above 128 need > 160 registers per thread: one 8-warp CTA per SM would be slower -- measured)
  const bool wide = Nq >= 1024 && Nk <= 64 && head_dim <= 128 && g_attn_mma_wide;
  int rc;
  static int trg_env = -1;
  if (trg_env < 0) {
    const char* e = getenv("CWM_XATTN_TRG");
    trg_env = (e == nullptr) ? 1 : atoi(e);
    e = getenv("CWM_XATTN_PIPE");
    if (e != nullptr) g_attn_mma_pipe = atoi(e);
  }
  // many queries against one resident K / V tile (cross-attention "trg" direction): the tile-walking kernel
  const bool trg = trg_env != 0 && Nq >= 512 && Nk <= 64 && p.n_splits == 1 && ldo % 8 == 0 &&
                   reinterpret_cast<uintptr_t>(out) % 16 == 0;
  static int trg_warps = -1;
  if (trg_warps < 0) {
    const char* e = getenv("CWM_XATTN_TRG_WARPS");
    trg_warps = (e == nullptr) ? 4 : atoi(e);
  }
  const bool trg8 = trg_warps == 8 && head_dim <= 128;
#define CWM_TRG_CASE(HDV)                                                                                   \
  case HDV:                                                                                                 \
    if (trg8 && HDV <= 128)                                                                                 \
      rc = Nk <= 32 ? launch_attn_trg<HDV, (HDV <= 128 ? 8 : 4), 32>(p, B, s)                               \
                    : launch_attn_trg<HDV, (HDV <= 128 ? 8 : 4), 64>(p, B, s);                              \
    else                                                                                                    \
      rc = Nk <= 32 ? launch_attn_trg<HDV, 4, 32>(p, B, s) : launch_attn_trg<HDV, 4, 64>(p, B, s);          \
    break;
  if (trg && B <= 65535) {
    switch (head_dim) {
      CWM_TRG_CASE(32)
      CWM_TRG_CASE(64)
      CWM_TRG_CASE(96)
      CWM_TRG_CASE(128)
      CWM_TRG_CASE(192)
      default:
        return fail(CWM_ERR_UNSUPPORTED, "cwm_attention_generic_f16: head_dim %d (supported: 32, 64, 96, 128, 192)", head_dim);
    }
    return rc;
  }
#undef CWM_TRG_CASE
#define CWM_MMA_CASE(HDV)                                                            \
  case HDV:                                                                          \
    rc = narrow ? launch_attn_mma<HDV, 2>(p, B, s)                                   \
                : (wide ? launch_attn_mma<HDV, 8>(p, B, s) : launch_attn_mma<HDV, 4>(p, B, s)); \
    break;
  switch (head_dim) {
    CWM_MMA_CASE(32)
    CWM_MMA_CASE(64)
    CWM_MMA_CASE(96)
    CWM_MMA_CASE(128)
    CWM_MMA_CASE(192)
    default:
      return fail(CWM_ERR_UNSUPPORTED, "cwm_attention_generic_f16: head_dim %d (supported: 32, 64, 96, 128, 192)", head_dim);
  }
#undef CWM_MMA_CASE
  if (rc != CWM_OK) return rc;
  if (p.n_splits > 1) {
    dim3 grid(Nq, H, B);
    attn_combine_kernel<<<grid, 64, 0, s>>>(p.part_o, p.part_ml, Nq, H, p.n_splits, head_dim, p.out, p.ldo);
    CWM_LAUNCH_CHECK();
  }
  return CWM_OK;
}

extern "C" void cwm_debug_attn_mma_wide(int on) { g_attn_mma_wide = on; }
extern "C" void cwm_debug_attn_mma_split(int keys) { g_attn_mma_split = keys >= 64 ? (keys / 64) * 64 : 64; }
