// gemm.cu -- Y[M,N] = epilogue(A[M,K] . W[N,K]^T) on tcgen05 tensor cores (sm_100a).
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer: 128x64 A tile + BNx64 W tile per stage, 128B-swizzled, mbarrier pipelined
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//               fp32 accumulators in tensor memory, two accumulator stages so the epilogue of tile i overlaps
//               the main loop of tile i+1
//   warp 2      TMEM allocator
//   warps 4-11  epilogue, two warps per TMEM lane quadrant (they split the column chunks).  One thread owns one
//               output row of a 32-row x 128-byte chunk:
//                 tcgen05.ld -> registers -> (+bias, q-scale | GELU | +residual) -> 128B-swizzled smem -> TMA store.
//               The fp32 residual chunk is TMA-loaded into the same smem buffer one chunk ahead (double buffered),
//               so neither the residual read nor the output write occupies the LSU or stalls on global latency.
//   A generic (slower) epilogue handles the two row-remapped / row-gathered uses (patch-embed + pos[perm], and
//   encoder_to_decoder written into the decoder sequence): coalesced direct global accesses through a padded smem
//   transpose.
//
// Roofline: tensor bound for K >= 768; the fp32-residual GEMMs with K <= 512 are HBM bound (8 bytes of residual
// traffic per output element) -- see DESIGN.md.
#include <cstdlib>

#include "common.cuh"

namespace cwm {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 f16 = 128 bytes = one swizzle row
constexpr int kGemmThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kChunkBytes = 32 * 128;  // one epilogue chunk: 32 rows x 128 bytes

template <int BN, bool kRes, bool kConv = false, bool kCta2 = false>
struct GemmCfg {
  // epilogue smem: kRes (fp32 out, optional residual): 2 buffers / warp, else 1 buffer / warp
  // (+ 2 KB per warp for the f16 copy the LayerNorm-producer epilogue emits: 32 rows x 64 bytes, 64B-swizzled)
  static constexpr int kX16Bytes = 32 * 64;
  static constexpr int kEpiBytes = kEpiWarps * (kChunkBytes * (kRes ? 2 : 1) + (kRes ? kX16Bytes : 0));
  // per warp: f16-out epilogues stage 64 bias values + 64 LayerNorm column sums, fp32 epilogues 32 bias values
  static constexpr int kBiasWarpBytes = kRes ? 128 : 512;
  static constexpr int kBiasBytes = kEpiWarps * kBiasWarpBytes;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  // a CTA of a pair stages only its half of the W tile: the stage is 16 KB smaller and one more stage fits (fc2 with its
  // 824 MB DRAM-resident A operand ran at 67 % tensor activity on 3 stages = two k-steps of prefetch)
  static constexpr int kBStageBytes = kCta2 ? kBBytes / 2 : kBBytes;
  static constexpr int kStageBytes = kABytes + kBStageBytes;
  static constexpr int kBudget = 232448 - kEpiBytes - kBiasBytes - 512;
  static constexpr int kMaxStages = kBudget / kStageBytes;
  // plain GEMMs: 6 stages are enough to cover the TMA latency; the convolution instantiation takes the whole budget (its
  // halo mode carves the ring into up to 4 halo stages + the weight ring / the resident weight matrix)
  static constexpr int kStages = kConv ? (kMaxStages > 8 ? 8 : kMaxStages) : (kMaxStages > 6 ? 6 : kMaxStages);
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  // the operand ring: whole stages for the GEMM / per-tap modes; the convolution halo mode carves the entire budget (halo
  // stages + 8 / 16 KB weight stages), which for BN = 256 is 190 KB instead of 3 x 48 KB
  static constexpr int kRingBytes = kConv ? (kBudget / 1024) * 1024 : kStages * kStageBytes;
  static_assert(kRingBytes >= kStages * kStageBytes, "ring");
  static constexpr int kSmemBytes = kRingBytes + kEpiBytes + kBiasBytes + 512;
  static_assert(kStages >= 3, "pipeline too shallow");
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
  // ---- LayerNorm fusion ----
  // producer side (fp32 residual epilogue): f16 copy of the updated residual rows (the A operand of the next GEMM) and
  // per-row partial statistics (sum, sum of squares) of this CTA's columns: ln_stats_out[(n_blk * 2 + half) * M + row]
  __half* x16;
  int ldx16;
  float2* ln_stats_out;
  // consumer side (f16 epilogues): out = rstd_m * (acc - mean_m * ln_s[n]) + bias[n], with mean / rstd of row m
  // rebuilt from ln_parts partial statistics; the LayerNorm affine is folded into W (gamma), ln_s and bias (beta)
  const float2* ln_stats_in;
  int ln_parts;
  const float* ln_s;
  float ln_inv_c;
  float ln_eps;
  int relu;  // f16-out epilogues: max(v, 0) after the bias (convolution + relu)
  // convolution post-ops of RAFT's ConvGRU (cwm/models/raft/update.py:43-60), evaluated on the fp32 accumulators:
  //   1 = gate:   columns [0, C): z = sigmoid(v) -> out;  columns [C, 2C): sigmoid(v) * h -> out2 (column n - C)
  //   2 = update: h + z * (tanh(v) - h) -> out (h's own slot, in place) and, when out2 is given, a dense copy
  //   3 = tail:   the last two output columns are replaced by the 2 f16 at aux_h[row * ld_h] (the flow columns that close
  //               the GRU input rows; out2 = the second row buffer)
  // h / z are f16 pixel rows (aux_h, aux_z); out2 is the kernel's second output tensor map.
  int post;
  int post_c;
  const __half* aux_h;
  int ld_h;
  const __half* aux_z;
  int ld_z;
  int has_out2;
  int img_h, img_w, img_s;
};
// one MUFU each: tanh.approx.f32 (max relative error 2^-11, the rounding of the f16 the result is stored as) and
// sigmoid(x) = 0.5 tanh(x / 2) + 0.5
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ void h8_unpack(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// Implicit-GEMM convolution mode (stride 1, "same" zero padding, NHWC f16 rows; cwm_conv2d_f16).  The A operand of
// k-step kb = (tap, 64-channel slab) is ONE 4-D TMA box [1, hb, wb, 64] of the input image at the tap's spatial offset:
// nothing is ever im2col-ed, and the zero padding is the TMA unit's out-of-bounds fill.  A 128-row tile is hb image rows of
// wb pixel slots (wb = 16 or 32 >= image width; slots beyond the width are computed and clipped by the 4-D TMA store).
struct ConvDev {
  int taps;           // kh * kw; 0 = plain GEMM
  int kw;
  int pad_h, pad_w;
  int cin_slabs;      // 64-channel slabs per tap (K = taps * cin_slabs * 64, weights zero-padded to that)
  int tiles_per_img;  // tiles_x * ceil(Hout / hb)
  int tiles_x;        // column blocks of wb pixel slots per image row (1 for maps of <= 32 pixels; the encoders' 56 / 112
                      // pixel maps are tiled in both directions)
  int stride;         // 1 or 2: output pixel (y, x) reads input (stride * y + ky - pad_h, stride * x + kx - pad_w); the A
                      // tensor map then carries the traversal stride (one box still loads hb x wb pixels)
  int wo;             // output pixels per tile row that are kept (= the tile pitch along x): wb, except in the 2-D halo mode
                      // where a tile row of wb = 32 slots yields wb - 2 pad_w outputs (the halo of the row sits in the slots)
  int halo_x;         // halo mode: the staged box starts halo_x pixels left of the tile's first output pixel ...
  int tap_x;          // ... and tap (ky, kx) reads from ky * wb + kx + tap_x rows further down.  Full-row tiles (maps <= 32 px):
                      // halo_x = 2 pad_w, tap_x = pad_w (reads past a row's end land on the next row's leading zeros);
                      // 2-D tiles: halo_x = pad_w, tap_x = 0 (the kept outputs never read past their own row)
  int w_resident;     // halo mode: the whole weight matrix fits the weight ring -> staged once per CTA, never released
  int h_stages;       // halo mode: halo stages in flight (2..4: one staged halo is ~2 us of TMA latency away) ...
  int h_stride;       // ... and the bytes between them (multiple of 1024)
  int hb;             // image rows per 128-row tile
  int rows_per_warp;  // image rows per 32-row epilogue chunk (32 / wb)
  // halo mode (W + 2 pad_w <= wb): per channel slab ONE box of hb + kh - 1 image rows is staged, each row laid out as
  // [2 pad_w zero slots | W pixels | zero slots]; tap (ky, kx) then reads the SAME smem through an A descriptor whose start
  // is shifted by ky * wb + kx + pad_w rows (a read past a row's end lands on the next row's leading zeros), so the
  // shared-memory fill per tile drops from taps x 16 KB to one halo per slab.
  int halo;           // 0 = one A box per (tap, slab) k-step
  int halo_rows;      // hb + kh - 1 (+ 1, see cwm_conv2d_f16)
  int wb;
  int w_stages;       // depth of the weight-tile ring in halo mode
  int w_stride;       // bytes between its stages (a CTA of a pair holds half a tile)
  int base_off;       // timing-experiment switches (CWM_CONV_BASEOFF): bit 1 (2) = epilogue hands the accumulator straight back,
                      // bit 2 (4) = halo-mode MMA issuer skips the MMAs (pair kernels).  (Round 1 used the value 1 to set the
                      // descriptor's base-offset field for the row-shifted A starts: wrong -- the swizzle XOR follows the
                      // absolute shared-memory address -- and removed.)
};
constexpr int kHaloStageBytes = 33 * 1024;  // 256 rows x 128 B + slack for the shifted reads of discarded output rows

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)), branch-free with one MUFU:
//   erfc(z) = 2^(z * Q(z)) for z = min(|x| / sqrt 2, 4), Q = degree-4 minimax fit (max |erf error| 6.8e-7, max
//   |GELU error| 1.1e-6 -- 400x below the f16 rounding of the output; tools/fit_erf.py reproduces the fit).
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  const float z = fminf(ax * 0.70710678118654752440f, 4.0f);
  float q = fmaf(-0.0029442342929542065f, z, 0.029590291902422905f);
  q = fmaf(q, z, -0.1486659049987793f);
  q = fmaf(q, z, -0.9185092449188232f);
  q = fmaf(q, z, -1.6278890371322632f);
  float p;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(z * q));
  return fmaf(0.5f * ax, 1.0f - p, 0.5f * x);
}

// ---- packed dual-fp32 math (Blackwell fma/mul/add .f32x2: one issue slot for two lanes).  The f16-out epilogues
//      are issue-bound for K <= 1024 (8 epilogue warps have ~6k cycles per 128x256 tile), so bias / LayerNorm / GELU
//      run on register pairs; every lane computes exactly what the scalar formulas above compute. ----
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// gelu_erf on a pair (same operations, lane for lane)
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t x) {
  float x0, x1;
  upk2(x, x0, x1);
  const uint64_t ax = pk2(fabsf(x0), fabsf(x1));
  float z0, z1;
  upk2(mul2(ax, pk2(0.70710678118654752440f, 0.70710678118654752440f)), z0, z1);
  const uint64_t z = pk2(fminf(z0, 4.0f), fminf(z1, 4.0f));
  uint64_t q = fma2(pk2(-0.0029442342929542065f, -0.0029442342929542065f), z, pk2(0.029590291902422905f, 0.029590291902422905f));
  q = fma2(q, z, pk2(-0.1486659049987793f, -0.1486659049987793f));
  q = fma2(q, z, pk2(-0.9185092449188232f, -0.9185092449188232f));
  q = fma2(q, z, pk2(-1.6278890371322632f, -1.6278890371322632f));
  float t0, t1, p0, p1;
  upk2(mul2(z, q), t0, t1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(t0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(t1));
  const uint64_t omp = fma2(pk2(p0, p1), pk2(-1.0f, -1.0f), pk2(1.0f, 1.0f));  // 1 - p
  const uint64_t half2 = pk2(0.5f, 0.5f);
  return fma2(mul2(ax, half2), omp, mul2(x, half2));
}

// ---- explicit shared-space accesses (the smem pointers are carved from a uintptr_t, keep them out of the
//      generic address space) ----
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_s(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster on the two SMs of one TPC share a 256 x BN tile.
//      Each CTA stages its own 128 rows of A and HALF of the W tile (BN/2 rows); the leader's single
//      tcgen05.mma.cta_group::2 reads both shared memories, so the smem fill per CTA drops from (128 + BN) to
//      (128 + BN/2) rows per k-step for the same tensor work.  Barriers: the TMA loads of both CTAs complete on
//      the LEADER's full barrier (peer bit of the shared::cluster address cleared), the leader's commits are
//      multicast to the empty / accumulator-full barriers of both CTAs, and the epilogue warps of both CTAs
//      arrive on the leader's accumulator-empty barrier.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_ss2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (count 1) on the mbarrier at this offset in BOTH CTAs of the pair when the prior MMAs complete
__device__ __forceinline__ void umma_commit2(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

// kConv: the convolution instantiation (4-D TMA operands, halo mode, relu / ConvGRU post-ops).  It is a template flag and not
// a run-time mode because the f16-out epilogue of the plain GEMMs is issue-bound: the extra branches per 16-byte unit cost
// the qkv / fc1 GEMMs 15 % when they were run-time (measured: 1173 -> 996 TFLOP/s inside a large-4x4 step).
// kConv: 0 = plain GEMM; 1 + post = convolution with the compile-time post-op (0 none, 1 GRU gate, 2 GRU update, 3 tail)
template <int BN, bool kRes, bool kCta2, int kConv>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w,
                const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_res,
                const __grid_constant__ CUtensorMap tma_x16, int M, int N, int K, EpiDev ep, ConvDev cv) {
  constexpr bool kIsConv = kConv > 0;
  // kConv < 0: plain GEMM with a compile-time f16 epilogue: -1 = fused LayerNorm + chunk-uniform q-scale (qkv), -2 = fused
  // LayerNorm + GELU (fc1): the run-time variant checks per 16-byte unit disappear from the issue-bound epilogue
  constexpr int kEpi = kConv < 0 ? -kConv : 0;
  using Cfg = GemmCfg<BN, kRes, kIsConv, kCta2>;
  constexpr int kPost = kConv > 0 ? kConv - 1 : 0;   // the convolution's post-op, compile-time: one variant per kernel
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint8_t* smem_epi = smem + Cfg::kRingBytes;
  uint8_t* smem_bias = smem_epi + Cfg::kEpiBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bias + Cfg::kBiasBytes);
  constexpr int kBarSlots = 12;                     // >= kStages; the convolution halo mode uses up to 12 weight stages
  static_assert(Cfg::kStages <= kBarSlots, "barrier slots");
  uint64_t* full_bar = bars;                        // kBarSlots
  uint64_t* empty_bar = bars + kBarSlots;           // kBarSlots
  uint64_t* tfull_bar = bars + 2 * kBarSlots;       // 2
  uint64_t* tempty_bar = tfull_bar + 2;             // 2
  uint64_t* res_bar = tempty_bar + 2;               // kEpiWarps * 2 (residual chunk landed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * kEpiWarps);
  uint64_t* halo_full = res_bar + 2 * kEpiWarps + 1;   // 4 (convolution halo mode)
  uint64_t* halo_empty = halo_full + 4;                // 4
  uint8_t* smem_halo = smem;                           // halo mode: h_stages halo stages, then the weight-tile ring
  uint8_t* smem_wring = smem + cv.h_stages * cv.h_stride;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  // a (pair) tile covers TM rows; in CTA-pair mode this CTA owns rows [cta_rank * 128, +128) of it
  constexpr int TM = kCta2 ? 2 * BM : BM;
  const int cta_rank = kCta2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int tile_first = kCta2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_stride = kCta2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int tiles_m = (M + TM - 1) / TM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();  // swizzle-128B needs 1024-byte alignment
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_w);
    tma_prefetch_desc(&tma_out);
    if (kRes) tma_prefetch_desc(&tma_res);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kBarSlots; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], kCta2 ? 2 * kEpiWarps : kEpiWarps);
    }
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&halo_full[i], 1);
      mbar_init(&halo_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (kCta2) {
      tmem_alloc2(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (kCta2) cluster_sync_all(); else __syncthreads();  // barrier inits visible to the peer CTA as well
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, tensor-memory allocation, descriptor prefetch) touches
  // no global memory and may run while the previous kernel of the stream drains; nothing below may start before that
  // kernel has completed and flushed.  The dependents of THIS kernel are released at once: their CTAs cannot become
  // resident before ours exit (shared / tensor memory), so all they gain is their launch latency and prologue.
  grid_dep_sync();

  // NOTE: the producer and MMA roles run warp-uniformly (all 32 lanes execute the loops and the mbarrier waits) and
  // only the TMA / tcgen05 instructions themselves are issued by one elected lane.  Running the whole role under
  // `if (lane == 0)` makes every operand thread-private and the compiler wraps each TMA / MMA instruction in a
  // per-lane "waterfall" loop (R2UR + BRA.U.ANY), ~100 cycles of issue per instruction.
  if (kIsConv && warp == 0 && cv.halo) {
    // ===================== TMA producer, convolution halo mode =====================
    int stage = 0, hs = 0;
    uint32_t phase = 0, hphase = 0;
    const uint32_t halo_bytes = static_cast<uint32_t>(cv.halo_rows * cv.wb) * 128u;
    for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
      const int m_blk = tile / tiles_n;
      const int n_blk = tile - m_blk * tiles_n;
      const int ct = m_blk * (kCta2 ? 2 : 1) + cta_rank;
      const int cs = ct / cv.tiles_per_img;
      const int cti = ct - cs * cv.tiles_per_img, cty = cti / cv.tiles_x;
      const int cy0 = cty * cv.hb - cv.pad_h, cx0 = (cti - cty * cv.tiles_x) * cv.wo - cv.halo_x;
      const bool load_w = !(cv.w_resident && tile != tile_first);   // resident weights: staged with the first tile only
      for (int slab = 0; slab < cv.cin_slabs; ++slab) {
        mbar_wait(&halo_empty[hs], hphase ^ 1);
        if (elect_one()) {
          if constexpr (kCta2) {
            if (cta_rank == 0) mbar_arrive_expect_tx(&halo_full[hs], 2 * halo_bytes);
            tma_load_4d_2sm(smem_halo + hs * cv.h_stride, &tma_a, &halo_full[hs], slab * BK, cx0, cy0, cs);
          } else {
            mbar_arrive_expect_tx(&halo_full[hs], halo_bytes);
            tma_load_4d(smem_halo + hs * cv.h_stride, &tma_a, &halo_full[hs], slab * BK, cx0, cy0, cs);
          }
        }
        __syncwarp();
        if (++hs == cv.h_stages) {
          hs = 0;
          hphase ^= 1;
        }
        if (!load_w) continue;
        for (int tap = 0; tap < cv.taps; ++tap) {
          if (!cv.w_resident) mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            const int kcol = (tap * cv.cin_slabs + slab) * BK;
            if constexpr (kCta2) {
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], Cfg::kBBytes);
              tma_load_2d_2sm(smem_wring + stage * cv.w_stride, &tma_w, &full_bar[stage], kcol,
                              n_blk * BN + cta_rank * (BN / 2));
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::kBBytes);
              tma_load_2d(smem_wring + stage * cv.w_stride, &tma_w, &full_bar[stage], kcol, n_blk * BN);
            }
          }
          __syncwarp();
          if (++stage == cv.w_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
      const int m_blk = tile / tiles_n;
      const int n_blk = tile - m_blk * tiles_n;
      // convolution mode: this CTA's 128-row tile = image rows [cy0, cy0 + hb) of sample cs
      const int ct = m_blk * (kCta2 ? 2 : 1) + cta_rank;
      const int cs = (kIsConv && cv.taps) ? ct / cv.tiles_per_img : 0;
      const int cti = (kIsConv && cv.taps) ? ct - cs * cv.tiles_per_img : 0;
      const int cty = (kIsConv && cv.taps) ? cti / cv.tiles_x : 0;
      const int cy0 = cty * cv.hb * cv.stride - cv.pad_h;                 // input coordinates of the tile's first pixel
      const int cx0 = (cti - cty * cv.tiles_x) * cv.wo * cv.stride - cv.pad_w;
      int tap = 0, slab = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          if constexpr (kCta2) {
            // the leader expects the bytes of BOTH CTAs; both CTAs' loads complete on the leader's barrier
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (Cfg::kABytes + Cfg::kBBytes / 2));
            if (kIsConv && cv.taps) {
              const int dy = tap / cv.kw;
              tma_load_4d_2sm(smem_a + stage * Cfg::kABytes, &tma_a, &full_bar[stage], slab * BK,
                              cx0 + tap - dy * cv.kw, cy0 + dy, cs);
            } else {
              tma_load_2d_2sm(smem_a + stage * Cfg::kABytes, &tma_a, &full_bar[stage], kb * BK, m_blk * TM + cta_rank * BM);
            }
            tma_load_2d_2sm(smem_b + stage * Cfg::kBStageBytes, &tma_w, &full_bar[stage], kb * BK,
                            n_blk * BN + cta_rank * (BN / 2));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kABytes + Cfg::kBBytes);
            if (kIsConv && cv.taps) {
              const int dy = tap / cv.kw;
              tma_load_4d(smem_a + stage * Cfg::kABytes, &tma_a, &full_bar[stage], slab * BK, cx0 + tap - dy * cv.kw,
                          cy0 + dy, cs);
            } else {
              tma_load_2d(smem_a + stage * Cfg::kABytes, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
            }
            tma_load_2d(smem_b + stage * Cfg::kBStageBytes, &tma_w, &full_bar[stage], kb * BK, n_blk * BN);
          }
        }
        if (kIsConv && ++slab == cv.cin_slabs) {
          slab = 0;
          ++tap;
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (kIsConv && warp == 1 && cta_rank == 0 && cv.halo) {
    // ===================== MMA issuer, convolution halo mode =====================
    // One elected thread issues everything, so the loop body per tap IS the pace of narrow convolutions: ncu of the
    // encoders' 64-channel 3x3 convolution showed this warp busy (not waiting) for ~470 cycles per tap against 128 cycles of
    // tensor work -- ~75 uniform-datapath instructions of descriptor arithmetic per tap.  The descriptors are therefore
    // advanced incrementally (only their 14-bit start-address field changes: +8 = 128 bytes per tap along a halo row, one
    // add per row wrap, per halo stage, per weight stage) and the halo / accumulator commits sit outside the tap loop.
    // The A operand of tap (ky, kx) is the halo read from (ky * wb + kx + tap_x) rows further down: a multiple of 128 B but
    // not of the 1024-byte swizzle atom; the tensor core applies the 128B-swizzle XOR to the absolute shared-memory
    // address (like the TMA write did), so the descriptor's base-offset field stays 0.
    constexpr uint32_t idesc = umma_idesc_f16(TM, BN, 0, 0);
    const uint64_t bdesc0 = umma_desc_kmajor_sw128(smem_u32(smem_wring));
    const uint64_t adesc0 = umma_desc_kmajor_sw128(smem_u32(smem_halo) + static_cast<uint32_t>(cv.tap_x) * 128u);
    const uint32_t h_step = static_cast<uint32_t>(cv.h_stride) >> 4, w_step = static_cast<uint32_t>(cv.w_stride) >> 4;
    const uint32_t row_wrap = static_cast<uint32_t>(cv.wb - cv.kw) * 8u;   // from the end of one tap row to the next row's first tap
    const int kh = cv.taps / cv.kw;
    int stage = 0, hs = 0, as = 0;
    uint32_t phase = 0, aphase = 0, hphase = 0;
    uint64_t bdesc = bdesc0, adesc_h = adesc0;   // current weight stage / current halo stage
    for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      const bool wait_w = !(cv.w_resident && tile != tile_first);   // resident weights landed with the first tile
      uint32_t accumulate = 0;
      for (int slab = 0; slab < cv.cin_slabs; ++slab) {
        mbar_wait(&halo_full[hs], hphase);
        tc_fence_after();
        uint64_t adesc = adesc_h;
        for (int ky = 0; ky < kh; ++ky) {
          for (int kx = 0; kx < cv.kw; ++kx) {
            if (wait_w) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
            }
            if (elect_one()) {
              if constexpr (kCta2) {
                if (!(cv.base_off & 4)) {   // (timing experiment CWM_CONV_BASEOFF=4: no MMAs, only the commits)
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k) umma_ss2(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (accumulate | k) != 0 ? 1u : 0u);
                }
                if (!cv.w_resident) umma_commit2(&empty_bar[stage]);
              } else {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (accumulate | k) != 0 ? 1u : 0u);
                if (!cv.w_resident) umma_commit(&empty_bar[stage]);
              }
            }
            __syncwarp();
            accumulate = 1;
            adesc += 8;
            bdesc += w_step;
            if (++stage == cv.w_stages) {
              stage = 0;
              phase ^= 1;
              bdesc = bdesc0;
            }
          }
          adesc += row_wrap;
        }
        if (elect_one()) {   // this halo is consumed; after the last slab the accumulator is complete
          if constexpr (kCta2) {
            umma_commit2(&halo_empty[hs]);
            if (slab == cv.cin_slabs - 1) umma_commit2(&tfull_bar[as]);
          } else {
            umma_commit(&halo_empty[hs]);
            if (slab == cv.cin_slabs - 1) umma_commit(&tfull_bar[as]);
          }
        }
        __syncwarp();
        adesc_h += h_step;
        if (++hs == cv.h_stages) {
          hs = 0;
          hphase ^= 1;
          adesc_h = adesc0;
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ===================== MMA issuer (the leader CTA only in CTA-pair mode) =====================
    constexpr uint32_t idesc = umma_idesc_f16(TM, BN, 0, 0);
    const uint64_t adesc0 = umma_desc_kmajor_sw128(smem_u32(smem_a));
    const uint64_t bdesc0 = umma_desc_kmajor_sw128(smem_u32(smem_b));
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        // descriptors: only the start-address field (>> 4 units) changes with the stage and the k-step
        const uint64_t adesc = adesc0 + static_cast<uint64_t>(stage * (Cfg::kABytes >> 4));
        const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(stage * (Cfg::kBStageBytes >> 4));
        if (elect_one()) {
          if constexpr (kCta2) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_ss2(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit2(&empty_bar[stage]);                      // frees the slot in both CTAs
            if (kb == num_kb - 1) umma_commit2(&tfull_bar[as]);   // accumulator complete, both CTAs
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);  // frees the smem slot when the MMAs above have read it
            if (kb == num_kb - 1) umma_commit(&tfull_bar[as]);  // accumulator complete
          }
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;   // 0..7
    const int quad = ew & 3;   // TMEM lane quadrant == warp % 4
    const int half = ew >> 2;  // which column chunks this warp takes
    const uint32_t buf0 = smem_u32(smem_epi) + ew * kChunkBytes * (kRes ? 2 : 1);
    const uint32_t x16_s = smem_u32(smem_epi) + kEpiWarps * kChunkBytes * 2 + ew * Cfg::kX16Bytes;  // kRes only
    const uint32_t bias_s = smem_u32(smem_bias) + ew * Cfg::kBiasWarpBytes;
    const uint32_t lns_s = bias_s + 256;
    const uint32_t row_s = lane * 128;      // this thread's row inside a chunk buffer
    const uint32_t sw = lane & 7;           // 128B swizzle: 16-byte unit index ^= row % 8
    int as = 0;
    uint32_t aphase = 0;

    if constexpr (!kRes) {
      // ---------------- f16 output: 64-column chunks, TMA store ----------------
      constexpr int kChunks = BN / 64;
      // (the convolution instantiation has neither GELU, LayerNorm folding nor the q-scale: compiled out there -- its
      // epilogue code was 140 KB of SASS and ncu showed 24 % of the warp samples waiting for instruction fetch)
      const bool gelu = kEpi == 2 || (kEpi == 0 && !kIsConv && (ep.mode == CWM_EPI_GELU_F16));
      const bool ln = kEpi != 0 || (!kIsConv && (ep.ln_stats_in != nullptr));
      float2 ln_t[8];  // partial statistics of this thread's row in the tile being prepared (all loads back to back:
                       // a rolled loop serialises one L2 round trip per plane, measured +30 us per launch)
      auto load_ln_stats = [&](int row) {
#pragma unroll
        for (int pp = 0; pp < 8; ++pp)
          ln_t[pp] = (pp < ep.ln_parts && row < M) ? __ldg(ep.ln_stats_in + static_cast<long long>(pp) * M + row)
                                                   : make_float2(0.f, 0.f);
      };
      if (ln && tile_first < num_tiles) load_ln_stats((tile_first / tiles_n) * TM + cta_rank * BM + quad * 32 + lane);
      int tile_iter = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_stride, ++tile_iter) {
        const int m_blk = tile / tiles_n;
        const int n_blk = tile - m_blk * tiles_n;
        const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
        const int row0 = m_blk * TM + cta_rank * BM + quad * 32;
        // fused LayerNorm: statistics of this thread's row from the producer's partial sums, summed in plane order
        // (deterministic).  The loads for the NEXT tile are issued here, one tile ahead, so that their L2 round trip
        // never sits on the epilogue's critical path.
        float ln_mean = 0.f, ln_rstd = 1.f;
        if (ln) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) {
            s1 += ln_t[pp].x;
            s2 += ln_t[pp].y;
          }
          const int row = row0 + lane;
          for (int pp = 8; pp < ep.ln_parts; ++pp) {  // (more than 4 N tiles: not reached by any CWM model)
            if (row < M) {
              const float2 u = __ldg(ep.ln_stats_in + static_cast<long long>(pp) * M + row);
              s1 += u.x;
              s2 += u.y;
            }
          }
          ln_mean = s1 * ep.ln_inv_c;
          ln_rstd = rsqrtf(fmaxf(s2 * ep.ln_inv_c - ln_mean * ln_mean, 0.f) + ep.ln_eps);
          const int next_tile = tile + tile_stride;
          if (next_tile < num_tiles) load_ln_stats((next_tile / tiles_n) * TM + cta_rank * BM + quad * 32 + lane);
        }
        // convolution post-ops: the pixel row this thread's tile row stands for (-1: a padding slot / beyond the image)
        long long post_row = -1;
        if (kIsConv && kPost) {
          const int ct = m_blk * (kCta2 ? 2 : 1) + cta_rank;
          const int cs = ct / cv.tiles_per_img;
          const int tr = quad * 32 + lane;
          const int cti = ct - cs * cv.tiles_per_img, cty = cti / cv.tiles_x;
          const int py = cty * cv.hb + tr / cv.wb, px = (cti - cty * cv.tiles_x) * cv.wo + tr % cv.wb;
          if (cs < ep.img_s && py < ep.img_h && px < ep.img_w)
            post_row = (static_cast<long long>(cs) * ep.img_h + py) * ep.img_w + px;
        }
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        if (kIsConv && (cv.base_off & 2)) {   // timing experiment (CWM_CONV_BASEOFF=2): hand the accumulator straight back
          tc_fence_before();
          if (lane == 0) { if constexpr (kCta2) mbar_arrive_leader(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]); }
          if (++as == 2) {
            as = 0;
            aphase ^= 1;
          }
          continue;
        }
        // one 64-column chunk per tile (BN == 64: the narrow convolutions, where the epilogue and not the MMA paces the
        // kernel): the two warp halves take alternate tiles instead of one half idling (ncu of the encoders' 64-channel
        // convolutions: ~4500 cycles per 128x64 tile for 600 cycles of MMA work)
        const int c_first = (kChunks == 1) ? (((tile_iter & 1) == half) ? 0 : kChunks) : half;
        for (int c = c_first; c < kChunks; c += 2) {
          const int n0 = n_blk * BN + c * 64;
          // bias of the 64 columns -> per-warp smem (broadcast reads below)
          float b_lo = 0.f, b_hi = 0.f, s_lo = 0.f, s_hi = 0.f;
          if (ep.bias != nullptr) {
            if (n0 + lane < N) b_lo = __ldg(ep.bias + n0 + lane);
            if (n0 + 32 + lane < N) b_hi = __ldg(ep.bias + n0 + 32 + lane);
          }
          if (ln) {
            if (n0 + lane < N) s_lo = __ldg(ep.ln_s + n0 + lane);
            if (n0 + 32 + lane < N) s_hi = __ldg(ep.ln_s + n0 + 32 + lane);
          }
          uint32_t acc[64];
          tmem_ld_x32(tmem_acc + c * 64, acc);
          tmem_ld_x32(tmem_acc + c * 64 + 32, acc + 32);
          __syncwarp();  // previous chunk's broadcast reads of bias_s are done
          sts32(bias_s + lane * 4, __float_as_uint(b_lo));
          sts32(bias_s + 128 + lane * 4, __float_as_uint(b_hi));
          if (ln) {
            sts32(lns_s + lane * 4, __float_as_uint(s_lo));
            sts32(lns_s + 128 + lane * 4, __float_as_uint(s_hi));
          }
          // ConvGRU post-ops: h (and z) of this thread's pixel row for the chunk's 64 columns, fetched as a 4-deep ring of
          // 16-byte loads that is primed here, before the accumulator wait, and refilled 4 units ahead in the loop
          uint4 h_raw[8], z_raw[4];   // gate: all 8 units of h up front; update: 4-deep rings of h and z
          const uint4* h_ptr = nullptr;
          const uint4* z_ptr = nullptr;
          bool post_ld_h = false;
          if (kIsConv && (kPost == 1 || kPost == 2)) {
            const int hc = (kPost == 1) ? n0 - ep.post_c : n0;   // gate: only the r columns [C, 2C) read h
            post_ld_h = post_row >= 0 && hc >= 0;
            if (post_ld_h) {
              h_ptr = reinterpret_cast<const uint4*>(ep.aux_h + post_row * ep.ld_h + hc);
#pragma unroll
              for (int q = 0; q < 4; ++q) h_raw[q] = __ldg(h_ptr + q);
              if (kPost == 1) {
#pragma unroll
                for (int q = 4; q < 8; ++q) h_raw[q] = __ldg(h_ptr + q);
              }
              if (kPost == 2) {
                z_ptr = reinterpret_cast<const uint4*>(ep.aux_z + post_row * ep.ld_z + n0);
#pragma unroll
                for (int q = 0; q < 4; ++q) z_raw[q] = __ldg(z_ptr + q);
              }
            }
          }
          tmem_ld_wait();
          const bool last_chunk = (c + 2 >= kChunks);
          if (last_chunk) {  // all TMEM reads of this tile by this warp are complete -> release the stage early
            tc_fence_before();
            if (lane == 0) { if constexpr (kCta2) mbar_arrive_leader(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]); }
          }
          __syncwarp();
          if (elect_one()) bulk_wait_read0();  // the previous TMA store has finished reading the buffer
          __syncwarp();
          // chunk-uniform q-scale (the boundary is a multiple of 64 for every model: heads * 64); a chunk that
          // straddles it falls back to per-element selection
          const bool sc_all = kEpi == 2 ? false : (!kIsConv && !gelu && (n0 + 64 <= ep.scale_cols));
          const bool sc_mixed = kEpi != 0 ? false : (!kIsConv && !gelu && !sc_all && (n0 < ep.scale_cols));
          const uint64_t sc2 = pk2(ep.scale, ep.scale);
          const uint64_t rstd2 = pk2(ln_rstd, ln_rstd);
          const uint64_t nmr2 = pk2(-ln_mean * ln_rstd, -ln_mean * ln_rstd);
#pragma unroll
          for (int u = 0; u < 8; ++u) {  // 16-byte unit u = columns 8u .. 8u+7
            uint32_t bb[8];
            lds128(bias_s + u * 32, bb[0], bb[1], bb[2], bb[3]);
            lds128(bias_s + u * 32 + 16, bb[4], bb[5], bb[6], bb[7]);
            uint64_t v2[4];
            if (ln) {  // rstd * acc + (c - rstd * mean * s)
              uint32_t ss[8];
              lds128(lns_s + u * 32, ss[0], ss[1], ss[2], ss[3]);
              lds128(lns_s + u * 32 + 16, ss[4], ss[5], ss[6], ss[7]);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                v2[q] = fma2(rstd2, pk2u(acc[u * 8 + 2 * q], acc[u * 8 + 2 * q + 1]),
                             fma2(nmr2, pk2u(ss[2 * q], ss[2 * q + 1]), pk2u(bb[2 * q], bb[2 * q + 1])));
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                v2[q] = add2(pk2u(acc[u * 8 + 2 * q], acc[u * 8 + 2 * q + 1]), pk2u(bb[2 * q], bb[2 * q + 1]));
            }
            float v[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (gelu) {
                v2[q] = gelu_erf2(v2[q]);
              } else if (sc_all) {
                v2[q] = mul2(v2[q], sc2);
              }
              upk2(v2[q], v[2 * q], v[2 * q + 1]);
            }
            if (sc_mixed) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (n0 + u * 8 + q < ep.scale_cols) v[q] *= ep.scale;
            }
            if (kIsConv && ep.relu) {
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
            }
            if (kIsConv && kPost == 1) {          // GRU gate
              float hh[8];
              if (post_ld_h) {
                h8_unpack(h_raw[u], hh);
              } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) hh[q] = (n0 >= ep.post_c) ? 0.f : 1.f;
              }
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = sigmoid_fast(v[q]) * hh[q];
            } else if (kIsConv && kPost == 2) {   // GRU update
              float hh[8], zz[8];
              if (post_ld_h) {
                h8_unpack(h_raw[u & 3], hh);
                h8_unpack(z_raw[u & 3], zz);
                if (u + 4 < 8) {
                  h_raw[u & 3] = __ldg(h_ptr + u + 4);
                  z_raw[u & 3] = __ldg(z_ptr + u + 4);
                }
              } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) hh[q] = zz[q] = 0.f;
              }
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = fmaf(zz[q], tanh_fast(v[q]) - hh[q], hh[q]);   // (1 - z) h + z tanh(v)
            }
            uint32_t last2 = pack_half2(v[6], v[7]);
            if (kIsConv && kPost == 3 && n0 + u * 8 + 8 == N)   // tail: the last two columns carry 2 f16 of another row buffer
              last2 = post_row >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(ep.aux_h + post_row * ep.ld_h)) : 0u;
            sts128(buf0 + row_s + ((u ^ sw) << 4), pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), last2);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            if (kIsConv && cv.taps) {  // this warp's 32 rows = rows_per_warp image rows of wb pixel slots; slots past the width are clipped
              const int ct = m_blk * (kCta2 ? 2 : 1) + cta_rank;
              const int cs = ct / cv.tiles_per_img;
              const int cti = ct - cs * cv.tiles_per_img, cty = cti / cv.tiles_x;
              const int sy = cty * cv.hb + quad * cv.rows_per_warp, sx = (cti - cty * cv.tiles_x) * cv.wo;
              if (kPost == 1 && n0 >= ep.post_c) {
                tma_store_4d(&tma_res, buf0, n0 - ep.post_c, sx, sy, cs);      // r * h -> the q convolution's input slot
              } else {
                tma_store_4d(&tma_out, buf0, n0, sx, sy, cs);
                if (kPost != 1 && ep.has_out2) tma_store_4d(&tma_res, buf0, n0, sx, sy, cs);   // second destination (GRU update: the dense copy of h)
              }
            } else {
              tma_store_2d(&tma_out, buf0, n0, row0);
            }
            bulk_commit();
          }
        }
        if (c_first >= kChunks) {  // this warp had no chunk in this tile (BN == 64): still release the stage
          tc_fence_before();
          if (lane == 0) { if constexpr (kCta2) mbar_arrive_leader(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]); }
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
      if (elect_one()) bulk_wait_all();
    } else if (ep.grp_rows <= 0) {
      // ---------------- fp32 output (+ residual): 32-column chunks, TMA load of the residual one chunk ahead,
      //                  TMA store ----------------
      constexpr int kChunks = BN / 32;
      const bool has_res = (ep.mode == CWM_EPI_RES_F32);
      uint64_t* my_res_bar = res_bar + ew * 2;
      // flat list of this warp's work items: (tile, chunk) with chunk = half, half+2, ...
      const int my_chunks = (kChunks - half + 1) / 2;
      const int my_tiles = (num_tiles > tile_first) ? (num_tiles - 1 - tile_first) / tile_stride + 1 : 0;
      const int n_items = my_tiles * my_chunks;
      auto item_coords = [&](int it, int& row0, int& n0, int& c, bool& first, bool& last) {
        const int ti = it / my_chunks;
        const int ci = it - ti * my_chunks;
        const int tile = tile_first + ti * tile_stride;
        const int m_blk = tile / tiles_n;
        const int n_blk = tile - m_blk * tiles_n;
        c = half + 2 * ci;
        row0 = m_blk * TM + cta_rank * BM + quad * 32;
        n0 = n_blk * BN + c * 32;
        first = (ci == 0);
        last = (ci == my_chunks - 1);
      };
      const bool emit = (ep.x16 != nullptr);  // LayerNorm fusion: f16 copy + partial row statistics
      float ps1 = 0.f, ps2 = 0.f;
      if (has_res && n_items > 0 && elect_one()) {  // prefetch the residual chunk of the first item
        int row0, n0, c;
        bool f, l;
        item_coords(0, row0, n0, c, f, l);
        mbar_arrive_expect_tx(&my_res_bar[0], kChunkBytes);
        tma_load_2d_s(buf0, &tma_res, smem_u32(&my_res_bar[0]), n0, row0);
      }
      for (int it = 0; it < n_items; ++it) {
        int row0, n0, c;
        bool first, last;
        item_coords(it, row0, n0, c, first, last);
        const uint32_t buf = buf0 + (it & 1) * kChunkBytes;
        const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
        float b_lo = 0.f;
        if (ep.bias != nullptr && n0 + lane < N) b_lo = __ldg(ep.bias + n0 + lane);
        if (first) {
          mbar_wait(&tfull_bar[as], aphase);
          tc_fence_after();
        }
        uint32_t acc[32];
        tmem_ld_x32(tmem_acc + c * 32, acc);
        __syncwarp();
        sts32(bias_s + lane * 4, __float_as_uint(b_lo));
        tmem_ld_wait();
        if (last) {
          tc_fence_before();
          if (lane == 0) { if constexpr (kCta2) mbar_arrive_leader(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]); }
        }
        __syncwarp();
        if (elect_one()) {
          // buffer (it+1)&1 was last used by item it-1: its store must have finished reading smem
          bulk_wait_read0();
          if (has_res && it + 1 < n_items) {
            int r1, n1, c1;
            bool f1, l1;
            item_coords(it + 1, r1, n1, c1, f1, l1);
            uint64_t* nb = &my_res_bar[(it + 1) & 1];
            mbar_arrive_expect_tx(nb, kChunkBytes);
            tma_load_2d_s(buf0 + ((it + 1) & 1) * kChunkBytes, &tma_res, smem_u32(nb), n1, r1);
          }
        }
        if (has_res) mbar_wait(&my_res_bar[it & 1], (it >> 1) & 1);
        __syncwarp();
        if (first) ps1 = ps2 = 0.f;
        uint64_t acc1 = pk2(0.f, 0.f), acc2 = pk2(0.f, 0.f);  // packed partial sums of this chunk
        uint32_t h16[16];
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // 16-byte unit u = columns 4u .. 4u+3
          uint32_t bb[4];
          lds128(bias_s + u * 16, bb[0], bb[1], bb[2], bb[3]);
          float v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = __uint_as_float(acc[u * 4 + q]) + __uint_as_float(bb[q]);
          const uint32_t addr = buf + row_s + ((u ^ sw) << 4);
          if (has_res) {
            uint32_t r[4];
            lds128(addr, r[0], r[1], r[2], r[3]);
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] += __uint_as_float(r[q]);
          }
          sts128(addr, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
          if (emit) {
            const uint64_t a2 = pk2(v[0], v[1]), b2 = pk2(v[2], v[3]);
            acc1 = add2(acc1, add2(a2, b2));
            acc2 = fma2(a2, a2, fma2(b2, b2, acc2));
            h16[2 * u] = pack_half2(v[0], v[1]);
            h16[2 * u + 1] = pack_half2(v[2], v[3]);
          }
        }
        if (emit) {
          float e0, e1;
          upk2(acc1, e0, e1);
          ps1 += e0 + e1;
          upk2(acc2, e0, e1);
          ps2 += e0 + e1;
          const int row = row0 + lane;
          // f16 copy: 64-byte rows in shared memory, 64B swizzle (16-byte unit ^= (row >> 1) & 3), one TMA store per chunk
          // (its buffer is free: the bulk_wait_read0 of this item also covered the previous chunk's f16 store)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(x16_s + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4), h16[4 * q], h16[4 * q + 1], h16[4 * q + 2], h16[4 * q + 3]);
          if (last && row < M) {
            const int n_blk_t = n0 / BN;
            ep.ln_stats_out[static_cast<long long>(n_blk_t * 2 + half) * M + row] = make_float2(ps1, ps2);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_2d(&tma_out, buf, n0, row0);
          if (emit) tma_store_2d(&tma_x16, x16_s, n0, row0);
          bulk_commit();
        }
        if (last) {
          if (++as == 2) {
            as = 0;
            aphase ^= 1;
          }
        }
      }
      if (elect_one()) bulk_wait_all();
    } else {
      // ---------------- fp32 epilogue with row remap / gathered residual (patch embed + pos[perm], and
      //                  encoder_to_decoder written into the decoder sequence): one thread owns one output row; its
      //                  destination / residual rows are resolved ONCE per tile (one division, one coalesced gather
      //                  load), then every 32-column chunk is 8 x 16-byte residual loads + 8 x 16-byte stores ----
      constexpr int kChunks = BN / 32;
      const bool has_res = (ep.mode == CWM_EPI_RES_F32);
      float* outp = reinterpret_cast<float*>(ep.out);
      const bool vec_ok = (ep.ldo % 4 == 0) && (!has_res || ep.ldr % 4 == 0) &&
                          (reinterpret_cast<uintptr_t>(ep.out) % 16 == 0) &&
                          (!has_res || reinterpret_cast<uintptr_t>(ep.res) % 16 == 0);
      for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
        const int m_blk = tile / tiles_n;
        const int n_blk = tile - m_blk * tiles_n;
        const int row0 = m_blk * TM + cta_rank * BM + quad * 32;
        const int m = row0 + lane;
        const bool valid = m < M;
        long long orow = 0, rrow = 0;
        if (valid) {
          const int g = m / ep.grp_rows;
          const int j = m - g * ep.grp_rows;
          orow = static_cast<long long>(g) * ep.grp_out_stride + j;
          rrow = orow;
          if (has_res && ep.res_gather != nullptr) rrow = __ldg(ep.res_gather + static_cast<long long>(g) * ep.gather_stride + j);
        }
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
        for (int c = half; c < kChunks; c += 2) {
          const int n0 = n_blk * BN + c * 32;
          float b_lo = 0.f;
          if (ep.bias != nullptr && n0 + lane < N) b_lo = __ldg(ep.bias + n0 + lane);
          uint32_t acc[32];
          tmem_ld_x32(tmem_acc + c * 32, acc);
          __syncwarp();  // previous chunk's broadcast reads of bias_s are done
          sts32(bias_s + lane * 4, __float_as_uint(b_lo));
          tmem_ld_wait();
          __syncwarp();
          if (!valid) continue;
          if (vec_ok && n0 + 32 <= N) {
            const float4* rp = has_res ? reinterpret_cast<const float4*>(ep.res + rrow * ep.ldr + n0) : nullptr;
            float4* op = reinterpret_cast<float4*>(outp + orow * ep.ldo + n0);
            float4 rr[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) rr[u] = has_res ? __ldg(rp + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              uint32_t bb[4];
              lds128(bias_s + u * 16, bb[0], bb[1], bb[2], bb[3]);
              float4 v;
              v.x = __uint_as_float(acc[4 * u]) + __uint_as_float(bb[0]) + rr[u].x;
              v.y = __uint_as_float(acc[4 * u + 1]) + __uint_as_float(bb[1]) + rr[u].y;
              v.z = __uint_as_float(acc[4 * u + 2]) + __uint_as_float(bb[2]) + rr[u].z;
              v.w = __uint_as_float(acc[4 * u + 3]) + __uint_as_float(bb[3]) + rr[u].w;
              op[u] = v;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              const int n = n0 + q;
              if (n < N) {
                float v = __uint_as_float(acc[q]) + __uint_as_float(lds32(bias_s + q * 4));
                if (has_res) v += __ldg(ep.res + rrow * ep.ldr + n);
                outp[orow * ep.ldo + n] = v;
              }
            }
          }
        }
        tc_fence_before();
        if (lane == 0) { if constexpr (kCta2) mbar_arrive_leader(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]); }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (kCta2) cluster_sync_all(); else __syncthreads();  // the peer may not exit while its smem / TMEM is in use
  if (warp == 2) {
    tc_fence_after();
    if constexpr (kCta2) tmem_dealloc2(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

static int g_gemm_cta2 = 1;  // CTA-pair kernels: on by default (cwm_debug_gemm_cta2(0) / CWM_GEMM_CTA2=0 switch them off)

template <int BN, bool kRes, bool kCta2, int kConv>
static int launch_gemm_impl(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to, const CUtensorMap& tr,
                            const CUtensorMap& tx, int M, int N, int K, const EpiDev& ep, cudaStream_t stream,
                            const ConvDev& cv_in) {
  using Cfg = GemmCfg<BN, kRes, (kConv > 0), kCta2>;
  static bool attr_set = false;
  if (!attr_set) {
    CWM_CUDA_CHECK(cudaFuncSetAttribute(gemm_f16_kernel<BN, kRes, kCta2, kConv>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_set = true;
  }
  ConvDev cv = cv_in;
  if (cv.halo) {  // carve the operand ring into halo stages + the weight ring (or the resident weight matrix)
    const int ring = Cfg::kRingBytes;
    cv.w_stride = Cfg::kBStageBytes;
    // a halo stage: the staged box + 1 KB of slack for the shifted reads of discarded output rows, 1024-byte aligned
    cv.h_stride = ((cv.halo_rows * cv.wb * 128 + 1024 + 1023) / 1024) * 1024;
    if (cv.h_stride > kHaloStageBytes) cv.h_stride = kHaloStageBytes;
    const int k_steps = cv.taps * cv.cin_slabs;
    static int resident_env = -1, hstage_env = -1;
    if (resident_env < 0) {
      const char* v = getenv("CWM_CONV_RESIDENT");
      resident_env = (v == nullptr) ? 1 : atoi(v);
      v = getenv("CWM_CONV_HALO_STAGES");
      hstage_env = (v == nullptr) ? 3 : atoi(v);
      if (hstage_env < 2) hstage_env = 2;
      if (hstage_env > 4) hstage_env = 4;
    }
    // the whole weight matrix fits beside two halo stages (the encoders' 64-channel 3x3 convolutions: 9 x 8 KB): stage it
    // once per CTA and give the rest of the ring to halo stages -- one staged halo is ~2 us of TMA latency away, and with
    // one halo per tile two stages leave a single load in flight (measured: 111 us per 112x112x64 convolution either way)
    cv.w_resident = (resident_env && N <= BN && k_steps <= 12 && 2 * cv.h_stride + k_steps * cv.w_stride <= ring) ? 1 : 0;
    const int w_min = cv.w_resident ? k_steps : 4;          // weight stages that must remain
    int hst = (ring - w_min * cv.w_stride) / cv.h_stride;
    hst = hst > hstage_env ? hstage_env : hst;
    cv.h_stages = hst < 2 ? 2 : hst;
    const int ws = (ring - cv.h_stages * cv.h_stride) / cv.w_stride;
    cv.w_stages = cv.w_resident ? k_steps : (ws > 12 ? 12 : ws);   // <= kBarSlots
    if (cv.w_stages < 2) return fail(CWM_ERR_INVALID, "convolution halo mode: no room for the weight ring (BN=%d)", BN);
  }
  if constexpr (kCta2) {
    const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
    int pairs = num_sms() / 2;
    if (tiles < pairs) pairs = tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    CWM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_f16_kernel<BN, kRes, true, kConv>, ta, tw, to, tr, tx, M, N, K, ep, cv));
    count_launch();
    return CWM_OK;
  } else {
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    CWM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_f16_kernel<BN, kRes, false, kConv>, ta, tw, to, tr, tx, M, N, K, ep, cv));
    count_launch();
    return CWM_OK;
  }
}

template <int BN, bool kRes, int kConv = 0>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to, const CUtensorMap& tr,
                       const CUtensorMap& tx, int M, int N, int K, const EpiDev& ep, cudaStream_t stream, bool cta2,
                       const ConvDev& cv = ConvDev{0, 1, 0, 0, 1 << 30, 1, 1, 1, 32, 0, 0, 0, 2, 33 * 1024, 1, 1, 0, 0, 32, 1, 0, 0}) {
  if (cta2) return launch_gemm_impl<BN, kRes, true, kConv>(ta, tw, to, tr, tx, M, N, K, ep, stream, cv);
  return launch_gemm_impl<BN, kRes, false, kConv>(ta, tw, to, tr, tx, M, N, K, ep, stream, cv);
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
  CWM_REQUIRE(!(f16_out && e->grp_rows > 0), "cwm_gemm_f16: row remap is only available for fp32 outputs");
  const int elt = f16_out ? 2 : 4;
  CWM_REQUIRE(reinterpret_cast<uintptr_t>(e->out) % 16 == 0 && (static_cast<long long>(e->ldo) * elt) % 16 == 0 &&
                  e->ldo >= N,
              "cwm_gemm_f16: out must be 16-byte aligned with ldo*elt %% 16 == 0 and ldo >= N (ldo=%d)", e->ldo);
  if (M == 0) return CWM_OK;
  EpiDev ep;
  ep.mode = e->mode; ep.bias = e->bias; ep.scale = e->scale; ep.scale_cols = (e->mode == CWM_EPI_F16) ? e->scale_cols : 0;
  ep.res = e->res; ep.ldr = e->ldr; ep.res_gather = e->res_gather; ep.gather_stride = e->gather_stride;
  ep.grp_rows = e->grp_rows; ep.grp_out_stride = e->grp_out_stride; ep.out = e->out; ep.ldo = e->ldo;
  ep.x16 = nullptr; ep.ldx16 = 0; ep.ln_stats_out = nullptr;
  ep.ln_stats_in = nullptr; ep.ln_parts = 0; ep.ln_s = nullptr; ep.ln_inv_c = 0.f; ep.ln_eps = 0.f; ep.relu = 0;
  ep.post = 0; ep.post_c = 0; ep.aux_h = nullptr; ep.ld_h = 0; ep.aux_z = nullptr; ep.ld_z = 0; ep.has_out2 = 0;
  ep.img_h = ep.img_w = ep.img_s = 0;
  if (e->ln_x16 != nullptr) {
    CWM_REQUIRE(e->mode == CWM_EPI_RES_F32 && e->grp_rows <= 0 && e->ln_stats_out != nullptr && N % 32 == 0 &&
                    reinterpret_cast<uintptr_t>(e->ln_x16) % 16 == 0 && e->ln_ldx16 % 8 == 0 && e->ln_ldx16 >= N,
                "cwm_gemm_f16: LayerNorm-producer outputs need the plain fp32-residual epilogue, N %% 32 == 0 and a "
                "16-byte aligned f16 copy");
    ep.x16 = reinterpret_cast<__half*>(e->ln_x16); ep.ldx16 = e->ln_ldx16;
    ep.ln_stats_out = reinterpret_cast<float2*>(e->ln_stats_out);
  }
  if (e->ln_stats_in != nullptr) {
    CWM_REQUIRE((e->mode == CWM_EPI_F16 || e->mode == CWM_EPI_GELU_F16) && e->ln_parts > 0 && e->ln_colsum != nullptr &&
                    e->ln_width > 0,
                "cwm_gemm_f16: the fused-LayerNorm epilogue needs an f16 output mode, partial statistics and column sums");
    ep.ln_stats_in = reinterpret_cast<const float2*>(e->ln_stats_in); ep.ln_parts = e->ln_parts; ep.ln_s = e->ln_colsum;
    ep.ln_inv_c = 1.0f / static_cast<float>(e->ln_width); ep.ln_eps = e->ln_eps;
  }
  const int bn = pick_bn(N);
  const bool remap = e->grp_rows > 0;
  // CTA pairs: non-remapped epilogues, BN >= 128 (each CTA stages BN/2 rows of W, a multiple of 8-row swizzle atoms),
  // enough rows for 256-row pair tiles
  static bool env_checked = false;
  if (!env_checked) {
    env_checked = true;
    const char* v = getenv("CWM_GEMM_CTA2");
    if (v != nullptr) g_gemm_cta2 = atoi(v);
  }
  const bool cta2 = g_gemm_cta2 != 0 && !remap && bn >= 128 && M >= 2 * BM;
  CUtensorMap ta, tw, to, tr;
  int rc = make_tmap_2d(&ta, A, CWM_TMAP_F16, M, K, K, BM, BK);
  if (rc) return rc;
  rc = make_tmap_2d(&tw, W, CWM_TMAP_F16, N, K, K, cta2 ? bn / 2 : bn, BK);
  if (rc) return rc;
  if (f16_out) {
    rc = make_tmap_2d(&to, e->out, CWM_TMAP_F16, M, N, e->ldo, 32, 64);
    tr = to;
  } else if (!remap) {
    rc = make_tmap_2d(&to, e->out, CWM_TMAP_F32, M, N, e->ldo, 32, 32);
    if (rc) return rc;
    if (e->mode == CWM_EPI_RES_F32) {
      CWM_REQUIRE(reinterpret_cast<uintptr_t>(e->res) % 16 == 0 && (static_cast<long long>(e->ldr) * 4) % 16 == 0,
                  "cwm_gemm_f16: residual must be 16-byte aligned");
      rc = make_tmap_2d(&tr, e->res, CWM_TMAP_F32, M, N, e->ldr, 32, 32);
    } else {
      tr = to;
    }
  } else {
    to = ta;  // unused by the generic epilogue
    tr = ta;
  }
  if (rc) return rc;
  CUtensorMap tx = to;  // f16 copy of the output rows (LayerNorm-producer epilogue): 32 x 32 boxes, 64B swizzle
  if (ep.x16 != nullptr) {
    rc = make_tmap_2d(&tx, ep.x16, CWM_TMAP_F16, M, N, ep.ldx16, 32, 32, 64);
    if (rc) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const char* kNames[4] = {"gemm_f16_out", "gemm_gelu_f16_out", "gemm_residual_f32", "gemm_f32_out"};
  const double out_bytes = static_cast<double>(M) * N * (f16_out ? 2.0 : (e->mode == CWM_EPI_RES_F32 ? 8.0 : 4.0));
  ProfileScope prof(s, kNames[e->mode], 2.0 * M * N * K,
                    (static_cast<double>(M) * K + static_cast<double>(N) * K) * 2.0 + out_bytes);
  if (f16_out) {
    // the two hot f16 epilogues of the VMAE blocks get compile-time variants (fused LayerNorm + q-scale on whole 64-column
    // chunks: qkv; fused LayerNorm + GELU: fc1)
    static int epi_env = -1;
    if (epi_env < 0) {
      const char* v = getenv("CWM_GEMM_EPI_VARIANTS");
      epi_env = (v == nullptr) ? 1 : atoi(v);
    }
    if (epi_env && ep.ln_stats_in != nullptr && ep.bias != nullptr && (bn == 256 || bn == 192)) {
      if (e->mode == CWM_EPI_GELU_F16) {
        if (bn == 256) return launch_gemm<256, false, -2>(ta, tw, to, tr, tx, M, N, K, ep, s, cta2);
        return launch_gemm<192, false, -2>(ta, tw, to, tr, tx, M, N, K, ep, s, cta2);
      }
      if (ep.scale_cols % 64 == 0) {
        if (bn == 256) return launch_gemm<256, false, -1>(ta, tw, to, tr, tx, M, N, K, ep, s, cta2);
        return launch_gemm<192, false, -1>(ta, tw, to, tr, tx, M, N, K, ep, s, cta2);
      }
    }
    switch (bn) {
      case 64: return launch_gemm<64, false>(ta, tw, to, tr, tx, M, N, K, ep, s, cta2);
      case 128: return launch_gemm<128, false>(ta, tw, to, tr, tx, M, N, K, ep, s, cta2);
      case 192: return launch_gemm<192, false>(ta, tw, to, tr, tx, M, N, K, ep, s, cta2);
      default: return launch_gemm<256, false>(ta, tw, to, tr, tx, M, N, K, ep, s, cta2);
    }
  }
  switch (bn) {
    case 64: return launch_gemm<64, true>(ta, tw, to, tr, tx, M, N, K, ep, s, cta2);
    case 128: return launch_gemm<128, true>(ta, tw, to, tr, tx, M, N, K, ep, s, cta2);
    case 192: return launch_gemm<192, true>(ta, tw, to, tr, tx, M, N, K, ep, s, cta2);
    default: return launch_gemm<256, true>(ta, tw, to, tr, tx, M, N, K, ep, s, cta2);
  }
}

// ---- convolution as an implicit GEMM on the same kernel (RAFT's recurrent block) ----
extern "C" int cwm_conv2d_weight_k(int Cin, int kh, int kw) { return kh * kw * ((Cin + BK - 1) / BK) * BK; }

struct ConvPost {
  int post = 0;            // 0 none, 1 GRU gate, 2 GRU update
  int C = 0;
  const uint16_t* h = nullptr;
  int ldh = 0;
  const uint16_t* z = nullptr;
  int ldz = 0;
  uint16_t* out2 = nullptr;
  int ldo2 = 0;
  int out2_cols = 0;
};

static int conv2d_impl(const uint16_t* x, int ldx, int S, int H, int W, int Cin, const uint16_t* w_packed, int Cout, int kh,
                       int kw, int pad_h, int pad_w, const float* bias, int relu, uint16_t* out, int ldo, int out_cols,
                       const ConvPost& post, cwm_stream_t stream, int stride = 1) {
  CWM_REQUIRE(x && w_packed && out, "cwm_conv2d_f16: null pointer");
  CWM_REQUIRE(S >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && kh > 0 && kw > 0, "cwm_conv2d_f16: bad shape");
  CWM_REQUIRE(stride == 1 || stride == 2, "cwm_conv2d_f16: stride %d (1 or 2)", stride);
  CWM_REQUIRE(kh == 2 * pad_h + 1 && kw == 2 * pad_w + 1, "cwm_conv2d_f16: only 'same'-padded convolutions (k = 2 pad + 1)");
  CWM_REQUIRE(Cin % 8 == 0 && ldx % 8 == 0 && ldx >= Cin && ldo % 8 == 0 && ldo >= out_cols && Cout % 8 == 0,
              "cwm_conv2d_f16: channel counts and row strides must be multiples of 8 (16-byte rows)");
  if (S == 0) return CWM_OK;
  // output map (torch.nn.Conv2d with padding = k // 2): (H + 2 pad - k) / stride + 1 = (H - 1) / stride + 1
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  CWM_REQUIRE(post.post == 0 || (Wo <= 32 && stride == 1), "cwm_conv2d_gru_*: maps of at most 32 pixels, stride 1");
  // a 128-row tile = hb image rows of wb pixel slots.  Maps of <= 32 pixels: one column block (wb = 16 / 32 >= Wo).  Wider
  // maps are tiled in both directions: stride-1 convolutions with more than one tap as 4 x 32-slot halo tiles that keep
  // 32 - 2 pad_w outputs per row (2-D halo mode), everything else with the wb in {8, 16, 32} that wastes the fewest slots
  static int halo_env = -1;
  if (halo_env < 0) {
    const char* v = getenv("CWM_CONV_HALO");
    halo_env = (v == nullptr) ? 1 : atoi(v);
  }
  int wb = Wo <= 16 ? 16 : 32;
  const bool halo2d = halo_env != 0 && Wo > 32 && stride == 1 && kh * kw > 1 && 32 - 2 * pad_w >= 16 &&
                      (BM / 32 + kh - 1) * 32 * 128 <= kHaloStageBytes - 1024;
  if (Wo > 32 && !halo2d) {
    long long best = -1;
    for (int cand = 32; cand >= 8; cand >>= 1) {
      const int hbc = BM / cand;
      const long long area = static_cast<long long>((Wo + cand - 1) / cand) * cand * ((Ho + hbc - 1) / hbc) * hbc;
      if (best < 0 || area < best) { best = area; wb = cand; }
    }
  }
  const int hb = BM / wb;
  ConvDev cv;
  cv.taps = kh * kw; cv.kw = kw; cv.pad_h = pad_h; cv.pad_w = pad_w; cv.cin_slabs = (Cin + BK - 1) / BK;
  cv.wo = halo2d ? wb - 2 * pad_w : wb;
  cv.tiles_x = (Wo + cv.wo - 1) / cv.wo; cv.stride = stride;
  cv.tiles_per_img = cv.tiles_x * ((Ho + hb - 1) / hb); cv.hb = hb; cv.rows_per_warp = 32 / wb;
  cv.w_resident = 0;
  cv.h_stages = 2;
  cv.h_stride = kHaloStageBytes;
  const int bn = pick_bn(Cout);
  // halo mode: every image row with its zero padding fits one row of wb pixel slots, and there is more than one tap
  cv.wb = wb;
  if (halo2d) {
    cv.halo = 1;
    cv.halo_rows = hb + kh - 1;
    cv.halo_x = pad_w;
    cv.tap_x = 0;
  } else {
    // one more row when the right-most taps of the right-most pixels read past their row's end (slot W - 1 + kw - 1 + pad_w
    // >= wb): they land on the NEXT row's leading zero slots, which must exist for the last row of the box too
    cv.halo_rows = hb + kh - 1 + ((W + 3 * pad_w > wb) ? 1 : 0);
    cv.halo = (halo_env != 0 && cv.taps > 1 && stride == 1 && cv.tiles_x == 1 && W + 2 * pad_w <= wb &&
               cv.halo_rows * wb * 128 <= kHaloStageBytes) ? 1 : 0;
    cv.halo_x = 2 * pad_w;
    cv.tap_x = pad_w;
  }
  cv.w_stages = 2;
  cv.w_stride = 0;
  {
    const char* v = getenv("CWM_CONV_BASEOFF");
    cv.base_off = (v == nullptr) ? 0 : atoi(v);  // 0 in production; 2 / 4 / 6: the timing experiments of DESIGN.md section 8
  }
  const int K = cv.taps * cv.cin_slabs * BK;
  const long long Mp = static_cast<long long>(S) * cv.tiles_per_img * BM;
  CWM_REQUIRE(Mp < (1ll << 31), "cwm_conv2d_f16: too many rows");
  const int M = static_cast<int>(Mp);
  EpiDev ep;
  ep.mode = CWM_EPI_F16; ep.bias = bias; ep.scale = 1.f; ep.scale_cols = 0; ep.res = nullptr; ep.ldr = 0;
  ep.res_gather = nullptr; ep.gather_stride = 0; ep.grp_rows = 0; ep.grp_out_stride = 0; ep.out = out; ep.ldo = ldo;
  ep.x16 = nullptr; ep.ldx16 = 0; ep.ln_stats_out = nullptr; ep.ln_stats_in = nullptr; ep.ln_parts = 0; ep.ln_s = nullptr;
  ep.ln_inv_c = 0.f; ep.ln_eps = 0.f; ep.relu = relu ? 1 : 0;
  ep.post = post.post; ep.post_c = post.C; ep.aux_h = reinterpret_cast<const __half*>(post.h); ep.ld_h = post.ldh;
  ep.aux_z = reinterpret_cast<const __half*>(post.z); ep.ld_z = post.ldz; ep.has_out2 = post.out2 != nullptr;
  ep.img_h = Ho; ep.img_w = Wo; ep.img_s = S;
  const bool cta2 = g_gemm_cta2 != 0 && bn >= 128 && M >= 2 * BM;
  CUtensorMap ta, tw, to;
  // stride 2: the box spans 2 hb x 2 wb input pixels and the map's traversal stride picks every second one
  int rc = make_tmap_nhwc(&ta, x, S, H, W, Cin, ldx, cv.halo ? cv.halo_rows : hb * stride, wb * stride, BK, stride);
  if (rc) return rc;
  rc = make_tmap_2d(&tw, w_packed, CWM_TMAP_F16, Cout, K, K, cta2 ? bn / 2 : bn, BK);
  if (rc) return rc;
  rc = make_tmap_nhwc(&to, out, S, Ho, Wo, out_cols, ldo, cv.rows_per_warp, cv.wo, 64);
  if (rc) return rc;
  CUtensorMap to2 = to;
  if (post.out2 != nullptr) {
    rc = make_tmap_nhwc(&to2, post.out2, S, Ho, Wo, post.out2_cols, post.ldo2, cv.rows_per_warp, cv.wo, 64);
    if (rc) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProfileScope prof(s, post.post == 1 ? "conv2d_gru_gate" : (post.post == 2 ? "conv2d_gru_update" :
                                                                (Wo > 32 ? (Cout <= 64 ? (cv.taps > 1 ? "conv2d_wide64_3x3" : "conv2d_wide64_1x1")
                                                                                       : (stride > 1 ? "conv2d_wide_s2" : "conv2d_wide"))
                                                                         : (stride > 1 ? "conv2d_s2" : "conv2d_f16"))),
                    2.0 * S * Ho * Wo * static_cast<double>(Cout) * cv.taps * Cin,
                    static_cast<double>(S) * (static_cast<double>(H) * W * Cin + static_cast<double>(Ho) * Wo * Cout) * 2.0 +
                        static_cast<double>(Cout) * K * 2.0);
  // one kernel per (N tile, post-op): the post-ops that exist are tied to their N tiles (gate: 2C = 256 columns; update and
  // tail: C = 128 columns), everything else is the plain bias / relu epilogue
  if (post.post == 1) {
    CWM_REQUIRE(bn == 256, "cwm_conv2d_gru_gate_f16: hidden width %d (2C must be a multiple of 256)", post.C);
    return launch_gemm<256, false, 2>(ta, tw, to, to2, to, M, Cout, K, ep, s, cta2, cv);
  }
  if (post.post == 2) {
    CWM_REQUIRE(bn == 128, "cwm_conv2d_gru_update_f16: hidden width %d (C must be a multiple of 128, not of 256)", post.C);
    return launch_gemm<128, false, 3>(ta, tw, to, to2, to, M, Cout, K, ep, s, cta2, cv);
  }
  if (post.post == 3) {
    CWM_REQUIRE(bn == 128, "cwm_conv2d_dual_f16 with a tail: %d output channels (a multiple of 128, not of 256)", Cout);
    return launch_gemm<128, false, 4>(ta, tw, to, to2, to, M, Cout, K, ep, s, cta2, cv);
  }
  switch (bn) {
    case 64: return launch_gemm<64, false, 1>(ta, tw, to, to2, to, M, Cout, K, ep, s, cta2, cv);
    case 128: return launch_gemm<128, false, 1>(ta, tw, to, to2, to, M, Cout, K, ep, s, cta2, cv);
    case 192: return launch_gemm<192, false, 1>(ta, tw, to, to2, to, M, Cout, K, ep, s, cta2, cv);
    default: return launch_gemm<256, false, 1>(ta, tw, to, to2, to, M, Cout, K, ep, s, cta2, cv);
  }
}

extern "C" int cwm_conv2d_f16(const uint16_t* x, int ldx, int S, int H, int W, int Cin, const uint16_t* w_packed, int Cout,
                              int kh, int kw, int pad_h, int pad_w, const float* bias, int relu, uint16_t* out, int ldo,
                              cwm_stream_t stream) {
  return conv2d_impl(x, ldx, S, H, W, Cin, w_packed, Cout, kh, kw, pad_h, pad_w, bias, relu, out, ldo, Cout, ConvPost(), stream);
}

extern "C" int cwm_conv2d_strided_f16(const uint16_t* x, int ldx, int S, int H, int W, int Cin, const uint16_t* w_packed,
                                      int Cout, int kh, int kw, int pad_h, int pad_w, int stride, const float* bias, int relu,
                                      uint16_t* out, int ldo, cwm_stream_t stream) {
  return conv2d_impl(x, ldx, S, H, W, Cin, w_packed, Cout, kh, kw, pad_h, pad_w, bias, relu, out, ldo, Cout, ConvPost(), stream,
                     stride);
}

extern "C" int cwm_conv2d_dual_f16(const uint16_t* x, int ldx, int S, int H, int W, int Cin, const uint16_t* w_packed, int Cout,
                                   int kh, int kw, int pad_h, int pad_w, const float* bias, int relu, const uint16_t* tail,
                                   int ld_tail, uint16_t* out, int ldo, uint16_t* out2, int ldo2, cwm_stream_t stream) {
  CWM_REQUIRE(out2 == nullptr || (ldo2 % 8 == 0 && ldo2 >= Cout), "cwm_conv2d_dual_f16: bad second row stride %d", ldo2);
  CWM_REQUIRE(tail == nullptr || (ld_tail >= 2 && ld_tail % 2 == 0 && (reinterpret_cast<uintptr_t>(tail) & 3) == 0),
              "cwm_conv2d_dual_f16: tail rows must be 4-byte aligned (ld %d)", ld_tail);
  ConvPost p;
  p.out2 = out2; p.ldo2 = ldo2; p.out2_cols = Cout;
  if (tail != nullptr) {
    p.post = 3; p.C = Cout; p.h = tail; p.ldh = ld_tail;
  }
  return conv2d_impl(x, ldx, S, H, W, Cin, w_packed, Cout, kh, kw, pad_h, pad_w, bias, relu, out, ldo, Cout, p, stream);
}

extern "C" int cwm_conv2d_gru_gate_f16(const uint16_t* x, int ldx, int S, int H, int W, int Cin, const uint16_t* w_zr, int C,
                                       int kh, int kw, int pad_h, int pad_w, const float* bias_zr, const uint16_t* h, int ldh,
                                       uint16_t* z_out, int ldz, uint16_t* rh_out, int ldrh, cwm_stream_t stream) {
  CWM_REQUIRE(h && z_out && rh_out && C % 64 == 0 && ldh % 8 == 0 && ldh >= C && ldz % 8 == 0 && ldrh % 8 == 0,
              "cwm_conv2d_gru_gate_f16: null pointer, or C not a multiple of 64 / rows not 16-byte aligned");
  ConvPost p;
  p.post = 1; p.C = C; p.h = h; p.ldh = ldh; p.out2 = rh_out; p.ldo2 = ldrh; p.out2_cols = C;
  return conv2d_impl(x, ldx, S, H, W, Cin, w_zr, 2 * C, kh, kw, pad_h, pad_w, bias_zr, 0, z_out, ldz, C, p, stream);
}

extern "C" int cwm_conv2d_gru_update_f16(const uint16_t* x, int ldx, int S, int H, int W, int Cin, const uint16_t* w_q, int C,
                                         int kh, int kw, int pad_h, int pad_w, const float* bias_q, const uint16_t* z, int ldz,
                                         uint16_t* h, int ldh, uint16_t* h_dense, cwm_stream_t stream) {
  CWM_REQUIRE(h && z && C % 64 == 0 && ldh % 8 == 0 && ldh >= C && ldz % 8 == 0 && ldz >= C,
              "cwm_conv2d_gru_update_f16: null pointer, or C not a multiple of 64 / rows not 16-byte aligned");
  ConvPost p;
  p.post = 2; p.C = C; p.h = h; p.ldh = ldh; p.z = z; p.ldz = ldz; p.out2 = h_dense; p.ldo2 = C; p.out2_cols = C;
  return conv2d_impl(x, ldx, S, H, W, Cin, w_q, C, kh, kw, pad_h, pad_w, bias_q, 0, h, ldh, C, p, stream);
}

extern "C" int cwm_debug_gemm_cta2(int enable) {
  g_gemm_cta2 = enable;
  return CWM_OK;
}

// number of partial-statistics planes a LayerNorm-producer GEMM with N output columns writes (2 per N tile)
extern "C" int cwm_gemm_ln_parts(int N) {
  const int bn = pick_bn(N);
  return 2 * ((N + bn - 1) / bn);
}
