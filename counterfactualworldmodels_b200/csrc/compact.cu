// compact.cu -- a4: visible-token compaction from the temporally-factored mask.
//
// Reference semantics (cwm/models/VideoMAE/vmae.py:166-167, :555-557): `x[~mask]` keeps the visible tokens
// of every row in ascending token order; the decoder sequence is [visible ascending..., masked ascending...].
// One CTA per sample: each thread owns a contiguous chunk of tokens, a block-wide exclusive prefix sum of the
// per-thread visible counts gives every token its rank.  Integer work, bit-exact by construction.
//
// Algorithmic bytes per sample: Ntot (mask) read + 4*Ntot (perm) [+ 4*Ntot (inv_perm)] written.
#include "common.cuh"

namespace cwm {

constexpr int kCompactThreads = 256;

__global__ void __launch_bounds__(kCompactThreads)
compact_mask_kernel(const uint8_t* __restrict__ mask, int Ntot, int32_t* __restrict__ perm,
                    int32_t* __restrict__ inv_perm, int32_t* __restrict__ n_visible) {
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const uint8_t* row = mask + static_cast<size_t>(b) * Ntot;
  const int chunk = (Ntot + kCompactThreads - 1) / kCompactThreads;
  const int t0 = min(tid * chunk, Ntot);
  const int t1 = min(t0 + chunk, Ntot);

  int vis = 0;
  for (int t = t0; t < t1; ++t) vis += (row[t] == 0);

  // block exclusive scan of `vis`
  __shared__ int warp_sums[kCompactThreads / 32];
  __shared__ int total_vis;
  const int lane = tid & 31, warp = tid >> 5;
  int incl = vis;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < kCompactThreads / 32) ? warp_sums[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    if (lane < kCompactThreads / 32) warp_sums[lane] = wi - w;  // exclusive
    if (lane == kCompactThreads / 32 - 1) total_vis = wi;
  }
  __syncthreads();
  const int nvis = total_vis;
  int vrank = warp_sums[warp] + incl - vis;  // visible tokens before t0
  int32_t* prow = perm + static_cast<size_t>(b) * Ntot;
  int32_t* irow = inv_perm ? inv_perm + static_cast<size_t>(b) * Ntot : nullptr;
  for (int t = t0; t < t1; ++t) {
    const bool v = (row[t] == 0);
    const int pos = v ? vrank : nvis + (t - vrank);  // masked rank = tokens before t that are masked
    prow[pos] = t;
    if (irow) irow[t] = pos;
    vrank += v;
  }
  if (tid == 0) n_visible[b] = nvis;
}

}  // namespace cwm

extern "C" int cwm_compact_mask(const uint8_t* mask, int B, int Ntot, int32_t* perm, int32_t* inv_perm,
                                int32_t* n_visible, cwm_stream_t stream) {
  CWM_REQUIRE(B >= 0 && Ntot >= 0, "cwm_compact_mask: negative size (B=%d, Ntot=%d)", B, Ntot);
  if (B == 0 || Ntot == 0) return CWM_OK;
  CWM_REQUIRE(mask && perm && n_visible, "cwm_compact_mask: null pointer");
  cwm::ProfileScope prof(static_cast<cudaStream_t>(stream), "compact_mask", 0.0,
                         static_cast<double>(B) * Ntot * (1.0 + 4.0 + (inv_perm ? 4.0 : 0.0)));
  cwm::compact_mask_kernel<<<B, cwm::kCompactThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      mask, Ntot, perm, inv_perm, n_visible);
  CWM_LAUNCH_CHECK();
  return CWM_OK;
}
