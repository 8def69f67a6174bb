// chain_kernel.cuh -- the warp-per-sequence fitting kernel (K5).
//
// One warp walks one sequence frame by frame (chain_core.cuh); a CTA holds up to 12 warps next to the
// tables they share: the eight symmetric 69 x 69 precision matrices (159 KB), the mixture means and the
// rest-offset table.  Per-warp shared memory (~4 KB): the evaluation point, the double-buffered GMM
// difference vectors, the per-lane partial sums of the eight components, the arg-min component's gradient,
// four L-BFGS gradient slots and the rho / alpha arrays.  The L-BFGS (y, s) history lives in global scratch
// (one block per walking warp, L2-resident), read one pair ahead.
// With few sequences up to three helper warps per sequence scan the mixture prior (bound by the FMA pipe of
// one scheduler per warp) while the first warp walks the kinematic tree; they meet through named barriers.
// Used for the reference's default sequence schedule (serial in t) and for small batches, where the
// one-thread-per-frame kernel is latency-bound.
#pragma once

#include <cuda_runtime.h>

#include "chain_core.cuh"

namespace k2b {

constexpr int kChainMaxWarps = 12;

struct ChainTables {       // global-memory copies owned by k2b_model
  const float* P;          // [8][69][72]
  const float* mu;         // [8][72]
  const float* nlw;        // [8]
  const float* rel;        // [24][1+NS][4]
};

template <int NS>
constexpr size_t chain_table_floats() {
  return (size_t)kGmmM * wc::kPFloats + kGmmM * kMuStride + kGmmM + kMaxFitJoints * (1 + NS) * 4;
}
inline size_t chain_smem_bytes(int ns, int warps, int hmax) {
  const size_t tab = ns == 20 ? chain_table_floats<20>() : chain_table_floats<10>();
  return sizeof(float) * (tab + (size_t)warps * wc::warp_mem_floats(hmax));
}

template <int NS, int K>
__global__ void __launch_bounds__(32 * kChainMaxWarps, 1)
chain_kernel(const __grid_constant__ wc::ChainParams p, const ChainTables tab) {
  extern __shared__ __align__(16) float smem[];
  float* s_P = smem;
  float* s_mu = s_P + kGmmM * wc::kPFloats;
  float* s_nlw = s_mu + kGmmM * kMuStride;
  float* s_rel = s_nlw + kGmmM;
  float* s_warp = s_rel + kMaxFitJoints * (1 + NS) * 4;
  const int tid = threadIdx.x, nthr = blockDim.x;
  {
    const float4* src = reinterpret_cast<const float4*>(tab.P);
    float4* dst = reinterpret_cast<float4*>(s_P);
    for (int i = tid; i < kGmmM * wc::kPFloats / 4; i += nthr) dst[i] = src[i];
    src = reinterpret_cast<const float4*>(tab.mu);
    dst = reinterpret_cast<float4*>(s_mu);
    for (int i = tid; i < kGmmM * kMuStride / 4; i += nthr) dst[i] = src[i];
    if (tid < kGmmM) s_nlw[tid] = tab.nlw[tid];
    src = reinterpret_cast<const float4*>(tab.rel);
    dst = reinterpret_cast<float4*>(s_rel);
    for (int i = tid; i < kMaxFitJoints * (1 + NS); i += nthr) dst[i] = src[i];
  }
  __syncthreads();

  const int warp = tid >> 5, nwarps = nthr >> 5;
  const wc::WarpTables tb{s_P, s_mu, s_nlw, reinterpret_cast<const float4*>(s_rel)};
  const int wmf = wc::warp_mem_floats(p.hmax);
  // a group = the warp that walks the sequence + p.helpers warps that scan the mixture components for it
  const int G = 1 + p.helpers, ngroups = nwarps / G;
  // the walking warp's position inside its group rotates with the group index, so the (latency-critical) walkers
  // of neighbouring groups sit on different schedulers (warp id mod 4)
  const int group = warp / G, role = (warp - group * G - group % G + G) % G;
  float* w = s_warp + (size_t)(group * G + role) * wmf;     // shared-memory blocks are laid out by role
  wc::WarpMem wm = wc::make_warp_mem(w);
  const int bar = 1 + 2 * group;          // named barriers bar (work posted) and bar + 1 (results ready)
  if (role == 0) {
    wm.helpers = p.helpers;
    wm.bar_id = bar;
    wm.helper_mem = w + wmf;
    wm.helper_stride = wmf;
    float* ro = w + wc::kWarpMemFloats;
    float* al = ro + p.hmax;
    const long slot = (long)blockIdx.x * ngroups + group;
    float* hist = p.hist ? p.hist + slot * wc::hist_floats(p.hmax) : nullptr;
    for (long seq = slot; seq < p.num_seq; seq += (long)gridDim.x * ngroups)
      wc::run_chain_warp<NS, K>(p, seq, tb, wm, hist, ro, al);
    if (p.helpers > 0) {                  // release the helpers
      if ((tid & 31) == 0) wm.dbuf[0] = 0.f;
      wc::bar_arrive(bar, 32 * G);
    }
  } else {
    const float* lead = w - (size_t)role * wmf;       // the leader's block: xs at 0, command word at dbuf[0]
    const int lane = tid & 31;
    while (true) {
      wc::bar_sync(bar, 32 * G);
      const float cmd = *reinterpret_cast<const volatile float*>(lead + 96);
      if (cmd == 0.f) break;
      float xr[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) xr[c] = *reinterpret_cast<const volatile float*>(lead + 3 * lane + c);
      float best;
      int bm;
      wc::gmm_scan(tb, wm, xr, role - 1, p.helpers, cmd == 2.f, best, bm);
      if (lane == 0) {
        w[0] = best;
        w[1] = (float)bm;
      }
      wc::bar_arrive(bar + 1, 32 * G);
    }
  }
}

// NS: shape coefficients (10 | 20); K: observed joints (22 | 24).  Specialised in chain_inst.cu.
template <int NS, int K>
cudaError_t launch_chain(const wc::ChainParams& p, const ChainTables& tab, int grid, int warps, cudaStream_t st);

}  // namespace k2b
