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
// shared memory of a launch: tables + `teams` teams of E evaluators x (1 + H) warps each
inline size_t chain_smem_bytes(int ns, int teams, int E, int H, int hmax) {
  const size_t tab = ns == 20 ? chain_table_floats<20>() : chain_table_floats<10>();
  return sizeof(float) * (tab + (size_t)teams * wc::team_floats(E, H, hmax));
}

template <int NS, int K, bool LB, bool CAM, bool FIN>
__global__ void __launch_bounds__(32 * kChainMaxWarps, 1)
chain_kernel(const __grid_constant__ wc::ChainParams p, const ChainTables tab) {
  extern __shared__ __align__(16) float smem[];
  float* s_P = smem;
  float* s_mu = s_P + kGmmM * wc::kPFloats;
  float* s_nlw = s_mu + kGmmM * kMuStride;
  float* s_rel = s_nlw + kGmmM;
  float* s_team = s_rel + kMaxFitJoints * (1 + NS) * 4;
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
  // a team = E evaluator warps (the first one leads the sequence) + E x H helper warps that scan the mixture
  // components for them; warps of a team are consecutive, evaluators first, so that the (latency-critical)
  // evaluators of a team sit on different schedulers (warp id mod 4)
  const int E = p.team, H = p.helpers, TW = E * (1 + H), nteams = nwarps / TW;
  const int team = warp / TW, member = warp - team * TW;
  float* base = s_team + (size_t)team * wc::team_floats(E, H, p.hmax);
  wc::TeamMem tm = wc::make_team_mem(base, E, H, p.hmax);
  // named barriers (ids 1..15): a lone evaluator with helpers (Adam) uses the helper pair, ids 1 + 2 team, + 1 (up to
  // 7 teams per CTA); a team of several evaluators (L-BFGS, no helpers: k2b_fit_chain never combines the two) uses three:
  // round posted, tables ready, results ready (up to 5 teams per CTA)
  const int bar_b = 1 + 2 * team;
  tm.bar_go = 1 + 3 * team;
  tm.bar_tab = 2 + 3 * team;
  tm.bar_done = 3 + 3 * team;
  if (LB || member < E) {      // helper warps exist for Adam only (capi.cu, chain_geometry)
    wc::WarpMem wm = wc::make_warp_mem(base + (size_t)member * wc::kEvalMemFloats, tm.gs);
    wm.helpers = H;
    wm.bar_id = bar_b;
    wm.bar_threads = 32 * TW;
    wm.helper_mem = base + (size_t)(E + member * H) * wc::kEvalMemFloats;
    wm.helper_stride = wc::kEvalMemFloats;
    const long slot = (long)blockIdx.x * nteams + team;
    float* hist = p.hist ? p.hist + slot * wc::hist_floats(p.hmax) : nullptr;
    wc::run_evaluator<NS, K, LB, CAM, FIN>(p, tb, wm, tm, member, slot, (long)gridDim.x * nteams, hist);
  } else {
    const int e = (member - E) / H, h = (member - E) % H;
    float* own = base + (size_t)member * wc::kEvalMemFloats;
    const wc::WarpMem wm = wc::make_warp_mem(own, nullptr);
    wc::team_helper(tb, wm, own, base + (size_t)e * wc::kEvalMemFloats, h, H, bar_b, 32 * TW);
  }
}

// NS: shape coefficients (10 | 20); K: observed joints (22 | 24).  Specialised in chain_inst.cu.
template <int NS, int K>
cudaError_t launch_chain(const wc::ChainParams& p, const ChainTables& tab, int grid, int teams, cudaStream_t st);

}  // namespace k2b
