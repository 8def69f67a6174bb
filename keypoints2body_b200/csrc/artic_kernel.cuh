// artic_kernel.cuh -- the general articulated fit as a kernel: one thread per frame (artic_core.cuh).
#pragma once

#include <cuda_runtime.h>

#include "artic_core.cuh"

namespace k2b {
namespace ar {

constexpr int kArticThreads = 64;
constexpr int kMaxObs = 128;
constexpr int kArticAdamTable = 64;
enum { kArticEval = 0, kArticAdam = 1, kArticLbfgs = 2 };

struct ArticFitParams {
  ArticModel M;
  long num_frames;
  int K, mode, iters, conf_per_frame, hmax;
  float lr, joint_w2, keep_scale;      // keep_scale = pose_preserve_weight^2 when the temporal term is on, else 0
  const int* obs_idx;                  // [K]
  const float* targets;                // [B][K][3]
  const float* conf;                   // [K] | [B][K] | null
  const float* init_x;                 // [B][n]
  const float* keep_x;                 // [B][n] or null = init_x
  const unsigned char* frozen;         // [n] or null
  float* out_x; float* out_loss; float* out_grad; float* out_points; int* out_evals; int* out_comp;
  float* ws;                           // L-BFGS vectors, floats_per_frame(n, hmax) rows of `slots` floats
  float adam_step[kArticAdamTable], adam_bc2[kArticAdamTable];
};

// The fit of one frame: WorldSpaceFitter.fit_frame / MANOFitter.fit_frame / FLAMEFitter.fit_frame semantics --
// Adam: torch single-tensor steps, returned loss = the last iteration's, before its step; L-BFGS: torch's machine
// (lbfgs_core.cuh), returned loss re-evaluated at the returned parameters.
template <bool WARP>
K2B_AR_FN void artic_fit_frame(const ArticFitParams& p, long f, float* lb_base, long lb_stride) {
  const ArticModel& M = p.M;
  const int n = M.n, K = p.K;
  float x[kMaxParams], g[kMaxParams], keepv[kMaxParams], tgt[3 * kMaxObs], wgt[kMaxObs];
  for (int i = 0; i < n; ++i) {
    x[i] = p.init_x[f * n + i];
    keepv[i] = p.keep_x ? p.keep_x[f * n + i] : x[i];
  }
  for (int k = 0; k < K; ++k) {
    for (int c = 0; c < 3; ++c) tgt[3 * k + c] = p.targets[(f * K + k) * 3 + c];
    const float cf = p.conf ? (p.conf_per_frame ? p.conf[f * K + k] : p.conf[k]) : 1.f;
    wgt[k] = p.joint_w2 * cf * cf;
  }
  float* pts = p.out_points ? p.out_points + f * K * 3 : nullptr;
  float loss = 0.f;
  int evals = 0, comp = 0;
  if (p.mode == kArticEval) {
    loss = artic_eval<WARP>(M, x, p.obs_idx, K, tgt, wgt, keepv, p.keep_scale, true, g, pts, &comp);
    for (int i = 0; i < n; ++i) p.out_grad[f * n + i] = g[i];
    if (p.out_comp) p.out_comp[f] = comp;
  } else if (p.mode == kArticAdam) {
    float m1[kMaxParams], m2[kMaxParams];
    for (int i = 0; i < n; ++i) m1[i] = m2[i] = 0.f;
    for (int k = 1; k <= p.iters; ++k) {
      float step_k, bc2_k;
      if (k <= kArticAdamTable) {
        step_k = p.adam_step[k - 1];
        bc2_k = p.adam_bc2[k - 1];
      } else {
        step_k = (float)((double)p.lr / (1.0 - pow(0.9, (double)k)));
        bc2_k = (float)sqrt(1.0 - pow(0.999, (double)k));
      }
      loss = artic_eval<WARP>(M, x, p.obs_idx, K, tgt, wgt, keepv, p.keep_scale, true, g, nullptr, nullptr);
      ++evals;
      for (int i = 0; i < n; ++i)
        if (!(p.frozen && p.frozen[i])) adam_update(x[i], m1[i], m2[i], g[i], step_k, bc2_k);
    }
    if (pts) artic_eval<WARP>(M, x, p.obs_idx, K, tgt, wgt, keepv, p.keep_scale, false, g, pts, nullptr);
  } else {
    Vecs v{lb_base, lb_stride, n, p.hmax};
    Cols c{x, g, 1, 1};
    Lbfgs<0, ThreadOps<0>> st;
    st.init();
    bool first = true;
    while (true) {
      const float l = artic_eval<WARP>(M, x, p.obs_idx, K, tgt, wgt, keepv, p.keep_scale, true, g, nullptr, nullptr);
      const Cols ce = st.eval_cols(c, v);
      for (int i = 0; i < n; ++i) ce.G(i) = (p.frozen && p.frozen[i]) ? 0.f : g[i];
      st.advance_now(c, v, l, first, p.iters, p.lr);
      first = false;
      if (st.done) break;
    }
    evals = st.evals;
    for (int i = 0; i < n; ++i) x[i] = v.at(i);
    loss = artic_eval<WARP>(M, x, p.obs_idx, K, tgt, wgt, keepv, p.keep_scale, false, g, pts, nullptr);
  }
  for (int i = 0; i < n; ++i) p.out_x[f * n + i] = x[i];
  p.out_loss[f] = loss;
  if (p.out_evals) p.out_evals[f] = evals;
}

#if defined(__CUDACC__)
// one thread per frame (throughput: large batches)
__global__ void __launch_bounds__(kArticThreads) artic_fit_kernel(const __grid_constant__ ArticFitParams p) {
  const long slots = (long)gridDim.x * blockDim.x;
  const long slot = (long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long f = slot; f < p.num_frames; f += slots) artic_fit_frame<false>(p, f, p.ws ? p.ws + slot : nullptr, slots);
}
// one warp per frame (latency: the reference's B = 1 calls); every lane keeps its own L-BFGS scratch column
__global__ void __launch_bounds__(kArticThreads) artic_fit_warp_kernel(const __grid_constant__ ArticFitParams p) {
  const long slots = (long)gridDim.x * blockDim.x;
  const long slot = (long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long f = slot >> 5; f < p.num_frames; f += slots >> 5) artic_fit_frame<true>(p, f, p.ws ? p.ws + slot : nullptr, slots);
}
#endif

}  // namespace ar
}  // namespace k2b
