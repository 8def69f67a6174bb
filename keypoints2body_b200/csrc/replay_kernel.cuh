// replay_kernel.cuh -- line-search conformance on the device (gate G3).
//
// Every strong-Wolfe line search torch performed in a recorded reference run is replayed with the objective
// replaced by the recorded (f, g.d) responses: the device machine (lbfgs_core.cuh) must propose the same trial
// steps, use the same number of evaluations and return the same (t, f) as torch's _strong_wolfe
// (torch/optim/lbfgs.py:40-209).  Both vector policies are covered: ThreadOps (one thread per line search, what
// fit_kernel instantiates) and WarpOps (one warp per line search, lane-distributed vectors, what chain_kernel
// instantiates).  The same hook runs on the CPU in tests/host_emul.
#pragma once

#include <cuda_runtime.h>

#include "chain_core.cuh"
#include "lbfgs_core.cuh"

namespace k2b {

struct ReplayParams {
  int num_searches, max_resp;
  const double* t0; const double* f0; const float* gtd0; const double* d_norm;
  const int* max_ls; const unsigned char* t_is_f32; const int* n_resp;
  const double* resp_f; const float* resp_gtd;      // [N][max_resp]
  double* out_t;                                     // [N][max_resp] proposed trial steps
  double* out_final;                                 // [N][3] returned t, returned f, evaluations used
  int* out_k;                                        // [N] responses consumed
};

constexpr int kReplayThreads = 128;

__global__ void __launch_bounds__(kReplayThreads) replay_thread_kernel(const ReplayParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.num_searches) return;
  float x = 0.f, g = 0.f;
  float scratch[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) scratch[k] = 0.f;
  Cols c{&x, &g, 1, 1};
  Vecs v{scratch, 1, 1, 1};
  Lbfgs<1> st;
  st.init();
  st.ls_replay_begin(c, v, p.t0[i], p.f0[i], p.gtd0[i], p.d_norm[i], p.max_ls[i], p.t_is_f32[i] != 0);
  int k = 0;
  const long row = (long)i * p.max_resp;
  while (!st.ls_replay_finished && k < p.n_resp[i]) {
    p.out_t[row + k] = st.t;
    ThreadOps<1>::replay_response(v, st.cur, p.resp_gtd[row + k]);
    st.after_eval(c, v, (float)p.resp_f[row + k]);
    ++k;
  }
  p.out_final[3 * i + 0] = st.t;
  p.out_final[3 * i + 1] = st.loss;
  p.out_final[3 * i + 2] = (double)st.ls_evals;
  p.out_k[i] = k;
}

// one warp per line search: the machine runs on lane-distributed vectors exactly as in chain_kernel
__global__ void __launch_bounds__(kReplayThreads) replay_warp_kernel(const ReplayParams p) {
  __shared__ float s_gs[kReplayThreads / 32][4 * wc::kWarpVec];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * (kReplayThreads / 32) + warp;
  if (i >= p.num_searches) return;
  wc::WVec v;
  v.gs = s_gs[warp];
  v.hist = nullptr;
  v.ro = nullptr;
  v.al = nullptr;
  v.hmax = 1;
  Lbfgs<85, wc::WarpOps> st;
  st.init();
  st.ls_replay_begin(v, v, p.t0[i], p.f0[i], p.gtd0[i], p.d_norm[i], p.max_ls[i], p.t_is_f32[i] != 0);
  int k = 0;
  const long row = (long)i * p.max_resp;
  while (!st.ls_replay_finished && k < p.n_resp[i]) {
    if (lane == 0) p.out_t[row + k] = st.t;
    wc::WarpOps::replay_response(v, st.cur, p.resp_gtd[row + k]);
    st.after_eval(v, v, (float)p.resp_f[row + k]);
    ++k;
  }
  if (lane == 0) {
    p.out_final[3 * i + 0] = st.t;
    p.out_final[3 * i + 1] = st.loss;
    p.out_final[3 * i + 2] = (double)st.ls_evals;
    p.out_k[i] = k;
  }
}

}  // namespace k2b
