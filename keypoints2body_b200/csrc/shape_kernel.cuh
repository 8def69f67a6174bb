// shape_kernel.cuh -- K4: shared-betas pre-pass over the first frames of each sequence.
//
// Restates optimize_shape_multi_frame (/root/reference/keypoints2body/core/shape.py:10-115):
//   loss(betas) = sum_t [ sum_k conf_k^2 |J_k(betas; pose_t) + (y_t0 - J_0) - y_tk|^2 + w^2 |betas|^2 ]
// minimised by torch.optim.LBFGS(lr=0.1, max_iter=num_iters, strong_wolfe) (shape.py:105-109).
// Mapping: one warp per sequence; lane l evaluates frames l, l+32, ...; loss and the 10-vector
// gradient are warp-reduced with shuffles; lane 0 advances the same L-BFGS machine the frame
// fitter uses (lbfgs_core.cuh).  Expression stays zero (the reference calls the model without
// it, shape.py:75-77).
#pragma once

#include <cuda_runtime.h>

#include "fit_core.cuh"
#include "lbfgs_core.cuh"

namespace k2b {

constexpr int kShapeWarps = 4;   // sequences per CTA

// One frame of the shape objective; adds d loss / d betas into grad[10]; returns the loss
// WITHOUT the prior term.  `rel` has row stride (1 + ns) float4 per joint.
K2B_HD float shape_frame_eval(const float4* rel, int ns, const int* parents, int K, const float* pose,
                              const float* tgt, const float* conf, const float* betas, float* grad) {
  M3 Rw[kMaxFitJoints];
  V3 t[kMaxFitJoints], sb[kMaxFitJoints];
  for (int j = 0; j < K; ++j) {
    const float4* e = rel + j * (1 + ns);
    const float4 r0 = e[0];
    V3 off = v3(r0.x, r0.y, r0.z);
    for (int s = 0; s < 10; ++s) {
      const float4 d = e[1 + s];
      off.x = fmaf(d.x, betas[s], off.x);
      off.y = fmaf(d.y, betas[s], off.y);
      off.z = fmaf(d.z, betas[s], off.z);
    }
    Rod o;
    const M3 R = rodrigues(v3(pose[3 * j], pose[3 * j + 1], pose[3 * j + 2]), o);
    const int pj = parents[j];
    if (pj < 0) {
      Rw[j] = R;
      t[j] = off;
    } else {
      Rw[j] = matmul(Rw[pj], R);
      t[j] = matvec(Rw[pj], off) + t[pj];
    }
  }
  // residuals after root alignment: e_k = (t_k - t_0) - (y_k - y_0)
  float loss = 0.f;
  V3 g0 = v3(0.f, 0.f, 0.f);
  sb[0] = g0;
  for (int k = 1; k < K; ++k) {
    const float c2 = conf ? conf[k] * conf[k] : 1.f;
    const V3 e = v3((t[k].x - t[0].x) - (tgt[3 * k] - tgt[0]), (t[k].y - t[0].y) - (tgt[3 * k + 1] - tgt[1]),
                    (t[k].z - t[0].z) - (tgt[3 * k + 2] - tgt[2]));
    loss = fmaf(c2, dot(e, e), loss);
    sb[k] = v3(2.f * c2 * e.x, 2.f * c2 * e.y, 2.f * c2 * e.z);
    g0 = g0 - sb[k];
  }
  sb[0] = g0;
  // subtree sums (children have larger indices), then d/d rel_j = Rw_p^T s_j
  for (int j = K - 1; j >= 1; --j) {
    const int pj = parents[j];
    const V3 rb = matvec_t(Rw[pj], sb[j]);
    const float4* e = rel + j * (1 + ns);
    for (int s = 0; s < 10; ++s) {
      const float4 d = e[1 + s];
      grad[s] = fmaf(d.x, rb.x, fmaf(d.y, rb.y, fmaf(d.z, rb.z, grad[s])));
    }
    sb[pj] = sb[pj] + sb[j];
  }
  // root: sb[0] is (numerically) zero -- the root position cancels under root alignment
  {
    const float4* e = rel;
    for (int s = 0; s < 10; ++s) {
      const float4 d = e[1 + s];
      grad[s] = fmaf(d.x, sb[0].x, fmaf(d.y, sb[0].y, fmaf(d.z, sb[0].z, grad[s])));
    }
  }
  return loss;
}

struct ShapeParams {
  const float* rel;        // [24][1+ns][4]
  const int* parents;      // [>=24]
  int ns, K;
  int num_seq, frames_per_seq;   // frames used per sequence
  long seq_stride_frames;        // frames between consecutive sequences in targets / poses
  int pose_per_frame;            // 0: one pose per sequence ([S][72]); 1: [S][T][72]
  int conf_per_seq;              // 0: conf [K]; 1: [S][K]
  int num_iters;
  float lr, w2;
  const float* targets;    // [S][seq_stride][K][3]
  const float* poses;
  const float* conf;       // or nullptr
  const float* init_betas; // [S][10]
  float* out_betas;        // [S][10]
  float* out_loss;         // [S]
  int* out_evals;          // [S] or nullptr
  float* scratch;          // [S][Vecs::floats_per_frame(10, hmax)]
  int hmax;
};

__global__ void __launch_bounds__(32 * kShapeWarps)
shape_pass_kernel(const __grid_constant__ ShapeParams p) {
  extern __shared__ __align__(16) float sm[];
  float4* s_rel = reinterpret_cast<float4*>(sm);                       // [24][1+ns]
  int* s_par = reinterpret_cast<int*>(sm + kMaxFitJoints * (1 + p.ns) * 4);
  float* s_xg = reinterpret_cast<float*>(s_par + kMaxFitJoints);       // [warps][2][10]
  for (int i = threadIdx.x; i < kMaxFitJoints * (1 + p.ns); i += blockDim.x)
    s_rel[i] = reinterpret_cast<const float4*>(p.rel)[i];
  for (int i = threadIdx.x; i < kMaxFitJoints; i += blockDim.x) s_par[i] = p.parents[i];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.x * kShapeWarps + warp;
  if (seq >= p.num_seq) return;
  float* x = s_xg + warp * 20;
  float* g = x + 10;
  if (lane < 10) x[lane] = p.init_betas[seq * 10 + lane];
  __syncwarp();

  Cols c{x, g, 1, 1};
  Vecs v{p.scratch + (long)seq * Vecs::floats_per_frame(10, p.hmax), 1, 10, p.hmax};
  Lbfgs<10> st;
  st.init();
  const float* conf = p.conf ? (p.conf_per_seq ? p.conf + (long)seq * p.K : p.conf) : nullptr;
  const float T = (float)p.frames_per_seq;
  int stage = 0;
  float loss_out = 0.f;
  while (true) {
    float betas[10], grad[10];
#pragma unroll
    for (int s = 0; s < 10; ++s) {
      betas[s] = x[s];
      grad[s] = 0.f;
    }
    float loss = 0.f;
    for (int t = lane; t < p.frames_per_seq; t += 32) {
      const long fr = (long)seq * p.seq_stride_frames + t;
      const float* pose = p.poses + (p.pose_per_frame ? fr * kPoseDim : (long)seq * kPoseDim);
      loss += shape_frame_eval(s_rel, p.ns, s_par, p.K, pose, p.targets + fr * p.K * 3, conf, betas, grad);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      loss += __shfl_xor_sync(0xffffffffu, loss, o);
#pragma unroll
      for (int s = 0; s < 10; ++s) grad[s] += __shfl_xor_sync(0xffffffffu, grad[s], o);
    }
    float bb = 0.f;
#pragma unroll
    for (int s = 0; s < 10; ++s) bb = fmaf(betas[s], betas[s], bb);
    loss = fmaf(T * p.w2, bb, loss);
    int done = 0;
    if (lane == 0) {
      const Cols ce = st.eval_cols(c, v);
#pragma unroll
      for (int s = 0; s < 10; ++s) ce.G(s) = fmaf(2.f * T * p.w2, betas[s], grad[s]);
      st.advance_now(c, v, loss, stage == 0, p.num_iters, p.lr);
      done = st.done;
      if (done) {
        for (int s = 0; s < 10; ++s) x[s] = v.at(s);
        loss_out = (float)st.loss;
      }
    }
    stage = 1;
    done = __shfl_sync(0xffffffffu, done, 0);
    __syncwarp();
    if (done) break;
  }
  if (lane < 10) p.out_betas[seq * 10 + lane] = x[lane];
  if (lane == 0) {
    p.out_loss[seq] = loss_out;
    if (p.out_evals) p.out_evals[seq] = st.evals;
  }
}

inline size_t shape_smem_bytes(int ns) {
  return sizeof(float) * (size_t)(kMaxFitJoints * (1 + ns) * 4 + kMaxFitJoints + kShapeWarps * 20);
}

}  // namespace k2b
