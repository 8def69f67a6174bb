// fit_kernel.cuh -- the fused, persistent fitting kernel (K1).
//
// Mapping: one thread = one frame, one CTA = fit_threads<NS>() frames (384: 12 warps, the most
// that fits next to the tables), grid = #SMs, CTAs loop over frame tiles.  Shared memory holds
// (a) the tables every frame shares -- the panel-packed Cholesky factors of the 8 mixture
// precisions (85 KB), the mixture means, the rest-pose joint offset table -- loaded once per
// CTA, and (b) the per-frame parameter vector x as columns [element][thread] (bank-conflict
// free; no block-level synchronisation after the table load, so warps drift apart and overlap
// each other's latencies).  Iterations run inside the kernel, so x never leaves shared memory
// between optimiser steps.  The gradient, Adam moments / L-BFGS vectors and the per-frame
// observations live in a transposed global scratch with one column per RESIDENT thread
// (coalesced; ~25 MB per GPU for Adam, L2-resident), not per frame.
#pragma once

#include <cuda_runtime.h>

#include "fit_core.cuh"
#include "lbfgs_core.cuh"

#ifndef K2B_LBFGS_HINTS
#define K2B_LBFGS_HINTS false
#endif

namespace k2b {

template <int NS>
struct FitThreads { static constexpr int value = NS == 20 ? 320 : 384; };
template <int NS>
__host__ __device__ constexpr int fit_threads() { return FitThreads<NS>::value; }
inline int fit_threads_rt(int ns) { return ns == 20 ? 320 : 384; }
constexpr int kAdamTable = 256;

// kModeAdamWorld / kModeLbfgsWorld: the same optimisers for launches that are plain world-space fits without a forward
// pass at the returned parameters (loss_kind 0, final_loss_mode 0, no out_joints -- what the frame-parallel schedule
// launches): the camera stage and the final round are compiled out (Adam 128.7 -> 117.6 ms per 1 M-frame two-sweep
// step, L-BFGS 304.4 -> 293.7 ms in a same-box A/B; bit-identical results).
enum FitMode { kModeEval = 0, kModeAdam = 1, kModeLbfgs = 2, kModeAdamWorld = 3, kModeLbfgsWorld = 4 };

struct DeviceTables {      // global-memory copies owned by k2b_model
  const float* chol;       // [8][kCholStride]
  const float* mu;         // [8][kMuStride]
  const float* nlw;        // [8]
  const float* rel;        // [24][1+NS][4]
};

struct FitParams {
  DeviceTables tab;
  long num_frames;
  long rows;               // scratch rows per CTA; layout [CTA][row][thread] so the row stride is compile-time
  int num_obs;
  int num_iters;
  int freeze_betas;
  int conf_per_frame;
  int preserve_all;
  float lr, joint_w2, keep_w2;
  const float* targets; const float* conf;
  const float* init_pose; const float* init_betas; const float* init_transl; const float* init_expr;
  const float* preserve_pose;
  const int* frame_iters; const unsigned char* frame_preserve;
  float* out_pose; float* out_betas; float* out_transl; float* out_expr;
  float* out_loss; float* out_joints; int* out_evals;
  // evaluate-only outputs
  float* out_grad_pose; float* out_grad_betas; float* out_grad_transl; float* out_grad_expr;
  int* out_gmm_component;
  float* scratch;          // transposed per-frame scratch, see scratch_floats_per_frame
  int lbfgs_hmax;
  int loss_kind;           // 0 body_fitting_loss_3d (world / camera stage 2), 1 camera_fitting_loss_3d (camera stage 1)
  int final_mode;          // 1: returned loss = priors + joints at the final parameters, no preserve (camera_space.py:316-326)
  const float* depth_ref;  // [B][3] initial camera translation (loss_kind 1)
  float depth_w2;
  int debug_rounds;        // -DK2B_DIAG builds only (K2B_DEBUG_ROUNDS): pack the warp's round count into out_evals
  int outer_quorum;        // lanes waiting at an outer-iteration boundary that trigger the direction update
  int adam_fuse;           // 1: body-pose Adam steps inside the gradient pass (K2B_ADAM_FUSE=0 turns it off; same results)
};

struct AdamTable {
  float step[kAdamTable];  // lr / (1 - 0.9^k), k = 1..
  float bc2[kAdamTable];   // sqrt(1 - 0.999^k)
};

// scratch rows (each row is `stride` floats): targets 72, weights 24, preserve 69, gradient
// 95, then optimiser state.
constexpr int kScrTgt = 0;
constexpr int kScrWgt = 72;
constexpr int kScrKeep = 96;
constexpr int kScrGrad = 165;
constexpr int kScrOpt = 260;

inline long scratch_rows(int ns, int mode, int hmax) {
  const int n = 75 + ns;
  if (mode == kModeAdam) return kScrOpt + 2 * n;
  if (mode == kModeLbfgs) return kScrOpt + Vecs::floats_per_frame(n, hmax);
  return kScrOpt;
}

template <int NS>
constexpr size_t fit_smem_bytes() {
  return sizeof(float) * (size_t)(kGmmM * kCholStride + kGmmM * kMuStride + kGmmM +
                                  kMaxFitJoints * (1 + NS) * 4 + (75 + NS) * fit_threads<NS>());
}

template <int NS, int K, int MODE>
__global__ void __launch_bounds__(fit_threads<NS>(), 1)
fit_kernel(const __grid_constant__ FitParams p, const __grid_constant__ AdamTable at) {
  constexpr bool kWorldOnly = MODE >= kModeAdamWorld;
  constexpr int OPT = kWorldOnly ? MODE - 2 : MODE;
  constexpr int NX = 75 + NS;
  constexpr int kFitThreads = fit_threads<NS>();
  extern __shared__ __align__(16) float smem[];
  float* s_chol = smem;
  float* s_mu = s_chol + kGmmM * kCholStride;
  float* s_nlw = s_mu + kGmmM * kMuStride;
  float* s_rel = s_nlw + kGmmM;
  float* s_x = s_rel + kMaxFitJoints * (1 + NS) * 4;

  const int tid = threadIdx.x;
  // ---- tables: one cooperative, vectorised copy per CTA --------------------------------
  {
    const float4* src = reinterpret_cast<const float4*>(p.tab.chol);
    float4* dst = reinterpret_cast<float4*>(s_chol);
    for (int i = tid; i < kGmmM * kCholStride / 4; i += kFitThreads) dst[i] = src[i];
    src = reinterpret_cast<const float4*>(p.tab.mu);
    dst = reinterpret_cast<float4*>(s_mu);
    for (int i = tid; i < kGmmM * kMuStride / 4; i += kFitThreads) dst[i] = src[i];
    if (tid < kGmmM) s_nlw[tid] = p.tab.nlw[tid];
    src = reinterpret_cast<const float4*>(p.tab.rel);
    dst = reinterpret_cast<float4*>(s_rel);
    for (int i = tid; i < kMaxFitJoints * (1 + NS); i += kFitThreads) dst[i] = src[i];
  }
  __syncthreads();

  FitTables tb;
  tb.chol = s_chol;
  tb.mu = s_mu;
  tb.nlw = s_nlw;
  tb.rel = reinterpret_cast<const float4*>(s_rel);
  // this thread's scratch column: rows are kFitThreads floats apart (compile-time -> immediate offsets)
  constexpr long kStride = kFitThreads;
  float* const scr = p.scratch + (long)blockIdx.x * p.rows * kFitThreads + tid;
  Cols c{s_x + tid, scr + kScrGrad * kStride, kFitThreads, kStride};

  const long num_tiles = (p.num_frames + kFitThreads - 1) / kFitThreads;
  for (long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const long f = tile * kFitThreads + tid;
    const bool valid = f < p.num_frames;
    const long fr = valid ? f : p.num_frames - 1;   // padding lanes mirror the last frame

    // ---- load this frame: parameters -> smem columns, observations -> scratch ---------
    {
      const float* ip = p.init_pose + fr * kPoseDim;
#pragma unroll 4
      for (int i = 0; i < kPoseDim; ++i) c.X(i) = ip[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) c.X(kTranslOff + i) = p.init_transl[fr * 3 + i];
#pragma unroll
      for (int i = 0; i < 10; ++i) c.X(kShapeOff + i) = p.init_betas[fr * 10 + i];
      if (NS == 20) {
#pragma unroll
        for (int i = 0; i < 10; ++i) c.X(kShapeOff + 10 + i) = p.init_expr[fr * 10 + i];
      }
      const float* tg = p.targets + fr * K * 3;
#pragma unroll 2
      for (int i = 0; i < K * 3; ++i) scr[(kScrTgt + i) * kStride] = tg[i];
#pragma unroll 2
      for (int j = 0; j < K; ++j) {
        const float cf = p.conf ? (p.conf_per_frame ? p.conf[fr * K + j] : p.conf[j]) : 1.f;
        float wj = p.joint_w2 * cf * cf;
        // camera stage 1 looks at RHip, LHip, RShoulder, LShoulder only, unweighted (losses.py:80-92)
        if (!kWorldOnly && p.loss_kind == 1) wj = (j == 1 || j == 2 || j == 16 || j == 17) ? 1.f : 0.f;
        scr[(kScrWgt + j) * kStride] = wj;
      }
      const float* kp = p.preserve_pose ? p.preserve_pose + fr * kBodyDim : p.init_pose + fr * kPoseDim + 3;
#pragma unroll 3
      for (int i = 0; i < kBodyDim; ++i) scr[(kScrKeep + i) * kStride] = kp[i];
    }
    FrameConsts fc;
    fc.tgt = scr + kScrTgt * kStride;
    fc.wgt = scr + kScrWgt * kStride;
    fc.keep = scr + kScrKeep * kStride;
    fc.stride = kStride;
    const bool keep_on = p.frame_preserve ? (p.frame_preserve[fr] != 0) : (p.preserve_all != 0);
    fc.keep_w2 = keep_on ? p.keep_w2 : 0.f;
    const bool stage1 = !kWorldOnly && p.loss_kind == 1;       // only global_orient and the translation are optimised
    fc.plain_sq = stage1;
    fc.depth_w2 = stage1 ? p.depth_w2 : 0.f;
    fc.dref[0] = fc.dref[1] = fc.dref[2] = 0.f;
    if (stage1) {
      fc.dref[0] = p.depth_ref[fr * 3]; fc.dref[1] = p.depth_ref[fr * 3 + 1]; fc.dref[2] = p.depth_ref[fr * 3 + 2];
    }
    const bool priors = !stage1;
    const bool freeze_betas = (p.freeze_betas & 1) != 0;    // bit 0: betas, bit 1: expression (NS == 20)
    const bool freeze_expr = NS == 20 && (p.freeze_betas & 2) != 0;

    float out_loss = 0.f;
    int evals = 0;

    int iters = p.frame_iters ? p.frame_iters[fr] : p.num_iters;
    if (!valid) iters = 0;
    float* jout = (!kWorldOnly && p.out_joints && valid) ? p.out_joints + f * K * 3 : nullptr;
    const int final_mode = kWorldOnly ? 0 : p.final_mode;
    const bool want_final = !kWorldOnly && (p.out_joints || p.final_mode);

    // Every mode funnels through ONE eval_frame call site (one copy of the evaluation code):
    // each round evaluates at the current x, then the mode-specific (cheap) update runs.
    if (OPT == kModeEval) {
      int comp = 0;
      out_loss = eval_frame<NS, K>(c, tb, fc, true, priors, jout, &comp);
      if (valid) {
        for (int i = 0; i < kPoseDim; ++i) p.out_grad_pose[f * kPoseDim + i] = c.G(i);
        for (int i = 0; i < 3; ++i) p.out_grad_transl[f * 3 + i] = c.G(kTranslOff + i);
        for (int i = 0; i < 10; ++i) p.out_grad_betas[f * 10 + i] = c.G(kShapeOff + i);
        if (NS == 20 && p.out_grad_expr)
          for (int i = 0; i < 10; ++i) p.out_grad_expr[f * 10 + i] = c.G(kShapeOff + 10 + i);
        if (p.out_gmm_component) p.out_gmm_component[f] = comp;
        p.out_loss[f] = out_loss;
      }
      continue;
    }

    if (OPT == kModeAdam) {
      float* m1 = scr + kScrOpt * kStride;
      float* m2 = m1 + (long)NX * kStride;
#pragma unroll 5
      for (int i = 0; i < NX; ++i) {
        m1[i * kStride] = 0.f;
        m2[i * kStride] = 0.f;
      }
      const int warp_iters = __reduce_max_sync(0xffffffffu, iters);
      // rounds 1..warp_iters: loss + gradient + Adam step; round warp_iters+1: joints-only
      // forward at the final parameters (world_space.py:258-278)
      const int rounds = warp_iters + (want_final ? 1 : 0);
      for (int k = 1; k <= rounds; ++k) {
        const bool last = k > warp_iters;
        if (last) fc.keep_w2 = 0.f;
        const bool stepping = !last && k <= iters;
        float step_k = 0.f, bc2_k = 1.f;
        if (stepping) {
          if (k <= kAdamTable) {
            step_k = at.step[k - 1];
            bc2_k = at.bc2[k - 1];
          } else {
            step_k = (float)((double)p.lr / (1.0 - pow(0.9, (double)k)));
            bc2_k = (float)sqrt(1.0 - pow(0.999, (double)k));
          }
        }
        // with the priors on, the body-pose entries take their Adam step inside the gradient pass
        const bool fused = stepping && priors && p.adam_fuse;
        fc.adam_m = fused ? m1 : nullptr;
        fc.adam_v = m2;
        fc.adam_step = step_k;
        fc.adam_bc2 = bc2_k;
        const float loss = eval_frame<NS, K, true>(c, tb, fc, !last, priors && (!last || final_mode), last ? jout : nullptr, nullptr);
        if (last && final_mode) out_loss = loss;
        if (stepping) {
          out_loss = loss;   // loss of the last iteration, before its step (world_space.py:250-256)
          ++evals;
#pragma unroll 4
          for (int i = 0; i < NX; ++i) {
            if (fused && i >= 3 && i < kTranslOff) continue;
            if (freeze_betas && i >= kShapeOff && i < kShapeOff + 10) continue;
            if (freeze_expr && i >= kShapeOff + 10) continue;
            if (stage1 && !(i < 3 || (i >= kTranslOff && i < kShapeOff))) continue;
            float mm = m1[i * kStride], vv = m2[i * kStride], x = c.X(i);
            adam_update(x, mm, vv, c.G(i), step_k, bc2_k);
            c.X(i) = x;
            m1[i * kStride] = mm;
            m2[i * kStride] = vv;
          }
        }
      }
    }

    if (OPT == kModeLbfgs) {
      Vecs v;
      v.base = scr + kScrOpt * kStride;
      v.stride = kStride;
      v.n = NX;
      v.hmax = p.lbfgs_hmax;
      Lbfgs<NX> st;
      st.init();
      // round 0: closure at the initial parameters; rounds while any lane is searching: one
      // closure per round; final round: loss (+ joints) at the returned parameters
      // (world_space.py:246-247).  Each closure writes its gradient into the slot the machine
      // designates, so nothing is copied afterwards.
      int stage = 0;   // 0 first closure, 1 searching, 2 final loss
      int rounds = 0;
      while (true) {
        ++rounds;
        const bool fin = stage == 2;
        const Cols ce = st.eval_cols(c, v);
        if (fin && final_mode) fc.keep_w2 = 0.f;
        const float loss = eval_frame<NS, K, K2B_LBFGS_HINTS>(ce, tb, fc, !fin, priors, fin ? jout : nullptr, nullptr);
        if (fin) {
          out_loss = loss;
          break;
        }
        if (freeze_betas)
          for (int i = 0; i < 10; ++i) ce.G(kShapeOff + i) = 0.f;
        if (freeze_expr)
          for (int i = 10; i < NS; ++i) ce.G(kShapeOff + i) = 0.f;
        if (stage1) {
#pragma unroll 3
          for (int i = 3; i < kTranslOff; ++i) ce.G(i) = 0.f;
          for (int i = kShapeOff; i < NX; ++i) ce.G(i) = 0.f;
        }
        if (stage == 0 || (!st.done && !st.need_outer)) st.advance(c, v, loss, stage == 0, iters, p.lr);
        stage = 1;
        // Lanes reach their outer-iteration boundary on different rounds.  The direction update
        // (two-loop recursion) is the expensive divergent part, so a lane that has finished its line
        // search idles (its evaluations are discarded, its slots protected) until every live lane of
        // the warp is at the boundary; then all of them run it together.  Per-frame arithmetic is
        // unchanged -- only the round in which it happens.
        {
          const bool waiting = st.need_outer && !st.done;
          const int n_wait = __popc(__ballot_sync(0xffffffffu, waiting));
          const int n_live = __popc(__ballot_sync(0xffffffffu, !st.done));
          if (n_wait > 0 && (n_wait == n_live || n_wait >= p.outer_quorum)) {
            if (waiting) st.start_outer(c, v);
          }
        }
        if (!__any_sync(0xffffffffu, !st.done)) {
          stage = 2;
#pragma unroll kVecUnroll
          for (int i = 0; i < NX; ++i) c.X(i) = v.at(i);
          // The returned parameters ARE the accepted trial point (lbfgs.py:488-493 adds t d to the iterate the way the
          // trial was formed), so the loss re-evaluated there (world_space.py:246-247) is the machine's own, bit for bit;
          // the extra forward pass is only run for its joints or for the camera stage's loss without the temporal term.
          if (!want_final) {
            out_loss = (float)st.loss;
            break;
          }
        }
      }
      evals = st.evals;
#ifdef K2B_DIAG
      if (p.debug_rounds) evals |= rounds << 16;
#endif
    }

    if (valid) {
      float* op = p.out_pose + f * kPoseDim;
#pragma unroll 4
      for (int i = 0; i < kPoseDim; ++i) op[i] = c.X(i);
      for (int i = 0; i < 3; ++i) p.out_transl[f * 3 + i] = c.X(kTranslOff + i);
      for (int i = 0; i < 10; ++i) p.out_betas[f * 10 + i] = c.X(kShapeOff + i);
      if (NS == 20 && p.out_expr)
        for (int i = 0; i < 10; ++i) p.out_expr[f * 10 + i] = c.X(kShapeOff + 10 + i);
      p.out_loss[f] = out_loss;
      if (p.out_evals) p.out_evals[f] = evals;
    }
  }
}

}  // namespace k2b
