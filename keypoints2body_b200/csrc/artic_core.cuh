// artic_core.cuh -- general articulated-model fit: any kinematic tree, observations of kinematic joints AND of
// vertex-picked joints (finger tips, face landmarks), any set of parameter blocks.
//
// Rows f2 / f4 of the scope table: what the reference does when
//   * WorldSpaceFitter gets `target_model_indices` beyond the body joints -- SMPL-H / SMPL-X hand and face blocks
//     (/root/reference/keypoints2body/core/joints/adapters.py:224-380, core/fitters/world_space.py:198-201), and
//   * MANOFitter / FLAMEFitter fit a hand / a head (core/fitters/misc_models.py:18-359, generic_keypoint_loss_3d
//     core/losses.py:96-112).
// One thread owns one frame; per-joint state lives in the thread's local memory (the models have 5 .. 55 joints and
// up to ~100 picked vertices, far beyond registers).  This kernel is about capability and parity, the body-keypoint
// kernels (fit_kernel, chain_kernel) remain the fast path for the layouts they cover.
//
// Forward, per frame (smplx lbs [smplx-from-memory], restricted to what the loss reads):
//   J_j = J0_j + JS_j . shape;  R_j = rodrigues(r_j);  Rw_j = Rw_p R_j;  t_j = Rw_p (J_j - J_p) + t_p
//   kinematic joint j: t_j + transl
//   picked vertex p:   vp = vt_p + S_p . shape + Pd_p . pose_feature,  pose_feature = (R_j - I) for j >= 1
//                      v  = sum_k w_k (Rw_jk (vp - J_jk) + t_jk) + transl
// Backward: the world-frame reverse pass of SURVEY.md Appendix C with three more seeds per joint (world-rotation
// gradient from skinned vertices, local-rotation gradient from the pose blend, rest-joint gradient), then
// rodrigues_bwd.  Loss: GMoF joint term, per-element quadratic regularisers, per-element temporal term, and
// (SMPL-family) the max-mixture pose prior + angle prior on the 69 body-pose entries (core/losses.py:24-67).
#pragma once

#include "fit_core.cuh"
#include "lbfgs_core.cuh"

namespace k2b {
namespace ar {

constexpr int kMaxJoints = 56;
constexpr int kMaxShape = 20;
constexpr int kMaxSkin = 8;
constexpr int kMaxParams = 200;

struct ArticModel {            // device pointers (host pointers in tests/host_emul)
  int nj, ns, n, npick, npf;   // joints, shape coefficients, parameters, picked vertices, pose-feature length 9 (nj - 1)
  const int* parents;          // [nj]
  const float* J0;             // [nj][3]
  const float* JS;             // [nj][3][ns]
  const int* pose_src;         // [3 nj] index into the parameter vector, -1 = fixed zero
  const int* shape_src;        // [ns]
  int transl_src;              // index of the translation's first element, -1 = none
  const float* pv_t;           // [P][3]
  const float* pv_S;           // [P][3][ns]
  const float* pv_P;           // [P][3][npf]
  const int* pv_idx;           // [P][kMaxSkin]
  const float* pv_w;           // [P][kMaxSkin] (0 = unused slot)
  const float* reg_w;          // [n] coefficient of x_i^2
  const float* keep_w;         // [n] coefficient of (x_i - keep_i)^2 (times keep_scale)
  int body_off;                // index of body_pose[0] (69 entries) for the SMPL priors, -1 = none
  const float* gmm_P;          // [8][69][72] symmetric precisions
  const float* gmm_mu;         // [8][72]
  const float* gmm_nlw;        // [8]
};

K2B_HD void m3_add_outer(float* M, V3 a, V3 b, float s) {      // M += s a b^T
  M[0] = fmaf(s * a.x, b.x, M[0]); M[1] = fmaf(s * a.x, b.y, M[1]); M[2] = fmaf(s * a.x, b.z, M[2]);
  M[3] = fmaf(s * a.y, b.x, M[3]); M[4] = fmaf(s * a.y, b.y, M[4]); M[5] = fmaf(s * a.y, b.z, M[5]);
  M[6] = fmaf(s * a.z, b.x, M[6]); M[7] = fmaf(s * a.z, b.y, M[7]); M[8] = fmaf(s * a.z, b.z, M[8]);
}
K2B_HD M3 m3_load(const float* p) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.m[i] = p[i];
  return r;
}
K2B_HD void m3_store(float* p, const M3& r) {
#pragma unroll
  for (int i = 0; i < 9; ++i) p[i] = r.m[i];
}

#if defined(__CUDACC__)
#define K2B_AR_FN __host__ __device__ inline
#else
#define K2B_AR_FN inline
#endif

// One evaluation for one frame.  x [n]; obs_idx [K] model-point indices (j < nj kinematic joint, nj + p picked vertex);
// tgt [K][3]; wgt [K] = joint_w^2 conf^2; keep [n] (temporal anchor) with keep_scale = pose_preserve_weight^2 or 0.
// with_grad fills g [n].  pts_out [K][3] (optional) receives the model points incl. the translation.
// WARP: the 32 lanes of a warp work on ONE frame (small batches: a lone frame on one thread leaves the GPU idle for
// 75 ms).  Every lane runs the whole evaluation redundantly -- identical values in every lane -- except the three loops
// that carry the work, which are dealt out by lane and combined with shuffles: the pose blend of the picked vertices
// (forward: lane-strided partial sums + butterfly; backward: each lane accumulates its own entries of dL/dR_local, one
// all-gather after the observation loop) and the scan of the mixture prior (rows by lane + butterfly).
K2B_AR_FN float ar_wsum(float v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
#endif
  return v;
}
K2B_AR_FN float ar_bcast(float v, int src) {
#if defined(__CUDA_ARCH__)
  return __shfl_sync(0xffffffffu, v, src);
#else
  return v;
#endif
}

template <bool WARP>
K2B_AR_FN float artic_eval(const ArticModel& M, const float* x, const int* obs_idx, int K, const float* tgt,
                           const float* wgt, const float* keep, float keep_scale, bool with_grad, float* g,
                           float* pts_out, int* comp_out) {
  const int nj = M.nj, ns = M.ns, n = M.n;
#if defined(__CUDA_ARCH__)
  const int lane = WARP ? (int)(threadIdx.x & 31) : 0;
#else
  const int lane = 0;
#endif
  constexpr int NL = WARP ? 32 : 1;
  float shape[kMaxShape];
  for (int s = 0; s < ns; ++s) shape[s] = M.shape_src[s] >= 0 ? x[M.shape_src[s]] : 0.f;
  float J[kMaxJoints][3], rv[kMaxJoints][3], Rl[kMaxJoints][9], Rw[kMaxJoints][9], tw[kMaxJoints][3];
  Rod rod[kMaxJoints];
  for (int j = 0; j < nj; ++j) {
    for (int c = 0; c < 3; ++c) {
      float v = M.J0[3 * j + c];
      const float* js = M.JS + (size_t)(3 * j + c) * ns;
      for (int s = 0; s < ns; ++s) v = fmaf(js[s], shape[s], v);
      J[j][c] = v;
      const int src = M.pose_src[3 * j + c];
      rv[j][c] = src >= 0 ? x[src] : 0.f;
    }
    m3_store(Rl[j], rodrigues(v3(rv[j][0], rv[j][1], rv[j][2]), rod[j]));
  }
  for (int j = 0; j < nj; ++j) {
    const int p = M.parents[j];
    if (p < 0) {
      for (int i = 0; i < 9; ++i) Rw[j][i] = Rl[j][i];
      for (int c = 0; c < 3; ++c) tw[j][c] = J[j][c];
    } else {
      const M3 Rp = m3_load(Rw[p]);
      const V3 t = matvec(Rp, v3(J[j][0] - J[p][0], J[j][1] - J[p][1], J[j][2] - J[p][2]));
      tw[j][0] = t.x + tw[p][0]; tw[j][1] = t.y + tw[p][1]; tw[j][2] = t.z + tw[p][2];
      m3_store(Rw[j], matmul(Rp, m3_load(Rl[j])));
    }
  }
  V3 transl = v3(0.f, 0.f, 0.f);
  if (M.transl_src >= 0) transl = v3(x[M.transl_src], x[M.transl_src + 1], x[M.transl_src + 2]);

  float gpos[kMaxJoints][3], gRw[kMaxJoints][9], gRl[kMaxJoints][9], gJ[kMaxJoints][3], gshape[kMaxShape];
  V3 gtr = v3(0.f, 0.f, 0.f);
  if (with_grad) {
    for (int j = 0; j < nj; ++j) {
      for (int c = 0; c < 3; ++c) gpos[j][c] = gJ[j][c] = 0.f;
      for (int i = 0; i < 9; ++i) gRw[j][i] = gRl[j][i] = 0.f;
    }
    for (int s = 0; s < ns; ++s) gshape[s] = 0.f;
  }
  float loss = 0.f;
  // ---- observed model points: residuals (gmof, losses.py:6-10) and the seeds of the reverse pass -------------------
  for (int k = 0; k < K; ++k) {
    const int idx = obs_idx[k];
    V3 p, vp = v3(0.f, 0.f, 0.f);
    const int pi = idx - nj;
    if (idx < nj) {
      p = v3(tw[idx][0], tw[idx][1], tw[idx][2]) + transl;
    } else {
      float vpa[3];
      {
        // pose blend: sum over the 9 (nj - 1) pose features, entries dealt by lane (three independent chains)
        const float* pd0 = M.pv_P + (size_t)(3 * pi) * M.npf;
        const float* pd1 = pd0 + M.npf;
        const float* pd2 = pd1 + M.npf;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        if (WARP) {
          for (int je = lane; je < M.npf; je += NL) {
            const int j = 1 + je / 9, e = je - 9 * (j - 1);
            const float feat = Rl[j][e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
            a0 = fmaf(pd0[je], feat, a0);
            a1 = fmaf(pd1[je], feat, a1);
            a2 = fmaf(pd2[je], feat, a2);
          }
          a0 = ar_wsum(a0); a1 = ar_wsum(a1); a2 = ar_wsum(a2);
        } else {
          for (int j = 1; j < nj; ++j) {
            const float* r = Rl[j];
            const int o = 9 * (j - 1);
#pragma unroll
            for (int e = 0; e < 9; ++e) {
              const float feat = r[e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
              a0 = fmaf(pd0[o + e], feat, a0);
              a1 = fmaf(pd1[o + e], feat, a1);
              a2 = fmaf(pd2[o + e], feat, a2);
            }
          }
        }
        const float ap[3] = {a0, a1, a2};
        for (int c = 0; c < 3; ++c) {
          float v = M.pv_t[3 * pi + c];
          const float* sd = M.pv_S + (size_t)(3 * pi + c) * ns;
          for (int s = 0; s < ns; ++s) v = fmaf(sd[s], shape[s], v);
          vpa[c] = v + ap[c];
        }
      }
      vp = v3(vpa[0], vpa[1], vpa[2]);
      p = transl;
      for (int s = 0; s < kMaxSkin; ++s) {
        const float w = M.pv_w[kMaxSkin * pi + s];
        if (w == 0.f) continue;
        const int js = M.pv_idx[kMaxSkin * pi + s];
        const V3 loc = matvec(m3_load(Rw[js]), v3(vp.x - J[js][0], vp.y - J[js][1], vp.z - J[js][2]));
        p = v3(fmaf(w, loc.x + tw[js][0], p.x), fmaf(w, loc.y + tw[js][1], p.y), fmaf(w, loc.z + tw[js][2], p.z));
      }
    }
    if (pts_out) { pts_out[3 * k] = p.x; pts_out[3 * k + 1] = p.y; pts_out[3 * k + 2] = p.z; }
    const float w = wgt[k];
    const float ex = p.x - tgt[3 * k], ey = p.y - tgt[3 * k + 1], ez = p.z - tgt[3 * k + 2];
    const float ix = fdiv(1.f, kSigma2 + ex * ex), iy = fdiv(1.f, kSigma2 + ey * ey), iz = fdiv(1.f, kSigma2 + ez * ez);
    loss = fmaf(w, (kSigma2 * ex * ex * ix + kSigma2 * ey * ey * iy) + kSigma2 * ez * ez * iz, loss);
    if (!with_grad) continue;
    const float c2 = 2.f * kSigma2 * kSigma2 * w;
    const V3 gp = v3(c2 * ex * ix * ix, c2 * ey * iy * iy, c2 * ez * iz * iz);
    gtr = gtr + gp;
    if (idx < nj) {
      gpos[idx][0] += gp.x; gpos[idx][1] += gp.y; gpos[idx][2] += gp.z;
      continue;
    }
    V3 gvp = v3(0.f, 0.f, 0.f);
    for (int s = 0; s < kMaxSkin; ++s) {
      const float ws = M.pv_w[kMaxSkin * pi + s];
      if (ws == 0.f) continue;
      const int js = M.pv_idx[kMaxSkin * pi + s];
      gpos[js][0] = fmaf(ws, gp.x, gpos[js][0]); gpos[js][1] = fmaf(ws, gp.y, gpos[js][1]); gpos[js][2] = fmaf(ws, gp.z, gpos[js][2]);
      m3_add_outer(gRw[js], gp, v3(vp.x - J[js][0], vp.y - J[js][1], vp.z - J[js][2]), ws);
      const V3 back = matvec_t(m3_load(Rw[js]), gp);         // Rw^T gp
      gJ[js][0] = fmaf(-ws, back.x, gJ[js][0]); gJ[js][1] = fmaf(-ws, back.y, gJ[js][1]); gJ[js][2] = fmaf(-ws, back.z, gJ[js][2]);
      gvp = v3(fmaf(ws, back.x, gvp.x), fmaf(ws, back.y, gvp.y), fmaf(ws, back.z, gvp.z));
    }
    const float gv[3] = {gvp.x, gvp.y, gvp.z};
    for (int c = 0; c < 3; ++c) {
      const float* sd = M.pv_S + (size_t)(3 * pi + c) * ns;
      for (int s = 0; s < ns; ++s) gshape[s] = fmaf(sd[s], gv[c], gshape[s]);
    }
    {
      // WARP: a lane accumulates only its own entries of dL/dR_local here; they are gathered after the loop
      const float* pd0 = M.pv_P + (size_t)(3 * pi) * M.npf;
      const float* pd1 = pd0 + M.npf;
      const float* pd2 = pd1 + M.npf;
      if (WARP) {
        for (int je = lane; je < M.npf; je += NL) {
          const int j = 1 + je / 9, e = je - 9 * (j - 1);
          gRl[j][e] = fmaf(pd0[je], gv[0], fmaf(pd1[je], gv[1], fmaf(pd2[je], gv[2], gRl[j][e])));
        }
      } else {
        for (int j = 1; j < nj; ++j) {
          const int o = 9 * (j - 1);
#pragma unroll
          for (int e = 0; e < 9; ++e)
            gRl[j][e] = fmaf(pd0[o + e], gv[0], fmaf(pd1[o + e], gv[1], fmaf(pd2[o + e], gv[2], gRl[j][e])));
        }
      }
    }
  }
  if (WARP && with_grad && M.npick > 0) {      // all-gather of the lane-owned entries (entry je lives in lane je % 32)
    for (int je = 0; je < M.npf; ++je) {
      const int j = 1 + je / 9, e = je - 9 * (j - 1);
      gRl[j][e] = ar_bcast(gRl[j][e], je & 31);
    }
  }
  if (with_grad) {
    for (int i = 0; i < n; ++i) g[i] = 0.f;
    // ---- reverse pass over the tree, children before parents ---------------------------------
    for (int j = nj - 1; j >= 1; --j) {
      const int p = M.parents[j];
      const V3 gj = v3(gpos[j][0], gpos[j][1], gpos[j][2]);
      const V3 rel = v3(J[j][0] - J[p][0], J[j][1] - J[p][1], J[j][2] - J[p][2]);
      gpos[p][0] += gj.x; gpos[p][1] += gj.y; gpos[p][2] += gj.z;
      m3_add_outer(gRw[p], gj, rel, 1.f);
      const M3 Rp = m3_load(Rw[p]);
      const V3 relb = matvec_t(Rp, gj);
      gJ[j][0] += relb.x; gJ[j][1] += relb.y; gJ[j][2] += relb.z;
      gJ[p][0] -= relb.x; gJ[p][1] -= relb.y; gJ[p][2] -= relb.z;
      const M3 Gw = m3_load(gRw[j]);
      const M3 a = matmul_tn(Rp, Gw);                            // Rw_p^T  dL/dRw_j
      for (int i = 0; i < 9; ++i) gRl[j][i] += a.m[i];
      const M3 b = matmul_nt(Gw, m3_load(Rl[j]));                // dL/dRw_j  R_j^T
      for (int i = 0; i < 9; ++i) gRw[p][i] += b.m[i];
    }
    for (int i = 0; i < 9; ++i) gRl[0][i] += gRw[0][i];
    gJ[0][0] += gpos[0][0]; gJ[0][1] += gpos[0][1]; gJ[0][2] += gpos[0][2];
    for (int j = 0; j < nj; ++j) {
      for (int c = 0; c < 3; ++c) {
        const float* js = M.JS + (size_t)(3 * j + c) * ns;
        for (int s = 0; s < ns; ++s) gshape[s] = fmaf(js[s], gJ[j][c], gshape[s]);
      }
      const V3 rb = rodrigues_bwd(m3_load(gRl[j]), v3(rv[j][0], rv[j][1], rv[j][2]), rod[j]);
      const float rbv[3] = {rb.x, rb.y, rb.z};
      for (int c = 0; c < 3; ++c) {
        const int src = M.pose_src[3 * j + c];
        if (src >= 0) g[src] += rbv[c];
      }
    }
    for (int s = 0; s < ns; ++s)
      if (M.shape_src[s] >= 0) g[M.shape_src[s]] += gshape[s];
    if (M.transl_src >= 0) { g[M.transl_src] += gtr.x; g[M.transl_src + 1] += gtr.y; g[M.transl_src + 2] += gtr.z; }
  }
  // ---- quadratic regularisers and the temporal term ------------------------------------------------------------------
  for (int i = 0; i < n; ++i) {
    const float xi = x[i], rw = M.reg_w[i];
    loss = fmaf(rw * xi, xi, loss);
    float gi = 2.f * rw * xi;
    if (keep_scale != 0.f) {
      const float kw = keep_scale * M.keep_w[i], d = xi - keep[i];
      loss = fmaf(kw * d, d, loss);
      gi = fmaf(2.f * kw, d, gi);
    }
    if (with_grad) g[i] += gi;
  }
  // ---- SMPL-family priors on the 69 body-pose entries (prior.py:182-195, losses.py:13-21) --------------------------------
  if (M.body_off >= 0) {
    const float* xb = x + M.body_off;
    float best = INFINITY;
    int bm = 0;
    for (int m = 0; m < kGmmM; ++m) {
      const float* P = M.gmm_P + (size_t)m * kBodyDim * 72;
      const float* mu = M.gmm_mu + m * kMuStride;
      float q = 0.f;
      for (int i = lane; i < kBodyDim; i += NL) {
        float y = 0.f;
        for (int j = 0; j < kBodyDim; ++j) y = fmaf(P[i * 72 + j], xb[j] - mu[j], y);
        q = fmaf(y, xb[i] - mu[i], q);
      }
      if (WARP) q = ar_wsum(q);
      const float ll = fmaf(0.5f, q, M.gmm_nlw[m]);
      if (ll < best) { best = ll; bm = m; }
    }
    if (comp_out) *comp_out = bm;
    loss = fmaf(kPosePriorW2, best, loss);
    if (with_grad) {
      const float* P = M.gmm_P + (size_t)bm * kBodyDim * 72;
      const float* mu = M.gmm_mu + bm * kMuStride;
      for (int i = 0; i < kBodyDim; ++i) {
        float y = 0.f;
        for (int j = 0; j < kBodyDim; ++j) y = fmaf(P[i * 72 + j], xb[j] - mu[j], y);
        g[M.body_off + i] = fmaf(kPosePriorW2, y, g[M.body_off + i]);
      }
    }
    for (int q = 0; q < 4; ++q) {
      const int i = q == 0 ? 52 : (q == 1 ? 55 : (q == 2 ? 9 : 12));
      const float sgn = q == 0 ? 1.f : -1.f;
      const float e = expf(xb[i] * sgn);
      loss = fmaf(kAnglePriorW2, e * e, loss);
      if (with_grad) g[M.body_off + i] = fmaf(2.f * kAnglePriorW2 * sgn, e * e, g[M.body_off + i]);
    }
  }
  return loss;
}

}  // namespace ar
}  // namespace k2b
