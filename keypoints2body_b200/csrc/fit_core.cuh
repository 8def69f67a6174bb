// fit_core.cuh -- per-frame evaluation (forward + analytic backward + priors) and
// optimiser updates of the fused fitting kernel.
//
// One CUDA thread owns one frame.  Everything here is written against a small
// "column accessor" so that the same code runs (a) inside the kernel with the
// per-frame state in shared memory laid out [element][thread], and (b) on the
// host inside tests/host_emul (a debugging harness, never shipped in the
// library) with plain arrays.
//
// Maths restated from (paths into /root/reference/keypoints2body):
//   core/fitters/world_space.py:173-212  compute_loss (which joints, which terms)
//   core/losses.py:6-67                  gmof, angle_prior, body_fitting_loss_3d
//   core/prior.py:182-195                MaxMixturePrior.merged_log_likelihood
//   smplx lbs / batch_rodrigues / batch_rigid_transform [smplx-from-memory], joint branch only
// The backward is SURVEY.md Appendix C in world-frame form: for joint j with
// parent p, s_j = sum of residual gradients in the subtree, S_j = sum g_k p_k^T,
// M_j = S_j - s_j t_j^T, dL/dR_j = Rw_p^T M_j Rw_j, dL/d(rel_j) = Rw_p^T s_j.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define K2B_HD __host__ __device__ __forceinline__
#else
#define K2B_HD inline
#endif

namespace k2b {

constexpr int kPoseDim = 72;
constexpr int kBodyDim = 69;
constexpr int kGmmM = 8;
constexpr int kMaxFitJoints = 24;
constexpr int kTranslOff = 72;   // x layout: [go 3 | body 69 | transl 3 | shape NS]
constexpr int kShapeOff = 75;    // shape = [betas 10 | expression 10 (NS == 20)]

// Packed lower-triangular Cholesky factor: row j holds L[j][0..j], padded to a
// multiple of 4 floats so every row starts 16-byte aligned.
K2B_HD constexpr int chol_row_off(int j) {
  return 8 * (j / 4) * (j / 4 + 1) + 4 * (j % 4) * (j / 4 + 1);
}
constexpr int kCholStride = 2520;  // chol_row_off(68) + 72
constexpr int kMuStride = 72;

constexpr float kSigma2 = 100.f * 100.f;               // gmof sigma^2 (losses.py:33)
constexpr float kPosePriorW2 = (4.78f * 1.5f) * (4.78f * 1.5f);
constexpr float kAnglePriorW2 = 15.2f * 15.2f;
constexpr float kShapePriorW2 = 25.f;                   // frame loss always uses 5.0 (losses.py:35)

struct V3 {
  float x, y, z;
};
struct M3 {
  float m[9];  // row-major
};
struct Acc {   // subtree sums of the world-frame backward
  M3 S;
  V3 s;
};

K2B_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
K2B_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
K2B_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
K2B_HD float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }

K2B_HD M3 eye3() { return M3{{1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}}; }
K2B_HD M3 zero3() { return M3{{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}}; }

K2B_HD M3 matmul(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C.m[3 * i + j] = fmaf(A.m[3 * i], B.m[j], fmaf(A.m[3 * i + 1], B.m[3 + j], A.m[3 * i + 2] * B.m[6 + j]));
  return C;
}
// A^T B
K2B_HD M3 matmul_tn(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C.m[3 * i + j] = fmaf(A.m[i], B.m[j], fmaf(A.m[3 + i], B.m[3 + j], A.m[6 + i] * B.m[6 + j]));
  return C;
}
K2B_HD V3 matvec(const M3& A, V3 v) {
  return V3{fmaf(A.m[0], v.x, fmaf(A.m[1], v.y, A.m[2] * v.z)),
            fmaf(A.m[3], v.x, fmaf(A.m[4], v.y, A.m[5] * v.z)),
            fmaf(A.m[6], v.x, fmaf(A.m[7], v.y, A.m[8] * v.z))};
}
K2B_HD V3 matvec_t(const M3& A, V3 v) {
  return V3{fmaf(A.m[0], v.x, fmaf(A.m[3], v.y, A.m[6] * v.z)),
            fmaf(A.m[1], v.x, fmaf(A.m[4], v.y, A.m[7] * v.z)),
            fmaf(A.m[2], v.x, fmaf(A.m[5], v.y, A.m[8] * v.z))};
}
K2B_HD void acc_add(Acc& a, const Acc& b) {
#pragma unroll
  for (int i = 0; i < 9; ++i) a.S.m[i] += b.S.m[i];
  a.s = a.s + b.s;
}
// a.S += g p^T ; a.s += g
K2B_HD void acc_point(Acc& a, V3 g, V3 p) {
  a.S.m[0] = fmaf(g.x, p.x, a.S.m[0]); a.S.m[1] = fmaf(g.x, p.y, a.S.m[1]); a.S.m[2] = fmaf(g.x, p.z, a.S.m[2]);
  a.S.m[3] = fmaf(g.y, p.x, a.S.m[3]); a.S.m[4] = fmaf(g.y, p.y, a.S.m[4]); a.S.m[5] = fmaf(g.y, p.z, a.S.m[5]);
  a.S.m[6] = fmaf(g.z, p.x, a.S.m[6]); a.S.m[7] = fmaf(g.z, p.y, a.S.m[7]); a.S.m[8] = fmaf(g.z, p.z, a.S.m[8]);
  a.s = a.s + g;
}

// Division without the IEEE slow path on the device (<= 2 ulp); plain division on the host.
K2B_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdividef(a, b);
#else
  return a / b;
#endif
}

// ---- Rodrigues, smplx flavour: angle = ||r + 1e-8||, axis = r / angle ----------
struct Rod {
  float inv, s, c;   // 1/theta, sin, cos  (axis k = r * inv is recomputed where needed)
};

// sin/cos without libdevice's inlined Payne-Hanek slow path: 3-term Cody-Waite reduction by
// pi/2 and minimax polynomials on [-pi/4, pi/4] (|error| ~ 1 ulp for |x| < ~1e4, far beyond any
// axis-angle magnitude).  Identical code on host and device, so tests/host_emul reproduces
// the device numerics.
K2B_HD void k2b_sincos(float x, float* sn, float* cs) {
  const float q = rintf(x * 0.636619772f);
  float r = fmaf(q, -1.57079601e+00f, x);
  r = fmaf(q, -3.13916473e-07f, r);
  r = fmaf(q, -5.39030253e-15f, r);
  const float r2 = r * r;
  float ps = fmaf(-1.95152959e-4f, r2, 8.33216087e-3f);
  ps = fmaf(ps, r2, -1.66666546e-1f);
  const float s = fmaf(ps * r2, r, r);
  float pc = fmaf(2.44331571e-5f, r2, -1.38873163e-3f);
  pc = fmaf(pc, r2, 4.16666457e-2f);
  pc = fmaf(pc, r2, -0.5f);
  const float c = fmaf(pc, r2, 1.f);
  const int n = (int)q;
  const float s1 = (n & 1) ? c : s;
  const float c1 = (n & 1) ? s : c;
  *sn = (n & 2) ? -s1 : s1;
  *cs = ((n + 1) & 2) ? -c1 : c1;
}

K2B_HD M3 rodrigues(V3 r, Rod& o) {
  const V3 a = v3(r.x + 1e-8f, r.y + 1e-8f, r.z + 1e-8f);
  const float theta = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
  o.inv = fdiv(1.f, theta);
  k2b_sincos(theta, &o.s, &o.c);
  const float oc = 1.f - o.c;
  const V3 k = v3(r.x * o.inv, r.y * o.inv, r.z * o.inv);
  const float xx = k.x * k.x, yy = k.y * k.y, zz = k.z * k.z;
  const float xy = k.x * k.y, xz = k.x * k.z, yz = k.y * k.z;
  M3 R;
  R.m[0] = 1.f - oc * (yy + zz); R.m[1] = fmaf(oc, xy, -o.s * k.z); R.m[2] = fmaf(oc, xz, o.s * k.y);
  R.m[3] = fmaf(oc, xy, o.s * k.z); R.m[4] = 1.f - oc * (xx + zz); R.m[5] = fmaf(oc, yz, -o.s * k.x);
  R.m[6] = fmaf(oc, xz, -o.s * k.y); R.m[7] = fmaf(oc, yz, o.s * k.x); R.m[8] = 1.f - oc * (xx + yy);
  return R;
}

// dL/dr from G = dL/dR (see header comment / SURVEY.md Appendix C)
K2B_HD V3 rodrigues_bwd(const M3& G, V3 r, const Rod& o) {
  const V3 k = v3(r.x * o.inv, r.y * o.inv, r.z * o.inv);
  const V3 w = v3(G.m[7] - G.m[5], G.m[2] - G.m[6], G.m[3] - G.m[1]);
  const float trG = G.m[0] + G.m[4] + G.m[8];
  // (G + G^T) k
  const V3 gk = v3(2.f * G.m[0] * k.x + (G.m[1] + G.m[3]) * k.y + (G.m[2] + G.m[6]) * k.z,
                   (G.m[1] + G.m[3]) * k.x + 2.f * G.m[4] * k.y + (G.m[5] + G.m[7]) * k.z,
                   (G.m[2] + G.m[6]) * k.x + (G.m[5] + G.m[7]) * k.y + 2.f * G.m[8] * k.z);
  const float kk = dot(k, k);
  const float A = dot(k, w);                         // <G, K>
  const float B = 0.5f * dot(k, gk) - trG * kk;      // <G, K^2> = k^T G k - tr(G) k.k
  const float oc = 1.f - o.c;
  const V3 kb = v3(fmaf(o.s, w.x, oc * (gk.x - 2.f * trG * k.x)),
                   fmaf(o.s, w.y, oc * (gk.y - 2.f * trG * k.y)),
                   fmaf(o.s, w.z, oc * (gk.z - 2.f * trG * k.z)));
  const float inv = o.inv;
  const float tb = (o.c * A + o.s * B - dot(kb, r) * inv * inv) * inv;  // theta_bar / theta
  const V3 a = v3(r.x + 1e-8f, r.y + 1e-8f, r.z + 1e-8f);
  return v3(fmaf(kb.x, inv, tb * a.x), fmaf(kb.y, inv, tb * a.y), fmaf(kb.z, inv, tb * a.z));
}

// ---- tables shared by all frames of a CTA (shared memory on the device) ---------
struct FitTables {
  const float* chol;   // [kGmmM][kCholStride]
  const float* mu;     // [kGmmM][kMuStride]
  const float* nlw;    // [kGmmM]
  const float4* rel;   // [kMaxFitJoints][1 + NS]: entry 0 = rest offset J0_j - J0_p, then d/d shape_s
};

// Per-frame observations (targets, weights, preserve pose) -- device: transposed global
// scratch (coalesced across the warp); host emulation: plain arrays with stride 1.
struct FrameConsts {
  const float* tgt;      // element (j*3+c) at tgt[(j*3+c)*stride]
  const float* wgt;      // joint weight joint_w^2 * conf_j^2 at wgt[j*stride]
  const float* keep;     // preserve pose element i at keep[i*stride]
  long stride;
  float keep_w2;         // pose_preserve_weight^2 or 0
};

// Column accessor: element i of this frame's x / g vector.
struct Cols {
  float* x;
  float* g;
  int stride;
  K2B_HD float& X(int i) const { return x[i * stride]; }
  K2B_HD float& G(int i) const { return g[i * stride]; }
};

// ---------------------------------------------------------------------------------
// GMM max-mixture prior on body pose (prior.py:182-195) with P_m = L_m L_m^T:
//   q_m = ||L_m^T d||^2, d = x - mu_m;  m* = argmin(0.5 q_m + nlw_m)
//   gradient = L_m* (L_m*^T d)
// The best candidate's z = L^T d is parked in the (not yet used) gradient column
// G(3..71) so only one 72-register accumulator set is live.  With with_grad the
// final G(3..71) = kPosePriorW2 * gradient.  Returns kPosePriorW2 * min.
// ---------------------------------------------------------------------------------
K2B_HD float gmm_prior(const Cols& c, const FitTables& tb, bool with_grad, int* best_m_out) {
  float best = INFINITY;
  int best_m = 0;
#pragma unroll 1
  for (int m = 0; m < kGmmM; ++m) {
    const float* __restrict__ Lm = tb.chol + m * kCholStride;
    const float* __restrict__ mum = tb.mu + m * kMuStride;
    float z[72];
#pragma unroll
    for (int i = 0; i < 72; ++i) z[i] = 0.f;
#pragma unroll
    for (int j = 0; j < kBodyDim; ++j) {
      const float dj = c.X(3 + j) - mum[j];
      const float4* row = reinterpret_cast<const float4*>(Lm + chol_row_off(j));
#pragma unroll
      for (int ib = 0; ib <= j / 4; ++ib) {
        const float4 l = row[ib];
        if (4 * ib + 0 <= j) z[4 * ib + 0] = fmaf(l.x, dj, z[4 * ib + 0]);
        if (4 * ib + 1 <= j) z[4 * ib + 1] = fmaf(l.y, dj, z[4 * ib + 1]);
        if (4 * ib + 2 <= j) z[4 * ib + 2] = fmaf(l.z, dj, z[4 * ib + 2]);
        if (4 * ib + 3 <= j) z[4 * ib + 3] = fmaf(l.w, dj, z[4 * ib + 3]);
      }
    }
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int i = 0; i < 68; i += 4) {
      q0 = fmaf(z[i], z[i], q0); q1 = fmaf(z[i + 1], z[i + 1], q1);
      q2 = fmaf(z[i + 2], z[i + 2], q2); q3 = fmaf(z[i + 3], z[i + 3], q3);
    }
    q0 = fmaf(z[68], z[68], q0);
    const float ll = fmaf(0.5f, (q0 + q1) + (q2 + q3), tb.nlw[m]);
    if (ll < best) {   // strict: first minimum wins, like torch.min
      best = ll;
      best_m = m;
      if (with_grad) {
#pragma unroll
        for (int i = 0; i < kBodyDim; ++i) c.G(3 + i) = z[i];
      }
    }
  }
  if (best_m_out) *best_m_out = best_m;
  if (with_grad) {
    float zb[72];
#pragma unroll
    for (int i = 0; i < kBodyDim; ++i) zb[i] = c.G(3 + i);
    const float* __restrict__ Lb = tb.chol + best_m * kCholStride;
#pragma unroll
    for (int i = 0; i < kBodyDim; ++i) {
      const float4* row = reinterpret_cast<const float4*>(Lb + chol_row_off(i));
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int ib = 0; ib <= i / 4; ++ib) {
        const float4 l = row[ib];
        if (4 * ib + 0 <= i) a0 = fmaf(l.x, zb[4 * ib + 0], a0);
        if (4 * ib + 1 <= i) a1 = fmaf(l.y, zb[4 * ib + 1], a1);
        if (4 * ib + 2 <= i) a2 = fmaf(l.z, zb[4 * ib + 2], a2);
        if (4 * ib + 3 <= i) a3 = fmaf(l.w, zb[4 * ib + 3], a3);
      }
      c.G(3 + i) = kPosePriorW2 * ((a0 + a1) + (a2 + a3));
    }
  }
  return kPosePriorW2 * best;
}

// ---------------------------------------------------------------------------------
// Kinematic chain: forward, residuals, world-frame backward.
// ---------------------------------------------------------------------------------
template <int NS>
K2B_HD V3 rel_offset(const FitTables& tb, int j, const float (&shape)[NS]) {
  const float4* e = tb.rel + j * (1 + NS);
  const float4 r = e[0];
  float x = r.x, y = r.y, z = r.z;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const float4 d = e[1 + s];
    x = fmaf(d.x, shape[s], x);
    y = fmaf(d.y, shape[s], y);
    z = fmaf(d.z, shape[s], z);
  }
  return v3(x, y, z);
}

// d loss / d shape accumulates straight into the gradient column (keeps registers free)
template <int NS>
K2B_HD void rel_offset_bwd(const Cols& c, const FitTables& tb, int j, V3 rb) {
  const float4* e = tb.rel + j * (1 + NS);
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const float4 d = e[1 + s];
    c.G(kShapeOff + s) = fmaf(d.x, rb.x, fmaf(d.y, rb.y, fmaf(d.z, rb.z, c.G(kShapeOff + s))));
  }
}

struct KinAcc {
  V3 transl;
  float loss;
};

// Residual of joint j at position t (transl not yet added): accumulates the loss, returns
// the gradient w.r.t. the joint position (zero when !with_grad).
K2B_HD V3 residual(const FrameConsts& fc, int j, V3 t, KinAcc& ks, bool with_grad, float* joints_out) {
  const V3 p = t + ks.transl;
  if (joints_out) {
    joints_out[3 * j + 0] = p.x;
    joints_out[3 * j + 1] = p.y;
    joints_out[3 * j + 2] = p.z;
  }
  const float w = fc.wgt[j * fc.stride];
  const float ex = p.x - fc.tgt[(3 * j + 0) * fc.stride];
  const float ey = p.y - fc.tgt[(3 * j + 1) * fc.stride];
  const float ez = p.z - fc.tgt[(3 * j + 2) * fc.stride];
  const float ix = fdiv(1.f, kSigma2 + ex * ex), iy = fdiv(1.f, kSigma2 + ey * ey), iz = fdiv(1.f, kSigma2 + ez * ez);
  const float gx = kSigma2 * ex * ex * ix, gy = kSigma2 * ey * ey * iy, gz = kSigma2 * ez * ez * iz;
  ks.loss = fmaf(w, (gx + gy) + gz, ks.loss);
  if (!with_grad) return v3(0.f, 0.f, 0.f);
  const float c2 = 2.f * kSigma2 * kSigma2 * w;
  return v3(c2 * ex * ix * ix, c2 * ey * iy * iy, c2 * ez * iz * iz);
}

// M = S - s t^T ; returns Rp^T M Rw
K2B_HD M3 rot_grad(const Acc& a, const M3& Rp, const M3& Rw, V3 t) {
  M3 M;
  M.m[0] = fmaf(-a.s.x, t.x, a.S.m[0]); M.m[1] = fmaf(-a.s.x, t.y, a.S.m[1]); M.m[2] = fmaf(-a.s.x, t.z, a.S.m[2]);
  M.m[3] = fmaf(-a.s.y, t.x, a.S.m[3]); M.m[4] = fmaf(-a.s.y, t.y, a.S.m[4]); M.m[5] = fmaf(-a.s.y, t.z, a.S.m[5]);
  M.m[6] = fmaf(-a.s.z, t.x, a.S.m[6]); M.m[7] = fmaf(-a.s.z, t.y, a.S.m[7]); M.m[8] = fmaf(-a.s.z, t.z, a.S.m[8]);
  return matmul(matmul_tn(Rp, M), Rw);
}

K2B_HD V3 load_rot(const Cols& c, int j) { return v3(c.X(3 * j), c.X(3 * j + 1), c.X(3 * j + 2)); }
K2B_HD void add_rot_grad(const Cols& c, int j, V3 rb) {
  c.G(3 * j + 0) += rb.x;
  c.G(3 * j + 1) += rb.y;
  c.G(3 * j + 2) += rb.z;
}

// ---------------------------------------------------------------------------------
// One function evaluation for one frame (K observed joints: 22 AMASS / 24 SMPL24).
// with_grad fills G(0 .. 75+NS) completely; with_priors = false skips every prior term
// (joints-only final forward).  Returns the total loss (losses.py:41-67).
// Tree walk: root -> spine (3,6,9) forward -> neck (12,15) -> four limbs in one rolled
// loop (legs hang from the root, arms from joint 9) -> spine backward -> root backward.
// ---------------------------------------------------------------------------------
template <int NS, int K>
K2B_HD float eval_frame(const Cols& c, const FitTables& tb, const FrameConsts& fc, bool with_grad,
                        bool with_priors, float* joints_out, int* gmm_component) {
  constexpr int LEN = (K == 24) ? 5 : 4;   // joints per limb chain (arms gain the hand joint)
  float loss = 0.f;
  float shape[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) shape[s] = c.X(kShapeOff + s);

  if (with_priors) {
    loss = gmm_prior(c, tb, with_grad, gmm_component);
    // angle prior (losses.py:13-21) on body-pose entries 52, 55, 9, 12 with signs +,-,-,-
    const int idx[4] = {52, 55, 9, 12};
    const float sgn[4] = {1.f, -1.f, -1.f, -1.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float e = expf(c.X(3 + idx[q]) * sgn[q]);
      loss = fmaf(kAnglePriorW2, e * e, loss);
      if (with_grad) c.G(3 + idx[q]) += 2.f * kAnglePriorW2 * sgn[q] * e * e;
    }
    // temporal pose-preserve term (losses.py:57-59), active when seq_ind > 0
    if (fc.keep_w2 != 0.f) {
      float acc = 0.f;
#pragma unroll 3
      for (int i = 0; i < kBodyDim; ++i) {
        const float d = c.X(3 + i) - fc.keep[i * fc.stride];
        acc = fmaf(d, d, acc);
        if (with_grad) c.G(3 + i) = fmaf(2.f * fc.keep_w2, d, c.G(3 + i));
      }
      loss = fmaf(fc.keep_w2, acc, loss);
    }
    // shape prior on betas only (losses.py:56)
    float acc = 0.f;
#pragma unroll
    for (int s = 0; s < 10; ++s) acc = fmaf(shape[s], shape[s], acc);
    loss = fmaf(kShapePriorW2, acc, loss);
  }
  if (with_grad) {
#pragma unroll
    for (int s = 0; s < NS; ++s) c.G(kShapeOff + s) = (with_priors && s < 10) ? 2.f * kShapePriorW2 * shape[s] : 0.f;
  }

  KinAcc ks;
  ks.transl = v3(c.X(kTranslOff), c.X(kTranslOff + 1), c.X(kTranslOff + 2));
  ks.loss = 0.f;

  // ---- root ---------------------------------------------------------------------------
  Rod o0;
  const M3 R0 = rodrigues(load_rot(c, 0), o0);
  const V3 t0 = rel_offset<NS>(tb, 0, shape);
  const V3 g0 = residual(fc, 0, t0, ks, with_grad, joints_out);
  Acc a0{zero3(), v3(0.f, 0.f, 0.f)};

  // ---- spine forward: 3 -> 6 -> 9 ---------------------------------------------------------
  M3 Rs[3];
  V3 ts[3], gs[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int j = 3 + 3 * k;
    const M3& Rp = k == 0 ? R0 : Rs[k - 1];
    const V3 tp = k == 0 ? t0 : ts[k - 1];
    ts[k] = matvec(Rp, rel_offset<NS>(tb, j, shape)) + tp;
    Rod o;
    Rs[k] = matmul(Rp, rodrigues(load_rot(c, j), o));
    gs[k] = residual(fc, j, ts[k], ks, with_grad, joints_out);
  }
  Acc a9{zero3(), v3(0.f, 0.f, 0.f)};

  // ---- neck 12 -> head 15 (leaf) ----------------------------------------------------------
  {
    const V3 t12 = matvec(Rs[2], rel_offset<NS>(tb, 12, shape)) + ts[2];
    Rod o12;
    const V3 r12 = load_rot(c, 12);
    const M3 R12 = matmul(Rs[2], rodrigues(r12, o12));
    const V3 g12 = residual(fc, 12, t12, ks, with_grad, joints_out);
    const V3 t15 = matvec(R12, rel_offset<NS>(tb, 15, shape)) + t12;
    const V3 g15 = residual(fc, 15, t15, ks, with_grad, joints_out);
    if (with_grad) {
      Acc a{zero3(), v3(0.f, 0.f, 0.f)};
      acc_point(a, g15, t15);
      rel_offset_bwd<NS>(c, tb, 15, matvec_t(R12, a.s));
      acc_point(a, g12, t12);
      rel_offset_bwd<NS>(c, tb, 12, matvec_t(Rs[2], a.s));
      add_rot_grad(c, 12, rodrigues_bwd(rot_grad(a, Rs[2], R12, t12), r12, o12));
      acc_add(a9, a);
    }
  }

  // ---- limbs: legs {1,4,7,10}+side from the root, arms {13,16,18,20,(22)}+side from joint 9 ---
#pragma unroll 1
  for (int limb = 0; limb < 4; ++limb) {
    const bool arm = limb >= 2;
    const int side = limb & 1;
    int id[LEN];
    id[0] = (arm ? 13 : 1) + side;
    id[1] = id[0] + 3;
    id[2] = id[0] + (arm ? 5 : 6);
    id[3] = id[0] + (arm ? 7 : 9);
    if (LEN == 5) id[4] = arm ? id[0] + 9 : id[3];   // legs have no 5th joint: masked below
    M3 Rp0;
    V3 tp0;
#pragma unroll
    for (int i = 0; i < 9; ++i) Rp0.m[i] = arm ? Rs[2].m[i] : R0.m[i];
    tp0 = arm ? ts[2] : t0;

    M3 Rw[LEN - 1];
    V3 t[LEN], gb[LEN];
    Rod rod[LEN - 1];
#pragma unroll
    for (int k = 0; k < LEN; ++k) {
      const bool live = (LEN == 4) || k < 4 || arm;   // 5th joint exists on arms only
      const M3& Rp = k == 0 ? Rp0 : Rw[k - 1];
      const V3 tp = k == 0 ? tp0 : t[k - 1];
      t[k] = matvec(Rp, rel_offset<NS>(tb, id[k], shape)) + tp;
      if (k < LEN - 1) Rw[k] = matmul(Rp, rodrigues(load_rot(c, id[k]), rod[k]));
      gb[k] = live ? residual(fc, id[k], t[k], ks, with_grad, joints_out) : v3(0.f, 0.f, 0.f);
    }
    if (with_grad) {
      Acc a{zero3(), v3(0.f, 0.f, 0.f)};
#pragma unroll
      for (int k = LEN - 1; k >= 0; --k) {
        const bool live = (LEN == 4) || k < 4 || arm;
        const M3& Rp = k == 0 ? Rp0 : Rw[k - 1];
        acc_point(a, gb[k], t[k]);
        if (live) rel_offset_bwd<NS>(c, tb, id[k], matvec_t(Rp, a.s));
        if (k < LEN - 1 && (LEN == 4 || k < 3 || arm))
          add_rot_grad(c, id[k], rodrigues_bwd(rot_grad(a, Rp, Rw[k], t[k]), load_rot(c, id[k]), rod[k]));
      }
      if (arm) acc_add(a9, a); else acc_add(a0, a);
    }
  }
  loss += ks.loss;

  if (with_grad) {
    // ---- spine backward 9 -> 6 -> 3 (Rodrigues terms recomputed: cheaper than holding them) ----
    Acc a = a9;
#pragma unroll
    for (int k = 2; k >= 0; --k) {
      const int j = 3 + 3 * k;
      const M3& Rp = k == 0 ? R0 : Rs[k - 1];
      acc_point(a, gs[k], ts[k]);
      rel_offset_bwd<NS>(c, tb, j, matvec_t(Rp, a.s));
      const V3 r = load_rot(c, j);
      Rod o;
      (void)rodrigues(r, o);
      add_rot_grad(c, j, rodrigues_bwd(rot_grad(a, Rp, Rs[k], ts[k]), r, o));
    }
    acc_add(a0, a);
    // ---- root: Rw_p = I, rel_0 = J_0 ----------------------------------------------------
    acc_point(a0, g0, t0);
    rel_offset_bwd<NS>(c, tb, 0, a0.s);
    const V3 rb = rodrigues_bwd(rot_grad(a0, eye3(), R0, t0), load_rot(c, 0), o0);
    c.G(0) = rb.x;
    c.G(1) = rb.y;
    c.G(2) = rb.z;
    c.G(kTranslOff) = a0.s.x;
    c.G(kTranslOff + 1) = a0.s.y;
    c.G(kTranslOff + 2) = a0.s.z;
  }
  return loss;
}

// ---------------------------------------------------------------------------------
// Adam, torch single-tensor semantics (torch/optim/adam.py:346-548, non-capturable CPU path):
//   m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g g; p -= step_k * m / (sqrt(v)/bc2_k + eps)
// step_k = lr / (1 - b1^k) and bc2_k = sqrt(1 - b2^k) are computed in double on the host.
// ---------------------------------------------------------------------------------
K2B_HD void adam_update(float& p, float& m, float& v, float g, float step_k, float bc2_k) {
  m = fmaf(0.1f, g - m, m);                 // 1 - 0.9 as float32
  v = fmaf(0.001f, g * g, v * 0.999f);      // float32(1 - 0.999) == 0.001f
  const float denom = fdiv(sqrtf(v), bc2_k) + 1e-8f;
  p = p - fdiv(step_k * m, denom);
}

}  // namespace k2b
