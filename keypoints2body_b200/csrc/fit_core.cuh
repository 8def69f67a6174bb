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

// Packed lower-triangular Cholesky factor in 9 row panels of 8 rows (the last has 5): every
// row of panel p stores L[j][0 .. 8(p+1)-1] (structural zeros above the diagonal kept), so a
// panel is a dense 8 x 8(p+1) block and the kernels walk it with a rolled row loop and a
// statically unrolled column loop.  2664 floats per component (2415 non-zeros).
K2B_HD constexpr int chol_panel_off(int p) { return 32 * p * (p + 1); }
K2B_HD constexpr int chol_row_off(int j) { return chol_panel_off(j / 8) + (j % 8) * 8 * (j / 8 + 1); }
constexpr int kCholStride = 2664;
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

// Packed pair of FP32 multiply-adds: one FFMA2 instruction on sm_100 (Blackwell's two-wide FP32
// FMA, half the issue slots of two FFMAs); two fmaf on the host.
K2B_HD float2 fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
  return __ffma2_rn(a, b, c);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
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

// one rounding each, never contracted into a neighbouring operation
K2B_HD float mul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float p = a * b;
  return p;
#endif
}
K2B_HD float add_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  volatile float p = a + b;
  return p;
#endif
}

K2B_HD M3 rodrigues(V3 r, Rod& o) {
  const V3 a = v3(r.x + 1e-8f, r.y + 1e-8f, r.z + 1e-8f);
  const float theta = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
  o.inv = fdiv(1.f, theta);
  k2b_sincos(theta, &o.s, &o.c);
  const float oc = 1.f - o.c;
  const V3 k = v3(r.x * o.inv, r.y * o.inv, r.z * o.inv);
  // The diagonal is written with explicit roundings: left to the compiler, `1 - oc (yy + zz)` was contracted one way
  // in one build of a kernel and another way in the next (fma(k.y, k.y, zz) or not, depending on unrelated code around
  // it), which moves every fit by an ulp and, along a chain of warm-started L-BFGS fits, by millimetres.
  // (the form below is the one the shipped warp-per-sequence kernel had: yy + zz in two roundings, xx + zz and xx + yy
  // as fma(k.x, k.x, .))
  const float yy = mul_rn(k.y, k.y), zz = mul_rn(k.z, k.z);
  const float xy = k.x * k.y, xz = k.x * k.z, yz = k.y * k.z;
  M3 R;
  R.m[0] = fmaf(-oc, add_rn(yy, zz), 1.f); R.m[1] = fmaf(oc, xy, -o.s * k.z); R.m[2] = fmaf(oc, xz, o.s * k.y);
  R.m[3] = fmaf(oc, xy, o.s * k.z); R.m[4] = fmaf(-oc, fmaf(k.x, k.x, zz), 1.f); R.m[5] = fmaf(oc, yz, -o.s * k.x);
  R.m[6] = fmaf(oc, xz, -o.s * k.y); R.m[7] = fmaf(oc, yz, o.s * k.x); R.m[8] = fmaf(-oc, fmaf(k.x, k.x, yy), 1.f);
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

struct FrameConsts {
  const float* tgt;      // element (j*3+c) at tgt[(j*3+c)*stride]
  const float* wgt;      // joint weight joint_w^2 * conf_j^2 at wgt[j*stride]
  const float* keep;     // preserve pose element i at keep[i*stride]
  long stride;
  float keep_w2;         // pose_preserve_weight^2 or 0
  // camera-space stage 1 (camera_fitting_loss_3d, core/losses.py:70-93): plain squared joint error
  // instead of GMoF, plus depth_w2 * |transl - depth_ref|^2
  bool plain_sq;
  float depth_w2;
  float dref[3];
  // Adam fused into the gradient pass (Adam kernel only): when adam_m is set, the 69 body-pose entries are
  // updated where their gradient is produced (moments at adam_m / adam_v [(3 + i) * stride]) and their
  // gradient is not written back; the caller updates the remaining entries.
  float* adam_m;
  float* adam_v;
  float adam_step, adam_bc2;
};

// Column accessor: element i of this frame's parameter vector x (shared memory on the device,
// column stride xs) and gradient vector g (coalesced global scratch, column stride gs).
struct Cols {
  float* x;
  float* g;
  int xs;
  long gs;
  K2B_HD float& X(int i) const { return x[i * xs]; }
  K2B_HD float& G(int i) const { return g[(long)i * gs]; }
};

// ---------------------------------------------------------------------------------
// GMM max-mixture prior on body pose (prior.py:182-195) with P_m = L_m L_m^T:
//   q_m = ||L_m^T d||^2, d = x - mu_m;  m* = argmin(0.5 q_m + nlw_m)
//   gradient = L_m* (L_m*^T d)
// z = L^T d is accumulated in 72 registers by a rolled loop over the rows of each panel; the
// arg-min component's z is recomputed in a 9th pass (cheaper than keeping a second register
// set alive), then G(3..71) = kPosePriorW2 * L z.  Returns kPosePriorW2 * min.
// (Tried in round 2: scanning the previous evaluation's winner last so that its z is still in registers and the
// ninth pass can be skipped.  The extra live state pushed the Adam kernel from 68 to 520 bytes of spills and cost 25 %.)
// ---------------------------------------------------------------------------------
template <int P>
K2B_HD void gmm_zpanel(const Cols& c, const float* __restrict__ Lm, const float* __restrict__ mum, float2 (&z)[36]) {
  constexpr int W = 8 * (P + 1);
  constexpr int ROWS = P < 8 ? 8 : 5;
  const float* row = Lm + chol_panel_off(P);
  // narrow panels: unroll rows so the next row's loads overlap this row's multiply-adds
#pragma unroll (P <= 1 ? 4 : (P <= 4 ? 2 : 1))
  for (int r = 0; r < ROWS; ++r, row += W) {
    const int j = 8 * P + r;
    const float dj = c.X(3 + j) - mum[j];
    const float2 dj2 = make_float2(dj, dj);
    const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
      const float4 l = r4[q];
      z[2 * q + 0] = fma2(make_float2(l.x, l.y), dj2, z[2 * q + 0]);
      z[2 * q + 1] = fma2(make_float2(l.z, l.w), dj2, z[2 * q + 1]);
    }
  }
}

// Rows of panel P of g = L z.  Adds, in the same pass, the other body-pose terms so the
// gradient column is touched once: G(3+i) = kinematic part (already there) + pose-prior part
// + temporal pose-preserve part (losses.py:57-59) + angle-prior part (losses.py:13-21).
// Ask for the kinematic gradients (and preserve poses) of panel P ahead of their use: they sit in global
// scratch (L2), and nothing else is in flight to hide that latency.  No registers are held.
template <int P>
K2B_HD void gmm_gprefetch(const Cols& c, const FrameConsts& fc) {
#if defined(__CUDA_ARCH__)
  if (P <= 8) {
    constexpr int ROWS = P < 8 ? 8 : 5;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(&c.G(3 + 8 * P + r)));
      if (fc.keep_w2 != 0.f) asm volatile("prefetch.global.L1 [%0];" ::"l"(fc.keep + (8 * P + r) * fc.stride));
      if (fc.adam_m) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(fc.adam_m + (3 + 8 * P + r) * fc.stride));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(fc.adam_v + (3 + 8 * P + r) * fc.stride));
      }
    }
  }
#endif
}

template <int P, bool HINTS>
K2B_HD void gmm_gpanel(const Cols& c, const FrameConsts& fc, const float* __restrict__ Lb, const float2 (&z)[36],
                       float& extra_loss) {
  constexpr int W = 8 * (P + 1);
  constexpr int ROWS = P < 8 ? 8 : 5;
  if (HINTS) gmm_gprefetch<P + 2>(c, fc);
  const float* row = Lb + chol_panel_off(P);
#pragma unroll (P <= 1 ? 4 : (P <= 4 ? 2 : 1))
  for (int r = 0; r < ROWS; ++r, row += W) {
    const int i = 8 * P + r;
    const float gk = c.G(3 + i);                      // issued early, consumed after the dot product
    const bool fuse = HINTS && fc.adam_m != nullptr;
    float am = 0.f, av = 0.f;
    if (fuse) {
      am = fc.adam_m[(3 + i) * fc.stride];
      av = fc.adam_v[(3 + i) * fc.stride];
    }
    const float xi = c.X(3 + i);
    const float keep = fc.keep_w2 != 0.f ? fc.keep[i * fc.stride] : xi;
    const float4* r4 = reinterpret_cast<const float4*>(row);
    float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
      const float4 l = r4[q];
      a01 = fma2(make_float2(l.x, l.y), z[2 * q + 0], a01);
      a23 = fma2(make_float2(l.z, l.w), z[2 * q + 1], a23);
    }
    float g = fmaf(kPosePriorW2, (a01.x + a01.y) + (a23.x + a23.y), gk);
    const float d = xi - keep;
    extra_loss = fmaf(fc.keep_w2 * d, d, extra_loss);
    g = fmaf(2.f * fc.keep_w2, d, g);
    if (i == 9 || i == 12 || i == 52 || i == 55) {
      const float sgn = i == 52 ? 1.f : -1.f;
      const float e = expf(xi * sgn);
      extra_loss = fmaf(kAnglePriorW2, e * e, extra_loss);
      g = fmaf(2.f * kAnglePriorW2 * sgn, e * e, g);
    }
    if (fuse) {   // Adam step of this entry, right where its gradient exists
      float xn = xi;
      adam_update(xn, am, av, g, fc.adam_step, fc.adam_bc2);
      c.X(3 + i) = xn;
      fc.adam_m[(3 + i) * fc.stride] = am;
      fc.adam_v[(3 + i) * fc.stride] = av;
    } else {
      c.G(3 + i) = g;
    }
  }
}

// with_grad: also folds the preserve / angle terms in (see gmm_gpanel) and returns their loss in
// `extra_loss`; without it the caller adds those terms itself.
template <bool HINTS>
K2B_HD float gmm_prior(const Cols& c, const FitTables& tb, const FrameConsts& fc, bool with_grad, int* best_m_out,
                       float& extra_loss) {
  float best = INFINITY;
  int best_m = 0;
  float2 z[36];
  const int passes = with_grad ? kGmmM + 1 : kGmmM;
#pragma unroll 1
  for (int it = 0; it < passes; ++it) {
    const int m = it < kGmmM ? it : best_m;   // last pass: recompute z of the arg-min component
    const float* __restrict__ Lm = tb.chol + m * kCholStride;
    const float* __restrict__ mum = tb.mu + m * kMuStride;
#pragma unroll
    for (int i = 0; i < 36; ++i) z[i] = make_float2(0.f, 0.f);
    gmm_zpanel<0>(c, Lm, mum, z); gmm_zpanel<1>(c, Lm, mum, z); gmm_zpanel<2>(c, Lm, mum, z);
    gmm_zpanel<3>(c, Lm, mum, z); gmm_zpanel<4>(c, Lm, mum, z); gmm_zpanel<5>(c, Lm, mum, z);
    gmm_zpanel<6>(c, Lm, mum, z); gmm_zpanel<7>(c, Lm, mum, z); gmm_zpanel<8>(c, Lm, mum, z);
    if (it < kGmmM) {
      float2 q01 = make_float2(0.f, 0.f), q23 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 34; i += 2) {
        q01 = fma2(z[i], z[i], q01);
        q23 = fma2(z[i + 1], z[i + 1], q23);
      }
      q01.x = fmaf(z[34].x, z[34].x, q01.x);    // element 68; 69..71 are padding
      const float ll = fmaf(0.5f, (q01.x + q01.y) + (q23.x + q23.y), tb.nlw[m]);
      if (ll < best) {   // strict: first minimum wins, like torch.min
        best = ll;
        best_m = m;
      }
    }
  }
  if (best_m_out) *best_m_out = best_m;
  if (with_grad) {
    const float* __restrict__ Lb = tb.chol + best_m * kCholStride;
    if (HINTS) {
      gmm_gprefetch<0>(c, fc);
      gmm_gprefetch<1>(c, fc);
    }
    gmm_gpanel<0, HINTS>(c, fc, Lb, z, extra_loss); gmm_gpanel<1, HINTS>(c, fc, Lb, z, extra_loss);
    gmm_gpanel<2, HINTS>(c, fc, Lb, z, extra_loss); gmm_gpanel<3, HINTS>(c, fc, Lb, z, extra_loss);
    gmm_gpanel<4, HINTS>(c, fc, Lb, z, extra_loss); gmm_gpanel<5, HINTS>(c, fc, Lb, z, extra_loss);
    gmm_gpanel<6, HINTS>(c, fc, Lb, z, extra_loss); gmm_gpanel<7, HINTS>(c, fc, Lb, z, extra_loss);
    gmm_gpanel<8, HINTS>(c, fc, Lb, z, extra_loss);
  }
  return kPosePriorW2 * best;
}

// ---------------------------------------------------------------------------------
// Kinematic tree: rolled walk with O(1) state.
//
// A chain is walked forward to its tail keeping only the running (Rw, t); on the way back the
// parent's state is reconstructed from the child's -- Rw_p = Rw_j R_j^T, t_p = t_j - Rw_p rel_j --
// so no per-joint state is stored.  Residuals (loss and gradient) are evaluated on the way
// back, where every joint is visited exactly once.
// Tree (SMPL body, identical in SMPL / SMPL-H / SMPL-X): root 0; legs {1,4,7,10}+side from the
// root; spine 3 -> 6 -> 9; neck {12,15} and arms {13,16,18,20,(22)}+side from joint 9.
// ---------------------------------------------------------------------------------
enum ChainType { kSpine = 0, kNeck = 1, kArm = 2, kLeg = 3 };

K2B_HD int chain_joint(int type, int side, int k) {
  if (type == kSpine) return 3 + 3 * k;
  if (type == kNeck) return 12 + 3 * k;
  if (type == kLeg) return 1 + side + 3 * k;
  return 13 + side + (k ? 1 + 2 * k : 0);
}

struct KinState {
  M3 R;   // world rotation of the current joint (of its parent while standing on a leaf)
  V3 t;   // position of the current joint, translation not yet added
};

template <int NS>
struct KinCtx {
  const Cols& c;
  const FitTables& tb;
  const FrameConsts& fc;
  float shape[NS];
  float shape_bar[NS];
  V3 transl;
  float loss;
  bool with_grad;
  float* joints_out;
};

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

template <int NS>
K2B_HD void rel_offset_bwd(const FitTables& tb, int j, V3 rb, float (&shape_bar)[NS]) {
  const float4* e = tb.rel + j * (1 + NS);
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const float4 d = e[1 + s];
    shape_bar[s] = fmaf(d.x, rb.x, fmaf(d.y, rb.y, fmaf(d.z, rb.z, shape_bar[s])));
  }
}

K2B_HD V3 load_rot(const Cols& c, int j) { return v3(c.X(3 * j), c.X(3 * j + 1), c.X(3 * j + 2)); }

// Observation of one joint (weight, target), read from global scratch: requested one joint ahead of its
// use so the L2 latency overlaps the previous joint's arithmetic.
struct Obs {
  float w, x, y, z;
};
K2B_HD Obs load_obs(const FrameConsts& fc, int j) {
  return Obs{fc.wgt[j * fc.stride], fc.tgt[(3 * j + 0) * fc.stride], fc.tgt[(3 * j + 1) * fc.stride],
             fc.tgt[(3 * j + 2) * fc.stride]};
}

// Residual of joint j at position t: accumulates the loss, returns d loss / d position.
template <int NS>
K2B_HD V3 residual(KinCtx<NS>& k, int j, V3 t, const Obs& ob) {
  const V3 p = t + k.transl;
  if (k.joints_out) {
    k.joints_out[3 * j + 0] = p.x;
    k.joints_out[3 * j + 1] = p.y;
    k.joints_out[3 * j + 2] = p.z;
  }
  const FrameConsts& fc = k.fc;
  const float w = ob.w;
  const float ex = p.x - ob.x;
  const float ey = p.y - ob.y;
  const float ez = p.z - ob.z;
  if (k.fc.plain_sq) {
    k.loss = fmaf(w, fmaf(ex, ex, fmaf(ey, ey, ez * ez)), k.loss);
    return v3(2.f * w * ex, 2.f * w * ey, 2.f * w * ez);
  }
  const float ix = fdiv(1.f, kSigma2 + ex * ex), iy = fdiv(1.f, kSigma2 + ey * ey), iz = fdiv(1.f, kSigma2 + ez * ez);
  const float gx = kSigma2 * ex * ex * ix, gy = kSigma2 * ey * ey * iy, gz = kSigma2 * ez * ez * iz;
  k.loss = fmaf(w, (gx + gy) + gz, k.loss);
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

// A B^T
K2B_HD M3 matmul_nt(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C.m[3 * i + j] = fmaf(A.m[3 * i], B.m[3 * j], fmaf(A.m[3 * i + 1], B.m[3 * j + 1], A.m[3 * i + 2] * B.m[3 * j + 2]));
  return C;
}

// Forward: advance `s` from the chain's base to its tail joint.
template <int NS>
K2B_HD void chain_fwd(KinCtx<NS>& k, int type, int side, int len, bool leaf_tail, KinState& s) {
#pragma unroll 1
  for (int i = 0; i < len; ++i) {
    const int j = chain_joint(type, side, i);
    s.t = matvec(s.R, rel_offset<NS>(k.tb, j, k.shape)) + s.t;
    if (!(leaf_tail && i == len - 1)) {
      Rod o;
      s.R = matmul(s.R, rodrigues(load_rot(k.c, j), o));
    }
  }
}

// Backward: `s` stands on the tail; `a` holds the subtree sums of the tail's children (zero for a
// leaf).  Walks back to the base, emitting losses / gradients; returns with `a` = chain sums.
template <int NS, bool HINTS>
K2B_HD void chain_bwd(KinCtx<NS>& k, int type, int side, int len, bool leaf_tail, KinState& s, Acc& a) {
  Obs ob = load_obs(k.fc, chain_joint(type, side, len - 1));
#pragma unroll 1
  for (int i = len - 1; i >= 0; --i) {
    const int j = chain_joint(type, side, i);
    // HINTS: the next joint's observation is requested now; otherwise each one is loaded where it is used
    const Obs ob_next = HINTS ? load_obs(k.fc, chain_joint(type, side, i > 0 ? i - 1 : 0)) : Obs{};
    const bool has_rot = !(leaf_tail && i == len - 1);
    M3 Rp = s.R;
    V3 r = v3(0.f, 0.f, 0.f);
    Rod o;
    if (has_rot) {
      r = load_rot(k.c, j);
      Rp = matmul_nt(s.R, rodrigues(r, o));    // Rw_p = Rw_j R_j^T
    }
    const V3 g = residual<NS>(k, j, s.t, HINTS || i == len - 1 ? ob : load_obs(k.fc, j));
    ob = ob_next;
    if (k.with_grad) {
      acc_point(a, g, s.t);
      rel_offset_bwd<NS>(k.tb, j, matvec_t(Rp, a.s), k.shape_bar);
      V3 rb = v3(0.f, 0.f, 0.f);                      // a leaf's own rotation moves nothing observed
      if (has_rot) rb = rodrigues_bwd(rot_grad(a, Rp, s.R, s.t), r, o);
      k.c.G(3 * j + 0) = rb.x;                          // each entry is written exactly once per evaluation
      k.c.G(3 * j + 1) = rb.y;
      k.c.G(3 * j + 2) = rb.z;
    }
    s.t = s.t - matvec(Rp, rel_offset<NS>(k.tb, j, k.shape));
    s.R = Rp;
  }
}

// ---------------------------------------------------------------------------------
// One function evaluation for one frame (K observed joints: 22 AMASS / 24 SMPL24).
// with_grad fills G(0 .. 75+NS) completely; with_priors = false skips every prior term
// (joints-only final forward).  Returns the total loss (losses.py:41-67).
// ---------------------------------------------------------------------------------
// HINTS (the Adam kernel's build): software latency hints (observations one joint ahead, L1 prefetch of the
// gradient rows; +3 % there, -2 % in the L-BFGS kernel) and the option of fusing the Adam step of the
// body-pose entries into the gradient pass (FrameConsts::adam_m).
template <int NS, int K, bool HINTS = false>
K2B_HD float eval_frame(const Cols& c, const FitTables& tb, const FrameConsts& fc, bool with_grad,
                        bool with_priors, float* joints_out, int* gmm_component) {
  constexpr int ARM_LEN = (K == 24) ? 5 : 4;
  float loss = 0.f;

  KinCtx<NS> k{c, tb, fc};
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    k.shape[s] = c.X(kShapeOff + s);
    k.shape_bar[s] = 0.f;
  }
  k.transl = v3(c.X(kTranslOff), c.X(kTranslOff + 1), c.X(kTranslOff + 2));
  k.loss = 0.f;
  k.with_grad = with_grad;
  k.joints_out = joints_out;
  if (with_priors) {   // shape prior on betas only (losses.py:56)
    float acc = 0.f;
#pragma unroll
    for (int s = 0; s < 10; ++s) acc = fmaf(k.shape[s], k.shape[s], acc);
    loss = fmaf(kShapePriorW2, acc, loss);
  }

  // ---- root ---------------------------------------------------------------------------
  Rod o0;
  const V3 r0 = load_rot(c, 0);
  KinState root;
  root.R = rodrigues(r0, o0);
  root.t = rel_offset<NS>(tb, 0, k.shape);

  // ---- spine forward to joint 9, then everything hanging from it ---------------------------
  KinState j9 = root;
  chain_fwd<NS>(k, kSpine, 0, 3, false, j9);
  Acc a9{zero3(), v3(0.f, 0.f, 0.f)};
  Acc a0{zero3(), v3(0.f, 0.f, 0.f)};
#pragma unroll 1
  for (int ch = 0; ch < 5; ++ch) {          // neck, left arm, right arm, left leg, right leg
    const bool from9 = ch < 3;
    const int type = ch == 0 ? kNeck : (ch < 3 ? kArm : kLeg);
    const int side = ch == 0 ? 0 : ((ch - 1) & 1);
    const int len = ch == 0 ? 2 : (ch < 3 ? ARM_LEN : 4);
    KinState s;
#pragma unroll
    for (int i = 0; i < 9; ++i) s.R.m[i] = from9 ? j9.R.m[i] : root.R.m[i];
    s.t = from9 ? j9.t : root.t;
    chain_fwd<NS>(k, type, side, len, true, s);
    Acc a{zero3(), v3(0.f, 0.f, 0.f)};
    chain_bwd<NS, HINTS>(k, type, side, len, true, s, a);
    if (with_grad) {
      if (from9) acc_add(a9, a); else acc_add(a0, a);
    }
  }
  // ---- spine backward 9 -> 6 -> 3, then the root -------------------------------------------
  const Obs ob0 = load_obs(fc, 0);          // in flight during the spine's backward walk
  chain_bwd<NS, HINTS>(k, kSpine, 0, 3, false, j9, a9);
  const V3 g0 = residual<NS>(k, 0, root.t, ob0);
  loss += k.loss;
  V3 dgrad = v3(0.f, 0.f, 0.f);
  if (fc.depth_w2 != 0.f) {   // camera-space stage 1: keep the translation near its initial estimate
    const V3 d = v3(k.transl.x - fc.dref[0], k.transl.y - fc.dref[1], k.transl.z - fc.dref[2]);
    loss = fmaf(fc.depth_w2, dot(d, d), loss);
    dgrad = v3(2.f * fc.depth_w2 * d.x, 2.f * fc.depth_w2 * d.y, 2.f * fc.depth_w2 * d.z);
  }
  if (with_grad) {
    acc_add(a0, a9);
    acc_point(a0, g0, root.t);
    rel_offset_bwd<NS>(tb, 0, a0.s, k.shape_bar);           // Rw_p = I, rel_0 = J_0
    const V3 rb = rodrigues_bwd(rot_grad(a0, eye3(), root.R, root.t), r0, o0);
    c.G(0) = rb.x;
    c.G(1) = rb.y;
    c.G(2) = rb.z;
    c.G(kTranslOff) = a0.s.x + dgrad.x;
    c.G(kTranslOff + 1) = a0.s.y + dgrad.y;
    c.G(kTranslOff + 2) = a0.s.z + dgrad.z;
#pragma unroll
    for (int s = 0; s < NS; ++s)
      c.G(kShapeOff + s) = (with_priors && s < 10) ? fmaf(2.f * kShapePriorW2, k.shape[s], k.shape_bar[s]) : k.shape_bar[s];
    if (K == 22) {   // body-pose entries of joints 22, 23: prior-only
#pragma unroll
      for (int i = 66; i < 72; ++i) c.G(i) = 0.f;
    }
  }

  // ---- priors on the body pose (after the kinematic part of the gradient is in place) -------
  if (with_priors) {
    float extra = 0.f;
    loss += gmm_prior<HINTS>(c, tb, fc, with_grad, gmm_component, extra);
    if (!with_grad) {
      // angle prior (losses.py:13-21) on body-pose entries 52, 55, 9, 12 with signs +,-,-,-
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int idx = q == 0 ? 52 : (q == 1 ? 55 : (q == 2 ? 9 : 12));
        const float e = expf(c.X(3 + idx) * (q == 0 ? 1.f : -1.f));
        extra = fmaf(kAnglePriorW2, e * e, extra);
      }
      // temporal pose-preserve term (losses.py:57-59), active when seq_ind > 0
      if (fc.keep_w2 != 0.f) {
        float acc = 0.f;
#pragma unroll 23
        for (int i = 0; i < kBodyDim; ++i) {
          const float d = c.X(3 + i) - fc.keep[i * fc.stride];
          acc = fmaf(d, d, acc);
        }
        extra = fmaf(fc.keep_w2, acc, extra);
      }
    }
    loss += extra;
  }
  return loss;
}

}  // namespace k2b
