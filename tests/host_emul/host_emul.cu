// host_emul.cu -- DEBUGGING HARNESS, NOT PART OF THE PRODUCT.
//
// Runs the exact __host__ __device__ evaluation / optimiser code of the fitting kernel
// (keypoints2body_b200/csrc/fit_core.cuh, lbfgs_core.cuh) on the CPU, one frame at a time,
// so the maths can be checked against the oracle in the authoring container (which has no
// GPU) before GPU time is spent.  It is compiled into tests/host_emul/libk2b_host_emul.so by
// tests/host_emul/build.sh and loaded only by tests/test_host_emul.py.  The shipped library
// (libk2b_b200.so) contains none of this and has no CPU path.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../keypoints2body_b200/csrc/fit_core.cuh"
#include "../../keypoints2body_b200/csrc/lbfgs_core.cuh"
#include "../../keypoints2body_b200/csrc/shape_kernel.cuh"

using namespace k2b;

struct EmuModel {
  int ns;
  std::vector<float> chol, mu, nlw, rel;
};

extern "C" void* emu_model_create(int ns, const float* chol69, const float* means, const float* nlw,
                                  const double* J0, const double* JS, const int* parents, int nj) {
  EmuModel* m = new EmuModel();
  m->ns = ns;
  m->chol.assign((size_t)kGmmM * kCholStride, 0.f);
  m->mu.assign((size_t)kGmmM * kMuStride, 0.f);
  m->nlw.assign(nlw, nlw + kGmmM);
  for (int c = 0; c < kGmmM; ++c)
    for (int j = 0; j < kBodyDim; ++j) {
      for (int i = 0; i <= j; ++i)
        m->chol[(size_t)c * kCholStride + chol_row_off(j) + i] = chol69[((size_t)c * kBodyDim + j) * kBodyDim + i];
      m->mu[(size_t)c * kMuStride + j] = means[(size_t)c * kBodyDim + j];
    }
  m->rel.assign((size_t)kMaxFitJoints * (1 + ns) * 4, 0.f);
  const int nfit = nj < kMaxFitJoints ? nj : kMaxFitJoints;
  for (int j = 0; j < nfit; ++j) {
    const int pj = parents[j];
    for (int k = 0; k < 3; ++k) {
      m->rel[((size_t)j * (1 + ns)) * 4 + k] = (float)(J0[j * 3 + k] - (pj >= 0 ? J0[pj * 3 + k] : 0.0));
      for (int s = 0; s < ns; ++s)
        m->rel[((size_t)j * (1 + ns) + 1 + s) * 4 + k] =
            (float)(JS[((size_t)j * 3 + k) * ns + s] - (pj >= 0 ? JS[((size_t)pj * 3 + k) * ns + s] : 0.0));
    }
  }
  return m;
}
extern "C" void emu_model_destroy(void* m) { delete (EmuModel*)m; }

// runtime dispatch over the two vector lengths the fit kernel instantiates (85 / 95)
struct LbfgsAny {
  int n;
  Lbfgs<85> a;
  Lbfgs<95> b;
  explicit LbfgsAny(int n_) : n(n_) { a.init(); b.init(); }
  Cols eval_cols(const Cols& c, const Vecs& v) const { return n == 85 ? a.eval_cols(c, v) : b.eval_cols(c, v); }
  void begin(const Cols& c, const Vecs& v, float l, int it, float lr) { n == 85 ? a.advance_now(c, v, l, true, it, lr) : b.advance_now(c, v, l, true, it, lr); }
  void after_eval(const Cols& c, const Vecs& v, float l) { n == 85 ? a.advance_now(c, v, l, false, 0, 0.f) : b.advance_now(c, v, l, false, 0, 0.f); }
  bool done_() const { return n == 85 ? a.done : b.done; }
  double t_() const { return n == 85 ? a.t : b.t; }
  int evals_() const { return n == 85 ? a.evals : b.evals; }
  float dot_cur_d(const Vecs& v) const { return n == 85 ? a.dot_cur_d(v) : b.dot_cur_d(v); }
};

template <int NS, int K, bool ADAM>
static float run_eval(const Cols& c, const FitTables& tb, const FrameConsts& fc, bool g, bool pr, float* j, int* comp) {
  return eval_frame<NS, K, ADAM>(c, tb, fc, g, pr, j, comp);
}
typedef float (*EvalFn)(const Cols&, const FitTables&, const FrameConsts&, bool, bool, float*, int*);
// adam_build = the instantiation the Adam kernel uses (fused body-pose step inside the gradient pass)
static EvalFn pick(int ns, int K, bool adam_build = false) {
  if (adam_build) {
    if (K == 24) return run_eval<10, 24, true>;
    return ns == 20 ? run_eval<20, 22, true> : run_eval<10, 22, true>;
  }
  if (K == 24) return run_eval<10, 24, false>;
  return ns == 20 ? run_eval<20, 22, false> : run_eval<10, 22, false>;
}

// mode 0: evaluate (out_x = gradient), 1: Adam, 2: L-BFGS.  x layout [go3|body69|transl3|betas10|expr10?]
extern "C" int emu_fit(void* model, int mode, int B, int K, int iters, int freeze_betas, float lr, float joint_w,
                       float keep_w, const unsigned char* keep_on, const float* targets, const float* conf,
                       int conf_per_frame, const float* x0, const float* keep_pose, float* out_x, float* out_loss,
                       float* out_joints, int* out_evals, int* out_comp, float* out_trace /* [B][64][3] or null */,
                       int loss_kind, int final_mode, float depth_weight, const float* depth_ref) {
  EmuModel* m = (EmuModel*)model;
  const int NS = m->ns, NX = 75 + NS;
  FitTables tb{m->chol.data(), m->mu.data(), m->nlw.data(), (const float4*)m->rel.data()};
  EvalFn ev = pick(NS, K);
  for (int f = 0; f < B; ++f) {
    std::vector<float> x(x0 + (size_t)f * NX, x0 + (size_t)(f + 1) * NX), g(NX, 0.f), w(K);
    const bool stage1 = loss_kind == 1;
    for (int j = 0; j < K; ++j) {
      const float cf = conf ? (conf_per_frame ? conf[f * K + j] : conf[j]) : 1.f;
      w[j] = joint_w * joint_w * cf * cf;
      if (stage1) w[j] = (j == 1 || j == 2 || j == 16 || j == 17) ? 1.f : 0.f;
    }
    FrameConsts fc{targets + (size_t)f * K * 3, w.data(), keep_pose + (size_t)f * kBodyDim, 1,
                   (keep_on && keep_on[f]) ? keep_w * keep_w : 0.f, stage1, stage1 ? 4.f * depth_weight * depth_weight : 0.f,
                   {0.f, 0.f, 0.f}};
    if (stage1) for (int i = 0; i < 3; ++i) fc.dref[i] = depth_ref[f * 3 + i];
    const bool priors = !stage1;
    Cols c{x.data(), g.data(), 1, 1};
    float loss = 0.f;
    int evals = 0, comp = 0;
    float* jout = out_joints ? out_joints + (size_t)f * K * 3 : nullptr;
    if (mode == 0) {
      loss = ev(c, tb, fc, true, true, jout, &comp);
      evals = 1;
      memcpy(out_x + (size_t)f * NX, g.data(), NX * sizeof(float));
    } else if (mode == 1) {
      std::vector<float> m1(NX, 0.f), m2(NX, 0.f);
      EvalFn eva = pick(NS, K, true);
      for (int k = 1; k <= iters; ++k) {
        const float step_k = (float)((double)lr / (1.0 - std::pow(0.9, (double)k)));
        const float bc2_k = (float)std::sqrt(1.0 - std::pow(0.999, (double)k));
        // same split as fit_kernel: with the priors on, body-pose entries step inside the gradient pass
        const bool fused = priors;
        fc.adam_m = fused ? m1.data() : nullptr;
        fc.adam_v = m2.data();
        fc.adam_step = step_k;
        fc.adam_bc2 = bc2_k;
        loss = eva(c, tb, fc, true, priors, nullptr, nullptr);
        ++evals;
        for (int i = 0; i < NX; ++i) {
          if (fused && i >= 3 && i < kTranslOff) continue;
          if (freeze_betas && i >= kShapeOff && i < kShapeOff + 10) continue;
          if (stage1 && !(i < 3 || (i >= kTranslOff && i < kShapeOff))) continue;
          adam_update(x[i], m1[i], m2[i], g[i], step_k, bc2_k);
        }
      }
      fc.keep_w2 = 0.f;
      fc.adam_m = nullptr;
      {
        const float fl = ev(c, tb, fc, false, priors && final_mode, jout, nullptr);
        if (final_mode) loss = fl;
      }
      memcpy(out_x + (size_t)f * NX, x.data(), NX * sizeof(float));
    } else {
      const int hmax = lbfgs_history_capacity(iters);
      std::vector<float> scratch(Vecs::floats_per_frame(NX, hmax), 0.f);
      Vecs v{scratch.data(), 1, NX, hmax};
      LbfgsAny st(NX);
      {
        const Cols ce = st.eval_cols(c, v);
        loss = ev(ce, tb, fc, true, true, nullptr, nullptr);
        if (freeze_betas) for (int i = 0; i < 10; ++i) ce.G(kShapeOff + i) = 0.f;
      }
      st.begin(c, v, loss, iters, lr);
      int ntr = 0;
      while (!st.done_()) {
        const double t_trial = st.t_();
        const Cols ce = st.eval_cols(c, v);
        loss = ev(ce, tb, fc, true, true, nullptr, nullptr);
        if (freeze_betas) for (int i = 0; i < 10; ++i) ce.G(kShapeOff + i) = 0.f;
        if (out_trace && ntr < 64) {
          float* tr = out_trace + ((size_t)f * 64 + ntr) * 3;
          tr[0] = (float)t_trial; tr[1] = loss; tr[2] = st.dot_cur_d(v);
          ++ntr;
        }
        st.after_eval(c, v, loss);
      }
      evals = st.evals_();
      for (int i = 0; i < NX; ++i) x[i] = v.at(i);
      loss = ev(c, tb, fc, false, true, jout, nullptr);
      memcpy(out_x + (size_t)f * NX, x.data(), NX * sizeof(float));
    }
    out_loss[f] = loss;
    if (out_evals) out_evals[f] = evals;
    if (out_comp) out_comp[f] = comp;
  }
  return 0;
}

// sin/cos accuracy probe
extern "C" void emu_sincos(int n, const float* x, float* s, float* c) {
  for (int i = 0; i < n; ++i) k2b_sincos(x[i], &s[i], &c[i]);
}

// Line-search replay: drive the strong-Wolfe machine with recorded (f, gtd) responses.
// The objective is a table, so the machine's proposals t can be compared with torch's.
extern "C" int emu_linesearch_replay(double t0, double f0, float gtd0, double d_norm, int max_ls, int t_is_f32, int n_resp,
                                     const double* resp_f, const float* resp_gtd, double* out_t, double* out_final) {
  // 1-D surrogate: n = 1, direction d = 1, gradient slot values = gtd.
  float x = 0.f, g = gtd0;
  Cols c{&x, &g, 1, 1};
  std::vector<float> scratch(Vecs::floats_per_frame(1, 1) + 8, 0.f);
  Vecs v{scratch.data(), 1, 1, 1};
  Lbfgs<1> st;
  st.init();
  st.ls_replay_begin(c, v, t0, f0, gtd0, d_norm, max_ls, t_is_f32 != 0);
  int k = 0;
  while (!st.ls_replay_finished && k < n_resp) {
    out_t[k] = st.t;
    ThreadOps<1>::replay_response(v, st.cur, resp_gtd[k]);
    st.after_eval(c, v, (float)resp_f[k]);
    ++k;
  }
  out_final[0] = st.t;
  out_final[1] = st.loss;
  out_final[2] = (double)st.ls_evals;
  return k;
}

// Shape pre-pass for ONE sequence: frames evaluated sequentially, same machine as the kernel.
extern "C" int emu_shape_pass(void* model, const int* parents, int K, int T, int iters, float lr, float w,
                              const float* targets, const float* poses, const float* conf, const float* betas0,
                              float* out_betas, float* out_loss) {
  EmuModel* m = (EmuModel*)model;
  float x[10], g[10];
  for (int s = 0; s < 10; ++s) x[s] = betas0[s];
  Cols c{x, g, 1, 1};
  const int hmax = lbfgs_history_capacity(iters);
  std::vector<float> scratch(Vecs::floats_per_frame(10, hmax), 0.f);
  Vecs v{scratch.data(), 1, 10, hmax};
  Lbfgs<10> st;
  st.init();
  int stage = 0;
  while (true) {
    float grad[10] = {0};
    float loss = 0.f;
    for (int t = 0; t < T; ++t)
      loss += shape_frame_eval((const float4*)m->rel.data(), m->ns, parents, K, poses + (size_t)t * 72,
                               targets + (size_t)t * K * 3, conf, x, grad);
    float bb = 0.f;
    for (int s = 0; s < 10; ++s) bb += x[s] * x[s];
    loss += (float)T * w * w * bb;
    { const Cols ce = st.eval_cols(c, v); for (int s = 0; s < 10; ++s) ce.G(s) = grad[s] + 2.f * (float)T * w * w * x[s]; }
    st.advance_now(c, v, loss, stage == 0, iters, lr);
    stage = 1;
    if (st.done) break;
  }
  for (int s = 0; s < 10; ++s) out_betas[s] = v.at(s);
  *out_loss = (float)st.loss;
  return st.evals;
}
