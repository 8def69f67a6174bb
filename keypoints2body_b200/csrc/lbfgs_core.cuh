// lbfgs_core.cuh -- torch.optim.LBFGS(line_search_fn="strong_wolfe") as a per-frame
// state machine that asks for exactly one closure evaluation per round.
//
// Restates torch/optim/lbfgs.py of the torch build the reference runs on
// (2.11): _cubic_interpolate :12-37, _strong_wolfe :40-209, LBFGS.step :332-537,
// as driven by /root/reference/keypoints2body/core/fitters/world_space.py:231-247
// (one optimizer.step(closure), max_iter = num_iters, max_eval = max_iter*5//4,
// history_size 100, tolerance_grad 1e-7, tolerance_change 1e-9, lr = step_size).
//
// All lanes of a warp evaluate the objective together; this machine is the cheap,
// possibly divergent, scalar part between evaluations.  Vectors (iterate, direction,
// three gradient slots, (y,s) history) live in per-frame scratch reached through
// `Vecs`; the trial point and its gradient live in the x / g columns (Cols).
//
// Precision: the loss and step length are carried in double where torch carries
// Python floats; dot products and the two-loop recursion run in float32 like the
// float32 tensors torch uses.
#pragma once

#include "fit_core.cuh"

namespace k2b {

constexpr double kTolGrad = 1e-7;
constexpr double kTolChange = 1e-9;
constexpr double kC1 = 1e-4;
constexpr double kC2 = 0.9;
constexpr int kHistorySize = 100;

// Per-frame scratch vectors: element e of this frame at base[e * stride].
struct Vecs {
  float* base;
  long stride;
  int n;     // optimised vector length (75 + NS)
  int hmax;  // history capacity
  K2B_HD float& at(int e) const { return base[(long)e * stride]; }
  K2B_HD int xk() const { return 0; }
  K2B_HD int d() const { return n; }
  K2B_HD int gslot(int s) const { return (2 + s) * n; }
  K2B_HD int y(int h) const { return (5 + h) * n; }
  K2B_HD int s(int h) const { return (5 + hmax + h) * n; }
  K2B_HD int ro(int h) const { return (5 + 2 * hmax) * n + h; }
  K2B_HD int al(int h) const { return (5 + 2 * hmax) * n + hmax + h; }
  static K2B_HD long floats_per_frame(int n, int hmax) { return (long)(5 + 2 * hmax) * n + 2 * hmax; }
};

K2B_HD int lbfgs_history_capacity(int max_iter) {
  int h = max_iter - 1;
  if (h < 1) h = 1;
  return h > kHistorySize ? kHistorySize : h;
}

// torch/optim/lbfgs.py:12-37.  g1, g2 are float32 tensors there, so the derived
// quantities are rounded to float32; positions and function values are doubles.
// x_f32: the positions are float32 tensors (first outer iteration, where t = lr / ||g||_1 is a
// tensor), so the quotient is formed in float32 instead of double.
K2B_HD double cubic_interpolate(double x1, double f1, float g1, double x2, double f2, float g2,
                                bool has_bounds, double lo, double hi, bool x_f32) {
  if (!has_bounds) {
    lo = x1 <= x2 ? x1 : x2;
    hi = x1 <= x2 ? x2 : x1;
  }
  const float quot = x_f32 ? (float)(3.0 * (f1 - f2)) / ((float)x1 - (float)x2)
                           : (float)(3.0 * (f1 - f2) / (x1 - x2));
  const float d1 = (g1 + g2) - quot;
  const float d2sq = d1 * d1 - g1 * g2;
  if (d2sq >= 0.f) {
    const float d2 = sqrtf(d2sq);
    float min_pos;
    if (x1 <= x2)
      min_pos = (float)x2 - (float)(x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2.f * d2));
    else
      min_pos = (float)x1 - (float)(x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2.f * d2));
    double r = (double)min_pos;
    r = (lo > r) ? lo : r;    // Python max(min_pos, lo)
    r = (hi < r) ? hi : r;    // Python min(.., hi)
    return r;
  }
  return (lo + hi) / 2.0;
}

struct Lbfgs {
  // configuration
  int max_iter, max_eval;
  float lr;
  // outer loop
  int n_iter, evals, num_old, head;
  bool done;
  double loss, prev_loss, t;
  float H_diag;
  int g0;  // gradient slot holding flat_grad of the current iterate
  // line search
  int phase;  // 0 bracket, 1 zoom
  bool first_eval, ls_done, insuf, t_f32;
  int ls_iter, max_ls, ls_evals;
  double f0, d_norm;
  float gtd0;
  double t_prev, f_prev;
  float gtd_prev;
  int slot_prev;
  double br_t[2], br_f[2];
  float br_gtd[2];
  int br_slot[2];
  int br_n, low, high;

  K2B_HD int free_slot(int keep1, int keep2) const {
#pragma unroll
    for (int s = 0; s < 3; ++s)
      if (s != g0 && s != keep1 && s != keep2) return s;
    return (g0 + 1) % 3;  // unreachable when keep1/keep2 follow the analysis in DESIGN.md
  }

  K2B_HD static void store_grad(const Cols& c, const Vecs& v, int slot) {
    const int o = v.gslot(slot);
#pragma unroll 5
    for (int i = 0; i < v.n; ++i) v.at(o + i) = c.G(i);
  }
  K2B_HD static float dot_g_d(const Cols& c, const Vecs& v) {
    float a = 0.f;
    const int od = v.d();
#pragma unroll 5
    for (int i = 0; i < v.n; ++i) a = fmaf(c.G(i), v.at(od + i), a);
    return a;
  }
  K2B_HD void set_trial(const Cols& c, const Vecs& v) const {
    const float tf = (float)t;
    const int od = v.d();
#pragma unroll 5
    for (int i = 0; i < v.n; ++i) c.X(i) = fmaf(tf, v.at(od + i), v.at(i));
  }

  // Called once after the first evaluation at the initial parameters (loss, G valid).
  K2B_HD void begin(const Cols& c, const Vecs& v, float loss0, int max_iter_, float lr_) {
    max_iter = max_iter_;
    max_eval = max_iter_ * 5 / 4;
    lr = lr_;
    n_iter = 0;
    evals = 1;
    num_old = 0;
    head = 0;
    done = false;
    loss = (double)loss0;
    prev_loss = loss;
    t = 0.0;
    H_diag = 1.f;
    g0 = 0;
    store_grad(c, v, 0);
    float gmax = 0.f;
    for (int i = 0; i < v.n; ++i) {
      v.at(i) = c.X(i);
      gmax = fmaxf(gmax, fabsf(c.G(i)));
    }
    if (max_iter_ <= 0 || (double)gmax <= kTolGrad) {
      done = true;
      return;
    }
    start_outer(c, v);
  }

  // lbfgs.py:388-476: direction update, initial step, line-search setup, first trial point.
  K2B_HD void start_outer(const Cols& c, const Vecs& v) {
    const int n = v.n, od = v.d(), og = v.gslot(g0);
    ++n_iter;
    if (n_iter == 1) {
#pragma unroll 5
      for (int i = 0; i < n; ++i) v.at(od + i) = -v.at(og + i);
      H_diag = 1.f;
      num_old = 0;
      head = 0;
    } else {
      // y = flat_grad - prev_flat_grad (still in slot g_prev_slot), s = d * t
      const int op = v.gslot(slot_prev_grad);
      const float tf = (float)t;
      float ys = 0.f, yy = 0.f;
#pragma unroll 5
      for (int i = 0; i < n; ++i) {
        const float yi = v.at(og + i) - v.at(op + i);
        const float si = v.at(od + i) * tf;
        ys = fmaf(yi, si, ys);
        yy = fmaf(yi, yi, yy);
      }
      if (ys > 1e-10f) {
        int h;
        if (num_old == v.hmax) {  // drop the oldest pair (ring buffer)
          h = head;
          head = (head + 1) % v.hmax;
        } else {
          h = (head + num_old) % v.hmax;
          ++num_old;
        }
        const int oy = v.y(h), os = v.s(h);
#pragma unroll 5
        for (int i = 0; i < n; ++i) {
          v.at(oy + i) = v.at(og + i) - v.at(op + i);
          v.at(os + i) = v.at(od + i) * tf;
        }
        v.at(v.ro(h)) = 1.f / ys;
        H_diag = ys / yy;
      }
      // two-loop recursion (lbfgs.py:430-442); q lives in the direction buffer
#pragma unroll 5
      for (int i = 0; i < n; ++i) v.at(od + i) = -v.at(og + i);
      for (int k = num_old - 1; k >= 0; --k) {
        const int h = (head + k) % v.hmax;
        const int oy = v.y(h), os = v.s(h);
        float a = 0.f;
#pragma unroll 5
        for (int i = 0; i < n; ++i) a = fmaf(v.at(os + i), v.at(od + i), a);
        a *= v.at(v.ro(h));
        v.at(v.al(h)) = a;
#pragma unroll 5
        for (int i = 0; i < n; ++i) v.at(od + i) = fmaf(-a, v.at(oy + i), v.at(od + i));
      }
#pragma unroll 5
      for (int i = 0; i < n; ++i) v.at(od + i) *= H_diag;
      for (int k = 0; k < num_old; ++k) {
        const int h = (head + k) % v.hmax;
        const int oy = v.y(h), os = v.s(h);
        float b = 0.f;
#pragma unroll 5
        for (int i = 0; i < n; ++i) b = fmaf(v.at(oy + i), v.at(od + i), b);
        b *= v.at(v.ro(h));
        const float coef = v.at(v.al(h)) - b;
#pragma unroll 5
        for (int i = 0; i < n; ++i) v.at(od + i) = fmaf(coef, v.at(os + i), v.at(od + i));
      }
    }
    slot_prev_grad = g0;  // prev_flat_grad.copy_(flat_grad)
    prev_loss = loss;

    float gsum = 0.f, gtd = 0.f, dmax = 0.f;
#pragma unroll 5
    for (int i = 0; i < n; ++i) {
      const float gi = v.at(og + i), di = v.at(od + i);
      gsum += fabsf(gi);
      gtd = fmaf(gi, di, gtd);
      dmax = fmaxf(dmax, fabsf(di));
    }
    if (n_iter == 1) {
      const float inv = 1.f / gsum;
      t_f32 = inv < 1.f;   // Python min(1.0, tensor) keeps the tensor only when it is smaller
      t = t_f32 ? (double)(inv * lr) : (double)lr;
    } else {
      t_f32 = false;
      t = (double)lr;
    }
    if ((double)gtd > -kTolChange) {
      done = true;
      return;
    }
    // _strong_wolfe prologue (lbfgs.py:43-56)
    d_norm = (double)dmax;
    f0 = loss;
    gtd0 = gtd;
    max_ls = max_eval - evals;
    ls_evals = 0;
    t_prev = 0.0;
    f_prev = f0;
    gtd_prev = gtd0;
    slot_prev = g0;
    ls_iter = 0;
    phase = 0;
    first_eval = true;
    ls_done = false;
    insuf = false;
    br_n = 0;
    set_trial(c, v);
  }

  // Process the evaluation at the current trial point (loss f_new, gradient in G).
  // Afterwards either `done` is set (parameters are in v.xk) or X holds the next trial point.
  K2B_HD void after_eval(const Cols& c, const Vecs& v, float f_new_f) {
    const double f_new = (double)f_new_f;
    const float gtd_new = dot_g_d(c, v);
    ++ls_evals;
    bool finished = false;

    if (phase == 0) {
      if (!first_eval) ++ls_iter;
      first_eval = false;
      if (ls_iter < max_ls) {
        const bool armijo_fail = (float)f_new > (float)(f0 + (double)((float)(kC1 * t) * gtd0));
        if (armijo_fail || (ls_iter > 1 && f_new >= f_prev) ) {
          make_bracket(c, v, f_new, gtd_new);
        } else if (fabsf(gtd_new) <= -(float)kC2 * gtd0) {
          br_n = 1;
          br_t[0] = t;
          br_f[0] = f_new;
          br_slot[0] = free_slot(slot_prev, -1);
          store_grad(c, v, br_slot[0]);
          br_gtd[0] = gtd_new;
          ls_done = true;
        } else if (gtd_new >= 0.f) {
          make_bracket(c, v, f_new, gtd_new);
        } else {
          // extrapolate (lbfgs.py:76-94)
          double min_step = t + 0.01 * (t - t_prev);
          double max_step = t * 10.0;
          if (t_f32) {  // tensor arithmetic: float32(0.01) * (t - t_prev), t * float32(10)
            const float tf = (float)t;
            min_step = (double)(tf + 0.01f * (tf - (float)t_prev));
            max_step = (double)(tf * 10.f);
          }
          const double tmp = t;
          t = cubic_interpolate(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, true, min_step, max_step, t_f32);
          t_prev = tmp;
          f_prev = f_new;
          const int s = free_slot(-1, -1);  // old g_prev is dropped unless it is g0
          store_grad(c, v, s);
          slot_prev = s;
          gtd_prev = gtd_new;
          set_trial(c, v);
          return;  // evaluate the extrapolated point
        }
      } else {
        // reached max_ls in the bracket phase (lbfgs.py:96-100)
        br_n = 2;
        br_t[0] = 0.0; br_t[1] = t;
        br_f[0] = f0;  br_f[1] = f_new;
        br_slot[0] = g0;
        br_slot[1] = free_slot(-1, -1);
        store_grad(c, v, br_slot[1]);
        br_gtd[0] = gtd0; br_gtd[1] = gtd_new;
      }
      phase = 1;
      low = (br_f[0] <= br_f[br_n - 1]) ? 0 : 1;
      high = 1 - low;
      if (br_n == 1) low = 0;
    } else {
      // zoom: bracket update after evaluating t (lbfgs.py:163-203)
      ++ls_iter;
      const bool armijo_fail = (float)f_new > (float)(f0 + (double)((float)(kC1 * t) * gtd0));
      if (armijo_fail || f_new >= br_f[low]) {
        const int s = (br_slot[high] != g0) ? br_slot[high] : free_slot(br_slot[low], -1);
        store_grad(c, v, s);
        br_t[high] = t; br_f[high] = f_new; br_slot[high] = s; br_gtd[high] = gtd_new;
        low = (br_f[0] <= br_f[1]) ? 0 : 1;
        high = 1 - low;
      } else {
        if (fabsf(gtd_new) <= -(float)kC2 * gtd0) {
          ls_done = true;
        } else if ((double)gtd_new * (br_t[high] - br_t[low]) >= 0.0) {
          br_t[high] = br_t[low]; br_f[high] = br_f[low];
          br_slot[high] = br_slot[low]; br_gtd[high] = br_gtd[low];
        }
        // the new point becomes the low end; it may reuse any slot that is neither g0 nor
        // the (possibly just reassigned) high end
        const int s = free_slot(br_slot[high], -1);
        store_grad(c, v, s);
        br_t[low] = t; br_f[low] = f_new; br_slot[low] = s; br_gtd[low] = gtd_new;
      }
    }

    // zoom loop head (lbfgs.py:108-160): propose the next trial or stop
    if (!ls_done && ls_iter < max_ls && br_n == 2) {
      const double width = fabs(br_t[1] - br_t[0]);
      if (!(width * d_norm < kTolChange)) {
        double tn = cubic_interpolate(br_t[0], br_f[0], br_gtd[0], br_t[1], br_f[1], br_gtd[1], false, 0.0, 0.0, t_f32);
        const double bmax = br_t[0] > br_t[1] ? br_t[0] : br_t[1];
        const double bmin = br_t[0] > br_t[1] ? br_t[1] : br_t[0];
        const double eps = 0.1 * (bmax - bmin);
        const double da = bmax - tn, db = tn - bmin;
        if ((da < db ? da : db) < eps) {
          if (insuf || tn >= bmax || tn <= bmin) {
            tn = (fabs(tn - bmax) < fabs(tn - bmin)) ? bmax - eps : bmin + eps;
            insuf = false;
          } else {
            insuf = true;
          }
        } else {
          insuf = false;
        }
        t = tn;
        set_trial(c, v);
        return;  // evaluate the zoom point
      }
    }
    finished = true;

    if (finished) {
      // lbfgs.py:205-209, 488-526
      t = br_t[low];
      loss = br_f[low];
      if (replay) {
        ls_replay_finished = true;
        return;
      }
      g0_next(c, v, br_slot[low]);
    }
  }

  // ---- test hook: run only the line search on a 1-D surrogate (tests/host_emul) ----------
  bool ls_replay_finished;
  K2B_HD void ls_replay_begin(const Cols& c, const Vecs& v, double t0, double f0_, float gtd0_, double d_norm_,
                              int max_ls_, bool t_is_f32) {
    max_iter = 1; max_eval = max_ls_ + 1; lr = 1.f;
    n_iter = 1; evals = 1; num_old = 0; head = 0; done = false;
    loss = f0_; prev_loss = f0_; t = t0; H_diag = 1.f; g0 = 0; slot_prev_grad = 0;
    v.at(v.xk()) = 0.f; v.at(v.d()) = 1.f; v.at(v.gslot(0)) = gtd0_;
    d_norm = d_norm_; f0 = f0_; gtd0 = gtd0_; max_ls = max_ls_; ls_evals = 0;
    t_prev = 0.0; f_prev = f0_; gtd_prev = gtd0_; slot_prev = 0; ls_iter = 0; phase = 0;
    first_eval = true; ls_done = false; insuf = false; br_n = 0; ls_replay_finished = false;
    replay = true;
    t_f32 = t_is_f32;
    set_trial(c, v);
  }
  bool replay = false;

 private:
  int slot_prev_grad;  // slot holding prev_flat_grad

  K2B_HD void make_bracket(const Cols& c, const Vecs& v, double f_new, float gtd_new) {
    br_n = 2;
    br_t[0] = t_prev; br_t[1] = t;
    br_f[0] = f_prev; br_f[1] = f_new;
    br_slot[0] = slot_prev;
    br_slot[1] = free_slot(slot_prev, -1);
    store_grad(c, v, br_slot[1]);
    br_gtd[0] = gtd_prev; br_gtd[1] = gtd_new;
  }

  // Line search returned: move the iterate, account evaluations, test termination,
  // and either stop or start the next outer iteration.
  K2B_HD void g0_next(const Cols& c, const Vecs& v, int new_g_slot) {
    const int n = v.n, od = v.d();
    const float tf = (float)t;
    float dtmax = 0.f;
#pragma unroll 5
    for (int i = 0; i < n; ++i) {
      const float di = v.at(od + i);
      v.at(i) = fmaf(tf, di, v.at(i));      // _add_grad(t, d)
      dtmax = fmaxf(dtmax, fabsf(di * tf));
    }
    evals += ls_evals;
    // keep prev_flat_grad readable for the next y = g - prev_g: it stays in slot_prev_grad
    g0 = new_g_slot;
    float gmax = 0.f;
    const int og = v.gslot(g0);
#pragma unroll 5
    for (int i = 0; i < n; ++i) gmax = fmaxf(gmax, fabsf(v.at(og + i)));
    if (n_iter == max_iter || evals >= max_eval || (double)gmax <= kTolGrad ||
        (double)dtmax <= kTolChange || fabs(loss - prev_loss) < kTolChange) {
      done = true;
      return;
    }
    // prev_flat_grad (slot_prev_grad) is consumed by start_outer before any store of the next
    // line search, and free_slot never hands out g0.
    start_outer(c, v);
  }

};

}  // namespace k2b
