// warp_emul.cu -- DEBUGGING HARNESS, NOT PART OF THE PRODUCT.
//
// Runs the warp-cooperative fitting code (keypoints2body_b200/csrc/chain_core.cuh) on the CPU: the 32
// lanes of a warp are 32 coroutines (ucontext), scheduled round-robin; a shuffle publishes the lane's
// value, yields once around the ring and reads the source lane's value; a warp sync is one trip around
// the ring.  This is valid for warp-uniform control flow, which is what the device code has.  Built into
// tests/host_emul/libk2b_warp_emul.so by build.sh, loaded only by tests/test_warp_emul.py.
#define K2B_WARP_EMUL 1
#include <ucontext.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../keypoints2body_b200/csrc/chain_core.cuh"

using namespace k2b;

namespace {
constexpr int kMaxWarps = 16;
constexpr int kMaxLanes = 32 * kMaxWarps;
constexpr size_t kStack = 1 << 20;
ucontext_t g_ctx[kMaxLanes], g_main;
char* g_stacks = nullptr;
int g_lane = 0, g_done = 0, kLanes = 32;     // kLanes: coroutines of the current run (32 per emulated warp)
float g_buf[2][kMaxLanes];
int g_par[kMaxLanes];
int g_bar_cnt[16], g_bar_gen[16];
void (*g_job)() = nullptr;

void yield_next() {
  const int cur = g_lane, nxt = (cur + 1) % kLanes;
  g_lane = nxt;
  swapcontext(&g_ctx[cur], &g_ctx[nxt]);
  g_lane = cur;
}
void lane_entry() {
  g_job();
  const int me = g_lane;
  ++g_done;
  while (true) {
    if (g_done == kLanes) swapcontext(&g_ctx[me], &g_main);
    yield_next();
  }
}
// runs `job` on `warps` emulated warps (32 coroutines each) that share named barriers
void run_warps(void (*job)(), int warps) {
  if (!g_stacks) g_stacks = (char*)malloc(kStack * kMaxLanes);
  kLanes = 32 * warps;
  g_job = job;
  g_done = 0;
  for (int i = 0; i < 16; ++i) g_bar_cnt[i] = g_bar_gen[i] = 0;
  for (int l = 0; l < kLanes; ++l) {
    g_par[l] = 0;
    getcontext(&g_ctx[l]);
    g_ctx[l].uc_stack.ss_sp = g_stacks + kStack * l;
    g_ctx[l].uc_stack.ss_size = kStack;
    g_ctx[l].uc_link = nullptr;
    makecontext(&g_ctx[l], lane_entry, 0);
  }
  g_lane = 0;
  swapcontext(&g_main, &g_ctx[0]);
}
void run_warp(void (*job)()) { run_warps(job, 1); }
int emul_warp() { return g_lane >> 5; }
}  // namespace

float* k2b_emul_trace = nullptr;
int k2b_emul_trace_cap = 0;
int* k2b_emul_trace_n = nullptr;
long k2b_emul_rounds = 0;
extern "C" long wemu_rounds(int reset) {
  const long r = k2b_emul_rounds;
  if (reset) k2b_emul_rounds = 0;
  return r;
}
// rows of (t, f, g.d) for every line-search trial of the next wemu_chain call(s); pass nulls to stop tracing
extern "C" void wemu_set_trace(float* rows, int cap, int* count) {
  k2b_emul_trace = rows;
  k2b_emul_trace_cap = cap;
  k2b_emul_trace_n = count;
}

int k2b_emul_lane() { return g_lane & 31; }
float k2b_emul_shfl(float v, int src) {
  const int me = g_lane, p = g_par[me];
  g_buf[p][me] = v;
  g_par[me] = p ^ 1;
  yield_next();
  return g_buf[p][(me & ~31) + src];
}
void k2b_emul_sync() { yield_next(); }
// bar.arrive / bar.sync among the emulated warps: a generation barrier counted in threads
void k2b_emul_bar(int id, int threads, int blocking) {
  const int gen = g_bar_gen[id];
  if (++g_bar_cnt[id] == threads) {
    g_bar_cnt[id] = 0;
    ++g_bar_gen[id];
  } else if (blocking) {
    while (g_bar_gen[id] == gen) yield_next();
  }
}

struct WEmuModel {
  int ns;
  std::vector<float> P, mu, nlw, rel;
};

// chol69: [8][69][69] lower-triangular L; the precision tables are rebuilt as L L^T in double, like
// k2b_model_create does for the device tables.
extern "C" void* wemu_model_create(int ns, const float* chol69, const float* means, const float* nlw,
                                   const double* J0, const double* JS, const int* parents, int nj) {
  WEmuModel* m = new WEmuModel();
  m->ns = ns;
  m->P.assign((size_t)kGmmM * wc::kPFloats, 0.f);
  m->mu.assign((size_t)kGmmM * kMuStride, 0.f);
  m->nlw.assign(nlw, nlw + kGmmM);
  for (int c = 0; c < kGmmM; ++c) {
    const float* L = chol69 + (size_t)c * kBodyDim * kBodyDim;
    for (int i = 0; i < kBodyDim; ++i) {
      for (int j = 0; j < kBodyDim; ++j) {
        double acc = 0.0;
        const int kmax = i < j ? i : j;
        for (int k = 0; k <= kmax; ++k) acc += (double)L[i * kBodyDim + k] * (double)L[j * kBodyDim + k];
        m->P[(size_t)c * wc::kPFloats + i * wc::kPStride + j] = (float)acc;
      }
      m->mu[(size_t)c * kMuStride + i] = means[(size_t)c * kBodyDim + i];
    }
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
extern "C" void wemu_model_destroy(void* m) { delete (WEmuModel*)m; }

namespace {
struct Job {
  const WEmuModel* m;
  wc::ChainParams p;
  int K;
  long seq;
  // evaluate-only
  bool eval_only;
  const float* x;      // [NX]
  float* grad;         // [NX]
  float* loss;
  float* joints;
  int* comp;
  std::vector<float> wmem, hist;
  int E = 1, H = 0;     // team: evaluators, helpers per evaluator
} g;

template <int NS, int K>
void lane_job() {
  wc::WarpTables tb{g.m->P.data(), g.m->mu.data(), g.m->nlw.data(), (const float4*)g.m->rel.data()};
  float* base = g.wmem.data();
  const int E = g.E, H = g.H, TW = E * (1 + H), member = emul_warp();
  wc::TeamMem tm = wc::make_team_mem(base, E, H, g.p.hmax > 0 ? g.p.hmax : 1);
  const int bar_b = E == 1 ? 1 : 2;          // the kernel's ids for team 0
  tm.bar_go = 1;
  tm.bar_tab = 4;
  tm.bar_done = 5;
  if (member >= E) {                          // helper warp
    const int e = (member - E) / H, h = (member - E) % H;
    float* own = base + (size_t)member * wc::kEvalMemFloats;
    wc::team_helper(tb, wc::make_warp_mem(own, nullptr), own, base + (size_t)e * wc::kEvalMemFloats, h, H, bar_b, 32 * TW);
    return;
  }
  wc::WarpMem wm = wc::make_warp_mem(base + (size_t)member * wc::kEvalMemFloats, tm.gs);
  wm.helpers = H;
  wm.bar_id = bar_b;
  wm.bar_threads = 32 * TW;
  wm.helper_mem = base + (size_t)(E + member * H) * wc::kEvalMemFloats;
  wm.helper_stride = wc::kEvalMemFloats;
  if (g.eval_only) {
    const int lane = wc::lane_id();
    float xr[3], gr[3];
    for (int c = 0; c < 3; ++c) xr[c] = 3 * lane + c < 75 + NS ? g.x[3 * lane + c] : 0.f;
    wc::FrameObs ob;
    ob.tx = ob.ty = ob.tz = ob.w = 0.f;
    if (lane < K) {
      ob.tx = g.p.targets[lane * 3]; ob.ty = g.p.targets[lane * 3 + 1]; ob.tz = g.p.targets[lane * 3 + 2];
      const float cf = g.p.conf ? g.p.conf[lane] : 1.f;
      ob.w = g.p.joint_w2 * cf * cf;
    }
    for (int c = 0; c < 3; ++c)
      ob.keep[c] = (lane >= 1 && lane < 24) ? g.p.preserve_pose[3 * lane - 3 + c] : 0.f;
    ob.keep_w2 = g.p.keep_w2;
    ob.plain_sq = false;
    ob.depth_w2 = 0.f;
    ob.dref[0] = ob.dref[1] = ob.dref[2] = 0.f;
    int comp = 0;
    const float loss = wc::eval_warp<NS, K>(tb, wm, ob, xr, true, true, gr, g.joints, &comp);
    for (int c = 0; c < 3; ++c)
      if (3 * lane + c < 75 + NS) g.grad[3 * lane + c] = gr[c];
    if (lane == 0) { *g.loss = loss; *g.comp = comp; }
    wc::team_release_helpers(wm);
    return;
  }
  if (g.p.camera_seq) {
    if (g.p.lbfgs) wc::run_evaluator<NS, K, true, true>(g.p, tb, wm, tm, member, g.seq, g.p.num_seq, g.hist.data());
    else wc::run_evaluator<NS, K, false, true>(g.p, tb, wm, tm, member, g.seq, g.p.num_seq, g.hist.data());
  } else if (g.p.lbfgs) wc::run_evaluator<NS, K, true>(g.p, tb, wm, tm, member, g.seq, g.p.num_seq, g.hist.data());      // one sequence per run
  else wc::run_evaluator<NS, K, false>(g.p, tb, wm, tm, member, g.seq, g.p.num_seq, g.hist.data());
}

void (*pick_job(int ns, int K))() {
  if (K == 24) return lane_job<10, 24>;
  return ns == 20 ? lane_job<20, 22> : lane_job<10, 22>;
}
}  // namespace

// One evaluation (loss, gradient, joints, arg-min component) of one frame; x layout [go3|body69|transl3|shape NS].
extern "C" int wemu_eval(void* model, int K, float joint_w, float keep_w, const float* targets, const float* conf,
                         const float* x, const float* keep_pose, float* out_grad, float* out_loss, float* out_joints,
                         int* out_comp) {
  g.m = (WEmuModel*)model;
  g.K = K;
  g.eval_only = true;
  g.p = wc::ChainParams{};
  g.p.targets = targets;
  g.p.conf = conf;
  g.p.joint_w2 = joint_w * joint_w;
  g.p.keep_w2 = keep_w * keep_w;
  g.p.preserve_pose = keep_pose;
  g.x = x; g.grad = out_grad; g.loss = out_loss; g.joints = out_joints; g.comp = out_comp;
  g.p.hmax = 1;
  g.E = 1;
  g.p.team = 1;
  g.p.helpers = g.H;
  g.wmem.assign(wc::team_floats(1, g.H, 1), 0.f);
  run_warps(pick_job(g.m->ns, K), 1 + g.H);
  return 0;
}

// Team of the following wemu_eval / wemu_chain calls: E evaluator warps (> 1: speculative line-search evaluation,
// L-BFGS chains only) and H helper warps per evaluator (mixture prior scanned by other warps).
// 1: the next wemu_chain calls run whole camera-space fits per frame (ChainParams::camera_seq)
static int g_camera_seq = 0;
extern "C" void wemu_set_camera_seq(int on) { g_camera_seq = on; }

extern "C" void wemu_set_team(int evaluators, int helpers) {
  g.E = evaluators < 1 ? 1 : evaluators;
  g.H = helpers < 0 ? 0 : helpers;
}

// S sequences of T frames each, walked by one (emulated) warp per sequence.
extern "C" int wemu_chain(void* model, int K, int S, int T, long first_seq_ind, int chain, int lbfgs, int iters_first,
                          int iters_follow, int freeze_betas, float lr, float joint_w, float keep_w, const float* targets,
                          const float* conf, int conf_mode, const float* init_pose, const float* init_betas,
                          const float* init_transl, const float* init_expr, const float* preserve_pose, float* out_pose,
                          float* out_betas, float* out_transl, float* out_expr, float* out_loss, float* out_joints,
                          int* out_evals, int loss_kind, int final_mode, float depth_weight, const float* depth_ref) {
  g.m = (WEmuModel*)model;
  g.K = K;
  g.eval_only = false;
  wc::ChainParams& p = g.p;
  p = wc::ChainParams{};
  p.num_seq = S; p.frames = T; p.first_seq_ind = first_seq_ind; p.chain = chain;
  p.in_seq_stride = T; p.out_seq_stride = T; p.out_frame_stride = 1;
  p.iters_first = iters_first; p.iters_follow = iters_follow; p.lbfgs = lbfgs; p.freeze_betas = freeze_betas;
  p.conf_mode = conf ? conf_mode : 0;
  p.lr = lr; p.joint_w2 = joint_w * joint_w; p.keep_w2 = keep_w * keep_w;
  p.targets = targets; p.conf = conf;
  p.init_pose = init_pose; p.init_betas = init_betas; p.init_transl = init_transl; p.init_expr = init_expr;
  p.preserve_pose = preserve_pose;
  p.loss_kind = loss_kind; p.final_mode = final_mode;
  p.depth_w2 = 4.f * depth_weight * depth_weight;   // the reference's (B,4,3)+(B,3) broadcast counts the depth term per joint row
  p.depth_ref = depth_ref;
  p.out_pose = out_pose; p.out_betas = out_betas; p.out_transl = out_transl; p.out_expr = out_expr;
  p.out_loss = out_loss; p.out_joints = out_joints; p.out_evals = out_evals;
  const int max_it = iters_first > iters_follow ? iters_first : iters_follow;
  p.hmax = lbfgs_history_capacity(max_it);
  for (int k = 1; k <= wc::kAdamTableW; ++k) {
    p.adam_step[k - 1] = (float)((double)lr / (1.0 - std::pow(0.9, (double)k)));
    p.adam_bc2[k - 1] = (float)std::sqrt(1.0 - std::pow(0.999, (double)k));
  }
  p.camera_seq = g_camera_seq;
  if (p.camera_seq) p.final_mode = 1;
  const int E = (lbfgs && !p.camera_seq) ? g.E : 1;
  p.team = E;
  p.helpers = g.H;
  const int keepE = g.E;
  g.E = E;
  g.wmem.assign(wc::team_floats(E, g.H, p.hmax), 0.f);
  g.hist.assign(wc::hist_floats(p.hmax), 0.f);
  p.hist = g.hist.data();
  for (long s = 0; s < S; ++s) {
    g.seq = s;
    run_warps(pick_job(g.m->ns, K), E * (1 + g.H));
  }
  g.E = keepE;
  return 0;
}

// Line-search replay through the lane-distributed vector policy (WarpOps): the same conformance hook as
// host_emul.cu's emu_linesearch_replay / the device's k2b_linesearch_replay(warp_policy = 1).
namespace {
struct ReplayJob {
  double t0, f0, d_norm;
  float gtd0;
  int max_ls, t_is_f32, n_resp;
  const double* resp_f;
  const float* resp_gtd;
  double* out_t;
  double* out_final;
  int k;
  std::vector<float> gs;
} r;

void replay_lane_job() {
  wc::WVec v;
  v.gs = r.gs.data();
  v.hist = nullptr;
  v.ro = nullptr;
  v.al = nullptr;
  v.hmax = 1;
  Lbfgs<85, wc::WarpOps> st;
  st.init();
  st.ls_replay_begin(v, v, r.t0, r.f0, r.gtd0, r.d_norm, r.max_ls, r.t_is_f32 != 0);
  int k = 0;
  while (!st.ls_replay_finished && k < r.n_resp) {
    if (wc::lane_id() == 0) r.out_t[k] = st.t;
    wc::WarpOps::replay_response(v, st.cur, r.resp_gtd[k]);
    st.after_eval(v, v, (float)r.resp_f[k]);
    ++k;
  }
  if (wc::lane_id() == 0) {
    r.out_final[0] = st.t;
    r.out_final[1] = st.loss;
    r.out_final[2] = (double)st.ls_evals;
    r.k = k;
  }
}
}  // namespace

extern "C" int wemu_linesearch_replay(double t0, double f0, float gtd0, double d_norm, int max_ls, int t_is_f32, int n_resp,
                                      const double* resp_f, const float* resp_gtd, double* out_t, double* out_final) {
  r.t0 = t0; r.f0 = f0; r.gtd0 = gtd0; r.d_norm = d_norm; r.max_ls = max_ls; r.t_is_f32 = t_is_f32; r.n_resp = n_resp;
  r.resp_f = resp_f; r.resp_gtd = resp_gtd; r.out_t = out_t; r.out_final = out_final; r.k = 0;
  r.gs.assign(4 * wc::kWarpVec, 0.f);
  run_warp(replay_lane_job);
  return r.k;
}
