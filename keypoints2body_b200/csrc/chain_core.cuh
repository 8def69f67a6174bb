// chain_core.cuh -- warp-cooperative evaluation and optimiser steps: one WARP owns one frame.
//
// The fused fitting kernel (fit_kernel.cuh) maps one thread to one frame: best throughput for
// thousands of independent frames, but a lone frame takes ~1.3 ms, and the reference's default
// sequence schedule (api/sequence.py:214-281: frame t starts from frame t-1's result) is serial in
// t.  Here the 32 lanes of a warp share ONE frame, so an evaluation takes microseconds and a whole
// sequence can be walked serially inside one launch (chain_kernel.cuh).
//
// Same maths as fit_core.cuh (paths into /root/reference/keypoints2body):
//   core/fitters/world_space.py:173-212  compute_loss        core/losses.py:6-67  gmof, angle_prior, loss
//   core/prior.py:182-195                MaxMixturePrior     smplx lbs (joint branch) [smplx-from-memory]
// and the same L-BFGS machine (lbfgs_core.cuh) through a lane-distributed vector policy (WarpOps).
//
// Ownership: element e of the parameter vector [go 3 | body 69 | transl 3 | shape NS] belongs to lane
// e / 3, register e % 3 -- lane j < 24 owns the rotation of joint j, lane 24 the translation, lanes
// 25.. the shape coefficients.
//   kinematic tree   lane = joint; world transforms by pointer doubling towards the root (3-4 rounds of
//                    shuffles), subtree sums of the world-frame backward by suffix scans along the
//                    first-child chains plus the four side-branch additions;
//   GMM prior        y = P_m d with the symmetric precision P_m = L_m L_m^T in shared memory: 27 lanes
//                    as 3 row groups x 9 column chunks of 8; q_m = d.y needs no cross-group exchange,
//                    only the arg-min component's y is combined (its gradient is P d itself).
// Host build (tests/host_emul/warp_emul.cu, a debugging harness): the lanes run as 32 coroutines and
// the shuffle / sync hooks below switch between them.
#pragma once

#include "fit_core.cuh"
#include "lbfgs_core.cuh"

#if !defined(__CUDA_ARCH__) && defined(K2B_WARP_EMUL)
int k2b_emul_lane();
float k2b_emul_shfl(float v, int src);
void k2b_emul_sync();
void k2b_emul_bar(int id, int threads, int blocking);   // named barrier among the emulated warps (lanes are coroutines)
// optional line-search trace of the emulated warp: rows of (t, f, g.d) per trial evaluation (tests only)
extern float* k2b_emul_trace;
extern int k2b_emul_trace_cap;
extern int* k2b_emul_trace_n;
extern long k2b_emul_rounds;      // team rounds (serial evaluation steps) of the emulated leader
#endif

// Code size matters here: eval_warp is ~2 k instructions, and a kernel that inlines it at every use (Adam loop, final
// forward, L-BFGS trial, team sibling, evaluate-only ...) outgrows the instruction cache and runs at half speed (measured:
// 22 k instructions, 2.3x slower; out of line it was no better, the arguments then live in local memory).  So every
// evaluator warp runs ONE loop with ONE call site (run_evaluator): what differs between the modes is how the next
// point is chosen and where the result goes.

namespace k2b {
namespace wc {

constexpr int kPStride = 72;                       // row stride of a precision matrix (69 + 3 zero columns)
constexpr int kPFloats = kBodyDim * kPStride;      // per component
constexpr int kWarpVec = 96;                       // 32 lanes x 3 elements
constexpr int kDbufStride = 80;                    // one staged difference vector (72 used)
constexpr int kEvalMemFloats = 96 + 2 * kDbufStride + 80 + 256;   // per-warp block: xs, dbuf x2, ybuf, qbuf
constexpr int kMaxCand = 6;                        // candidate steps a team evaluates per round (<= evaluators)

K2B_HD int lane_id() {
#if defined(__CUDA_ARCH__)
  return threadIdx.x & 31;
#elif defined(K2B_WARP_EMUL)
  return k2b_emul_lane();
#else
  return 0;
#endif
}
K2B_HD float shfl(float v, int src) {
#if defined(__CUDA_ARCH__)
  return __shfl_sync(0xffffffffu, v, src);
#elif defined(K2B_WARP_EMUL)
  return k2b_emul_shfl(v, src & 31);
#else
  return v;
#endif
}
K2B_HD void wsync() {
#if defined(__CUDA_ARCH__)
  __syncwarp();
#elif defined(K2B_WARP_EMUL)
  k2b_emul_sync();
#endif
}
K2B_HD float wsum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += shfl(v, lane_id() ^ m);
  return v;
}
K2B_HD float wmax(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = fmaxf(v, shfl(v, lane_id() ^ m));
  return v;
}

// SMPL body tree (identical in SMPL / SMPL-H / SMPL-X for the first 22 joints; 22, 23 = SMPL hands)
K2B_HD int tree_parent(int j) {
  if (j <= 3) return 0;
  if (j <= 12) return j - 3;
  if (j <= 14) return 9;
  if (j <= 17) return j - 3;
  return j - 2;
}
K2B_HD int tree_depth(int j) {
  // 0 | 1 1 1 | 2 2 2 | 3 3 3 | 4 4 4 4 4 | 5 5 5 | 6 6 | 7 7 | 8 8
  const unsigned long long lo = 0x5444443332221110ull, hi = 0x88776655ull;
  return (int)(((j < 16 ? lo >> (4 * j) : hi >> (4 * (j - 16)))) & 15ull);
}
template <int NJ>
K2B_HD int tree_first_child(int j) {
  int c = -1;
  if (j == 0) c = 1;
  else if (j <= 9) c = j + 3;
  else if (j >= 12 && j <= 14) c = j + 3;
  else if (j >= 16 && j <= 21) c = j + 2;
  return c < NJ ? c : -1;
}

struct WarpTables {    // shared by every warp of a CTA (shared memory on the device)
  const float* P;      // [8][69][kPStride] symmetric precisions
  const float* mu;     // [8][kMuStride]
  const float* nlw;    // [8]
  const float4* rel;   // [24][1 + NS]
};
struct WarpMem {       // per-warp shared memory, kWarpMemFloats floats
  float* xs;           // [96] evaluation point, readable by every lane
  float* dbuf;         // [2][kDbufStride] x_body - mu_m, double-buffered over the components
  float* ybuf;         // [72] P d of the arg-min component
  float* qbuf;         // [8][32] per-lane partial sums of d.P_m d
  float* gs;           // [4][96] gradient slots (L-BFGS; the sequence's leading warp only, in the team area)
  // GMM delegation: `helpers` > 0 means other warps scan the mixture components for this evaluator while it walks the
  // kinematic tree.  Named barriers bar_id (points posted) and bar_id + 1 (results ready) are shared by every evaluator
  // and helper of the team (bar_threads threads in all): the evaluators of a round run the same code, so they post
  // and collect together.  Helper h's shared-memory block starts at helper_mem + h * helper_stride.
  int helpers;
  int bar_id;
  int bar_threads;
  float* helper_mem;
  int helper_stride;
};
constexpr int kYbufOff = 96 + 2 * kDbufStride;
K2B_HD WarpMem make_warp_mem(float* w, float* gs) {
  return WarpMem{w, w + 96, w + kYbufOff, w + kYbufOff + 80, gs, 0, 0, 32, nullptr, 0};
}
K2B_HD void bar_arrive(int id, int threads) {
#if defined(__CUDA_ARCH__)
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
#elif defined(K2B_WARP_EMUL)
  k2b_emul_bar(id, threads, 0);
#endif
}
K2B_HD void bar_sync(int id, int threads) {
#if defined(__CUDA_ARCH__)
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
#elif defined(K2B_WARP_EMUL)
  k2b_emul_bar(id, threads, 1);
#endif
}

// ---- a sequence's team --------------------------------------------------------------------------------------------
// One sequence is served by E evaluator warps (the first one leads: it owns the iterate and runs the L-BFGS machine)
// and E x H helper warps (H per evaluator, scanning the mixture components).  With E > 1 the line search is evaluated
// speculatively: torch's bracket phase almost always walks t -> 10 t (or the mid-point 5.505 t) -> 10x that
// (lbfgs.py:76-94), so while the leader evaluates the step the machine asked for, the other evaluators evaluate the
// steps it would ask for next; when the machine then asks for one of them, bit for bit, the result is already there.
// Evaluations are pure functions of the point, so the fit is bit-identical to the one-evaluator kernel; only the
// number of serial rounds per frame drops (about 13 -> 7 at the reference's 10-iteration budget).
// Team shared memory: [E (1 + H) warp blocks of kEvalMemFloats][team area].
struct TeamMem {
  float* gs;      // [4][96] gradient slots of the leader
  float* ro;      // [hmax]
  float* al;      // [hmax]
  float* xk;      // [96] base point of the posted candidates (the iterate; the point itself when `point_is_base`)
  float* d;       // [96] search direction
  float* tval;    // [8] candidate steps
  int* cmd;       // [8] kind, n_c, flags, row lo, row hi
  float* res_f;   // [8] losses of candidates 1..
  float* res_g;   // [E][96] gradients of candidates 1..
  float* lu;      // [8][72] u_m = P_m (xk - mu_m)      (LineEval)
  float* lw;      // [8][72] w_m = P_m d
  float* labc;    // [8][4]  A_m, B_m, C_m
  int E, H;
  int bar_go, bar_tab, bar_done;   // named barriers: round posted / line tables ready / results ready
};
K2B_HD constexpr int team_area_floats(int E, int hmax) {
  return 4 * kWarpVec + 2 * ((hmax + 3) & ~3) + 96 + 96 + 8 + 8 + 8 + E * kWarpVec + 2 * kGmmM * kPStride + 4 * kGmmM;
}
K2B_HD constexpr int team_floats(int E, int H, int hmax) { return E * (1 + H) * kEvalMemFloats + team_area_floats(E, hmax); }
K2B_HD TeamMem make_team_mem(float* team_base, int E, int H, int hmax) {
  float* a = team_base + E * (1 + H) * kEvalMemFloats;
  const int hp = (hmax + 3) & ~3;
  TeamMem t;
  t.gs = a;
  t.ro = a + 4 * kWarpVec;
  t.al = t.ro + hp;
  t.xk = t.al + hp;
  t.d = t.xk + 96;
  t.tval = t.d + 96;
  t.cmd = reinterpret_cast<int*>(t.tval + 8);
  t.res_f = t.tval + 16;
  t.res_g = t.tval + 24;
  t.lu = t.res_g + E * kWarpVec;
  t.lw = t.lu + kGmmM * kPStride;
  t.labc = t.lw + kGmmM * kPStride;
  t.E = E;
  t.H = H;
  t.bar_go = 0;
  t.bar_tab = 0;
  t.bar_done = 0;
  return t;
}
enum { kCmdExit = 0, kCmdEval = 1 };
enum { kFlagGrad = 1, kFlagPriors = 2, kFlagKeep = 4, kFlagFinal = 8, kFlagBase = 16 };
// what a round does to the line tables before evaluating: nothing (same line search continues), rebuild u at the
// posted point (a frame's first evaluation), follow the iterate and take the new direction, or follow the iterate only
enum { kTabKeep = 0, kTabRefresh = 1, kTabLine = 2, kTabPoint = 3 };

// ---- GMM prior pieces ----------------------------------------------------------------------------
// Lane -> (row group rg, column chunk cc): lanes 0-7 / 8-15 / 16-23 are row groups 0 / 1 / 2 with chunks 0..7,
// lanes 24-26 are chunk 8 of the three groups.  Chunk cc covers columns 4cc..4cc+3 and 36+4cc..36+4cc+3, row
// group rg the rows 3 it + rg: every quarter-warp of an LDS.128 then touches 32 distinct banks.
K2B_HD void gmm_stage(const WarpTables& tb, float* dbuf, const float (&xr)[3], bool body_owner, int i0, int m) {
  if (body_owner) {
    const float* mum = tb.mu + m * kMuStride + i0;
#pragma unroll
    for (int c = 0; c < 3; ++c) dbuf[i0 + c] = xr[c] - mum[c];
  }
}
K2B_HD void gmm_rows(const WarpTables& tb, const float* dbuf, int m, int rg, int cc, float2 (&acc)[4]) {
  const float* row = tb.P + (size_t)m * kPFloats + rg * kPStride + 4 * cc;
  const float* dj = dbuf + rg;
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll 23
  for (int it = 0; it < 23; ++it, row += 3 * kPStride, dj += 3) {
    const float d1 = *dj;
    const float2 d2 = make_float2(d1, d1);
    const float4 l0 = *reinterpret_cast<const float4*>(row);
    const float4 l1 = *reinterpret_cast<const float4*>(row + 36);
    acc[0] = fma2(make_float2(l0.x, l0.y), d2, acc[0]);
    acc[1] = fma2(make_float2(l0.z, l0.w), d2, acc[1]);
    acc[2] = fma2(make_float2(l1.x, l1.y), d2, acc[2]);
    acc[3] = fma2(make_float2(l1.z, l1.w), d2, acc[3]);
  }
}

// Scans the components m0, m0 + step, ... of the max-mixture prior at the point whose owned elements are xr:
// best = min(0.5 d.P_m d + nlw_m) (first minimum wins, like torch.min), bm its component; with want_y the
// product P d of that component (its gradient) is left in wm.ybuf.  Uses wm.dbuf / qbuf / ybuf.
K2B_HD void gmm_scan(const WarpTables& tb, const WarpMem& wm, const float (&xr)[3], int m0, int step, bool want_y,
                     float& best, int& bm) {
  const int lane = lane_id();
  const bool body_owner = lane >= 1 && lane < 24;     // elements 3 .. 71
  const int i0 = body_owner ? 3 * lane - 3 : 0;       // body-pose index of xr[0]
  const int rg = lane < 24 ? lane >> 3 : lane - 24;
  const int cc = lane < 24 ? lane & 7 : 8;
  const bool act = lane < 27;
  float2 acc[4];
  if (lane < 6) wm.dbuf[(lane / 3) * kDbufStride + 69 + lane % 3] = 0.f;   // padding columns of both buffers
  if (step > 1) {
#pragma unroll
    for (int m = 0; m < kGmmM; ++m) wm.qbuf[m * 32 + lane] = 0.f;
  }
  gmm_stage(tb, wm.dbuf, xr, body_owner, i0, m0);
  int par = 0;
#pragma unroll 1
  for (int m = m0; m < kGmmM; m += step, par ^= 1) {
    wsync();           // d_m is staged; every lane is done with the buffer the next d goes to
    float* dm = wm.dbuf + par * kDbufStride;
    if (m + step < kGmmM) gmm_stage(tb, wm.dbuf + (par ^ 1) * kDbufStride, xr, body_owner, i0, m + step);
    float part = 0.f;
    if (act) {
      gmm_rows(tb, dm, m, rg, cc, acc);
      const float4 e0 = *reinterpret_cast<const float4*>(dm + 4 * cc);
      const float4 e1 = *reinterpret_cast<const float4*>(dm + 36 + 4 * cc);
      part = fmaf(acc[0].x, e0.x, fmaf(acc[0].y, e0.y, fmaf(acc[1].x, e0.z, acc[1].y * e0.w))) +
             fmaf(acc[2].x, e1.x, fmaf(acc[2].y, e1.y, fmaf(acc[3].x, e1.z, acc[3].y * e1.w)));
    }
    wm.qbuf[m * 32 + lane] = part;
  }
  wsync();
  // q_m = sum of the 32 partials of component m: lanes 4m..4m+3 add 8 each, two butterfly steps finish;
  // then the arg-min over the components by three more steps
  bm = lane >> 2;
  {
    const float4 a = *reinterpret_cast<const float4*>(wm.qbuf + 8 * lane);
    const float4 b = *reinterpret_cast<const float4*>(wm.qbuf + 8 * lane + 4);
    float q = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
    q += shfl(q, lane ^ 1);
    q += shfl(q, lane ^ 2);
    best = (bm - m0) % step == 0 && bm >= m0 ? fmaf(0.5f, q, tb.nlw[bm]) : INFINITY;
#pragma unroll
    for (int msk = 4; msk <= 16; msk <<= 1) {
      const float ol = shfl(best, lane ^ msk);
      const int om = (int)shfl((float)bm, lane ^ msk);
      if (ol < best || (ol == best && om < bm)) {
        best = ol;
        bm = om;
      }
    }
  }
  if (want_y) {
    // y = P d of the arg-min component, recomputed (cheaper than keeping eight results alive), then summed
    // over the three row groups: lanes 0..7 and 24 hold the totals of their chunk
    gmm_stage(tb, wm.dbuf, xr, body_owner, i0, bm);
    wsync();
    if (act) gmm_rows(tb, wm.dbuf, bm, rg, cc, acc);
    else {
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = make_float2(0.f, 0.f);
    }
    float y[8] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y, acc[2].x, acc[2].y, acc[3].x, acc[3].y};
    const int p1 = lane < 24 ? (lane + 8) % 24 : 24 + (lane - 23) % 3;
    const int p2 = lane < 24 ? (lane + 16) % 24 : 24 + (lane - 22) % 3;
#pragma unroll
    for (int k = 0; k < 8; ++k) y[k] = (y[k] + shfl(y[k], p1)) + shfl(y[k], p2);
    if (act && rg == 0) {
      *reinterpret_cast<float4*>(wm.ybuf + 4 * cc) = make_float4(y[0], y[1], y[2], y[3]);
      *reinterpret_cast<float4*>(wm.ybuf + 36 + 4 * cc) = make_float4(y[4], y[5], y[6], y[7]);
    }
    wsync();
  }
}

// out[0..71] = P_m dvec for a vector staged in shared memory (72 floats, entries 69..71 zero).  Same row / column
// ownership as gmm_scan; the three row groups are combined by shuffles.
K2B_HD void gmm_matvec(const WarpTables& tb, const float* dvec, int m, float* out) {
  const int lane = lane_id();
  const int rg = lane < 24 ? lane >> 3 : lane - 24;
  const int cc = lane < 24 ? lane & 7 : 8;
  const bool act = lane < 27;
  float2 acc[4];
  if (act) gmm_rows(tb, dvec, m, rg, cc, acc);
  else {
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = make_float2(0.f, 0.f);
  }
  float y[8] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y, acc[2].x, acc[2].y, acc[3].x, acc[3].y};
  const int p1 = lane < 24 ? (lane + 8) % 24 : 24 + (lane - 23) % 3;
  const int p2 = lane < 24 ? (lane + 16) % 24 : 24 + (lane - 22) % 3;
#pragma unroll
  for (int k = 0; k < 8; ++k) y[k] = (y[k] + shfl(y[k], p1)) + shfl(y[k], p2);
  if (act && rg == 0) {
    *reinterpret_cast<float4*>(out + 4 * cc) = make_float4(y[0], y[1], y[2], y[3]);
    *reinterpret_cast<float4*>(out + 36 + 4 * cc) = make_float4(y[4], y[5], y[6], y[7]);
  }
  wsync();
}

// The mixture prior along a line.  On x(t) = xk + t d the Mahalanobis form of component m is the quadratic
//   q_m(t) = A_m + 2 t B_m + t^2 C_m,   A_m = a.u_m, B_m = d.u_m, C_m = d.w_m,   a = xk - mu_m, u_m = P_m a, w_m = P_m d,
// and its gradient is u_m + t w_m.  Every evaluation of an L-BFGS line search lies on one line, so the eight
// precision-matrix products are done once per line search (w_m; u_m follows the iterate: u_m += t_accepted w_m, and
// is recomputed from scratch at every frame's first evaluation) instead of once per evaluation, and an evaluation at
// any step t costs 8 fused multiply-adds for the arg-min and one axpy for the gradient.
struct LineEval {
  const float* u;     // [8][72]
  const float* w;     // [8][72]
  const float* abc;   // [8][4] A, B, C
  float t;
};

struct FrameObs {      // this lane's share of the frame's observations
  float tx, ty, tz, w; // lane j < K: target and weight joint_w^2 conf_j^2 of joint j
  float keep[3];       // preserve pose of the owned body-pose entries
  float keep_w2;       // pose_preserve_weight^2 or 0
  // camera-space stage 1 (camera_fitting_loss_3d, core/losses.py:70-93): plain squared joint error instead of GMoF,
  // plus depth_w2 * |transl - dref|^2 (lane 24 holds dref)
  bool plain_sq;
  float depth_w2;
  float dref[3];
};

K2B_HD Acc shfl_acc(const Acc& a, int src) {
  Acc r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.S.m[i] = shfl(a.S.m[i], src);
  r.s = v3(shfl(a.s.x, src), shfl(a.s.y, src), shfl(a.s.z, src));
  return r;
}

// One function evaluation of one frame by the whole warp.  xr: owned elements of the evaluation point
// (zero where 3*lane + c >= 75 + NS).  Returns the total loss in every lane; with_grad fills gr with the
// gradient of the owned elements (zero for unowned ones).  joints_out: global [K][3] or null.
// LINE: the caller always evaluates the mixture prior in line form (`le` is set whenever with_priors is: the L-BFGS
// instantiation of the kernel), so the component scan and the helper-warp handshake are compiled out.
template <int NS, int K, bool LINE = false>
K2B_HD float eval_warp(const WarpTables& tb, const WarpMem& wm, const FrameObs& ob, const float (&xr)[3],
                       bool with_grad, bool with_priors, float (&gr)[3], float* joints_out, int* gmm_component,
                       const LineEval* le = nullptr) {
  constexpr int NJ = (K == 24) ? 24 : 22;
  constexpr int MAXD = (K == 24) ? 8 : 7;
  const int lane = lane_id();
#pragma unroll
  for (int c = 0; c < 3; ++c) wm.xs[3 * lane + c] = xr[c];
  if (!LINE && with_priors && wm.helpers > 0 && !le) {   // hand the point to the helper warps (they read xs and dbuf[0])
    if (lane == 0) wm.dbuf[0] = with_grad ? 2.f : 1.f;
    bar_arrive(wm.bar_id, wm.bar_threads);
  }
  wsync();
  float shape[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) shape[s] = wm.xs[kShapeOff + s];
  const V3 transl = v3(wm.xs[kTranslOff], wm.xs[kTranslOff + 1], wm.xs[kTranslOff + 2]);
  float lsum = 0.f;    // this lane's share of the loss
  float uni = 0.f;     // terms every lane computes identically
  if (with_priors) {   // shape prior on betas only (losses.py:56)
    float acc = 0.f;
#pragma unroll
    for (int s = 0; s < 10; ++s) acc = fmaf(shape[s], shape[s], acc);
    uni = kShapePriorW2 * acc;
  }

  // ---- kinematic tree, lane = joint ------------------------------------------------------------
  const bool isj = lane < NJ;
  const int j = isj ? lane : 0;
  const int par = tree_parent(j);
  const int c0 = isj ? tree_first_child<NJ>(j) : -1;
  V3 rel;
  {
    const float4* e = tb.rel + j * (1 + NS);
    const float4 r0 = e[0];
    float x = r0.x, y = r0.y, z = r0.z;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const float4 d = e[1 + s];
      x = fmaf(d.x, shape[s], x);
      y = fmaf(d.y, shape[s], y);
      z = fmaf(d.z, shape[s], z);
    }
    rel = v3(x, y, z);
  }
  Rod o;
  const V3 r = v3(xr[0], xr[1], xr[2]);
  const M3 R = rodrigues(r, o);
  // World transforms by pointer doubling: after round r a lane holds the transform from its ancestor 2^(r+1)
  // levels up (exclusive) down to itself, so 3 rounds (4 for the SMPL hands at depth 8) of 12 shuffles reach the
  // root, instead of one round per tree level.  The products associate differently from a root-to-leaf walk
  // (rounding-level differences, ~1e-7 m on the joints).
  M3 Rw = R;
  V3 t = rel;
  {
    int up = (isj && j > 0) ? par : -1;          // ancestor the accumulated transform hangs from; -1: reached the root
#pragma unroll
    for (int r = 0; r < ((MAXD > 7) ? 4 : 3); ++r) {
      const int src = up >= 0 ? up : lane;
      M3 Ru;
#pragma unroll
      for (int i = 0; i < 9; ++i) Ru.m[i] = shfl(Rw.m[i], src);
      const V3 tu = v3(shfl(t.x, src), shfl(t.y, src), shfl(t.z, src));
      const int upu = (int)shfl((float)up, src);
      if (up >= 0) {
        t = matvec(Ru, t) + tu;
        Rw = matmul(Ru, Rw);
        up = upu;
      }
    }
  }
  M3 Rp;                                          // the parent's world rotation (identity for the root)
#pragma unroll
  for (int i = 0; i < 9; ++i) Rp.m[i] = shfl(Rw.m[i], par);
  if (!isj || j == 0) Rp = eye3();

  // ---- residuals (gmof, losses.py:6-10) ----------------------------------------------------------
  V3 g = v3(0.f, 0.f, 0.f);
  if (lane < K) {
    const V3 p = t + transl;
    if (joints_out) {
      joints_out[3 * lane + 0] = p.x;
      joints_out[3 * lane + 1] = p.y;
      joints_out[3 * lane + 2] = p.z;
    }
    const float ex = p.x - ob.tx, ey = p.y - ob.ty, ez = p.z - ob.tz;
    if (ob.plain_sq) {
      lsum = ob.w * fmaf(ex, ex, fmaf(ey, ey, ez * ez));
      g = v3(2.f * ob.w * ex, 2.f * ob.w * ey, 2.f * ob.w * ez);
    } else {
      const float ix = fdiv(1.f, kSigma2 + ex * ex), iy = fdiv(1.f, kSigma2 + ey * ey), iz = fdiv(1.f, kSigma2 + ez * ez);
      const float gx = kSigma2 * ex * ex * ix, gy = kSigma2 * ey * ey * iy, gz = kSigma2 * ez * ez * iz;
      lsum = ob.w * ((gx + gy) + gz);
      const float c2w = 2.f * kSigma2 * kSigma2 * ob.w;
      g = v3(c2w * ex * ix * ix, c2w * ey * iy * iy, c2w * ez * iz * iz);
    }
  }
  gr[0] = gr[1] = gr[2] = 0.f;

  // ---- world-frame backward: subtree sums child -> parent, then per-joint gradients ---------------
  if (with_grad) {
    Acc a{zero3(), v3(0.f, 0.f, 0.f)};
    if (lane < K) acc_point(a, g, t);
    // Subtree sums.  Following first children the tree is five chains (0-1-4-7-10, 2-5-8-11, 3-6-9-12-15 and the
    // two arms from 13 / 14); three pointer-doubling rounds give every joint the sum over the rest of its chain,
    // then the arm chains are added to 9, 6, 3 and the chains headed by 2 and 3 to the root: 7 exchanges of 12
    // values instead of one per tree level plus the side branches (11).
    {
      int nx = c0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int src = nx >= 0 ? nx : lane;
        const Acc v = shfl_acc(a, src);
        const int nxn = (int)shfl((float)nx, src);
        if (nx >= 0) {
          acc_add(a, v);
          nx = nxn;
        }
      }
      const bool spine = lane == 3 || lane == 6 || lane == 9;
      Acc v = shfl_acc(a, 13);
      if (spine) acc_add(a, v);
      v = shfl_acc(a, 14);
      if (spine) acc_add(a, v);
      v = shfl_acc(a, 2);
      if (lane == 0) acc_add(a, v);
      v = shfl_acc(a, 3);
      if (lane == 0) acc_add(a, v);
    }
    V3 rb = v3(0.f, 0.f, 0.f);           // a leaf's own rotation moves nothing observed
    V3 drel = v3(0.f, 0.f, 0.f);
    if (isj) {
      if (c0 >= 0) rb = rodrigues_bwd(rot_grad(a, Rp, Rw, t), r, o);
      drel = matvec_t(Rp, a.s);         // d loss / d rel_j
    }
    float shape_bar[NS];
    {
      const float4* e = tb.rel + j * (1 + NS);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float4 d = e[1 + s];
        shape_bar[s] = wsum(fmaf(d.x, drel.x, fmaf(d.y, drel.y, d.z * drel.z)));
      }
    }
    const V3 s0 = v3(shfl(a.s.x, 0), shfl(a.s.y, 0), shfl(a.s.z, 0));
    if (lane < 24) {
      gr[0] = rb.x; gr[1] = rb.y; gr[2] = rb.z;
    } else if (lane == 24) {
      gr[0] = s0.x; gr[1] = s0.y; gr[2] = s0.z;
    } else {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float v = (with_priors && s < 10) ? fmaf(2.f * kShapePriorW2, shape[s], shape_bar[s]) : shape_bar[s];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (3 * lane + c == kShapeOff + s) gr[c] = v;
      }
    }
  }

  if (ob.depth_w2 != 0.f && lane == 24) {   // camera-space stage 1: keep the translation near its initial estimate
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = xr[c] - ob.dref[c];
      lsum = fmaf(ob.depth_w2 * d, d, lsum);
      if (with_grad) gr[c] = fmaf(2.f * ob.depth_w2, d, gr[c]);
    }
  }

  // ---- priors on the body pose -----------------------------------------------------------------
  if (with_priors) {
    const bool body_owner = lane >= 1 && lane < 24;     // elements 3 .. 71
    const int i0 = body_owner ? 3 * lane - 3 : 0;       // body-pose index of xr[0]
    float best = INFINITY;
    int bm = 0;
    const float* ysrc = wm.ybuf;
    const float* yw = nullptr;          // line form: gradient = ysrc + t yw
    if (LINE || le) {
      // every lane evaluates the eight quadratics (first minimum wins, like torch.min)
#pragma unroll
      for (int m = 0; m < kGmmM; ++m) {
        const float q = fmaf(le->t, fmaf(le->t, le->abc[4 * m + 2], 2.f * le->abc[4 * m + 1]), le->abc[4 * m]);
        const float ll = fmaf(0.5f, q, tb.nlw[m]);
        if (ll < best) {
          best = ll;
          bm = m;
        }
      }
      ysrc = le->u + bm * kPStride;
      yw = le->w + bm * kPStride;
    } else if (wm.helpers > 0) {
      // the helper warps were handed this point at the top of the evaluation; collect their results
      bar_sync(wm.bar_id + 1, wm.bar_threads);
      for (int h = 0; h < wm.helpers; ++h) {
        const float* hm = wm.helper_mem + h * wm.helper_stride;
        const float ll = hm[0];
        const int m = (int)hm[1];
        if (ll < best || (ll == best && m < bm)) {
          best = ll;
          bm = m;
          ysrc = hm + kYbufOff;
        }
      }
    } else {
      gmm_scan(tb, wm, xr, 0, 1, with_grad, best, bm);
    }
    if (gmm_component) *gmm_component = bm;
    uni = fmaf(kPosePriorW2, best, uni);
    if (body_owner) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int i = i0 + c;
        const float xi = xr[c];
        float gi = gr[c];
        if (with_grad) gi = fmaf(kPosePriorW2, yw ? fmaf(le->t, yw[i], ysrc[i]) : ysrc[i], gi);
        if (ob.keep_w2 != 0.f) {      // temporal pose-preserve term (losses.py:57-59)
          const float d = xi - ob.keep[c];
          lsum = fmaf(ob.keep_w2 * d, d, lsum);
          gi = fmaf(2.f * ob.keep_w2, d, gi);
        }
        if (i == 9 || i == 12 || i == 52 || i == 55) {   // angle prior (losses.py:13-21)
          const float sgn = i == 52 ? 1.f : -1.f;
          const float e = expf(xi * sgn);
          lsum = fmaf(kAnglePriorW2, e * e, lsum);
          gi = fmaf(2.f * kAnglePriorW2 * sgn, e * e, gi);
        }
        gr[c] = gi;
      }
    }
  }
  return wsum(lsum) + uni;
}

// ---------------------------------------------------------------------------------------------
// L-BFGS over lane-distributed vectors.  Every lane carries the same scalar state (all reductions
// are butterfly all-reduces, so the lanes agree bit for bit) and runs the same control flow.
// ---------------------------------------------------------------------------------------------
struct WVec {
  mutable float x[3];    // trial point; the two-loop's running vector between evaluations
  mutable float xk[3];   // iterate
  mutable float d[3];    // search direction
  float* gs;             // [4][96] gradient slots (shared memory)
  float* hist;           // this warp's (y, s) history in global memory: vector v of pair h at
                         // hist[((2 h + v) * 3 + c) * 32 + lane]
  float* ro;             // [hmax] (shared memory)
  float* al;             // [hmax]
  int hmax;
  K2B_HD float& G(int slot, int c) const { return gs[slot * kWarpVec + 3 * lane_id() + c]; }
  K2B_HD float& H(int h, int v, int c) const { return hist[((2 * h + v) * 3 + c) * 32 + lane_id()]; }
};

struct WarpOps {
  typedef WVec C;
  typedef WVec V;
  static K2B_HD float dot3(const float (&a)[3], const float (&b)[3]) {
    return wsum(fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2])));
  }
  static K2B_HD float dot_cur_d(const V& v, int cur) {
    const float g[3] = {v.G(cur, 0), v.G(cur, 1), v.G(cur, 2)};
    return dot3(g, v.d);
  }
  static K2B_HD void set_trial(const C&, const V& v, float tf) {
#pragma unroll
    for (int c = 0; c < 3; ++c) v.x[c] = fmaf(tf, v.d[c], v.xk[c]);
  }
  static K2B_HD float begin_copy(const C&, const V& v) {
    float gm = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v.xk[c] = v.x[c];
      gm = fmaxf(gm, fabsf(v.G(0, c)));
    }
    return wmax(gm);
  }
  static K2B_HD void neg_grad(const C&, const V& v, int g0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) v.x[c] = -v.G(g0, c);
  }
  static K2B_HD void update_direction(const C&, const V& v, int g0, int slot_prev_grad, float tf, int& num_old,
                                      int& head, float& H_diag) {
    int h = (head + num_old) % v.hmax;
    if (num_old == v.hmax) h = head;
    float y[3], s[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float gi = v.G(g0, c);
      y[c] = gi - v.G(slot_prev_grad, c);
      s[c] = v.d[c] * tf;
      v.x[c] = -gi;
      v.H(h, 0, c) = y[c];
      v.H(h, 1, c) = s[c];
    }
    const float ys = dot3(y, s), yy = dot3(y, y);
    if (ys > 1e-10f) {
      if (num_old == v.hmax) head = (head + 1) % v.hmax;
      else ++num_old;
      v.ro[h] = 1.f / ys;           // every lane stores the same value
      H_diag = ys / yy;
    }
    wsync();
    // two-loop recursion (lbfgs.py:430-442); the next pair is requested before this one is used
    float yn[3], sn[3];
    if (num_old > 0) {
      const int hk = (head + num_old - 1) % v.hmax;
#pragma unroll
      for (int c = 0; c < 3; ++c) { yn[c] = v.H(hk, 0, c); sn[c] = v.H(hk, 1, c); }
    }
#pragma unroll 1
    for (int k = num_old - 1; k >= 0; --k) {
      const int hk = (head + k) % v.hmax;
      float yk[3], sk[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) { yk[c] = yn[c]; sk[c] = sn[c]; }
      if (k > 0) {
        const int hn = (head + k - 1) % v.hmax;
#pragma unroll
        for (int c = 0; c < 3; ++c) { yn[c] = v.H(hn, 0, c); sn[c] = v.H(hn, 1, c); }
      }
      const float a = dot3(sk, v.x) * v.ro[hk];
      v.al[hk] = a;
#pragma unroll
      for (int c = 0; c < 3; ++c) v.x[c] = fmaf(-a, yk[c], v.x[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) v.x[c] *= H_diag;
    wsync();
    if (num_old > 0) {
      const int hk = head % v.hmax;
#pragma unroll
      for (int c = 0; c < 3; ++c) { yn[c] = v.H(hk, 0, c); sn[c] = v.H(hk, 1, c); }
    }
#pragma unroll 1
    for (int k = 0; k < num_old; ++k) {
      const int hk = (head + k) % v.hmax;
      float yk[3], sk[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) { yk[c] = yn[c]; sk[c] = sn[c]; }
      if (k + 1 < num_old) {
        const int hn = (head + k + 1) % v.hmax;
#pragma unroll
        for (int c = 0; c < 3; ++c) { yn[c] = v.H(hn, 0, c); sn[c] = v.H(hn, 1, c); }
      }
      const float coef = v.al[hk] - dot3(yk, v.x) * v.ro[hk];
#pragma unroll
      for (int c = 0; c < 3; ++c) v.x[c] = fmaf(coef, sk[c], v.x[c]);
    }
  }
  static K2B_HD void commit_direction(const C&, const V& v, int g0, float& gsum, float& gtd, float& dmax) {
    float g[3], ga = 0.f, dm = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      g[c] = v.G(g0, c);
      v.d[c] = v.x[c];
      ga += fabsf(g[c]);
      dm = fmaxf(dm, fabsf(v.d[c]));
    }
    gsum = wsum(ga);
    gtd = dot3(g, v.d);
    dmax = wmax(dm);
  }
  static K2B_HD void first_trial(const C&, const V& v, float tf) {
#pragma unroll
    for (int c = 0; c < 3; ++c) v.x[c] = fmaf(tf, v.x[c], v.xk[c]);
  }
  static K2B_HD float move_iterate(const C&, const V& v, float tf) {
    float dm = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v.xk[c] = fmaf(tf, v.d[c], v.xk[c]);
      dm = fmaxf(dm, fabsf(v.d[c] * tf));
    }
    return wmax(dm);
  }
  static K2B_HD float grad_max(const V& v, int slot) {
    float gm = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) gm = fmaxf(gm, fabsf(v.G(slot, c)));
    return wmax(gm);
  }
  // line-search replay (1-D surrogate): element 0 (lane 0, register 0) carries the problem, the rest is zero
  static K2B_HD void replay_seed(const C&, const V& v, float gtd0) {
    const bool e0 = lane_id() == 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v.x[c] = 0.f;
      v.xk[c] = 0.f;
      v.d[c] = (e0 && c == 0) ? 1.f : 0.f;
#pragma unroll
      for (int s = 0; s < 4; ++s) v.G(s, c) = 0.f;
    }
    if (e0) v.G(0, 0) = gtd0;
    wsync();
  }
  static K2B_HD void replay_response(const V& v, int cur, float gtd) {
    if (lane_id() == 0) v.G(cur, 0) = gtd;
    wsync();
  }
};

// ---------------------------------------------------------------------------------------------
// Team protocol (see TeamMem).  A round: the leader posts the candidate steps and, when they changed, the base point
// and direction; every evaluator computes its point x = xk + t d, evaluates it and publishes loss and gradient; the
// leader feeds the machine with its own result and then with every published result the machine asks for next.
// ---------------------------------------------------------------------------------------------
K2B_HD void team_post(const TeamMem& tm, int kind, int n_c, int flags, long row, int tab_mode, float t_step) {
  if (lane_id() == 0) {
    tm.cmd[0] = kind;
    tm.cmd[1] = n_c;
    tm.cmd[2] = flags;
    tm.cmd[3] = (int)(row & 0xffffffffl);
    tm.cmd[4] = (int)(row >> 32);
    tm.cmd[5] = tab_mode;
    tm.tval[7] = t_step;
  }
}

// This warp's share (components idx, idx + E, ..) of a line-table update; every evaluator of the team calls it in the
// same round, then the team meets at bar_tab.  dbuf: this warp's two staging buffers ([2][kDbufStride]).
// Components are taken two at a time and walked in lockstep: an update is a chain of dependent shared-memory round trips
// (stage, product, dot products, butterfly), ~2 300 cycles per component at IPC 0.2 when done one after the other
// (measured with K2B_CHAIN_PROF); two independent chains interleave.  Per component the arithmetic and its order are
// unchanged, so results are bit-identical to the one-at-a-time version.
K2B_HD void line_tables_update(const WarpTables& tb, float* dbuf, const TeamMem& tm, int mode, float t_step, int idx) {
  const int lane = lane_id();
  const bool refresh = mode == kTabRefresh, line = mode == kTabLine;
  const int rg = lane < 24 ? lane >> 3 : lane - 24;
  const int cc = lane < 24 ? lane & 7 : 8;
  const bool act = lane < 27;
  const int p1 = lane < 24 ? (lane + 8) % 24 : 24 + (lane - 23) % 3;
  const int p2 = lane < 24 ? (lane + 16) % 24 : 24 + (lane - 22) % 3;
#pragma unroll 1
  for (int m0 = idx; m0 < kGmmM; m0 += 2 * tm.E) {
    const int m1 = m0 + tm.E;
    const bool two = m1 < kGmmM;
    const int ms[2] = {m0, two ? m1 : m0};
    float* u[2] = {tm.lu + ms[0] * kPStride, tm.lu + ms[1] * kPStride};
    float* w[2] = {tm.lw + ms[0] * kPStride, tm.lw + ms[1] * kPStride};
    float* abc[2] = {tm.labc + 4 * ms[0], tm.labc + 4 * ms[1]};
    float* db[2] = {dbuf, dbuf + kDbufStride};
    float A[2], pa[2] = {0.f, 0.f}, pb[2] = {0.f, 0.f}, pc[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      A[k] = abc[k][0];
      const float B = abc[k][1], Cc = abc[k][2];
      if (!refresh) A[k] = fmaf(t_step, fmaf(t_step, Cc, 2.f * B), A[k]);   // follow the iterate: xk moved by t_step along d
    }
    wsync();
    if (lane < 24) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (k == 1 && !two) break;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int i = 3 * lane + c;
          if (!refresh) u[k][i] = fmaf(t_step, w[k][i], u[k][i]);
          float dv = 0.f;
          if (i < kBodyDim) dv = refresh ? tm.xk[3 + i] - tb.mu[ms[k] * kMuStride + i] : tm.d[3 + i];
          db[k][i] = dv;
          if (!line) w[k][i] = 0.f;
        }
      }
    }
    wsync();
    if (refresh || line) {        // the precision-matrix products of the pair
      float2 acc[2][4];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (act && (k == 0 || two)) gmm_rows(tb, db[k], ms[k], rg, cc, acc[k]);
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[k][j] = make_float2(0.f, 0.f);
        }
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float y[8] = {acc[k][0].x, acc[k][0].y, acc[k][1].x, acc[k][1].y, acc[k][2].x, acc[k][2].y, acc[k][3].x, acc[k][3].y};
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = (y[j] + shfl(y[j], p1)) + shfl(y[j], p2);
        if (act && rg == 0 && (k == 0 || two)) {
          float* out = refresh ? u[k] : w[k];
          *reinterpret_cast<float4*>(out + 4 * cc) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(out + 36 + 4 * cc) = make_float4(y[4], y[5], y[6], y[7]);
        }
      }
      wsync();
    }
    if (lane < 24) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int i = 3 * lane + c;
          if (refresh) pa[k] = fmaf(db[k][i], u[k][i], pa[k]);
          if (line) {
            pb[k] = fmaf(db[k][i], u[k][i], pb[k]);
            pc[k] = fmaf(db[k][i], w[k][i], pc[k]);
          }
        }
      }
    }
    if (refresh) {
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const float a0 = shfl(pa[0], lane ^ s), a1 = shfl(pa[1], lane ^ s);
        pa[0] += a0; pa[1] += a1;
      }
      A[0] = pa[0]; A[1] = pa[1];
    }
    if (line) {
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const float b0 = shfl(pb[0], lane ^ s), c0 = shfl(pc[0], lane ^ s);
        const float b1 = shfl(pb[1], lane ^ s), c1 = shfl(pc[1], lane ^ s);
        pb[0] += b0; pc[0] += c0; pb[1] += b1; pc[1] += c1;
      }
    }
    if (lane == 0) {
      abc[0][0] = A[0]; abc[0][1] = pb[0]; abc[0][2] = pc[0];
      if (two) { abc[1][0] = A[1]; abc[1][1] = pb[1]; abc[1][2] = pc[1]; }
    }
  }
  wsync();
}

// an evaluator without a candidate this round still takes part in the helpers' handshake
K2B_HD void team_idle_round(const WarpMem& wm) {
  if (wm.helpers > 0) {
    if (lane_id() == 0) wm.dbuf[0] = 3.f;      // helpers: nothing to scan
    bar_arrive(wm.bar_id, wm.bar_threads);
    bar_sync(wm.bar_id + 1, wm.bar_threads);
  }
}
K2B_HD void team_release_helpers(const WarpMem& wm) {
  if (wm.helpers > 0) {
    if (lane_id() == 0) wm.dbuf[0] = 0.f;
    bar_arrive(wm.bar_id, wm.bar_threads);
  }
}

// Helper warp: scans components h, h + H, .. of the mixture prior at the point its evaluator posted.
K2B_HD void team_helper(const WarpTables& tb, const WarpMem& wm, float* own, const float* lead, int h, int H, int bar_b,
                        int bar_threads) {
  const int lane = lane_id();
  while (true) {
    bar_sync(bar_b, bar_threads);
    const float cmd = *reinterpret_cast<const volatile float*>(lead + 96);
    if (cmd == 0.f) break;
    if (cmd != 3.f) {
      float xr[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) xr[c] = *reinterpret_cast<const volatile float*>(lead + 3 * lane + c);
      float best;
      int bm;
      gmm_scan(tb, wm, xr, h, H, cmd == 2.f, best, bm);
      if (lane == 0) {
        own[0] = best;
        own[1] = (float)bm;
      }
    }
    bar_arrive(bar_b + 1, bar_threads);
  }
}

// The steps the team evaluates this round: the one the machine asked for, then the ones it would ask for next.
template <class Machine>
K2B_HD int team_candidates(const Machine& st, int E, float (&tc)[kMaxCand]) {
  tc[0] = (float)st.t;
  int n = 1;
  if (E > 1 && st.phase == 0) {       // bracket phase (lbfgs.py:57-94); the zoom phase is not predictable
    double a, b, a2, b2, a3, b3;
    st.predict_extrapolation(st.t, st.t_prev, a, b);
    st.predict_extrapolation(a, st.t, a2, b2);
    st.predict_extrapolation(b, st.t, a3, b3);
    const double order[5] = {a, b, a2, a3, b3};       // by frequency in the reference's runs: 10 t, 5.5 t, 100 t, 55 t, 30 t
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float f = (float)order[k];
      bool dup = !(fabsf(f) <= 3.0e38f);
#pragma unroll
      for (int i = 0; i < kMaxCand; ++i) dup = dup || (i < n && tc[i] == f);
      if (!dup && n < E && n < kMaxCand) {
#pragma unroll
        for (int i = 0; i < kMaxCand; ++i)       // tc[n] = f without a run-time index (tc stays in registers)
          if (i == n) tc[i] = f;
        ++n;
      }
    }
  }
  return n;
}

// ---------------------------------------------------------------------------------------------
// A whole sequence by one warp: the reference's frame loop (api/sequence.py:214-281).
// ---------------------------------------------------------------------------------------------
constexpr int kAdamTableW = 64;

// Cycle accounting of the leading evaluator (diagnostic builds only: make EXTRA=-DK2B_CHAIN_PROF; read with
// k2b_chain_prof).  Slots: 0 rounds, 1 next point + post, 2 line tables, 3 evaluation, 4 wait for the team, 5 machine
// (advance + candidates), 6 outer update (two-loop recursion), 7 rest of the round, 8 frames.
#if defined(K2B_CHAIN_PROF) && defined(__CUDACC__)
__device__ unsigned long long k2b_chain_prof_slots[16];
#endif
#if defined(K2B_CHAIN_PROF) && defined(__CUDA_ARCH__)
// accumulators live in shared memory (the kernel has no registers to spare), one row per warp of the CTA
#define K2B_PROF_DECL __shared__ unsigned long long pf_sh[12][16]; long long pf_t = clock64(); \
  unsigned long long* pf_acc = pf_sh[threadIdx.x >> 5]; if ((threadIdx.x & 31) < 16) pf_acc[threadIdx.x & 31] = 0ull; __syncwarp();
#define K2B_PROF_MARK(slot) { const long long pf_n = clock64(); if ((threadIdx.x & 31) == 0) pf_acc[slot] += (unsigned long long)(pf_n - pf_t); pf_t = pf_n; }
#define K2B_PROF_COUNT(slot) { if ((threadIdx.x & 31) == 0) ++pf_acc[slot]; }
#define K2B_PROF_FLUSH __syncwarp(); if (leader && lane == 0) { for (int i = 0; i < 14; ++i) atomicAdd(&k2b_chain_prof_slots[i], pf_acc[i]); }
#else
#define K2B_PROF_DECL
#define K2B_PROF_MARK(slot)
#define K2B_PROF_COUNT(slot)
#define K2B_PROF_FLUSH
#endif

struct ChainParams {
  long num_seq;            // S sequences, one warp each
  int frames;              // T frames per sequence, walked serially
  long in_seq_stride;      // frames between consecutive sequences in targets / conf / preserve_pose (>= T)
  long out_seq_stride;     // output row of (sequence s, frame t) = s * out_seq_stride + t * out_frame_stride:
  long out_frame_stride;   //   (T, 1) sequence-major, (1, S) time-major
  long first_seq_ind;      // seq_ind of frame 0 (0: first-frame budget, no temporal term; world_space.py:211,214)
  const int* seq_first;    // optional [S]: per-sequence seq_ind of frame 0 (overrides first_seq_ind)
  int chain;               // 1: frame t starts from frame t-1's result (use_previous_frame_init); 0: from the init
  int iters_first, iters_follow;
  int lbfgs, freeze_betas;
  int conf_mode;           // 0 none, 1 shared [K], 2 per frame [S][T][K]
  float lr, joint_w2, keep_w2;
  const float* targets;    // [S][T][K][3]
  const float* conf;
  const float* init_pose;  // [S][72]
  const float* init_betas; // [S][10]
  const float* init_transl;// [S][3]
  const float* init_expr;  // [S][10] (NS == 20)
  const float* preserve_pose;   // [S][T][69] or null = the frame's initial body pose (world_space.py:159)
  float* out_pose; float* out_betas; float* out_transl; float* out_expr;   // [S][T][..]
  float* out_loss; float* out_joints; int* out_evals;
  float* hist;             // L-BFGS (y, s) history, hist_floats(hmax) per resident warp
  int hmax;
  int helpers;             // helper warps per evaluator that scan the mixture components (0 = the evaluator does it)
  int team;                // evaluator warps per sequence (>= 1; > 1: speculative line-search evaluation, L-BFGS only)
  // camera-space fitter (core/fitters/camera_space.py:81-339), see k2b_fit_args: loss_kind 1 = stage 1, final_mode 1 = stage 2
  int loss_kind, final_mode;
  float depth_w2;
  const float* depth_ref;  // [S][stride][3] initial camera translation (loss_kind 1)
  int camera_seq;          // 1: CameraSpaceFitter.fit_frame per frame inside the launch (camera_space.py:81-339): forward at
                           // the frame's initial parameters -> camera translation from the four torso joints (:16-41),
                           // stage 1 over [global_orient, camera translation], stage 2 over the body (final_mode 1
                           // semantics); iters_first = iters_follow = num_iters; freeze_betas bit 0 holds for seq_ind > 0
                           // only (:219-224).  Needs out_joints and team == 1.
  int eval_only;           // 1: no fit -- one evaluation at the initial parameters per frame; out_pose / out_betas / out_transl /
                           // out_expr receive the GRADIENT, out_evals the arg-min mixture component (k2b_evaluate_batch, warp evaluator)
  float adam_step[kAdamTableW], adam_bc2[kAdamTableW];
};
K2B_HD constexpr long hist_floats(int hmax) { return (long)hmax * 2 * kWarpVec; }

template <int NS>
K2B_HD float load_elem(const ChainParams& p, long s, int e) {
  if (e < kPoseDim) return p.init_pose[s * kPoseDim + e];
  if (e < kShapeOff) return p.init_transl[s * 3 + (e - kTranslOff)];
  if (e < kShapeOff + 10) return p.init_betas[s * 10 + (e - kShapeOff)];
  if (NS == 20 && e < kShapeOff + 20) return p.init_expr[s * 10 + (e - kShapeOff - 10)];
  return 0.f;
}
template <int NS>
K2B_HD void store_elem(const ChainParams& p, long f, int e, float v) {
  if (e < kPoseDim) p.out_pose[f * kPoseDim + e] = v;
  else if (e < kShapeOff) p.out_transl[f * 3 + (e - kTranslOff)] = v;
  else if (e < kShapeOff + 10) p.out_betas[f * 10 + (e - kShapeOff)] = v;
  else if (NS == 20 && e < kShapeOff + 20 && p.out_expr) p.out_expr[f * 10 + (e - kShapeOff - 10)] = v;
}

// This lane's share of the observations of input row f (targets, joint weights, camera-stage extras); the temporal
// anchor (keep) is filled in by the caller.
template <int K>
K2B_HD void load_frame_obs(const ChainParams& p, long f, bool stage1, FrameObs& ob) {
  const int lane = lane_id();
  ob.tx = ob.ty = ob.tz = ob.w = 0.f;
  if (lane < K) {
    const float* tg = p.targets + (f * K + lane) * 3;
    ob.tx = tg[0]; ob.ty = tg[1]; ob.tz = tg[2];
    const float cf = p.conf_mode == 0 ? 1.f : (p.conf_mode == 1 ? p.conf[lane] : p.conf[f * K + lane]);
    ob.w = p.joint_w2 * cf * cf;
    // camera stage 1 looks at RHip, LHip, RShoulder, LShoulder only, unweighted (losses.py:80-92)
    if (stage1) ob.w = (lane == 1 || lane == 2 || lane == 16 || lane == 17) ? 1.f : 0.f;
  }
  ob.plain_sq = stage1;
  ob.depth_w2 = stage1 ? p.depth_w2 : 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) ob.dref[c] = (stage1 && lane == 24 && p.depth_ref) ? p.depth_ref[f * 3 + c] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// An evaluator warp of a team.  idx 0 leads: it walks its sequences frame by frame (the reference's loop,
// api/sequence.py:214-281; every frame one WorldSpaceFitter.fit_frame, world_space.py:93-257) and owns the optimiser
// state -- Adam moments in registers, or the L-BFGS machine.  idx > 0 evaluates the speculative steps of the rounds the
// leader posts.  All of them pass through the same evaluation call; see the note on code size at the top.
// ---------------------------------------------------------------------------------------------
// LB: the optimiser is a compile-time choice, so that the other optimiser's state does not occupy registers.
// CAM: camera sequences (ChainParams::camera_seq) -- a separate instantiation, because the per-frame stage state costs the
// world-space kernel registers it does not have (168 of 168 used: the L-BFGS build went from 36 to 130 bytes of spills
// and 16 % slower with the stage as run-time state).
// FIN: the launch may want the forward pass at the returned parameters (joints out, the camera stage's loss, the
// evaluation-only launch); without it every evaluation has a gradient and the final phase is compiled out.
template <int NS, int K, bool LB, bool CAM = false, bool FIN = true>
K2B_HD void run_evaluator(const ChainParams& p, const WarpTables& tb, const WarpMem& wm, const TeamMem& tm, int idx,
                          long first_seq, long seq_stride, float* hist) {
  const int lane = lane_id();
  // teams of several evaluators exist for L-BFGS only (k2b_fit_chain never combines them with Adam): for the Adam
  // instantiation the team protocol and the line form of the prior are compile-time dead
  const bool leader = LB ? idx == 0 : true;
  const bool teamed = LB && tm.E > 1;
  // camera-space stage 1: only global_orient and the translation move, no priors.  A launch constant for the two-launch
  // camera fit (loss_kind), a per-frame state of the leader when both stages run inside the launch (camera_seq).
  // the instantiation without a final phase is the plain world-space fit: no camera stage either (chain_inst.cu)
  bool stage1 = FIN ? p.loss_kind == 1 : false;
  bool priors_on = !stage1;
  constexpr bool lbfgs = LB;
  const bool body_owner = lane >= 1 && lane < 24;
  bool frozen[3];
  auto set_frozen = [&](bool s1, bool betas_fixed) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int e = 3 * lane + c;
      frozen[c] = e >= 75 + NS || (betas_fixed && e >= kShapeOff && e < kShapeOff + 10) ||
                  ((p.freeze_betas & 2) && e >= kShapeOff + 10) || (s1 && !(e < 3 || (e >= kTranslOff && e < kShapeOff)));
    }
  };
  set_frozen(stage1, (p.freeze_betas & 1) != 0);
  int cstage = 0;            // camera_seq: 1 forward at the initial parameters, 2 stage 1, 3 stage 2
  int evals_prev = 0;        // camera_seq: evaluations of stage 1
  bool first_frame = false;
  enum { kNewFrame, kAdam, kAdamFinal, kRound, kFinal, kEvalOnly, kCamGuess };
  // ---- leader state ------------------------------------------------------------------------------------------
  long seq = first_seq, f = 0, frow = 0;
  int t = -1, phase = kNewFrame, iters = 0, evals = 0, k = 1;
  float x0[3] = {0.f, 0.f, 0.f}, xr[3] = {0.f, 0.f, 0.f};
  float out_loss = 0.f;
  float m1[3] = {0.f, 0.f, 0.f}, m2[3] = {0.f, 0.f, 0.f};
  WVec v;
  v.gs = tm.gs;
  v.hist = hist;
  v.ro = tm.ro;
  v.al = tm.al;
  v.hmax = p.hmax;
#pragma unroll
  for (int c = 0; c < 3; ++c) v.x[c] = v.xk[c] = v.d[c] = 0.f;
  Lbfgs<75 + NS, WarpOps> st;
  st.init();
  bool lfirst = true;
  int tab_next = kTabRefresh;       // what the next round does to the line tables
  float t_pending = 0.f;            // step the iterate took since the tables were last brought up to date
  float tc[kMaxCand];
  int n_c = 1;
#pragma unroll
  for (int i = 0; i < kMaxCand; ++i) tc[i] = 0.f;
  // ---- state of every evaluator ---------------------------------------------------------------------------------
  FrameObs ob;
  ob.tx = ob.ty = ob.tz = ob.w = 0.f;
  ob.keep[0] = ob.keep[1] = ob.keep[2] = 0.f;
  ob.keep_w2 = 0.f;
  ob.plain_sq = stage1;
  ob.depth_w2 = 0.f;
  ob.dref[0] = ob.dref[1] = ob.dref[2] = 0.f;
  long cur_row = -1;
  if (leader && seq < p.num_seq) {
#pragma unroll
    for (int c = 0; c < 3; ++c) xr[c] = x0[c] = load_elem<NS>(p, seq, 3 * lane + c);
  }
  // camera_seq: (re)start the optimiser on the current xr for camera stage 1 (s1) or stage 2
  auto begin_camera_stage = [&](bool s1) {
    stage1 = s1;
    priors_on = !s1;
    set_frozen(s1, (p.freeze_betas & 1) != 0 && !first_frame);      // betas move on the first frame (camera_space.py:219-224)
    float dref[3] = {ob.dref[0], ob.dref[1], ob.dref[2]};
    load_frame_obs<K>(p, f, s1, ob);                                 // joint weights of the stage; keeps ob.keep
    if (s1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) ob.dref[c] = dref[c];
    }
    ob.keep_w2 = (s1 || first_frame) ? 0.f : p.keep_w2;
    cstage = s1 ? 2 : 3;
    evals = 0;
    if (lbfgs) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { v.x[c] = xr[c]; v.xk[c] = xr[c]; v.d[c] = 0.f; }
      st.init();
      lfirst = true;
      tab_next = kTabRefresh;
      t_pending = 0.f;
      phase = kRound;
    } else {
      k = 1;
#pragma unroll
      for (int c = 0; c < 3; ++c) m1[c] = m2[c] = 0.f;
      phase = iters > 0 ? kAdam : kAdamFinal;
    }
  };
  K2B_PROF_DECL
#pragma unroll 1
  while (true) {
    K2B_PROF_MARK(7)
    K2B_PROF_COUNT(0)
    // ===== 1. the next point ======================================================================================
    float x[3] = {0.f, 0.f, 0.f};
    bool with_grad = true, with_priors = priors_on, final_obs = false, do_eval = true, use_line = false, want_comp = false;
    float* jout = nullptr;
    int tab_mode = kTabKeep;
    float t_step = 0.f, le_t = 0.f;
    if (leader) {
      if (phase == kNewFrame) {
        if (++t >= p.frames) {
          seq += seq_stride;
          t = 0;
          if (seq < p.num_seq) {
#pragma unroll
            for (int c = 0; c < 3; ++c) xr[c] = x0[c] = load_elem<NS>(p, seq, 3 * lane + c);
          }
        }
        if (seq >= p.num_seq) {       // all sequences done: let the other evaluators go
          if (teamed) {
            team_post(tm, kCmdExit, 0, 0, 0, kTabKeep, 0.f);
            wsync();
            bar_arrive(tm.bar_go, 32 * tm.E);
          }
          break;
        }
        f = seq * p.in_seq_stride + t;                                   // input row
        frow = seq * p.out_seq_stride + (long)t * p.out_frame_stride;    // output row
        if (!p.chain) {
#pragma unroll
          for (int c = 0; c < 3; ++c) xr[c] = x0[c];
        }
        load_frame_obs<K>(p, f, stage1, ob);
#pragma unroll
        for (int c = 0; c < 3; ++c)
          ob.keep[c] = (p.preserve_pose && body_owner) ? p.preserve_pose[f * kBodyDim + 3 * lane - 3 + c] : xr[c];
        const bool first = (p.seq_first ? (long)p.seq_first[seq] : p.first_seq_ind) + t == 0;
        if (CAM) first_frame = first;
        ob.keep_w2 = first ? 0.f : p.keep_w2;      // the temporal term is on for seq_ind > 0 (world_space.py:211)
        iters = first ? p.iters_first : p.iters_follow;
        evals = 0;
        if (CAM) evals_prev = 0;
        out_loss = 0.f;
        if (CAM) {
          cstage = 1;
          stage1 = false;
          priors_on = false;
          phase = kCamGuess;
        } else if (!LB && FIN && p.eval_only) {      // evaluation-only launches use the Adam instantiation (chain_inst.cu)
          phase = kEvalOnly;
        } else if (lbfgs) {
#pragma unroll
          for (int c = 0; c < 3; ++c) { v.x[c] = xr[c]; v.xk[c] = xr[c]; v.d[c] = 0.f; }
          st.init();
          lfirst = true;
          tab_next = kTabRefresh;
          t_pending = 0.f;
          phase = kRound;
        } else {
          k = 1;
#pragma unroll
          for (int c = 0; c < 3; ++c) m1[c] = m2[c] = 0.f;
          phase = iters > 0 ? kAdam : kAdamFinal;
        }
      }
      float* jframe = (FIN && p.out_joints) ? p.out_joints + frow * K * 3 : nullptr;
      if (!LB && FIN && phase == kEvalOnly) {
        jout = jframe;
        want_comp = true;
      } else if (CAM && phase == kCamGuess) {
        // model joints at the frame's initial parameters, without a translation (camera_space.py:113-117)
        with_grad = false;
        with_priors = false;
        jout = jframe;
      } else if (!LB && !FIN && phase == kAdamFinal) {
        do_eval = false;
      } else if (!LB && phase == kAdamFinal) {
        // joints at the final parameters (world_space.py:258-278); camera stage 2 also re-evaluates the loss there
        do_eval = jframe != nullptr || p.final_mode != 0;
        with_grad = false;
        with_priors = priors_on && p.final_mode != 0;
        final_obs = true;
        jout = jframe;
      } else if (LB && (phase == kRound || phase == kFinal)) {
        const bool fin = FIN && phase == kFinal;
#if !defined(__CUDA_ARCH__) && defined(K2B_WARP_EMUL)
        if (lane == 0 && !fin) ++k2b_emul_rounds;
#endif
        n_c = 1;
        tc[0] = 0.f;
        if (!fin && !lfirst) n_c = team_candidates(st, tm.E, tc);
        tab_mode = priors_on ? tab_next : kTabKeep;
        t_step = t_pending;
        if (tab_next != kTabKeep || fin) {       // base point and direction of this line search
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            tm.xk[3 * lane + c] = fin ? xr[c] : (lfirst ? v.x[c] : v.xk[c]);
            tm.d[3 * lane + c] = v.d[c];
          }
        }
        {
          float tl = 0.f;
#pragma unroll
          for (int i = 0; i < kMaxCand; ++i)
            if (i == lane && i < n_c) tl = tc[i];
          if (lane < kMaxCand) tm.tval[lane] = tl;
        }
        const bool keep_on = (fin && p.final_mode ? 0.f : ob.keep_w2) != 0.f;
        team_post(tm, kCmdEval, n_c, (fin ? 0 : kFlagGrad) | (priors_on ? kFlagPriors : 0) | (keep_on ? kFlagKeep : 0) |
                                         ((fin || lfirst) ? kFlagBase : 0), f, tab_mode, t_step);
        wsync();
        if (teamed) bar_arrive(tm.bar_go, 32 * tm.E);
        tab_next = kTabKeep;
        t_pending = 0.f;
        with_grad = !fin;
        final_obs = fin;
        jout = fin ? jframe : nullptr;
        use_line = priors_on;
        le_t = tc[0];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) x[c] = (LB && phase == kRound) ? v.x[c] : xr[c];
      if (CAM && phase == kCamGuess && lane == 24) x[0] = x[1] = x[2] = 0.f;
    } else {
      bar_sync(tm.bar_go, 32 * tm.E);
      if (*reinterpret_cast<const volatile int*>(tm.cmd) == kCmdExit) break;
      n_c = *reinterpret_cast<const volatile int*>(tm.cmd + 1);
      const int flags = *reinterpret_cast<const volatile int*>(tm.cmd + 2);
      const long row = (long)(unsigned)*reinterpret_cast<const volatile int*>(tm.cmd + 3) |
                       ((long)*reinterpret_cast<const volatile int*>(tm.cmd + 4) << 32);
      tab_mode = *reinterpret_cast<const volatile int*>(tm.cmd + 5);
      t_step = *reinterpret_cast<const volatile float*>(tm.tval + 7);
      if (row != cur_row) {       // first round of a frame: the posted point is the frame's initial parameters
        load_frame_obs<K>(p, row, stage1, ob);
#pragma unroll
        for (int c = 0; c < 3; ++c)
          ob.keep[c] = (p.preserve_pose && body_owner) ? p.preserve_pose[row * kBodyDim + 3 * lane - 3 + c]
                                                       : *reinterpret_cast<const volatile float*>(tm.xk + 3 * lane + c);
        cur_row = row;
      }
      ob.keep_w2 = (flags & kFlagKeep) ? p.keep_w2 : 0.f;
      with_grad = FIN ? (flags & kFlagGrad) != 0 : true;
      with_priors = FIN ? (flags & kFlagPriors) != 0 : true;
      use_line = with_priors;
      do_eval = idx < n_c;
      if (do_eval) {
        const float ti = *reinterpret_cast<const volatile float*>(tm.tval + idx);
        const bool base = (flags & kFlagBase) != 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float xk = *reinterpret_cast<const volatile float*>(tm.xk + 3 * lane + c);
          const float d = *reinterpret_cast<const volatile float*>(tm.d + 3 * lane + c);
          x[c] = base ? xk : fmaf(ti, d, xk);
        }
        le_t = base ? 0.f : ti;
      }
    }
    K2B_PROF_MARK(1)
    // ===== 2. line tables: every evaluator of the team takes its share of the mixture components ====================
    if (LB && tab_mode != kTabKeep) {
      K2B_PROF_COUNT(9)
      line_tables_update(tb, wm.dbuf, tm, tab_mode, t_step, idx);
      K2B_PROF_MARK(10)
      if (teamed) bar_sync(tm.bar_tab, 32 * tm.E);
      K2B_PROF_MARK(11)
    }
    K2B_PROF_MARK(2)
    // ===== 3. the evaluation (the kernel's only call site of eval_warp) =============================================
    float gr[3] = {0.f, 0.f, 0.f};
    float loss = 0.f;
    int comp = 0;
    if (do_eval) {
      FrameObs oe = ob;
      if (final_obs && p.final_mode) oe.keep_w2 = 0.f;      // camera_space.py:316-326
      const LineEval le{tm.lu, tm.lw, tm.labc, le_t};
      loss = eval_warp<NS, K, LB>(tb, wm, oe, x, with_grad, with_priors && (!LB || use_line), gr, jout,
                                  want_comp ? &comp : nullptr, (LB && use_line) ? &le : nullptr);
    }
    K2B_PROF_MARK(3)
    // ===== 4. where the result goes ==================================================================================
    if (!leader) {
      if (do_eval) {
#pragma unroll
        for (int c = 0; c < 3; ++c) tm.res_g[idx * kWarpVec + 3 * lane + c] = frozen[c] ? 0.f : gr[c];
        if (lane == 0) tm.res_f[idx] = loss;
      }
      wsync();
      bar_arrive(tm.bar_done, 32 * tm.E);
      continue;
    }
    bool frame_done = false;
    if (CAM && phase == kCamGuess) {
      // initial camera translation: mean offset of RHip, LHip, RShoulder, LShoulder (guess_init_3d, camera_space.py:16-41)
      wsync();
      const float* jg = p.out_joints + frow * K * 3;
      const float* tg = p.targets + f * K * 3;
      float ct[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d2 = tg[2 * 3 + c] - jg[2 * 3 + c], d1 = tg[1 * 3 + c] - jg[1 * 3 + c];
        const float d17 = tg[17 * 3 + c] - jg[17 * 3 + c], d16 = tg[16 * 3 + c] - jg[16 * 3 + c];
        ct[c] = (((d2 + d1) + d17) + d16) / 4.f;
      }
      wsync();
      if (lane == 24) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { xr[c] = ct[c]; ob.dref[c] = ct[c]; }
      }
      if constexpr (CAM) begin_camera_stage(true);
    } else if (!LB && FIN && phase == kEvalOnly) {
#pragma unroll
      for (int c = 0; c < 3; ++c) store_elem<NS>(p, frow, 3 * lane + c, gr[c]);
      if (lane == 0) {
        p.out_loss[frow] = loss;
        if (p.out_evals) p.out_evals[frow] = comp;
      }
      phase = kNewFrame;
    } else if (!LB && phase == kAdam) {
      out_loss = loss;        // the loss of the last iteration, before its step (world_space.py:250-256)
      ++evals;
      float step_k, bc2_k;
      if (k <= kAdamTableW) {
        step_k = p.adam_step[k - 1];
        bc2_k = p.adam_bc2[k - 1];
      } else {
        step_k = (float)((double)p.lr / (1.0 - pow(0.9, (double)k)));
        bc2_k = (float)sqrt(1.0 - pow(0.999, (double)k));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (!frozen[c]) adam_update(xr[c], m1[c], m2[c], gr[c], step_k, bc2_k);
      if (++k > iters) {
        if (CAM && cstage == 2) {          // camera stage 1 is over: stage 2 starts from its result
          evals_prev = evals;
          if constexpr (CAM) begin_camera_stage(false);
        } else {
          phase = kAdamFinal;
        }
      }
    } else if (!LB && phase == kAdamFinal) {
      if (do_eval && p.final_mode) out_loss = loss;
      frame_done = true;
    } else if (LB && phase == kRound) {
#pragma unroll
      for (int c = 0; c < 3; ++c) v.G(st.cur, c) = frozen[c] ? 0.f : gr[c];
      wsync();
      if (teamed) bar_sync(tm.bar_done, 32 * tm.E);
      K2B_PROF_MARK(4)
      // feed the machine: own result, then every published result it asks for
      unsigned used = 1u;
      float next_loss = loss;
#pragma unroll 1
      while (true) {
#if !defined(__CUDA_ARCH__) && defined(K2B_WARP_EMUL)
        if (k2b_emul_trace && !lfirst) {
          const float gtd = st.dot_cur_d(v);
          if (lane == 0 && *k2b_emul_trace_n < k2b_emul_trace_cap) {
            float* row = k2b_emul_trace + 3 * (*k2b_emul_trace_n)++;
            row[0] = (float)st.t; row[1] = next_loss; row[2] = gtd;
          }
        }
#endif
        st.advance(v, v, next_loss, lfirst, iters, p.lr);
        lfirst = false;
        if (st.done || st.need_outer) {       // the line search is over (or never started): the iterate moved by t
          t_pending = (float)st.t;
          tab_next = kTabPoint;
          if (st.need_outer && !st.done) {
            K2B_PROF_MARK(5)
            st.start_outer(v, v);
            K2B_PROF_MARK(6)
            if (!st.done) tab_next = kTabLine;
          }
          break;
        }
        const float want = (float)st.t;
        int j = -1;
#pragma unroll
        for (int i = 1; i < kMaxCand; ++i)
          if (i < n_c && !((used >> i) & 1u) && tc[i] == want) j = i;
        if (j < 0) break;
        used |= 1u << j;
#pragma unroll
        for (int c = 0; c < 3; ++c) v.G(st.cur, c) = tm.res_g[j * kWarpVec + 3 * lane + c];
        next_loss = tm.res_f[j];
        wsync();
      }
      K2B_PROF_MARK(5)
      if (st.done) {
#pragma unroll
        for (int c = 0; c < 3; ++c) xr[c] = v.xk[c];
        // loss (and joints) re-evaluated at the returned parameters (world_space.py:246-247).  The returned parameters
        // ARE the accepted trial point (lbfgs.py:488-493 adds t d to the iterate the same way the trial was formed), so
        // the loss is the machine's own, bit for bit; the extra forward pass is only run when joints are wanted or the
        // camera stage re-evaluates without the temporal term.
        if (CAM && cstage == 2) {          // camera stage 1 is over: stage 2 starts from its result
          evals_prev = st.evals;
          if constexpr (CAM) begin_camera_stage(false);
        } else if (FIN && (p.out_joints || p.final_mode)) {
          phase = kFinal;
        } else {
          out_loss = (float)st.loss;
          evals = st.evals;
          frame_done = true;
        }
      }
    } else if (FIN && LB) {        // kFinal
      if (teamed) bar_sync(tm.bar_done, 32 * tm.E);
      out_loss = loss;
      evals = st.evals;
      frame_done = true;
    }
    if (frame_done) {
      K2B_PROF_COUNT(8)
#pragma unroll
      for (int c = 0; c < 3; ++c) store_elem<NS>(p, frow, 3 * lane + c, xr[c]);
      if (lane == 0) {
        p.out_loss[frow] = out_loss;
        if (p.out_evals) p.out_evals[frow] = CAM ? evals + evals_prev : evals;
      }
      phase = kNewFrame;
    }
  }
  K2B_PROF_FLUSH
  team_release_helpers(wm);
}

}  // namespace wc
}  // namespace k2b
