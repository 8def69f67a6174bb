// artic_emul.cu -- DEBUGGING HARNESS, NOT PART OF THE PRODUCT.
//
// Runs the __host__ __device__ code of the general articulated fit (keypoints2body_b200/csrc/artic_core.cuh,
// artic_kernel.cuh: hands / face observations, MANO, FLAME) on the CPU, one frame at a time, with the SAME argument
// structures as the C ABI (include/k2b_b200.h: k2b_artic_desc, k2b_artic_fit_args) but host pointers, so
// tests/test_artic_emul.py can pin the maths to the reference's goldens in the authoring container (no GPU there).
// Built into tests/host_emul/libk2b_artic_emul.so by build.sh; the shipped library has no CPU path.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/k2b_b200.h"
#include "../../keypoints2body_b200/csrc/artic_kernel.cuh"

using namespace k2b;

// gmm_P [8][69][72] symmetric precisions, gmm_mu [8][72], gmm_nlw [8] (what k2b_model keeps on the device); may be
// null when desc->body_off < 0.
extern "C" int emu_artic_fit(const k2b_artic_desc* d, const float* gmm_P, const float* gmm_mu, const float* gmm_nlw,
                             const k2b_artic_fit_args* g) {
  if (!d || !g) return -1;
  if (d->num_joints > ar::kMaxJoints || d->num_shape > ar::kMaxShape || d->num_params > ar::kMaxParams) return -2;
  if (g->num_obs > ar::kMaxObs) return -3;
  ar::ArticFitParams p{};
  ar::ArticModel& M = p.M;
  M.nj = d->num_joints; M.ns = d->num_shape; M.n = d->num_params; M.npick = d->num_picked; M.npf = 9 * (d->num_joints - 1);
  M.parents = d->parents; M.J0 = d->J0; M.JS = d->JS; M.pose_src = d->pose_src; M.shape_src = d->shape_src;
  M.transl_src = d->transl_src;
  M.pv_t = d->pv_template; M.pv_S = d->pv_shapedirs; M.pv_P = d->pv_posedirs; M.pv_idx = d->pv_skin_idx; M.pv_w = d->pv_skin_w;
  M.reg_w = d->reg_w; M.keep_w = d->keep_w;
  M.body_off = d->body_off;
  M.gmm_P = gmm_P; M.gmm_mu = gmm_mu; M.gmm_nlw = gmm_nlw;
  if (M.body_off >= 0 && !(gmm_P && gmm_mu && gmm_nlw)) return -4;
  p.num_frames = g->num_frames;
  p.K = g->num_obs;
  p.mode = g->mode;
  p.iters = g->num_iters;
  p.conf_per_frame = g->conf_per_frame;
  p.hmax = lbfgs_history_capacity(g->num_iters);
  p.lr = g->lr;
  p.joint_w2 = g->joint_loss_weight * g->joint_loss_weight;
  p.keep_scale = g->keep_scale;
  p.obs_idx = g->obs_idx; p.targets = g->targets; p.conf = g->conf; p.init_x = g->init_x; p.keep_x = g->keep_x;
  p.frozen = g->frozen;
  p.out_x = g->out_x; p.out_loss = g->out_loss; p.out_grad = g->out_grad; p.out_points = g->out_points;
  p.out_evals = g->out_evals; p.out_comp = g->out_gmm_component;
  for (int k = 1; k <= ar::kArticAdamTable; ++k) {
    p.adam_step[k - 1] = (float)((double)g->lr / (1.0 - std::pow(0.9, (double)k)));
    p.adam_bc2[k - 1] = (float)std::sqrt(1.0 - std::pow(0.999, (double)k));
  }
  std::vector<float> ws;
  if (g->mode == K2B_ARTIC_LBFGS) ws.assign((size_t)Vecs::floats_per_frame(M.n, p.hmax), 0.f);
  for (long f = 0; f < p.num_frames; ++f) {
    if (!ws.empty()) std::fill(ws.begin(), ws.end(), 0.f);
    ar::artic_fit_frame<false>(p, f, ws.empty() ? nullptr : ws.data(), 1);
  }
  return 0;
}
