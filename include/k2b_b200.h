/*
 * k2b_b200.h -- C ABI of the B200-native keypoints2body fitting path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The
 * Python host (keypoints2body_b200/_native.py) binds it with ctypes; a
 * maintainer of the reference would bind the same symbols.
 *
 * Reference interfaces replaced (paths into /root/reference/keypoints2body):
 *   k2b_model_create      <- per-call setup in core/fitters/world_space.py:56-91
 *                            (index tables, MaxMixturePrior load, core/prior.py:101-176)
 *                            plus the body-model buffers smplx keeps
 *   k2b_fit_batch         <- WorldSpaceFitter.fit_frame, core/fitters/world_space.py:93-257
 *                            (compute_loss :173-212, L-BFGS :231-247, Adam :248-256),
 *                            called per frame from api/sequence.py:214-281
 *   k2b_fit_chain         <- the frame loop of optimize_params_sequence, api/sequence.py:214-281
 *                            (frame t starts from frame t-1's result), one warp per sequence;
 *                            also WorldSpaceFitter.fit_frame for small batches (low latency)
 *   k2b_evaluate_batch    <- one compute_loss()+backward(), world_space.py:173-212,239-243
 *   k2b_mesh_batch        <- the final body-model forward, world_space.py:258-278
 *   k2b_shape_pass        <- optimize_shape_multi_frame, core/shape.py:10-115
 *   k2b_mpjae             <- evaluate_pose_pair / compute_angular_error_deg, cli/eval.py:88-157
 *
 * Conventions
 *   - all arrays are dense row-major float32 unless stated; "dev" pointers live
 *     on the current CUDA device, "host" pointers in host memory;
 *   - every call returns 0 on success, a negative K2B_E* code otherwise, and
 *     k2b_last_error() returns a thread-local message;
 *   - calls taking a stream enqueue work on it and do not synchronise; the
 *     *_host variants copy in, run, copy out and synchronise the stream.
 */
#ifndef K2B_B200_H_
#define K2B_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define K2B_OK 0
#define K2B_EINVAL (-1)   /* bad argument */
#define K2B_ECUDA (-2)    /* CUDA runtime error */
#define K2B_ENOMEM (-3)   /* workspace too small */
#define K2B_EUNSUPPORTED (-4)

#define K2B_POSE_DIM 72   /* global_orient(3) + body_pose(69) */
#define K2B_BODY_POSE_DIM 69
#define K2B_NUM_BETAS 10
#define K2B_NUM_EXPR 10
#define K2B_GMM_COMPONENTS 8
#define K2B_MAX_FIT_JOINTS 24 /* kinematic joints a fit can observe (SMPL-24 / AMASS-22) */

typedef struct k2b_model k2b_model; /* opaque: device-resident constants of one body model */

/* Host-side description of a body model + pose prior; all pointers are HOST memory
 * and are copied (pre-contracted where useful) by k2b_model_create. */
typedef struct k2b_model_desc {
  int32_t num_joints;      /* n_j: 24 (SMPL) / 52 (SMPL-H) / 55 (SMPL-X) */
  int32_t num_vertices;    /* V: 6890 / 6890 / 10475 */
  int32_t num_shape;       /* 10 (betas) or 20 (betas + expression) */
  int32_t num_extra;       /* vertex-picked extra joints appended after the n_j kinematic ones */
  const int32_t* parents;  /* [n_j], parents[0] = -1, parents[j] < j */
  const float* v_template; /* [V][3] */
  const float* shapedirs;  /* [V][3][num_shape] */
  const float* posedirs;   /* [9*(n_j-1)][3V] */
  const float* J_regressor;/* [n_j][V] */
  const float* lbs_weights;/* [V][n_j] */
  const int32_t* extra_vertex_ids; /* [num_extra] */
  /* max-mixture pose prior (core/prior.py:101-176), 69-D, K2B_GMM_COMPONENTS comps */
  const float* gmm_means;      /* [8][69] */
  const float* gmm_chol;       /* [8][69][69] lower-triangular L with sym(P) = L L^T */
  const float* gmm_neg_log_w;  /* [8] = -log(nll_weights) */
} k2b_model_desc;
/* k2b_model_create pre-contracts the rest joints once, in float64:
 *   J0 = J_regressor . v_template,  JS = J_regressor . shapedirs,
 * so the fitting loop never touches vertices (the reference recomputes the full mesh on
 * every evaluation, world_space.py:192). */

int k2b_model_create(const k2b_model_desc* desc, k2b_model** out);
void k2b_model_destroy(k2b_model* m);

enum { K2B_OPT_ADAM = 0, K2B_OPT_LBFGS = 1 };
enum { K2B_FREEZE_BETAS = 1, K2B_FREEZE_EXPR = 2 };

/* One batched fit = B independent WorldSpaceFitter.fit_frame calls.
 * Flat parameter blocks follow the reference's optimiser order
 * (world_space.py:215-229) restricted to blocks that receive gradient from
 * body keypoints: global_orient, body_pose, transl, [expression], [betas]. */
typedef struct k2b_fit_args {
  int64_t num_frames;        /* B */
  int32_t num_obs;           /* K: 22 (AMASS) or 24 (SMPL24, SMPL only) */
  int32_t optimizer;         /* K2B_OPT_* */
  int32_t num_iters;         /* iteration budget (max_iter for L-BFGS) when frame_iters == NULL */
  int32_t freeze_betas;      /* bit 0 (K2B_FREEZE_BETAS): betas kept at their initial value; bit 1 (K2B_FREEZE_EXPR):
                                expression kept (the reference optimises it only when the caller supplied one,
                                world_space.py:222-223) */
  int32_t conf_per_frame;    /* conf is [B][K] (1) or [K] shared (0) */
  float lr;                  /* step_size, reference default 1e-2 */
  float joint_loss_weight;   /* reference default 600 */
  float pose_preserve_weight;/* reference default 5; applied where preserve is on */
  /* inputs (device) */
  const float* targets;      /* [B][K][3] */
  const float* conf;         /* [B][K] or [K]; NULL = ones */
  const float* init_pose;    /* [B][72] */
  const float* init_betas;   /* [B][10] */
  const float* init_transl;  /* [B][3] */
  const float* init_expr;    /* [B][10] or NULL (required iff model num_shape == 20) */
  const float* preserve_pose;/* [B][69] or NULL = init body pose (world_space.py:159) */
  const int32_t* frame_iters;   /* [B] per-frame budget or NULL */
  const uint8_t* frame_preserve;/* [B] 1 = temporal term on (seq_ind > 0), or NULL = preserve_all */
  int32_t preserve_all;      /* used when frame_preserve == NULL */
  /* outputs (device) */
  float* out_pose;           /* [B][72] */
  float* out_betas;          /* [B][10] */
  float* out_transl;         /* [B][3] */
  float* out_expr;           /* [B][10] or NULL */
  float* out_loss;           /* [B] (L-BFGS: loss at the returned params; Adam: loss of the
                                last iteration before its step, world_space.py:246-256) */
  float* out_joints;         /* [B][K][3] posed kinematic joints incl. transl, or NULL.  NULL (with final_loss_mode 0)
                              * also spares an L-BFGS fit the forward pass at the returned parameters
                              * (world_space.py:246-247): they are the accepted trial point, whose loss is returned */
  int32_t* out_evals;        /* [B] closure evaluations performed, or NULL */
  /* scratch (device) */
  void* workspace;
  size_t workspace_bytes;    /* >= k2b_fit_workspace_bytes(...) */
  /* camera-space fitter (core/fitters/camera_space.py:81-339); all zero / NULL for the world-space fitter.
   * loss_kind 1 = stage 1: camera_fitting_loss_3d (core/losses.py:70-93) over RHip/LHip/RShoulder/LShoulder,
   * only global_orient and the translation move, depth term depth_weight^2 |transl - depth_ref|^2 counted
   * four times (the reference's (B,4,3)+(B,3) broadcast).  final_loss_mode 1 = stage 2: the returned loss is
   * re-evaluated at the final parameters without the preserve term (camera_space.py:316-326). */
  int32_t loss_kind;
  int32_t final_loss_mode;
  float depth_weight;        /* reference: 100 */
  const float* depth_ref;    /* [B][3] initial camera translation, required for loss_kind 1 */
} k2b_fit_args;

size_t k2b_fit_workspace_bytes(const k2b_model* m, int64_t num_frames, int32_t optimizer,
                               int32_t max_iters);
int k2b_fit_batch(const k2b_model* m, const k2b_fit_args* args, void* cuda_stream);

/* Same arguments with HOST pointers for every input/output array; workspace may be
 * NULL (allocated internally and cached on the model).  Performs H2D copies, the fit,
 * D2H copies, then synchronises. */
int k2b_fit_batch_host(k2b_model* m, const k2b_fit_args* host_args, void* cuda_stream);

/* Sequences walked serially in t inside ONE launch, one warp per sequence: the reference's default
 * schedule (api/sequence.py:214-281, use_previous_frame_init=True).  Frame t of sequence s is one
 * WorldSpaceFitter.fit_frame(init, j3d[s][t], conf, seq_ind = first_seq_ind + t): seq_ind 0 gets
 * num_iters_first and no temporal term, later frames num_iters_followup and the pose-preserve term
 * against their own initial body pose (world_space.py:159,211,214).  chain_init 1: init = result of
 * frame t-1 (frame 0: the sequence's init); 0: every frame starts from the sequence's init
 * (use_previous_frame_init=False).  With frames_per_sequence = 1 this is a low-latency k2b_fit_batch
 * (world loss only).  Outputs are [S][T][..] (loss / evals semantics as k2b_fit_batch). */
typedef struct k2b_chain_args {
  int64_t num_sequences;        /* S */
  int32_t frames_per_sequence;  /* T */
  int32_t num_obs;              /* K: 22 or 24 */
  int32_t optimizer;            /* K2B_OPT_* */
  int32_t num_iters_first;      /* reference default 30 */
  int32_t num_iters_followup;   /* reference default 10 */
  int64_t first_seq_ind;        /* seq_ind of every sequence's frame 0 */
  int32_t chain_init;           /* 1 = use_previous_frame_init */
  int32_t freeze_betas;         /* K2B_FREEZE_* bits, as in k2b_fit_args */
  int32_t conf_mode;            /* 0 none, 1 conf is [K], 2 conf is [S][stride][K] */
  int32_t out_time_major;       /* 0: outputs are [S][T][..]; 1: [T][S][..] (a time chunk of every sequence is contiguous) */
  int64_t in_sequence_stride;   /* frames between consecutive sequences in targets / conf / preserve_pose; 0 = T.
                                   A launch can cover a time window of longer sequences: pass pointers to the
                                   window's first frame, the full length here and the window start as first_seq_ind */
  float lr;
  float joint_loss_weight;
  float pose_preserve_weight;
  const float* targets;         /* [S][stride][K][3] (device) */
  const float* conf;
  const float* init_pose;       /* [S][72] */
  const float* init_betas;      /* [S][10] */
  const float* init_transl;     /* [S][3] */
  const float* init_expr;       /* [S][10] or NULL (required iff model num_shape == 20) */
  const float* preserve_pose;   /* [S][stride][69] or NULL = each frame's initial body pose */
  const int32_t* seq_first_ind; /* [S] per-sequence seq_ind of frame 0, or NULL = first_seq_ind for all */
  float* out_pose;              /* [S][T][72] ([T][S][72] if out_time_major; likewise below) */
  float* out_betas;             /* [S][T][10] */
  float* out_transl;            /* [S][T][3] */
  float* out_expr;              /* [S][T][10] or NULL */
  float* out_loss;              /* [S][T] */
  float* out_joints;            /* [S][T][K][3] or NULL (NULL: no forward pass at the returned parameters, as in k2b_fit_args) */
  int32_t* out_evals;           /* [S][T] or NULL */
  void* workspace;              /* L-BFGS history; may be NULL for Adam */
  size_t workspace_bytes;       /* >= k2b_chain_workspace_bytes(m, S, optimizer, max(num_iters_*)) */
  /* camera-space fitter stages, same meaning as in k2b_fit_args; all zero / NULL for the world-space fitter */
  int32_t loss_kind;
  int32_t final_loss_mode;
  float depth_weight;
  const float* depth_ref;       /* [S][stride][3], required for loss_kind 1 */
  /* 1: every frame is a whole CameraSpaceFitter.fit_frame (core/fitters/camera_space.py:81-339) inside the launch, so a
   * camera-space sequence (api/sequence.py:214-281 with coordinate_mode="camera") is ONE launch: forward at the frame's
   * initial parameters -> camera translation from the four torso joints (guess_init_3d, :16-41), stage 1 over
   * [global_orient, camera translation] (camera_fitting_loss_3d), stage 2 over the body with the translation acting as
   * camera translation; the returned loss is stage 2's without the temporal term.  Both stages run num_iters_first
   * iterations (set num_iters_followup to the same value); init_transl is ignored; freeze_betas bit 0 applies to frames
   * with seq_ind > 0 only (:219-224); loss_kind / final_loss_mode / depth_ref must be 0 / 0 / NULL, depth_weight is
   * used; out_joints is required (the per-frame fit joints, camera translation included). */
  int32_t camera_sequence;
} k2b_chain_args;

size_t k2b_chain_workspace_bytes(const k2b_model* m, int64_t num_sequences, int32_t optimizer,
                                 int32_t max_iters);
/* Launch geometry k2b_fit_chain uses for S sequences: CTAs (one per SM, at most the SM count) and warps per CTA;
 * returns the number of SMs of the device (so a caller can size concurrent work for the SMs left free). */
int k2b_chain_geometry(const k2b_model* m, int64_t num_sequences, int32_t* out_ctas, int32_t* out_warps);
int k2b_fit_chain(const k2b_model* m, const k2b_chain_args* args, void* cuda_stream);

/* One evaluation of loss and gradient at given parameters (parity / debugging). */
typedef struct k2b_eval_args {
  int64_t num_frames;
  int32_t num_obs;
  int32_t conf_per_frame;
  int32_t preserve_all;
  float joint_loss_weight;
  float pose_preserve_weight;
  const float* targets;      /* [B][K][3] */
  const float* conf;         /* [B][K] or [K] or NULL */
  const float* pose;         /* [B][72] */
  const float* betas;        /* [B][10] */
  const float* transl;       /* [B][3] */
  const float* expr;         /* [B][10] or NULL */
  const float* preserve_pose;/* [B][69] or NULL = pose */
  float* out_loss;           /* [B] */
  float* out_grad_pose;      /* [B][72] */
  float* out_grad_betas;     /* [B][10] */
  float* out_grad_transl;    /* [B][3] */
  float* out_grad_expr;      /* [B][10] or NULL */
  float* out_joints;         /* [B][K][3] or NULL */
  int32_t* out_gmm_component;/* [B] arg-min mixture component, or NULL */
  void* workspace;
  size_t workspace_bytes;
  int32_t warp_evaluator;    /* 0: the one-thread-per-frame evaluator (the code k2b_fit_batch runs);
                                1: the warp-per-frame evaluator (the code k2b_fit_chain runs, with the helper-warp
                                   geometry k2b_fit_chain would pick for B sequences); workspace may be NULL */
} k2b_eval_args;

int k2b_evaluate_batch(const k2b_model* m, const k2b_eval_args* args, void* cuda_stream);

/* Final mesh: vertices [B][V][3] and joints [B][n_j + num_extra][3] for fitted parameters.
 * full_pose is [B][3*n_j] axis-angle in the model's own joint order. */
typedef struct k2b_mesh_args {
  int64_t num_frames;
  const float* full_pose;    /* [B][3*n_j] */
  const float* shape;        /* [B][num_shape] (betas, then expression) */
  const float* transl;       /* [B][3] or NULL */
  float* out_vertices;       /* [B][V][3] or NULL (joints only) */
  float* out_joints;         /* [B][n_j + num_extra][3] */
  void* workspace;
  size_t workspace_bytes;
  int32_t max_ctas;          /* 0 = one CTA per SM; otherwise the persistent blend / skinning kernel uses at most
                                this many CTAs (so it fits beside another resident kernel, e.g. k2b_fit_chain) */
} k2b_mesh_args;

size_t k2b_mesh_workspace_bytes(const k2b_model* m, int64_t num_frames);
int k2b_mesh_batch(const k2b_model* m, const k2b_mesh_args* args, void* cuda_stream);

/* Shared-betas pre-pass (core/shape.py:10-115): for each of S sequences, L-BFGS over betas on
 * the root-aligned squared joint error of its first T frames at fixed poses. */
typedef struct k2b_shape_args {
  int32_t num_sequences;       /* S */
  int32_t frames_per_sequence; /* T frames used per sequence (reference default 50) */
  int64_t sequence_stride;     /* frames between consecutive sequences in targets / poses */
  int32_t num_obs;             /* K: 22 or 24 */
  int32_t pose_per_frame;      /* poses are [S][stride][72] (1) or one [S][72] per sequence (0) */
  int32_t conf_per_sequence;   /* conf is [S][K] (1) or [K] (0) */
  int32_t num_iters;           /* L-BFGS max_iter (reference default 40) */
  float lr;                    /* reference: 0.1 */
  float shape_prior_weight;    /* reference default 5 */
  const float* targets;        /* [S][stride][K][3] (device) */
  const float* poses;          /* see pose_per_frame */
  const float* conf;           /* or NULL = ones */
  const float* init_betas;     /* [S][10] */
  float* out_betas;            /* [S][10] */
  float* out_loss;             /* [S] */
  int32_t* out_evals;          /* [S] or NULL */
  void* workspace;
  size_t workspace_bytes;      /* >= k2b_shape_workspace_bytes */
} k2b_shape_args;

size_t k2b_shape_workspace_bytes(const k2b_model* m, int32_t num_sequences, int32_t num_iters);
int k2b_shape_pass(const k2b_model* m, const k2b_shape_args* args, void* cuda_stream);

/* Mean per-joint angular error between two pose arrays (cli/eval.py:129-157).
 * pred_pose [num_frames][pred_dims] and gt_pose [num_frames][gt_dims] are device arrays of axis-angle
 * triples; the first min(pred_dims, gt_dims) / 3 joints of every frame are compared.  *out_sum_deg
 * (device, float64) receives the sum of the angles in degrees (the caller divides by
 * num_frames * joints); out_angles_deg (device, [num_frames][joints], may be null) the angles.
 * Enqueues a memset of *out_sum_deg and one kernel on the stream; does not synchronise. */
int k2b_mpjae(const float* pred_pose, int32_t pred_dims, const float* gt_pose, int32_t gt_dims, int64_t num_frames,
              float* out_angles_deg, double* out_sum_deg, void* cuda_stream);

/* Line-search conformance (gate G3): replays N recorded strong-Wolfe line searches of torch.optim.LBFGS
 * (torch/optim/lbfgs.py:40-209) through the device machine, the objective being a table of the recorded (f, g.d)
 * responses.  The machine must propose the same trial steps and return the same (t, f).  All pointers are DEVICE
 * arrays.  warp_policy 0: one thread per search (the vector policy of k2b_fit_batch); 1: one warp per search with
 * lane-distributed vectors (the policy of k2b_fit_chain). */
typedef struct k2b_replay_args {
  int32_t num_searches;      /* N */
  int32_t max_resp;          /* row length R of the response / output tables */
  int32_t warp_policy;
  const double* t0;          /* [N] first trial step */
  const double* f0;          /* [N] loss at the iterate */
  const float* gtd0;         /* [N] directional derivative at the iterate */
  const double* d_norm;      /* [N] max |d| */
  const int32_t* max_ls;     /* [N] line-search budget (max_eval - evaluations so far) */
  const uint8_t* t_is_f32;   /* [N] 1: t is a float32 tensor (first outer iteration, lbfgs.py:396-401) */
  const int32_t* n_resp;     /* [N] recorded evaluations */
  const double* resp_f;      /* [N][R] recorded losses */
  const float* resp_gtd;     /* [N][R] recorded directional derivatives */
  double* out_t;             /* [N][R] trial steps the machine proposed */
  double* out_final;         /* [N][3] returned t, returned f, evaluations used */
  int32_t* out_k;            /* [N] responses consumed */
} k2b_replay_args;

int k2b_linesearch_replay(const k2b_replay_args* args, void* cuda_stream);

/* ---- general articulated fit: hands / face observations, MANO, FLAME ------------------------------------------
 * Replaces, for inputs the body-keypoint kernels do not cover:
 *   WorldSpaceFitter.fit_frame with target_model_indices addressing hand joints / vertex-picked landmarks
 *       (core/fitters/world_space.py:198-201, core/joints/adapters.py:224-380, core/constants.py:65-71),
 *   MANOFitter.fit_frame and FLAMEFitter.fit_frame (core/fitters/misc_models.py:18-359; generic_keypoint_loss_3d,
 *       core/losses.py:96-112).
 * A model is any kinematic tree (<= 56 joints) with <= 20 shape coefficients; the parameter vector x [num_params] is
 * the concatenation of the caller's blocks, pose_src / shape_src / transl_src say where the model reads its pose,
 * shape and translation in it (-1 = fixed zero).  Observable points: kinematic joint j (index j) and picked vertex p
 * (index num_joints + p), for which the host passes the rows of v_template / shapedirs / posedirs and the non-zero
 * skinning weights.  Loss = joint_w^2 sum conf^2 gmof(point - target) + sum reg_w[i] x_i^2
 *   + [temporal] keep_scale sum keep_w[i] (x_i - keep_i)^2 + [body_off >= 0] the SMPL pose prior and angle prior on
 * x[body_off .. body_off + 69) (core/losses.py:24-67).  All host pointers are copied. */
typedef struct k2b_artic k2b_artic;
typedef struct k2b_artic_desc {
  int32_t num_joints;
  int32_t num_shape;
  int32_t num_params;
  int32_t num_picked;
  const int32_t* parents;        /* [n_j] */
  const float* J0;               /* [n_j][3]  J_regressor . v_template */
  const float* JS;               /* [n_j][3][num_shape]  J_regressor . shapedirs */
  const int32_t* pose_src;       /* [3 n_j] */
  const int32_t* shape_src;      /* [num_shape] */
  int32_t transl_src;
  const float* pv_template;      /* [P][3] */
  const float* pv_shapedirs;     /* [P][3][num_shape] */
  const float* pv_posedirs;      /* [P][3][9 (n_j - 1)] */
  const int32_t* pv_skin_idx;    /* [P][8] */
  const float* pv_skin_w;        /* [P][8], 0 = unused slot */
  const float* reg_w;            /* [num_params] */
  const float* keep_w;           /* [num_params] */
  int32_t body_off;              /* -1 = no SMPL body priors */
  const k2b_model* prior_model;  /* supplies the max-mixture prior when body_off >= 0 */
} k2b_artic_desc;

int k2b_artic_create(const k2b_artic_desc* desc, k2b_artic** out);
void k2b_artic_destroy(k2b_artic* m);

enum { K2B_ARTIC_EVAL = 0, K2B_ARTIC_ADAM = 1, K2B_ARTIC_LBFGS = 2 };
typedef struct k2b_artic_fit_args {
  int64_t num_frames;          /* B */
  int32_t num_obs;             /* K <= 128 */
  int32_t mode;                /* K2B_ARTIC_* */
  int32_t num_iters;
  int32_t conf_per_frame;
  float lr;
  float joint_loss_weight;
  float keep_scale;            /* pose_preserve_weight^2 when the temporal term is on (seq_ind > 0), else 0 */
  const int32_t* obs_idx;      /* [K] (device) */
  const float* targets;        /* [B][K][3] */
  const float* conf;           /* [K] | [B][K] | NULL */
  const float* init_x;         /* [B][num_params] */
  const float* keep_x;         /* [B][num_params] or NULL = init_x */
  const uint8_t* frozen;       /* [num_params] 1 = not optimised, or NULL */
  float* out_x;                /* [B][num_params] */
  float* out_loss;             /* [B] */
  float* out_grad;             /* [B][num_params], required for K2B_ARTIC_EVAL */
  float* out_points;           /* [B][K][3] model points at the returned parameters, or NULL */
  int32_t* out_evals;          /* [B] or NULL */
  int32_t* out_gmm_component;  /* [B] or NULL (K2B_ARTIC_EVAL) */
  void* workspace;
  size_t workspace_bytes;      /* >= k2b_artic_workspace_bytes(...) */
} k2b_artic_fit_args;

size_t k2b_artic_workspace_bytes(const k2b_artic* m, int64_t num_frames, int32_t mode, int32_t num_iters);
int k2b_artic_fit(const k2b_artic* m, const k2b_artic_fit_args* args, void* cuda_stream);

/* FP32-FMA micro-benchmark used as the roofline denominator of the fit kernel:
 * returns achieved TFLOP/s (2 flop per FMA) over `iters` dependent-chain rounds. */
int k2b_fma_peak(int iters, double* out_tflops, double* out_ms, void* cuda_stream);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t k2b_launch_count(void);

const char* k2b_last_error(void);
const char* k2b_version(void);

#ifdef __cplusplus
}
#endif
#endif /* K2B_B200_H_ */
