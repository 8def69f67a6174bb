// mesh_kernel.cuh -- final full-mesh output (K2 pose/shape blend + K3 LBS skinning).
//
// Restates the smplx forward the reference calls once per fitted frame
// (/root/reference/keypoints2body/core/fitters/world_space.py:258-278) [smplx-from-memory]:
//   v_posed = v_template + shapedirs.shape + (R[1:] - I).flatten() @ posedirs
//   verts   = (sum_j W_vj A_j) (v_posed; 1) + transl,  joints = [chain positions ; verts[extra ids]]
//
// Two paths:
//  * tensor-core path (blend depth 9(nj-1)+3ns <= kTcKpadMax = 576: SMPL, SMPL-H, SMPL-X):
//    mesh_pose_kernel -> blend_skin_tc_kernel (tcgen05 TF32 GEMM, blend_tc.cuh; 128 frames per pass
//    for SMPL, 64 for the deeper SMPL-H / SMPL-X blends) whose epilogue applies LBS to the
//    accumulators and writes the final vertices: v_posed never touches HBM.
//  * CUDA-core path (any deeper model, or joints-only calls): one fused FP32 kernel, register-tiled
//    over 16 frames per thread; the blend output never touches HBM.
#pragma once

#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/k2b_b200.h"
#include "blend_tc.cuh"
#include "fit_core.cuh"

namespace k2b {

constexpr int kMeshFT = 16;        // frames per CTA tile
constexpr int kMeshVT = 128;       // vertices per CTA tile (= threads)
constexpr int kMaxJoints = 64;

struct MeshModel {
  bool ready = false;
  int nj = 0, nv = 0, ns = 0, nextra = 0, npose = 0, ell = 0;
  int parents_h[kMaxJoints];
  int* parents = nullptr;      // [nj]
  float* rel = nullptr;        // [nj][1+ns][4] rest offsets to parent + shape derivatives
  float* J0S = nullptr;        // [nj][1+ns][4] absolute rest joints + shape derivatives
  float* v_template = nullptr; // [3V]
  float* shapedirs = nullptr;  // [ns][3V]
  float* posedirs = nullptr;   // [npose][3V]
  int* ell_idx = nullptr;      // [ell][V]
  float* ell_w = nullptr;      // [ell][V]
  int* extra_ids = nullptr;    // [nextra]
  // tensor-core blend (blend_tc.cuh)
  bool tc = false;
  bool fused = false;          // skinning fused into the blend epilogue (needs features + skinning rows in smem)
  bool vt = false;             // fused with lane = vertex tiles (blend_skin_vt_kernel); false: lane = coordinate tiles
  int fused_fr = 0;            // frames per pass of the fused kernel (64: SMPL; 32: SMPL-H / SMPL-X, whose skinning rows are larger)
  int kpad = 0, n_tiles = 0;
  __half* b_tiles = nullptr;   // [n_tiles][kpad/64][tc_b_bytes/2] pre-tiled FP16 image of dir_scale * [posedirs ; shapedirs]
  float dir_scale = 1.f;       // power of two that lifts the dirs into FP16's normal range
};

inline void mesh_model_free(MeshModel& m) {
  cudaFree(m.parents); cudaFree(m.rel); cudaFree(m.J0S); cudaFree(m.v_template); cudaFree(m.shapedirs);
  cudaFree(m.posedirs); cudaFree(m.ell_idx); cudaFree(m.ell_w); cudaFree(m.extra_ids); cudaFree(m.b_tiles);
  m = MeshModel();
}

// Rest joints J0 = J_regressor . v_template and JS = J_regressor . shapedirs in double.
inline void rest_joint_tables(const k2b_model_desc& d, std::vector<double>& J0, std::vector<double>& JS) {
  const int nj = d.num_joints, nv = d.num_vertices, ns = d.num_shape;
  J0.assign((size_t)nj * 3, 0.0);
  JS.assign((size_t)nj * 3 * ns, 0.0);
  for (int j = 0; j < nj; ++j) {
    const float* jr = d.J_regressor + (size_t)j * nv;
    for (int v = 0; v < nv; ++v) {
      const double w = jr[v];
      if (w == 0.0) continue;
      for (int k = 0; k < 3; ++k) {
        J0[j * 3 + k] += w * d.v_template[(size_t)v * 3 + k];
        const float* sd = d.shapedirs + ((size_t)v * 3 + k) * ns;
        double* o = &JS[((size_t)j * 3 + k) * ns];
        for (int s = 0; s < ns; ++s) o[s] += w * sd[s];
      }
    }
  }
}

template <class T>
inline bool mesh_upload(const std::vector<T>& h, T** dptr, std::string& err) {
  cudaError_t e = cudaMalloc((void**)dptr, h.size() * sizeof(T));
  if (e == cudaSuccess) e = cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    err = cudaGetErrorString(e);
    return false;
  }
  return true;
}

inline bool mesh_model_build(const k2b_model_desc& d, const std::vector<double>& J0, const std::vector<double>& JS,
                             MeshModel& m, std::string& err) {
  const int nj = d.num_joints, nv = d.num_vertices, ns = d.num_shape;
  if (nj > kMaxJoints) {
    err = "too many joints";
    return false;
  }
  m.nj = nj; m.nv = nv; m.ns = ns; m.nextra = d.num_extra; m.npose = 9 * (nj - 1);
  std::vector<int> parents(d.parents, d.parents + nj);
  for (int j = 0; j < nj; ++j) m.parents_h[j] = parents[j];
  std::vector<float> rel((size_t)nj * (1 + ns) * 4, 0.f), abs_((size_t)nj * (1 + ns) * 4, 0.f);
  for (int j = 0; j < nj; ++j) {
    const int pj = parents[j];
    for (int k = 0; k < 3; ++k) {
      abs_[((size_t)j * (1 + ns)) * 4 + k] = (float)J0[j * 3 + k];
      rel[((size_t)j * (1 + ns)) * 4 + k] = (float)(J0[j * 3 + k] - (pj >= 0 ? J0[pj * 3 + k] : 0.0));
      for (int s = 0; s < ns; ++s) {
        const double a = JS[((size_t)j * 3 + k) * ns + s];
        const double b = pj >= 0 ? JS[((size_t)pj * 3 + k) * ns + s] : 0.0;
        abs_[((size_t)j * (1 + ns) + 1 + s) * 4 + k] = (float)a;
        rel[((size_t)j * (1 + ns) + 1 + s) * 4 + k] = (float)(a - b);
      }
    }
  }
  std::vector<float> vt(d.v_template, d.v_template + (size_t)nv * 3);
  std::vector<float> sd((size_t)ns * nv * 3);
  for (int v = 0; v < nv; ++v)
    for (int k = 0; k < 3; ++k)
      for (int s = 0; s < ns; ++s) sd[(size_t)s * nv * 3 + v * 3 + k] = d.shapedirs[((size_t)v * 3 + k) * ns + s];
  std::vector<float> pd(d.posedirs, d.posedirs + (size_t)m.npose * nv * 3);
  // ELL skinning weights: width = max non-zeros per vertex
  int ell = 1;
  for (int v = 0; v < nv; ++v) {
    int c = 0;
    for (int j = 0; j < nj; ++j) c += d.lbs_weights[(size_t)v * nj + j] != 0.f;
    ell = c > ell ? c : ell;
  }
  m.ell = ell;
  std::vector<int> eidx((size_t)ell * nv, 0);
  std::vector<float> ew((size_t)ell * nv, 0.f);
  for (int v = 0; v < nv; ++v) {
    int c = 0;
    for (int j = 0; j < nj; ++j) {
      const float w = d.lbs_weights[(size_t)v * nj + j];
      if (w != 0.f) {
        eidx[(size_t)c * nv + v] = j;
        ew[(size_t)c * nv + v] = w;
        ++c;
      }
    }
  }
  std::vector<int> extra(d.extra_vertex_ids, d.extra_vertex_ids + d.num_extra);
  if (extra.empty()) extra.push_back(0);
  bool ok = mesh_upload(parents, &m.parents, err) && mesh_upload(rel, &m.rel, err) && mesh_upload(abs_, &m.J0S, err) &&
            mesh_upload(vt, &m.v_template, err) && mesh_upload(sd, &m.shapedirs, err) &&
            mesh_upload(pd, &m.posedirs, err) && mesh_upload(eidx, &m.ell_idx, err) &&
            mesh_upload(ew, &m.ell_w, err) && mesh_upload(extra, &m.extra_ids, err);
  // ---- tensor-core operand: dirs pre-tiled into the shared-memory image of every 48-column tile
  const int kdepth = m.npose + 3 * ns;     // shape rows split hi/lo: [S_hi ; S_hi ; S_lo]
  if (ok && kdepth <= kTcKpadMax) {
    m.kpad = tc_kpad(kdepth);
    const int ncols = nv * 3, kblocks = m.kpad / kTcBK;
    // SMPL: 64 frames of features (64 KB) + their skinning rows (72 KB) + the ring fit one SM -> fused epilogue
    // SMPL-H / SMPL-X (round 2): the lane = vertex kernel at 32 frames per pass (features 33 / 37 KB + ring 96 KB + the
    // skinning rows of 52 / 55 joints 80 / 84 KB = 211 / 216 KB), so that v_posed never reaches HBM there either
    const char* lay = getenv("K2B_MESH_LAYOUT");
    const bool coord = lay && std::string(lay) == "coord";       // round-1 lane = coordinate tiles, kept for A/B runs
    const char* unf = getenv("K2B_MESH_UNFUSED");                // A/B: blend + in-place skinning for SMPL-H / SMPL-X
    m.fused = false;
    if (ell <= 4 && nj == kTcFusedJoints && tc_smem_bytes(m.kpad, kTcMFused, kTcStagesFused, nj) <= 227 * 1024) {
      m.fused = true;
      m.fused_fr = kTcMFused;
    } else if (ell <= 4 && (nj == 52 || nj == 55) && !coord && !(unf && atoi(unf)) &&
               tc_smem_bytes(m.kpad, 32, kTcStagesFused, nj) <= 227 * 1024) {
      m.fused = true;
      m.fused_fr = 32;
    }
    // fused: lane = vertex tiles (128 vertices, three coordinate blocks each) unless K2B_MESH_LAYOUT=coord asks for the
    // round-1 lane = coordinate tiles (40 vertices x 3 coordinates per 128-row block; kept for A/B runs)
    m.vt = m.fused && !coord;
    const int per_tile = m.fused ? kTcVertsPerTile * 3 : kTcN;
    m.n_tiles = m.vt ? (nv + kTcVtVerts - 1) / kTcVtVerts
                     : ((ncols + per_tile - 1) / per_tile + 1) / 2 * 2;   // lane = coordinate tiles are consumed in pairs
    const int blocks_per_tile = m.vt ? 3 : 1;
    const size_t blk_halfs = (size_t)tc_b_bytes() / 2;
    // FP16 keeps TF32's 10 mantissa bits only in its normal range (>= 6.1e-5): scale the dirs by a power of two
    // so their largest entry sits near 2^10; the epilogue multiplies the accumulator by 1 / scale (exact).
    float dmax = 0.f;
    for (size_t i = 0; i < (size_t)m.npose * ncols; ++i) dmax = fmaxf(dmax, fabsf(d.posedirs[i]));
    for (size_t i = 0; i < (size_t)ncols * ns; ++i) dmax = fmaxf(dmax, fabsf(d.shapedirs[i]));
    int e = 0;
    if (dmax > 0.f) frexpf(dmax, &e);                 // dmax = f * 2^e, f in [0.5, 1)
    m.dir_scale = ldexpf(1.f, 10 - e);
    const float S = m.dir_scale;
    std::vector<__half> bt((size_t)m.n_tiles * blocks_per_tile * kblocks * blk_halfs, __float2half_rn(0.f));
    for (int nb = 0; nb < m.n_tiles * blocks_per_tile; ++nb)
      for (int n = 0; n < kTcN; ++n) {
        const int nt = nb;      // block index in the image ([tile][coordinate] for lane = vertex tiles)
        int col = nb * kTcN + n;
        if (m.vt) {
          const int vert = (nb / 3) * kTcVtVerts + n;
          col = vert < nv ? 3 * vert + nb % 3 : ncols;
        } else if (m.fused) {
          const int vert = tc_tile_vertex(nb, n);
          col = vert < 0 ? ncols : 3 * vert + (n % 32) % 3;
        }
        if (col >= ncols) continue;
        for (int k = 0; k < kdepth; ++k) {
          __half r;
          if (k < m.npose) {
            r = __float2half_rn(S * d.posedirs[(size_t)k * ncols + col]);
          } else {
            const int part = (k - m.npose) / ns, s = (k - m.npose) - part * ns;
            const float v = S * d.shapedirs[(size_t)col * ns + s];
            const __half hi = __float2half_rn(v);
            r = part == 2 ? __float2half_rn(v - __half2float(hi)) : hi;         // [S_hi ; S_hi ; S_lo]
          }
          bt[((size_t)nt * kblocks + k / kTcBK) * blk_halfs + tc_elem_off(n, k % kTcBK)] = r;   // swizzle-128B image
        }
      }
    ok = mesh_upload(bt, &m.b_tiles, err);
    m.tc = ok;
  }
  m.ready = ok;
  return ok;
}

// workspace: pose features [Bp][npose] + skinning matrices [Bp][nj][12], Bp = B padded to 128
inline long mesh_padded_frames(long B) { return (B + kTcM - 1) / kTcM * kTcM; }
inline size_t mesh_workspace_bytes(const MeshModel& m, long B) {
  const long Bp = mesh_padded_frames(B);
  return sizeof(float) * (size_t)Bp * (size_t)(m.npose + m.nj * 12);
}

// ---- kernel A: per-frame skeleton -> pose features, skinning matrices, posed joints ----------
__global__ void __launch_bounds__(128)
mesh_pose_kernel(int nj, int ns, const int* __restrict__ parents, const float4* __restrict__ rel,
                 const float4* __restrict__ J0S, const float* __restrict__ full_pose,
                 const float* __restrict__ shape, const float* __restrict__ transl, long B, long Bp,
                 float* __restrict__ posefeat, float* __restrict__ skin, float* __restrict__ out_joints,
                 int njout) {
  const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= Bp) return;
  const int npose = 9 * (nj - 1);
  float* pf = posefeat + f * npose;
  float* A = skin + f * (long)nj * 12;
  if (f >= B) {  // zero the padding frames so the tile kernels read finite values
    for (int i = 0; i < npose; ++i) pf[i] = 0.f;
    for (int i = 0; i < nj * 12; ++i) A[i] = 0.f;
    return;
  }
  float sh[20];
  for (int s = 0; s < ns; ++s) sh[s] = shape[f * ns + s];
  const V3 tr = transl ? v3(transl[f * 3], transl[f * 3 + 1], transl[f * 3 + 2]) : v3(0.f, 0.f, 0.f);
  // world transforms live in the skin buffer while the chain is walked: A[j] = [Rw | t]
  for (int j = 0; j < nj; ++j) {
    const float* r = full_pose + (f * nj + j) * 3;
    Rod o;
    const M3 R = rodrigues(v3(r[0], r[1], r[2]), o);
    if (j > 0) {
      float* q = pf + (j - 1) * 9;
#pragma unroll
      for (int i = 0; i < 9; ++i) q[i] = R.m[i] - ((i == 0 || i == 4 || i == 8) ? 1.f : 0.f);
    }
    const float4* e = rel + (long)j * (1 + ns);
    float4 r0 = e[0];
    V3 off = v3(r0.x, r0.y, r0.z);
    for (int s = 0; s < ns; ++s) {
      const float4 d = e[1 + s];
      off.x = fmaf(d.x, sh[s], off.x); off.y = fmaf(d.y, sh[s], off.y); off.z = fmaf(d.z, sh[s], off.z);
    }
    M3 Rw;
    V3 t;
    const int pj = parents[j];
    if (pj < 0) {
      Rw = R;
      t = off;
    } else {
      const float* P = A + pj * 12;
      M3 Rp;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        Rp.m[3 * i] = P[4 * i]; Rp.m[3 * i + 1] = P[4 * i + 1]; Rp.m[3 * i + 2] = P[4 * i + 2];
      }
      Rw = matmul(Rp, R);
      t = matvec(Rp, off) + v3(P[3], P[7], P[11]);
    }
    float* Q = A + j * 12;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Q[4 * i] = Rw.m[3 * i]; Q[4 * i + 1] = Rw.m[3 * i + 1]; Q[4 * i + 2] = Rw.m[3 * i + 2];
    }
    Q[3] = t.x; Q[7] = t.y; Q[11] = t.z;
    float* oj = out_joints + (f * njout + j) * 3;
    oj[0] = t.x + tr.x; oj[1] = t.y + tr.y; oj[2] = t.z + tr.z;
  }
  // A_j: translation reduced by Rw_j J_j, then + transl (so the skinned vertex lands in world space)
  for (int j = 0; j < nj; ++j) {
    const float4* e = J0S + (long)j * (1 + ns);
    float4 r0 = e[0];
    V3 J = v3(r0.x, r0.y, r0.z);
    for (int s = 0; s < ns; ++s) {
      const float4 d = e[1 + s];
      J.x = fmaf(d.x, sh[s], J.x); J.y = fmaf(d.y, sh[s], J.y); J.z = fmaf(d.z, sh[s], J.z);
    }
    float* Q = A + j * 12;
    Q[3] -= fmaf(Q[0], J.x, fmaf(Q[1], J.y, Q[2] * J.z));
    Q[7] -= fmaf(Q[4], J.x, fmaf(Q[5], J.y, Q[6] * J.z));
    Q[11] -= fmaf(Q[8], J.x, fmaf(Q[9], J.y, Q[10] * J.z));
  }
}

// ---- kernel A': the same, staged through shared memory (round 2) -----------------------------------------------------
// Kernel A gives every thread its own 0.8 + 1.2 KB rows of the two output arrays, so a warp's store instruction touches
// 32 sectors, and it walks the kinematic chain through global memory (the parent's transform is read back).  Here a
// block of kPoseStagedThreads frames builds its rows in shared memory (row strides padded to an odd number of words:
// conflict-free) and copies them out as one contiguous, coalesced run per array.  Used when the rows of 32 frames fit in
// 64 KB (SMPL, MANO, FLAME); same arithmetic in the same order as kernel A.
constexpr int kPoseStagedThreads = 32;
inline size_t mesh_pose_staged_smem(int nj) {
  const int npose = 9 * (nj - 1);
  return sizeof(float) * (size_t)kPoseStagedThreads * (size_t)((npose | 1) + ((nj * 12) | 1));
}
__global__ void __launch_bounds__(kPoseStagedThreads)
mesh_pose_staged_kernel(int nj, int ns, const int* __restrict__ parents, const float4* __restrict__ rel,
                        const float4* __restrict__ J0S, const float* __restrict__ full_pose,
                        const float* __restrict__ shape, const float* __restrict__ transl, long B, long Bp,
                        float* __restrict__ posefeat, float* __restrict__ skin, float* __restrict__ out_joints,
                        int njout) {
  extern __shared__ __align__(16) float ps_sm[];
  const int npose = 9 * (nj - 1), na = nj * 12;
  const int spf = npose | 1, sa = na | 1;          // odd row strides
  float* s_pf = ps_sm;
  float* s_A = ps_sm + (size_t)kPoseStagedThreads * spf;
  const int tid = threadIdx.x;
  const long f0 = (long)blockIdx.x * kPoseStagedThreads;
  const long f = f0 + tid;
  float* pf = s_pf + (size_t)tid * spf;
  float* A = s_A + (size_t)tid * sa;
  if (f >= B) {  // zero the padding frames so the tile kernels read finite values
    for (int i = 0; i < npose; ++i) pf[i] = 0.f;
    for (int i = 0; i < na; ++i) A[i] = 0.f;
  } else {
    float sh[20];
    for (int s = 0; s < ns; ++s) sh[s] = shape[f * ns + s];
    const V3 tr = transl ? v3(transl[f * 3], transl[f * 3 + 1], transl[f * 3 + 2]) : v3(0.f, 0.f, 0.f);
    for (int j = 0; j < nj; ++j) {
      const float* r = full_pose + (f * nj + j) * 3;
      Rod o;
      const M3 R = rodrigues(v3(r[0], r[1], r[2]), o);
      if (j > 0) {
        float* q = pf + (j - 1) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) q[i] = R.m[i] - ((i == 0 || i == 4 || i == 8) ? 1.f : 0.f);
      }
      const float4* e = rel + (long)j * (1 + ns);
      float4 r0 = e[0];
      V3 off = v3(r0.x, r0.y, r0.z);
      for (int s = 0; s < ns; ++s) {
        const float4 d = e[1 + s];
        off.x = fmaf(d.x, sh[s], off.x); off.y = fmaf(d.y, sh[s], off.y); off.z = fmaf(d.z, sh[s], off.z);
      }
      M3 Rw;
      V3 t;
      const int pj = parents[j];
      if (pj < 0) {
        Rw = R;
        t = off;
      } else {
        const float* P = A + pj * 12;
        M3 Rp;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          Rp.m[3 * i] = P[4 * i]; Rp.m[3 * i + 1] = P[4 * i + 1]; Rp.m[3 * i + 2] = P[4 * i + 2];
        }
        Rw = matmul(Rp, R);
        t = matvec(Rp, off) + v3(P[3], P[7], P[11]);
      }
      float* Q = A + j * 12;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        Q[4 * i] = Rw.m[3 * i]; Q[4 * i + 1] = Rw.m[3 * i + 1]; Q[4 * i + 2] = Rw.m[3 * i + 2];
      }
      Q[3] = t.x; Q[7] = t.y; Q[11] = t.z;
      float* oj = out_joints + (f * njout + j) * 3;
      oj[0] = t.x + tr.x; oj[1] = t.y + tr.y; oj[2] = t.z + tr.z;
    }
    for (int j = 0; j < nj; ++j) {
      const float4* e = J0S + (long)j * (1 + ns);
      float4 r0 = e[0];
      V3 J = v3(r0.x, r0.y, r0.z);
      for (int s = 0; s < ns; ++s) {
        const float4 d = e[1 + s];
        J.x = fmaf(d.x, sh[s], J.x); J.y = fmaf(d.y, sh[s], J.y); J.z = fmaf(d.z, sh[s], J.z);
      }
      float* Q = A + j * 12;
      Q[3] -= fmaf(Q[0], J.x, fmaf(Q[1], J.y, Q[2] * J.z));
      Q[7] -= fmaf(Q[4], J.x, fmaf(Q[5], J.y, Q[6] * J.z));
      Q[11] -= fmaf(Q[8], J.x, fmaf(Q[9], J.y, Q[10] * J.z));
    }
  }
  __syncwarp();
  // coalesced copy-out: the block's rows are one contiguous run in each global array (Bp is a multiple of 128)
  float* gp = posefeat + f0 * npose;
  for (int i = tid; i < kPoseStagedThreads * npose; i += kPoseStagedThreads) {
    const int fr = i / npose, e = i - fr * npose;
    gp[i] = s_pf[(size_t)fr * spf + e];
  }
  float* ga = skin + f0 * (long)na;
  for (int i = tid; i < kPoseStagedThreads * na; i += kPoseStagedThreads) {
    const int fr = i / na, e = i - fr * na;
    ga[i] = s_A[(size_t)fr * sa + e];
  }
}

// ---- kernel B: blend + skin a (vertex tile x frame tile) -------------------------------------
// thread = vertex; 16 frames per thread in registers; posefeat / shape / skin tiles in smem.
__global__ void __launch_bounds__(kMeshVT)
mesh_skin_kernel(int nj, int ns, int npose, int nv, int ell, const float* __restrict__ v_template,
                 const float* __restrict__ shapedirs, const float* __restrict__ posedirs,
                 const int* __restrict__ ell_idx, const float* __restrict__ ell_w,
                 const int* __restrict__ vlist, int nlist, const float* __restrict__ posefeat,
                 const float* __restrict__ skin, const float* __restrict__ shape,
                 const float* __restrict__ transl, long B, float* __restrict__ out, long out_stride,
                 long out_off) {
  extern __shared__ __align__(16) float sm[];
  float* s_pf = sm;                                  // [npose + ns][kMeshFT]
  float* s_A = sm + (size_t)(npose + ns) * kMeshFT;  // [kMeshFT][nj*12]
  const long f0 = (long)blockIdx.y * kMeshFT;
  const int tid = threadIdx.x;
  for (int i = tid; i < npose * kMeshFT; i += kMeshVT) {
    const int ft = i / npose, p = i - ft * npose;
    s_pf[p * kMeshFT + ft] = posefeat[(f0 + ft) * npose + p];
  }
  for (int i = tid; i < ns * kMeshFT; i += kMeshVT) {
    const int ft = i / ns, s = i - ft * ns;
    s_pf[(npose + s) * kMeshFT + ft] = (f0 + ft < B) ? shape[(f0 + ft) * ns + s] : 0.f;
  }
  for (int i = tid; i < nj * 12 * kMeshFT; i += kMeshVT) s_A[i] = skin[f0 * nj * 12 + i];
  __syncthreads();

  const int li = blockIdx.x * kMeshVT + tid;
  if (li >= nlist) return;
  const int v = vlist ? vlist[li] : li;
  float ax[kMeshFT], ay[kMeshFT], az[kMeshFT];
  {
    const float vx = v_template[3 * v], vy = v_template[3 * v + 1], vz = v_template[3 * v + 2];
#pragma unroll
    for (int ft = 0; ft < kMeshFT; ++ft) {
      ax[ft] = vx; ay[ft] = vy; az[ft] = vz;
    }
  }
  const long row = (long)nv * 3;
#pragma unroll 2
  for (int p = 0; p < npose + ns; ++p) {
    const float* src = (p < npose ? posedirs + (long)p * row : shapedirs + (long)(p - npose) * row) + 3 * v;
    const float dx = __ldg(src), dy = __ldg(src + 1), dz = __ldg(src + 2);
    const float4* w4 = reinterpret_cast<const float4*>(s_pf + p * kMeshFT);
#pragma unroll
    for (int q = 0; q < kMeshFT / 4; ++q) {
      const float4 w = w4[q];
      ax[4 * q] = fmaf(dx, w.x, ax[4 * q]); ay[4 * q] = fmaf(dy, w.x, ay[4 * q]); az[4 * q] = fmaf(dz, w.x, az[4 * q]);
      ax[4 * q + 1] = fmaf(dx, w.y, ax[4 * q + 1]); ay[4 * q + 1] = fmaf(dy, w.y, ay[4 * q + 1]); az[4 * q + 1] = fmaf(dz, w.y, az[4 * q + 1]);
      ax[4 * q + 2] = fmaf(dx, w.z, ax[4 * q + 2]); ay[4 * q + 2] = fmaf(dy, w.z, ay[4 * q + 2]); az[4 * q + 2] = fmaf(dz, w.z, az[4 * q + 2]);
      ax[4 * q + 3] = fmaf(dx, w.w, ax[4 * q + 3]); ay[4 * q + 3] = fmaf(dy, w.w, ay[4 * q + 3]); az[4 * q + 3] = fmaf(dz, w.w, az[4 * q + 3]);
    }
  }
  // skinning: T = sum_k w_k A_{j_k}; vertex = T (v_posed; 1) + transl
  int jk[8];
  float wk[8];
  const int ne = ell < 8 ? ell : 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    jk[k] = k < ne ? ell_idx[(long)k * nv + v] : 0;
    wk[k] = k < ne ? ell_w[(long)k * nv + v] : 0.f;
  }
#pragma unroll 4
  for (int ft = 0; ft < kMeshFT; ++ft) {
    if (f0 + ft >= B) break;
    float T[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < ne && wk[k] != 0.f) {
        const float4* Aj = reinterpret_cast<const float4*>(s_A + (ft * nj + jk[k]) * 12);
        const float4 a0 = Aj[0], a1 = Aj[1], a2 = Aj[2];
        T[0] = fmaf(wk[k], a0.x, T[0]); T[1] = fmaf(wk[k], a0.y, T[1]); T[2] = fmaf(wk[k], a0.z, T[2]); T[3] = fmaf(wk[k], a0.w, T[3]);
        T[4] = fmaf(wk[k], a1.x, T[4]); T[5] = fmaf(wk[k], a1.y, T[5]); T[6] = fmaf(wk[k], a1.z, T[6]); T[7] = fmaf(wk[k], a1.w, T[7]);
        T[8] = fmaf(wk[k], a2.x, T[8]); T[9] = fmaf(wk[k], a2.y, T[9]); T[10] = fmaf(wk[k], a2.z, T[10]); T[11] = fmaf(wk[k], a2.w, T[11]);
      }
    }
    // ELL rows wider than 8 (never for SMPL-family models) take the slow path
    for (int k = 8; k < ell; ++k) {
      const float w = ell_w[(long)k * nv + v];
      if (w == 0.f) continue;
      const float* Aj = s_A + (ft * nj + ell_idx[(long)k * nv + v]) * 12;
      for (int i = 0; i < 12; ++i) T[i] = fmaf(w, Aj[i], T[i]);
    }
    const long f = f0 + ft;
    const float tx = transl ? transl[f * 3] : 0.f, ty = transl ? transl[f * 3 + 1] : 0.f,
                tz = transl ? transl[f * 3 + 2] : 0.f;
    float* o = out + (f * out_stride + out_off + li) * 3;
    o[0] = fmaf(T[0], ax[ft], fmaf(T[1], ay[ft], fmaf(T[2], az[ft], T[3]))) + tx;
    o[1] = fmaf(T[4], ax[ft], fmaf(T[5], ay[ft], fmaf(T[6], az[ft], T[7]))) + ty;
    o[2] = fmaf(T[8], ax[ft], fmaf(T[9], ay[ft], fmaf(T[10], az[ft], T[11]))) + tz;
  }
}

// ---- in-place LBS on the tensor-core path: out[f][v] holds v_posed, becomes the skinned vertex ---
__global__ void __launch_bounds__(kMeshVT)
skin_inplace_kernel(int nj, int nv, int ell, const int* __restrict__ ell_idx, const float* __restrict__ ell_w,
                    const float* __restrict__ skin, const float* __restrict__ transl, long B, float* __restrict__ verts) {
  extern __shared__ __align__(16) float s_A[];     // [kMeshFT][nj*12]
  const long f0 = (long)blockIdx.y * kMeshFT;
  for (int i = threadIdx.x; i < nj * 12 * kMeshFT; i += kMeshVT) s_A[i] = skin[f0 * nj * 12 + i];
  __syncthreads();
  const int v = blockIdx.x * kMeshVT + threadIdx.x;
  if (v >= nv) return;
  int jk[8];
  float wk[8];
  const int ne = ell < 8 ? ell : 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    jk[k] = k < ne ? ell_idx[(long)k * nv + v] : 0;
    wk[k] = k < ne ? ell_w[(long)k * nv + v] : 0.f;
  }
#pragma unroll 4
  for (int ft = 0; ft < kMeshFT; ++ft) {
    const long f = f0 + ft;
    if (f >= B) break;
    float* o = verts + (f * nv + v) * 3;
    const float px = o[0], py = o[1], pz = o[2];
    float T[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < ne && wk[k] != 0.f) {
        const float4* Aj = reinterpret_cast<const float4*>(s_A + (ft * nj + jk[k]) * 12);
        const float4 a0 = Aj[0], a1 = Aj[1], a2 = Aj[2];
        T[0] = fmaf(wk[k], a0.x, T[0]); T[1] = fmaf(wk[k], a0.y, T[1]); T[2] = fmaf(wk[k], a0.z, T[2]); T[3] = fmaf(wk[k], a0.w, T[3]);
        T[4] = fmaf(wk[k], a1.x, T[4]); T[5] = fmaf(wk[k], a1.y, T[5]); T[6] = fmaf(wk[k], a1.z, T[6]); T[7] = fmaf(wk[k], a1.w, T[7]);
        T[8] = fmaf(wk[k], a2.x, T[8]); T[9] = fmaf(wk[k], a2.y, T[9]); T[10] = fmaf(wk[k], a2.z, T[10]); T[11] = fmaf(wk[k], a2.w, T[11]);
      }
    }
    for (int k = 8; k < ell; ++k) {
      const float w = ell_w[(long)k * nv + v];
      if (w == 0.f) continue;
      const float* Aj = s_A + (ft * nj + ell_idx[(long)k * nv + v]) * 12;
      for (int i = 0; i < 12; ++i) T[i] = fmaf(w, Aj[i], T[i]);
    }
    const float tx = transl ? transl[f * 3] : 0.f, ty = transl ? transl[f * 3 + 1] : 0.f,
                tz = transl ? transl[f * 3 + 2] : 0.f;
    o[0] = fmaf(T[0], px, fmaf(T[1], py, fmaf(T[2], pz, T[3]))) + tx;
    o[1] = fmaf(T[4], px, fmaf(T[5], py, fmaf(T[6], pz, T[7]))) + ty;
    o[2] = fmaf(T[8], px, fmaf(T[9], py, fmaf(T[10], pz, T[11]))) + tz;
  }
}

// vertex-picked extra joints: joints[f][nj + e] = verts[f][extra_ids[e]]
__global__ void gather_extra_kernel(const float* __restrict__ verts, const int* __restrict__ ids, int nextra, int nv,
                                    int nj, long B, float* __restrict__ joints) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * nextra) return;
  const long f = i / nextra;
  const int e = (int)(i - f * nextra);
  const float* s = verts + (f * nv + ids[e]) * 3;
  float* o = joints + (f * (nj + nextra) + nj + e) * 3;
  o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
}

inline size_t mesh_skin_smem(const MeshModel& m) {
  return sizeof(float) * (size_t)((m.npose + m.ns) * kMeshFT + kMeshFT * m.nj * 12);
}

inline bool mesh_forward(const MeshModel& m, const k2b_mesh_args& a, cudaStream_t st, std::string& err,
                         int& launches) {
  const long B = a.num_frames;
  const long Bp = mesh_padded_frames(B);
  float* posefeat = (float*)a.workspace;
  float* skin = posefeat + Bp * m.npose;
  const int njout = m.nj + m.nextra;
  const size_t psm = mesh_pose_staged_smem(m.nj);
  if (psm <= 64 * 1024) {
    static size_t pose_configured = 0;
    if (psm > pose_configured) {
      cudaError_t e = cudaFuncSetAttribute(mesh_pose_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
      if (e != cudaSuccess) {
        err = cudaGetErrorString(e);
        return false;
      }
      pose_configured = psm;
    }
    mesh_pose_staged_kernel<<<(unsigned)(Bp / kPoseStagedThreads), kPoseStagedThreads, psm, st>>>(
        m.nj, m.ns, m.parents, (const float4*)m.rel, (const float4*)m.J0S, a.full_pose, a.shape, a.transl, B, Bp,
        posefeat, skin, a.out_joints, njout);
  } else {
    mesh_pose_kernel<<<(unsigned)((Bp + 127) / 128), 128, 0, st>>>(
        m.nj, m.ns, m.parents, (const float4*)m.rel, (const float4*)m.J0S, a.full_pose, a.shape, a.transl, B, Bp,
        posefeat, skin, a.out_joints, njout);
  }
  ++launches;
  const char* force_fp32 = getenv("K2B_MESH_FP32");   // diagnostics / tests: take the CUDA-core path
  if (m.tc && a.out_vertices && !(force_fp32 && atoi(force_fp32))) {
    // ---- tensor-core path: blend (tcgen05) -> in-place skinning -> extra-joint gather ----------
    // SMPL: fused blend + skinning, 64 frames per pass.  SMPL-H / SMPL-X: 128-frame blend, then in-place skinning.
    const int fr = m.fused ? m.fused_fr : kTcM;
    const size_t tsm = tc_smem_bytes(m.kpad, fr, m.fused ? kTcStagesFused : kTcStages, m.fused ? m.nj : 0);
    auto* kern = blend_skin_tc_kernel<kTcM, kTcStages, 0, 0>;
#define K2B_PICK_NE(KERN, FRV, NJV)                                                   \
  (m.ell == 1 ? KERN<FRV, kTcStagesFused, 1, NJV> : m.ell == 2 ? KERN<FRV, kTcStagesFused, 2, NJV> \
   : m.ell == 3 ? KERN<FRV, kTcStagesFused, 3, NJV> : KERN<FRV, kTcStagesFused, 4, NJV>)
    if (m.fused && m.vt) {
      kern = m.nj == 52 ? K2B_PICK_NE(blend_skin_vt_kernel, 32, 52)
           : m.nj == 55 ? K2B_PICK_NE(blend_skin_vt_kernel, 32, 55)
                        : K2B_PICK_NE(blend_skin_vt_kernel, kTcMFused, kTcFusedJoints);
    } else if (m.fused) {
      kern = K2B_PICK_NE(blend_skin_tc_kernel, kTcMFused, kTcFusedJoints);
    }
#undef K2B_PICK_NE
    static size_t tc_configured[17] = {0};
    const int variant = m.fused ? m.ell + (m.vt ? 4 : 0) + (m.nj == 52 ? 4 : (m.nj == 55 ? 8 : 0)) : 0;
    if (tsm > tc_configured[variant]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
      if (e != cudaSuccess) {
        err = cudaGetErrorString(e);
        return false;
      }
      tc_configured[variant] = tsm;
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#ifdef K2B_DIAG
    const char* dbg = getenv("K2B_TC_DEBUG");
#else
    const char* dbg = nullptr;
#endif
    BlendParams bp{posefeat, a.shape, m.b_tiles, 1.f / m.dir_scale, m.v_template, (const float4*)skin, a.transl, m.ell_idx, m.ell_w,
                   a.out_vertices, B, m.npose, m.ns, m.kpad, m.nv, m.nj, m.ell, m.n_tiles, dbg ? atoi(dbg) : 0};
    const long passes = Bp / fr;
    if (a.max_ctas > 0 && a.max_ctas < sms) sms = a.max_ctas;
    kern<<<(unsigned)(passes < sms ? passes : sms), kTcThreads, tsm, st>>>(bp);
    ++launches;
    if (!m.fused) {
      const size_t ssm = sizeof(float) * (size_t)kMeshFT * m.nj * 12;
      const long ft_total = (B + kMeshFT - 1) / kMeshFT;
      for (long y0 = 0; y0 < ft_total; y0 += 65535) {
        const unsigned ny = (unsigned)(ft_total - y0 < 65535 ? ft_total - y0 : 65535);
        const long fo = y0 * kMeshFT;
        dim3 grid((m.nv + kMeshVT - 1) / kMeshVT, ny);
        skin_inplace_kernel<<<grid, kMeshVT, ssm, st>>>(m.nj, m.nv, m.ell, m.ell_idx, m.ell_w, skin + fo * m.nj * 12,
                                                         a.transl ? a.transl + fo * 3 : nullptr, B - fo,
                                                         a.out_vertices + fo * m.nv * 3);
        ++launches;
      }
    }
    if (m.nextra > 0) {
      const long n = B * m.nextra;
      gather_extra_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.out_vertices, m.extra_ids, m.nextra, m.nv, m.nj, B,
                                                                         a.out_joints);
      ++launches;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      err = cudaGetErrorString(e);
      return false;
    }
    return true;
  }
  const size_t smem = mesh_skin_smem(m);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(mesh_skin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      err = cudaGetErrorString(e);
      return false;
    }
    configured = smem;
  }
  const unsigned ftiles = (unsigned)(Bp / kMeshFT);
  // grid.y is limited to 65535 tiles (1M frames): chunk the frame range
  for (unsigned y0 = 0; y0 < ftiles; y0 += 65535u) {
    const unsigned ny = ftiles - y0 < 65535u ? ftiles - y0 : 65535u;
    const long fo = (long)y0 * kMeshFT;
    const long Bc = B - fo;
    if (a.out_vertices) {
      dim3 grid((m.nv + kMeshVT - 1) / kMeshVT, ny);
      mesh_skin_kernel<<<grid, kMeshVT, smem, st>>>(
          m.nj, m.ns, m.npose, m.nv, m.ell, m.v_template, m.shapedirs, m.posedirs, m.ell_idx, m.ell_w, nullptr,
          m.nv, posefeat + fo * m.npose, skin + fo * m.nj * 12, a.shape + fo * m.ns,
          a.transl ? a.transl + fo * 3 : nullptr, Bc, a.out_vertices + fo * m.nv * 3, m.nv, 0);
      ++launches;
    }
    if (m.nextra > 0) {
      dim3 grid((m.nextra + kMeshVT - 1) / kMeshVT, ny);
      mesh_skin_kernel<<<grid, kMeshVT, smem, st>>>(
          m.nj, m.ns, m.npose, m.nv, m.ell, m.v_template, m.shapedirs, m.posedirs, m.ell_idx, m.ell_w, m.extra_ids,
          m.nextra, posefeat + fo * m.npose, skin + fo * m.nj * 12, a.shape + fo * m.ns,
          a.transl ? a.transl + fo * 3 : nullptr, Bc, a.out_joints + fo * njout * 3, njout, m.nj);
      ++launches;
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    err = cudaGetErrorString(e);
    return false;
  }
  return true;
}

}  // namespace k2b
