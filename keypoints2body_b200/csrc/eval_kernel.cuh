// eval_kernel.cuh -- MPJAE (mean per-joint angular error) between fitted and ground-truth poses.
//
// Restates /root/reference/keypoints2body/cli/eval.py:88-157 (rotvec_to_rotmat, compute_angular_error_deg,
// evaluate_pose_pair): per (frame, joint) the geodesic angle between the two axis-angle rotations,
//   cos = (tr(R_pred^T R_gt) - 1) / 2  clipped to [-1 + 1e-6, 1 - 1e-6],  angle = degrees(acos(cos)),
// summed in float64 (eval.py:153) over the first min(T) frames and min(D) // 3 joints.
// HBM-bound: reads 24 B per (frame, joint), writes nothing but one double per CTA (+ optional angles).
#pragma once

#include <cuda_runtime.h>

namespace k2b {

// eval.py:88-127 -- note it is NOT the smplx Rodrigues (no 1e-8 quirk): exact series below theta <= 1e-8
__device__ __forceinline__ void eval_rotmat(float x, float y, float z, float* r) {
  const float t2 = x * x + y * y + z * z;
  const float th = sqrtf(t2);
  float a, b;
  if (th > 1e-8f) {
    float s, c;
    sincosf(th, &s, &c);
    a = s / th;
    b = (1.f - c) / (th * th);
  } else {
    a = 1.f - t2 / 6.f + (t2 * t2) / 120.f;
    b = 0.5f - t2 / 24.f + (t2 * t2) / 720.f;
  }
  const float xy = x * y, xz = x * z, yz = y * z, xx = x * x, yy = y * y, zz = z * z;
  r[0] = 1.f - b * (yy + zz); r[1] = b * xy - a * z;     r[2] = b * xz + a * y;
  r[3] = b * xy + a * z;      r[4] = 1.f - b * (xx + zz); r[5] = b * yz - a * x;
  r[6] = b * xz - a * y;      r[7] = b * yz + a * x;      r[8] = 1.f - b * (xx + yy);
}

constexpr int kEvalThreads = 256;

// one thread per (frame, joint); grid-stride; per-CTA float64 partial sums -> atomicAdd
__global__ void __launch_bounds__(kEvalThreads)
mpjae_kernel(const float* __restrict__ pred, int pred_stride, const float* __restrict__ gt, int gt_stride, long frames,
             int joints, float* __restrict__ out_angles, double* __restrict__ out_sum) {
  __shared__ double s_part[kEvalThreads / 32];
  const long total = frames * joints;
  double acc = 0.0;
  for (long i = (long)blockIdx.x * kEvalThreads + threadIdx.x; i < total; i += (long)gridDim.x * kEvalThreads) {
    const long f = i / joints;
    const int j = (int)(i - f * joints);
    const float* a = pred + f * pred_stride + 3 * j;
    const float* b = gt + f * gt_stride + 3 * j;
    float ra[9], rb[9];
    eval_rotmat(a[0], a[1], a[2], ra);
    eval_rotmat(b[0], b[1], b[2], rb);
    float tr = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) tr += ra[k] * rb[k];
    float c = (tr - 1.f) * 0.5f;
    c = fminf(fmaxf(c, -1.f + 1e-6f), 1.f - 1e-6f);
    const float deg = acosf(c) * 57.29577951308232f;
    if (out_angles) out_angles[i] = deg;
    acc += (double)deg;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kEvalThreads / 32; ++w) t += s_part[w];
    atomicAdd(out_sum, t);
  }
}

}  // namespace k2b
