// chain_inst.cu -- one instantiation of the warp-per-sequence kernel per translation unit
// (compiled once per (K2B_NS, K2B_K) by the Makefile).
#include "chain_kernel.cuh"

#ifndef K2B_NS
#error "compile with -DK2B_NS=.. -DK2B_K=.."
#endif

namespace k2b {

template <>
cudaError_t launch_chain<K2B_NS, K2B_K>(const wc::ChainParams& p, const ChainTables& tab, int grid, int teams,
                                        cudaStream_t st) {
  // the optimiser is a template parameter of the kernel (the other optimiser's state would only cost registers)
  const bool lb = p.lbfgs != 0 && !p.eval_only;
  // plain world-space fits that want no forward pass at the returned parameters (no joints out, no camera stage, not an
  // evaluation-only launch) get the instantiation without the final phase: every evaluation has a gradient and priors
  const bool fin = p.out_joints != nullptr || p.final_mode != 0 || p.eval_only != 0 || p.camera_seq != 0 || p.loss_kind != 0;
  auto kern = lb ? (fin ? chain_kernel<K2B_NS, K2B_K, true, false, true> : chain_kernel<K2B_NS, K2B_K, true, false, false>)
                 : (fin ? chain_kernel<K2B_NS, K2B_K, false, false, true> : chain_kernel<K2B_NS, K2B_K, false, false, false>);
#if K2B_NS == 10
  // camera sequences (SMPL only, like the reference's CameraSpaceFitter) are their own instantiation
  if (p.camera_seq) kern = lb ? chain_kernel<K2B_NS, K2B_K, true, true, true> : chain_kernel<K2B_NS, K2B_K, false, true, true>;
#else
  if (p.camera_seq) return cudaErrorInvalidValue;
#endif
  const int warps = teams * p.team * (1 + p.helpers);
  const size_t smem = chain_smem_bytes(K2B_NS, teams, p.team, p.helpers, p.hmax);
  static size_t configured[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int variant = (lb ? 1 : 0) + (p.camera_seq ? 2 : 0) + (fin ? 0 : 4);
  if (smem > configured[variant]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[variant] = smem;
  }
  kern<<<grid, 32 * warps, smem, st>>>(p, tab);
  return cudaGetLastError();
}

#if defined(K2B_CHAIN_PROF) && K2B_NS == 10 && K2B_K == 22
// diagnostic builds only: read (and optionally clear) the leader's cycle accounting of the <10, 22> instantiation
extern "C" int k2b_chain_prof(unsigned long long* out16, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out16, wc::k2b_chain_prof_slots, sizeof(unsigned long long) * 16);
  if (e == cudaSuccess && reset) {
    unsigned long long z[16] = {0};
    e = cudaMemcpyToSymbol(wc::k2b_chain_prof_slots, z, sizeof(z));
  }
  return e == cudaSuccess ? 0 : -1;
}
#endif

}  // namespace k2b
