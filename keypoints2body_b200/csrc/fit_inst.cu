// fit_inst.cu -- one instantiation of the fused fitting kernel per translation unit
// (compiled once per (K2B_NS, K2B_K, K2B_MODE) by the Makefile so the builds run in parallel).
#include "fit_kernel.cuh"
#include "fit_launch.h"

#ifndef K2B_NS
#error "compile with -DK2B_NS=.. -DK2B_K=.. -DK2B_MODE=.."
#endif

namespace k2b {

template <>
cudaError_t launch_fit<K2B_NS, K2B_K, K2B_MODE>(const FitParams& p, const AdamTable& at, int grid, cudaStream_t st) {
  auto kern = fit_kernel<K2B_NS, K2B_K, K2B_MODE>;
  const size_t smem = fit_smem_bytes<K2B_NS>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<grid, fit_threads<K2B_NS>(), smem, st>>>(p, at);
  return cudaGetLastError();
}

}  // namespace k2b
