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
  auto kern = chain_kernel<K2B_NS, K2B_K>;
  const int warps = teams * p.team * (1 + p.helpers);
  const size_t smem = chain_smem_bytes(K2B_NS, teams, p.team, p.helpers, p.hmax);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  kern<<<grid, 32 * warps, smem, st>>>(p, tab);
  return cudaGetLastError();
}

}  // namespace k2b
