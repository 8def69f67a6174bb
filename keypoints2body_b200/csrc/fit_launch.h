// fit_launch.h -- launch entry points of the fit-kernel instantiations (one TU each).
#pragma once
#include <cuda_runtime.h>

namespace k2b {
struct FitParams;
struct AdamTable;
// NS: shape coefficients (10 betas | 20 betas+expression); K: observed joints (22 | 24);
// MODE: 0 evaluate, 1 Adam, 2 L-BFGS, 3 / 4 the same optimisers for plain world-space fits without a final forward
// pass (FitMode in fit_kernel.cuh).  Specialised in fit_inst.cu.
template <int NS, int K, int MODE>
cudaError_t launch_fit(const FitParams& p, const AdamTable& at, int grid, cudaStream_t st);
}  // namespace k2b
