cd $GRAFT_REPO_ROOT
for sp in 0 2 3 4 5; do
  echo "== K2B_CHAIN_SPEC=$sp"
  K2B_CHAIN_SPEC=$sp timeout 300 python tests/gpu_debug.py chain 1x256 148x64 256x64 2>&1 | grep "chain lbfgs"
done > gpurun_out/r2_spec.log 2>&1
cat gpurun_out/r2_spec.log
