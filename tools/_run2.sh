cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t2.log
timeout 300 python tools/chain_sweep.py > gpurun_out/r2_sweep.log 2>&1
tail -3 gpurun_out/r2_t2.log
