cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3.log
timeout 300 python tests/gpu_debug.py chain 1x256 256x64 > gpurun_out/r2_chain_time2.log 2>&1
timeout 300 python tools/chain_sweep.py 256x64 1x256 > gpurun_out/r2_sweep2.log 2>&1
python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_b2_default.json 2> gpurun_out/r2_b2.err
python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --optimizer adam > gpurun_out/r2_b2_adam.json 2>> gpurun_out/r2_b2.err
tail -2 gpurun_out/r2_t3.log; head -4 gpurun_out/r2_chain_time2.log; cat gpurun_out/r2_sweep2.log
