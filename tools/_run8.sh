cd $GRAFT_REPO_ROOT
timeout 300 python tests/gpu_debug.py chain 1x256 148x64 256x64 2>&1 | grep "chain " > gpurun_out/r2_lb_guard.log
cat gpurun_out/r2_lb_guard.log
timeout 600 python tools/window_tail.py 256 4096 1 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_lbfgs_parity.py -q -x 2>&1 | tail -3
