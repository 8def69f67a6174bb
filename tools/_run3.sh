cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/gpu_debug.py chain 1x256 256x64 > gpurun_out/r2_chain_time.log 2>&1
K2B_CHAIN_TEAM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 1 -c 1 -o gpurun_out/r2_chain_e1 python tests/gpu_debug.py chainprof lbfgs 148x16 > gpurun_out/r2_ncu_e1.log 2>&1
K2B_CHAIN_TEAM=6 timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 1 -c 1 -o gpurun_out/r2_chain_e6 python tests/gpu_debug.py chainprof lbfgs 148x16 > gpurun_out/r2_ncu_e6.log 2>&1
tail -3 gpurun_out/r2_chain_time.log
