cd $GRAFT_REPO_ROOT
python tests/gpu_debug.py chain 1x256 256x64 148x64 2>&1 | grep "chain lbfgs"
python -m pytest tests/test_gpu_lbfgs_parity.py -q -x -k "team or chains or teacher" 2>&1 | tail -2
