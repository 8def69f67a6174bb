cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_lbfgs_parity.py -q -x -s -k "final_forward" 2>&1 | tail -8
