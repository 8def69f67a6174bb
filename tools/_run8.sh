cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python tools/window_tail.py 256 4096 1 2>&1 | grep "S=256"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lbfgs_parity.py -q -x 2>&1 | tail -1
