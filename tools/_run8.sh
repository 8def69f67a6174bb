cd $GRAFT_REPO_ROOT
timeout 300 python tests/gpu_debug.py chain 1x256 256x64 2>&1 | grep "chain "
timeout 600 python tools/window_tail.py 256 4096 1 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6
