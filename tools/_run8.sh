cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -2
timeout 600 python tools/window_tail.py 256 4096 1 2>&1 | grep "S=256"
