cd $GRAFT_REPO_ROOT
for rep in 1 2; do
for lib in variants/libk2b_head.so keypoints2body_b200/libk2b_b200.so; do
  echo "== $lib"
  K2B_LIB=$GRAFT_REPO_ROOT/$lib timeout 600 python tools/window_tail.py 256 4096 1 2>&1 | grep "S=256"
done
done
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -2
