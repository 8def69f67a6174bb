cd $GRAFT_REPO_ROOT
for rep in 1 2; do
for v in head TABLES MACHINE; do
  echo "== $v"
  K2B_LIB=$GRAFT_REPO_ROOT/variants/libk2b_$v.so timeout 600 python tools/window_tail.py 256 4096 1 2>&1 | grep lbfgs
done
done
