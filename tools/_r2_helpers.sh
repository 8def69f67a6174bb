cd $GRAFT_REPO_ROOT
for H in 2 3 4 5; do
  echo "== helpers $H"
  K2B_CHAIN_HELPERS=$H python tests/gpu_debug.py chain 1x256 256x64 148x64 2>&1 | grep "chain adam"
done
