cd $GRAFT_REPO_ROOT
make -C keypoints2body_b200/csrc -j16 EXTRA=-DK2B_CHAIN_PROF OBJDIR=../../build/csrc_prof LIB=../libk2b_b200.so > gpurun_out/r2_cycles_build.log 2>&1
python tests/gpu_debug.py chaincycles 1x256 256x64 > gpurun_out/r2_cycles.log 2>&1
