cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_b3_default.json 2> gpurun_out/r2_b3.err; echo "rc=$?" >> gpurun_out/r2_b3.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_b3_ref.json 2>> gpurun_out/r2_b3.err; echo "rc=$?" >> gpurun_out/r2_b3.err
tail -5 gpurun_out/r2_b3.err
