set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_b1_default.json 2> gpurun_out/r2_b1.err
python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --schedule two_sweep > gpurun_out/r2_b1_s2_lbfgs.json 2>> gpurun_out/r2_b1.err
python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --schedule two_sweep --optimizer adam > gpurun_out/r2_b1_s2_adam.json 2>> gpurun_out/r2_b1.err
ncu --set full --clock-control none --import-source on -k regex:fit_kernel -s 2 -c 1 -o gpurun_out/r2_fit_adam python bench.py --schedule two_sweep --optimizer adam --steps 1 --warmup 1 --skip-cpu-baseline --frames-per-gpu 262144 > gpurun_out/r2_ncu1.log 2>&1
tail -5 gpurun_out/r2_t1.log
