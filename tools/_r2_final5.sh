cd $GRAFT_REPO_ROOT
O=gpurun_out/final5; mkdir -p $O
python bench.py > $O/r02_bench_default.json 2> $O/bench_default.err; echo "default rc=$?" >> $O/log.txt
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?" >> $O/log.txt
python bench.py --optimizer adam --skip-cpu-baseline > $O/r02_bench_s1_adam.json 2> $O/bench_adam.err; echo "adam rc=$?" >> $O/log.txt
python -m pytest tests -m gpu -q -s > $O/r02_gpu_tests.txt 2>&1; echo "pytest rc=$?" >> $O/log.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_kernel|fit_kernel|blend_skin|mesh_|gather_extra|skin_inplace|fma_peak|artic" --csv --log-file $O/r02_bench_launch_list_ncu.csv python bench.py --skip-cpu-baseline --no-e2e-vertices --steps 1 --warmup 1 --fp-steps 1 > $O/ncu_list.log 2>&1; echo "list rc=$?" >> $O/log.txt
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 20 -c 1 -o $O/r02_chain_bench -f python bench.py --skip-cpu-baseline --no-e2e-vertices --no-frame-parallel --steps 1 --warmup 1 > $O/ncu_chain.log 2>&1; echo "chain rc=$?" >> $O/log.txt
python tests/gpu_debug.py chain 1x256 148x64 256x64 > $O/r02_chain_timing.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.txt 2>&1; echo "smoke rc=$?" >> $O/log.txt
cat $O/log.txt; tail -2 $O/r02_gpu_tests.txt
