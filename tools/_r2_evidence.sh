cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default rc=$?" >> gpurun_out/r02_evidence.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?" >> gpurun_out/r02_evidence.log
python bench.py --optimizer adam --skip-cpu-baseline > gpurun_out/r02_bench_s1_adam.json 2> gpurun_out/r02_bench_s1_adam.err; echo "adam rc=$?" >> gpurun_out/r02_evidence.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_kernel|fit_kernel|blend_skin|mesh_|gather_extra|skin_inplace|fma_peak|artic" --csv --log-file gpurun_out/r02_bench_launch_list_ncu.csv python bench.py --skip-cpu-baseline --no-e2e-vertices --steps 1 --warmup 1 --fp-steps 1 > gpurun_out/r02_ncu_list.log 2>&1; echo "list rc=$?" >> gpurun_out/r02_evidence.log
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 20 -c 1 -o gpurun_out/r02_chain_bench -f python bench.py --skip-cpu-baseline --no-e2e-vertices --no-frame-parallel --steps 1 --warmup 1 > gpurun_out/r02_ncu_chain.log 2>&1; echo "chain rc=$?" >> gpurun_out/r02_evidence.log
