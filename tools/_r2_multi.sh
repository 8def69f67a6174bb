cd $GRAFT_REPO_ROOT
N=$1
O=gpurun_out/final3; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 --skip-cpu-baseline --no-e2e-vertices > $O/r02_bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "rc=$?"
tail -c 600 $O/r02_bench_${N}gpu.json
