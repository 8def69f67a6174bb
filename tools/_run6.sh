cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_b4_2gpu.json 2> gpurun_out/r2_b4.err; echo "rc=$?" >> gpurun_out/r2_b4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2_b4_ref2.json 2>> gpurun_out/r2_b4.err; echo "rc=$?" >> gpurun_out/r2_b4.err
tail -4 gpurun_out/r2_b4.err
