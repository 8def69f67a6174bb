cd $GRAFT_REPO_ROOT
timeout 300 python tests/gpu_debug.py chain 1x256 148x64 256x64 2>&1 | grep "chain lbfgs"
timeout 600 python tools/window_tail.py 256 4096 1 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lbfgs_parity.py -q -x 2>&1 | tail -2
timeout 600 python bench.py --skip-cpu-baseline --no-e2e-vertices > gpurun_out/r2_bench_slim.json 2> gpurun_out/r2_bench_slim.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_slim.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['mesh_overlap'], d['roofline']['frac'])
PY
