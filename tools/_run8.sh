cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "overlap or windows" 2>&1 | tail -2
for i in 1 2; do
timeout 600 python bench.py --skip-cpu-baseline --no-e2e-vertices --no-frame-parallel > gpurun_out/r2_bench_tuner2.json 2> gpurun_out/r2_bench_tuner2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_tuner2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['mesh_overlap'])
PY
done
timeout 600 python bench.py --optimizer adam --skip-cpu-baseline --no-e2e-vertices --no-frame-parallel > gpurun_out/r2_bench_tuner2_adam.json 2> gpurun_out/r2_bench_tuner2_adam.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_tuner2_adam.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['mesh_overlap'])
PY
