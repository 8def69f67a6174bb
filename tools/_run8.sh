cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -2
timeout 600 python bench.py --skip-cpu-baseline --no-e2e-vertices > gpurun_out/r2_x.json 2> gpurun_out/r2_x.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_x.json'))
print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'fit ms', round(d['roofline']['ms_per_step_in_kernel'],2), round(d['roofline']['frac'],4), d['mesh_overlap'])
for k in ('frame_parallel','frame_parallel_adam'):
    v=d[k]; print(k, round(v['value']), round(v['ms_per_step'],2), round(v['roofline']['frac'],4), v['roofline']['ms_per_step_in_kernel'])
PY
