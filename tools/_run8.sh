cd $GRAFT_REPO_ROOT
for fl in "" "--no-vertices" "" "--no-vertices"; do
  timeout 600 python bench.py --skip-cpu-baseline --no-e2e-vertices --no-frame-parallel $fl > gpurun_out/r2_x.json 2> gpurun_out/r2_x.err
  python - "$fl" <<'PY'
import json,sys
d=json.load(open('gpurun_out/r2_x.json'))
print('flags', sys.argv[1], round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'fit ms', round(d['roofline']['ms_per_step_in_kernel'],2), d.get('mesh_overlap'))
PY
done
