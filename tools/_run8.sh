cd $GRAFT_REPO_ROOT
for ppc in 0 1 2 4 0 2; do
  K2B_MESH_PASSES_PER_CTA=$ppc timeout 600 python bench.py --skip-cpu-baseline --no-e2e-vertices --no-frame-parallel > gpurun_out/r2_x.json 2> gpurun_out/r2_x.err
  python - "$ppc" <<'PY'
import json,sys
d=json.load(open('gpurun_out/r2_x.json'))
print('ppc', sys.argv[1], round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'fit ms', round(d['roofline']['ms_per_step_in_kernel'],2), 'mesh', round(d['roofline_mesh']['ms'],2), d['mesh_overlap'])
PY
done
