cd $GRAFT_REPO_ROOT
for F in 0.7 0.8 0.9 1.0; do
  K2B_MESH_CAPPED_FRACTION=$F python bench.py --skip-cpu-baseline --no-frame-parallel --no-e2e-vertices 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('lbfgs cap $F', round(d['value']), round(d['ms_per_step'],1), 'fit', round(r.get('ms_per_step_in_kernel') or 0,1), d.get('roofline_mesh',{}).get('note'))"
done
for F in 0.45 0.6 0.75 0.9; do
  K2B_MESH_CAPPED_FRACTION=$F python bench.py --optimizer adam --skip-cpu-baseline --no-frame-parallel --no-e2e-vertices 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('adam cap $F', round(d['value']), round(d['ms_per_step'],1), 'fit', round(r.get('ms_per_step_in_kernel') or 0,1), d.get('roofline_mesh',{}).get('note'))"
done
