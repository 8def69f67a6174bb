cd $GRAFT_REPO_ROOT
for args in "--no-vertices" "--chunks 1" ""; do
  python bench.py --optimizer adam --skip-cpu-baseline --no-frame-parallel --no-e2e-vertices $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$args', round(d['value']), round(d['ms_per_step'],1), 'fit', round(r.get('ms_per_step_in_kernel') or 0,1), d.get('roofline_mesh',{}).get('note'))"
done
for args in "--no-vertices" ""; do
  python bench.py --skip-cpu-baseline --no-frame-parallel --no-e2e-vertices $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('lbfgs $args', round(d['value']), round(d['ms_per_step'],1), 'fit', round(r.get('ms_per_step_in_kernel') or 0,1), d.get('roofline_mesh',{}).get('note'))"
done
