cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for opt in lbfgs adam; do
  python bench.py --schedule two_sweep --optimizer $opt --skip-cpu-baseline --no-e2e-vertices --steps 3 --warmup 2 > gpurun_out/r2_k1b_$opt.json 2>> gpurun_out/r2_k1.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_k1b_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), d['roofline']['frac'], d['roofline']['ms_per_step_in_kernel'])
    except Exception as e: print(f, e)
PY
