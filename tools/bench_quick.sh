# usage: bash tools/bench_quick.sh [workload]  -- value / ms_per_step for a few BatchNorm grid settings
W=${1:-cfg2}
for mr in 32 64; do
  echo -n "BN_MIN_ROWS=$mr: "
  FXN_BN_MIN_ROWS=$mr python bench.py --no-cpu --workload $W 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4))"
done
