# usage: bash tools/bench_quick.sh [workload]  -- prints value / ms_per_step / e2e / roofline launch time for a few env settings
W=${1:-cfg2}
for ahead in 0 4 6 10 16; do
  echo -n "L2_AHEAD=$ahead: "
  FXN_GEMM_L2_AHEAD=$ahead python bench.py --no-cpu --workload $W 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), 'roofline_us', round(r['avg_launch_us'],1), 'issued_frac', round(r.get('issued_frac') or 0,3))"
done
