# usage: bash tools/bench_quick.sh [workload]  -- value / ms_per_step for a few BatchNorm grid settings (FXN_BN_WAVES)
W=${1:-cfg4}
for wv in 16 48 128; do
  echo -n "BN_WAVES=$wv: "
  FXN_BN_WAVES=$wv python bench.py --no-cpu --workload $W --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4))"
done
