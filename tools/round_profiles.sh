# Round-end evidence run on ONE B200 (under gpurun): the five BASELINE workloads, the reference arm, an ncu launch list of a
# cfg2 step and one ncu --set full capture of every GEMM launch of that step. Outputs land in gpurun_out/.
set -x
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do
  python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_cfg2.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv \
    python bench.py --profile --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 60 -c 16 -f -o gpurun_out/gemm2_cfg2_step \
    python bench.py --profile --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/gemm2_cfg2_step.ncu-rep --page raw --csv > gpurun_out/ncu_full_gemm2_cfg2_step_raw.csv 2>/dev/null
rm -f gpurun_out/gemm2_cfg2_step.ncu-rep
python tools/graph_timeline.py cfg2 > gpurun_out/timeline_cfg2.log 2>&1
ls -la gpurun_out
