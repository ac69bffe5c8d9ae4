# Round-end evidence run on ONE B200 (under gpurun): GPU test suite, smoke, the five BASELINE workloads (cfg2 with its CPU
# baseline), the reference arm, an ncu launch list of a cfg2 step, one ncu --set full capture of every GEMM launch of that
# step, and the CUPTI graph timelines. Outputs land in gpurun_out/.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --workload cfg2 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
for w in cfg1 cfg3 cfg4 cfg5; do
  python bench.py --no-cpu --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_cfg2.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv \
    python bench.py --profile --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 60 -c 16 -f -o gpurun_out/gemm2_cfg2_step \
    python bench.py --profile --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/gemm2_cfg2_step.ncu-rep --page raw --csv > gpurun_out/ncu_full_gemm2_cfg2_step_raw.csv 2>/dev/null
rm -f gpurun_out/gemm2_cfg2_step.ncu-rep
python tools/graph_timeline.py cfg2 > gpurun_out/timeline_cfg2.log 2>&1
python tools/graph_timeline.py cfg3 > gpurun_out/timeline_cfg3.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log
