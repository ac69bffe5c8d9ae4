timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/gpu_tests.log
for w in cfg1 cfg3 cfg4 cfg5; do python bench.py --workload $w > gpurun_out/r02_bench_v6_$w.json 2> gpurun_out/r02_bench_v6_$w.err; done
python bench.py > gpurun_out/r02_bench_v6_cfg2.json 2> gpurun_out/r02_bench_v6_cfg2.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench_v6_reference_cfg2.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
