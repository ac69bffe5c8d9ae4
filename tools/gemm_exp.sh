export FXN_GEMM_TRACE=1
T=tools/gemm_selftest
for args in "512 5000 4096 1 1 3 0 -1" "4096 512 5000 0 0 3 128 0" "4096 1024 24000 0 0 3 0 0"; do
  $T one $args 2>&1 | grep -E "BENCH|cta start|gemm2\]" | tail -4
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
