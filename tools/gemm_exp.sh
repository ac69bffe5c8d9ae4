export FXN_GEMM_TRACE=1
T=tools/gemm_selftest
for args in "4096 512 5000 0 0 3 128 0" "4096 307 3000 0 0 3 0 0" "4096 384 3008 0 0 3 0 0" "4096 512 3000 0 0 3 128 0" "4096 307 5000 0 0 3 128 0" "4096 256 5000 0 0 3 128 0" "2048 512 5000 0 0 3 128 0" "4096 512 5000 0 0 3 64 0"; do
  $T one $args 2>&1 | grep -E "BENCH|trace|gemm2\]" | tail -13
done
