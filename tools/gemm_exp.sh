# GEMM pipeline traces for a list of shapes (under gpurun): plan chosen, cluster occupancy, per-CTA start/end times and the
# milestone stamps of CTA 0/1. Each line: M N K a_mn b_mn nterms block_n splitk (see tools/gemm_selftest.cu `one`).
export FXN_GEMM_TRACE=1
T=tools/gemm_selftest
for args in "4096 512 5000 0 0 3 128 0" "4096 512 5000 0 0 3 256 0" "512 5000 4096 1 1 3 0 -1" "4096 1024 24000 0 0 3 0 0" "4096 256 512 0 0 3 0 0"; do
  $T one $args 2>&1 | grep -E "BENCH|trace|gemm2\]|cta start" | cut -c1-400
done
