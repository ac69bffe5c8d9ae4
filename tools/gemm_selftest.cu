// Native self-test + micro-benchmark for fxn_gemm (tcgen05/TMA GEMM). Runs on a B200 without Python:
//   tools/gemm_selftest            correctness over all operand majorness / tail / split-K / epilogue cases
//   tools/gemm_selftest bench      adds timings of the BASELINE config-2 GEMM shapes
// The reference is a naive CUDA-core kernel with double accumulation on the original fp32 operands.
#include "../include/flexynesis_b200.h"
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

__global__ void fill_kernel(float* p, long long n, unsigned seed, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u ^ seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = ((x >> 8) * (1.0f / 8388608.0f) - 1.0f) * scale;  // uniform(-scale, scale)
  }
}

// A(m,k): K-major storage a[m*lda+k], MN-major storage a[k*lda+m]; same for B(n,k).
__global__ void ref_gemm_kernel(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn,
                                const float* bias, double* C, int M, int N, int K) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= M || n >= N) return;
  double acc = 0.0;
  for (int k = 0; k < K; ++k) {
    float a = a_mn ? A[(long long)k * lda + m] : A[(long long)m * lda + k];
    float b = b_mn ? B[(long long)k * ldb + n] : B[(long long)n * ldb + k];
    acc += (double)a * (double)b;
  }
  if (bias) acc += bias[n];
  C[(long long)m * N + n] = acc;
}

struct Case {
  int M, N, K, a_mn, b_mn, nterms, bias, splitk, stats, planes, block_n;
  int fix = 0;         // 1: give the call a fix-up workspace (stream-K with a fused epilogue)
  int groups = 0;      // fxn_gemm_desc.max_groups
};

static long long r8(long long x) { return (x + 7) / 8 * 8; }

static int run_case(const Case& c, bool verbose) {
  const long long a_rows = c.a_mn ? c.K : c.M, a_cols = c.a_mn ? c.M : c.K;
  const long long b_rows = c.b_mn ? c.K : c.N, b_cols = c.b_mn ? c.N : c.K;
  const long long lda32 = a_cols + 3, ldb32 = b_cols + 1;  // deliberately odd fp32 strides
  const long long lda = r8(a_cols), ldb = r8(b_cols), ldp = r8(c.N) + 8, ldc = c.N + (c.N % 4 ? 1 : 4);
  float *A, *B, *C, *bias = nullptr, *stats = nullptr;
  double* Cref;
  __nv_bfloat16 *Ah, *Al, *Bh, *Bl, *Ch = nullptr, *Cl = nullptr;
  CK(cudaMalloc(&A, a_rows * lda32 * 4));
  CK(cudaMalloc(&B, b_rows * ldb32 * 4));
  CK(cudaMalloc(&C, (long long)c.M * ldc * 4));
  CK(cudaMalloc(&Cref, (long long)c.M * c.N * 8));
  CK(cudaMalloc(&Ah, a_rows * lda * 2));
  CK(cudaMalloc(&Al, a_rows * lda * 2));
  CK(cudaMalloc(&Bh, b_rows * ldb * 2));
  CK(cudaMalloc(&Bl, b_rows * ldb * 2));
  CK(cudaMemset(C, 0xFF, (long long)c.M * ldc * 4));
  fill_kernel<<<512, 256>>>(A, a_rows * lda32, 1234u + c.M, 1.0f);
  fill_kernel<<<512, 256>>>(B, b_rows * ldb32, 99u + c.N, 0.05f);
  if (c.bias) {
    CK(cudaMalloc(&bias, c.N * 4));
    fill_kernel<<<8, 256>>>(bias, c.N, 7u, 0.5f);
  }
  const int mt = fxn_gemm_stat_tiles(c.M);
  if (c.stats) CK(cudaMalloc(&stats, (long long)mt * 2 * c.N * 4));
  if (c.planes) {
    CK(cudaMalloc(&Ch, (long long)c.M * ldp * 2));
    CK(cudaMalloc(&Cl, (long long)c.M * ldp * 2));
  }
  int rc = fxn_split_planes(A, lda32, a_rows, a_cols, Ah, Al, lda, 0);
  rc |= fxn_split_planes(B, ldb32, b_rows, b_cols, Bh, Bl, ldb, 0);
  if (rc) { printf("split failed: %s\n", fxn_last_error()); return 1; }

  fxn_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = c.M; d.N = c.N; d.K = c.K;
  d.a_hi = Ah; d.a_lo = Al; d.lda = lda; d.a_mn_major = c.a_mn;
  d.b_hi = Bh; d.b_lo = Bl; d.ldb = ldb; d.b_mn_major = c.b_mn;
  d.nterms = c.nterms;
  d.C = C; d.ldc = ldc; d.bias = bias;
  d.c_hi = Ch; d.c_lo = Cl; d.ldp = ldp;
  d.colstats = stats; d.stats_mode = 2;
  d.splitk = c.splitk; d.block_n = c.block_n;
  d.max_groups = c.groups;
  float* fix_ws = nullptr; unsigned* fix_flags = nullptr;
  const long long fix_bytes = fxn_gemm_fix_ws_bytes();
  const int fix_words = fxn_gemm_fix_flag_words();
  if (c.fix) {
    CK(cudaMalloc(&fix_ws, fix_bytes)); CK(cudaMemset(fix_ws, 0, fix_bytes));
    CK(cudaMalloc(&fix_flags, fix_words * 4)); CK(cudaMemset(fix_flags, 0, fix_words * 4));
    d.fix_ws = fix_ws; d.fix_ws_bytes = fix_bytes; d.fix_flags = fix_flags; d.fix_flags_count = fix_words;
    rc = fxn_gemm(&d, 0);             // a first launch: the checked one below must find the workspace handed back clean
    if (rc) { printf("fxn_gemm failed: %s\n", fxn_last_error()); return 1; }
    CK(cudaMemset(C, 0xFF, (long long)c.M * ldc * 4));
  }
  rc = fxn_gemm(&d, 0);
  if (rc) { printf("fxn_gemm failed: %s\n", fxn_last_error()); return 1; }
  dim3 rb(32, 8), rg((c.N + 31) / 32, (c.M + 7) / 8);
  ref_gemm_kernel<<<rg, rb>>>(A, lda32, c.a_mn, B, ldb32, c.b_mn, bias, Cref, c.M, c.N, c.K);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }

  std::vector<float> hC((size_t)c.M * ldc);
  std::vector<double> hR((size_t)c.M * c.N);
  CK(cudaMemcpy(hC.data(), C, hC.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hR.data(), Cref, hR.size() * 8, cudaMemcpyDeviceToHost));
  double num = 0, den = 0, maxerr = 0;
  for (int m = 0; m < c.M; ++m)
    for (int n = 0; n < c.N; ++n) {
      double r = hR[(size_t)m * c.N + n], g = hC[(size_t)m * ldc + n];
      double dd = g - r;
      if (!(std::fabs(dd) <= 1e30)) dd = 1e30;
      num += dd * dd; den += r * r;
      if (std::fabs(dd) > maxerr) maxerr = std::fabs(dd);
    }
  const double rms = std::sqrt(den / ((double)c.M * c.N));
  const double relrms = std::sqrt(num / (den + 1e-300));
  const double tol = c.nterms == 3 ? 2e-5 : 1.5e-2;
  bool ok = relrms < tol && maxerr / rms < tol * 20;
  // untouched padding of C must stay 0xFF bytes (NaN pattern)
  for (int m = 0; m < c.M && ok; ++m)
    for (long long n = c.N; n < ldc; ++n) {
      unsigned u; memcpy(&u, &hC[(size_t)m * ldc + n], 4);
      if (u != 0xFFFFFFFFu) { ok = false; printf("  C padding overwritten at (%d,%lld)\n", m, n); break; }
    }
  if (c.fix && ok) {
    std::vector<float> hw(fix_bytes / 4);
    std::vector<unsigned> hf(fix_words);
    CK(cudaMemcpy(hw.data(), fix_ws, fix_bytes, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hf.data(), fix_flags, fix_words * 4, cudaMemcpyDeviceToHost));
    (void)hw;
    for (unsigned v : hf) if (v != 0u) { ok = false; printf("  fix-up flags not reset\n"); break; }
  }
  double perr = 0, serr = 0;
  if (c.planes && ok) {
    std::vector<__nv_bfloat16> hh((size_t)c.M * ldp), hl((size_t)c.M * ldp);
    CK(cudaMemcpy(hh.data(), Ch, hh.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hl.data(), Cl, hl.size() * 2, cudaMemcpyDeviceToHost));
    for (int m = 0; m < c.M; ++m)
      for (int n = 0; n < c.N; ++n) {
        double v = (double)__bfloat162float(hh[(size_t)m * ldp + n]) + (double)__bfloat162float(hl[(size_t)m * ldp + n]);
        double dd = std::fabs(v - hC[(size_t)m * ldc + n]);
        if (dd > perr) perr = dd;
      }
    if (perr / rms > 3e-5) { ok = false; }
  }
  if (c.stats && ok) {
    std::vector<float> hs((size_t)mt * 2 * c.N);
    CK(cudaMemcpy(hs.data(), stats, hs.size() * 4, cudaMemcpyDeviceToHost));
    for (int t = 0; t < mt; ++t) {
      int r0 = t * 128, r1 = std::min(c.M, r0 + 128);
      for (int n = 0; n < c.N; ++n) {
        double s = 0, m2 = 0;
        for (int r = r0; r < r1; ++r) s += hC[(size_t)r * ldc + n];
        double mu = s / (r1 - r0);
        for (int r = r0; r < r1; ++r) { double q = hC[(size_t)r * ldc + n] - mu; m2 += q * q; }
        double e1 = std::fabs(hs[((size_t)t * 2) * c.N + n] - s) / (rms * (r1 - r0));
        double e2 = std::fabs(hs[((size_t)t * 2 + 1) * c.N + n] - m2) / (rms * rms * (r1 - r0));
        serr = std::max(serr, std::max(e1, e2));
      }
    }
    if (serr > 1e-5) ok = false;
  }
  if (verbose || !ok)
    printf("%s M=%d N=%d K=%d a_mn=%d b_mn=%d terms=%d bias=%d splitk=%d bn=%d fix=%d groups=%d  relrms=%.3e maxerr/rms=%.3e planes=%.2e stats=%.2e\n",
           ok ? "PASS" : "FAIL", c.M, c.N, c.K, c.a_mn, c.b_mn, c.nterms, c.bias, c.splitk, c.block_n, c.fix, c.groups,
           relrms, maxerr / rms, perr / rms, serr);
  if (fix_ws) { cudaFree(fix_ws); cudaFree(fix_flags); }
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(Cref); cudaFree(Ah); cudaFree(Al); cudaFree(Bh); cudaFree(Bl);
  if (bias) cudaFree(bias);
  if (stats) cudaFree(stats);
  if (Ch) { cudaFree(Ch); cudaFree(Cl); }
  return ok ? 0 : 1;
}

__global__ void fill_bf16_kernel(__nv_bfloat16* p, long long n, unsigned seed, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u ^ seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = __float2bfloat16(((x >> 8) * (1.0f / 8388608.0f) - 1.0f) * scale);
  }
}

static void bench_case(const char* name, int M, int N, int K, int a_mn, int b_mn, int nterms, int block_n, int splitk) {
  const long long a_rows = a_mn ? K : M, a_cols = a_mn ? M : K;
  const long long b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
  // FXN_TEST_LDALIGN (elements): leading-dimension alignment of the operand planes (8 = the minimum TMA accepts; 64 makes
  // every 128-byte box row start on a cache line)
  const long long al = getenv("FXN_TEST_LDALIGN") ? atoll(getenv("FXN_TEST_LDALIGN")) : 8;
  const long long lda = (a_cols + al - 1) / al * al, ldb = (b_cols + al - 1) / al * al;
  __nv_bfloat16 *Ah, *Al, *Bh, *Bl;
  float* C;
  CK(cudaMalloc(&Ah, a_rows * lda * 2)); CK(cudaMalloc(&Al, a_rows * lda * 2));
  CK(cudaMalloc(&Bh, b_rows * ldb * 2)); CK(cudaMalloc(&Bl, b_rows * ldb * 2));
  CK(cudaMalloc(&C, (long long)M * N * 4));
  // random operands: all-zero inputs draw less power and flatter the clocks
  fill_bf16_kernel<<<1024, 256>>>(Ah, a_rows * lda, 1u, 1.0f); fill_bf16_kernel<<<1024, 256>>>(Al, a_rows * lda, 2u, 0.004f);
  fill_bf16_kernel<<<1024, 256>>>(Bh, b_rows * ldb, 3u, 1.0f); fill_bf16_kernel<<<1024, 256>>>(Bl, b_rows * ldb, 4u, 0.004f);
  fxn_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = M; d.N = N; d.K = K;
  d.a_hi = Ah; d.a_lo = Al; d.lda = lda; d.a_mn_major = a_mn;
  d.b_hi = Bh; d.b_lo = Bl; d.ldb = ldb; d.b_mn_major = b_mn;
  d.nterms = nterms; d.C = C; d.ldc = N; d.block_n = block_n; d.splitk = splitk;
  float* fix_ws = nullptr; unsigned* fix_flags = nullptr;
  if (getenv("FXN_BENCH_FIX")) {
    CK(cudaMalloc(&fix_ws, fxn_gemm_fix_ws_bytes())); CK(cudaMemset(fix_ws, 0, fxn_gemm_fix_ws_bytes()));
    CK(cudaMalloc(&fix_flags, fxn_gemm_fix_flag_words() * 4)); CK(cudaMemset(fix_flags, 0, fxn_gemm_fix_flag_words() * 4));
    d.fix_ws = fix_ws; d.fix_ws_bytes = fxn_gemm_fix_ws_bytes(); d.fix_flags = fix_flags; d.fix_flags_count = fxn_gemm_fix_flag_words();
    if (getenv("FXN_BENCH_GROUPS")) d.max_groups = atoi(getenv("FXN_BENCH_GROUPS"));
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) fxn_gemm(&d, 0);
  CK(cudaDeviceSynchronize());
  const int iters = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) fxn_gemm(&d, 0);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double us = ms * 1000.0 / iters;
  double tf = 2.0 * M * N * K / (us * 1e-6) / 1e12;
  printf("BENCH %-28s M=%d N=%d K=%d terms=%d bn=%d splitk=%d : %.1f us  %.1f TFLOP/s algorithmic (%.1f issued)\n", name, M, N,
         K, nterms, block_n, splitk, us, tf, tf * nterms);
  if (getenv("FXN_GEMM_TRACE")) {
    long long t[32];
    if (fxn_debug_gemm_trace(t) == 0) {
      const char* names[11] = {"start", "setup done", "first load issued (last seg)", "all loads issued", "first stage landed",
                               "first tile MMAs issued", "all MMAs issued", "first accumulator ready", "first tile stored",
                               "all tiles stored", "end"};
      for (int i = 0; i < 11; ++i) printf("  trace %-30s cta0 %8lld  cta1 %8lld cycles\n", names[i], t[i], t[16 + i]);
    }
    long long st[512], en[512];
    if (fxn_debug_gemm_cta_times(st, en, 512) == 0) {
      printf("  cta start/end us:");
      for (int i = 0; i < 160; i += 2)
        if (st[i] >= 0) printf(" %d:%.1f-%.1f", i, st[i] * 1e-3, en[i] * 1e-3);
      printf("\n");
    }
  }
  cudaFree(Ah); cudaFree(Al); cudaFree(Bh); cudaFree(Bl); cudaFree(C);
  if (fix_ws) { cudaFree(fix_ws); cudaFree(fix_flags); }
}

int main(int argc, char** argv) {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device: %s sm_%d%d, %d SMs, lib version %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount,
         fxn_version());
  if (argc >= 10 && !strcmp(argv[1], "one")) {   // one M N K a_mn b_mn nterms bn splitk : time a single shape (for ncu)
    bench_case("one", atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]),
               atoi(argv[9]));
    return 0;
  }
  std::vector<Case> cases = {
      // M, N, K, a_mn, b_mn, terms, bias, splitk, stats, planes, block_n
      {128, 128, 64, 0, 0, 1, 0, 1, 0, 0, 0},      // smallest K-major, single term
      {128, 128, 64, 0, 0, 3, 0, 1, 0, 0, 0},
      {128, 256, 256, 0, 0, 3, 1, 1, 1, 1, 0},
      {256, 64, 512, 0, 0, 3, 1, 1, 1, 1, 0},
      {384, 32, 256, 0, 0, 3, 1, 1, 1, 0, 0},      // head layer_1 shape (N = 32)
      {300, 307, 1000, 0, 0, 3, 1, 1, 1, 1, 0},    // ragged M, N, K
      {512, 128, 1000, 0, 0, 3, 1, 1, 1, 1, 0},    // config 1 encoder
      {128, 128, 128, 0, 1, 3, 0, 1, 0, 0, 0},     // dgrad: B MN-major
      {300, 307, 256, 0, 1, 3, 0, 1, 0, 1, 0},
      {256, 512, 5, 0, 1, 3, 0, 1, 0, 0, 0},       // dgrad through a 5-class head (K = 5)
      {128, 128, 128, 1, 1, 3, 0, 1, 0, 0, 0},     // wgrad: both MN-major
      {307, 1000, 300, 1, 1, 3, 0, 1, 0, 0, 0},
      {256, 512, 4096, 1, 1, 3, 0, 8, 0, 0, 0},    // split-K wgrad of the fusion block
      {128, 128, 192, 1, 0, 3, 0, 1, 0, 0, 0},     // A MN-major, B K-major
      {1024, 512, 5000, 0, 0, 3, 1, 1, 1, 0, 0},   // config 2 encoder forward (M reduced)
      {1024, 512, 5000, 0, 0, 3, 1, 1, 1, 0, 128}, // same with 128-wide tiles
      {512, 5000, 1024, 1, 1, 3, 0, 1, 0, 0, 0},   // config 2 wgrad (K reduced)
      {1024, 256, 512, 0, 0, 1, 1, 1, 0, 0, 0},    // single-term mode
      {4096, 2048, 128, 0, 0, 3, 1, 1, 1, 1, 0},   // more tiles than CTA groups: persistent loop + TMEM double buffer
      {333, 2500, 96, 0, 1, 3, 1, 1, 0, 1, 0},     // ragged rows inside a CTA pair, many column tiles, tiny K
      {3000, 96, 200, 0, 0, 3, 1, 1, 1, 1, 0},     // narrow N, partial last pair
      {512, 5000, 4096, 1, 1, 3, 0, -1, 0, 0, 0},  // stream-K wgrad
      {307, 3000, 1000, 1, 1, 3, 1, -1, 0, 0, 0},  // stream-K, ragged, with bias
      {4096, 128, 8000, 0, 0, 3, 1, -1, 0, 0, 0},  // stream-K, K-major, long K
      {100, 1000, 4096, 1, 1, 3, 0, -1, 0, 0, 0},  // single-CTA groups (M <= 128) with stream-K
      // stream-K with fix-up: fused epilogues (bias + BN partials + planes) on split tiles
      {1024, 512, 5000, 0, 0, 3, 1, 1, 1, 1, 0, 1, 0},    // config 2 encoder forward (M reduced): 8 tiles over 74 groups
      {4096, 512, 5000, 0, 0, 3, 1, 1, 1, 0, 0, 1, 54},   // the full layer on 54 groups (as launched beside encoder 1)
      {4096, 307, 3000, 0, 0, 3, 1, 1, 1, 1, 0, 1, 20},   // encoder 1 on 20 groups, ragged N
      {333, 300, 2000, 0, 0, 3, 1, 1, 1, 1, 0, 1, 7},     // ragged M / N / K, tiles split three ways
      {512, 128, 1000, 0, 0, 3, 1, 1, 1, 1, 0, 1, 0},     // config 1 encoder
      {100, 96, 3000, 0, 0, 3, 1, 1, 1, 1, 0, 1, 0},      // single-CTA groups (M <= 128)
      {2048, 1024, 4096, 0, 1, 3, 0, 1, 0, 1, 0, 1, 37},  // B MN-major (dgrad), odd group count
  };
  int fails = 0;
  for (const Case& c : cases) fails += run_case(c, true);
  printf("%s: %d/%zu cases failed\n", fails ? "SELFTEST FAILED" : "SELFTEST OK", fails, cases.size());
  if (argc > 1 && !strcmp(argv[1], "bench") && !fails) {
    bench_case("cfg2 enc0 fwd", 4096, 512, 5000, 0, 0, 3, 256, 1);
    setenv("FXN_BENCH_FIX", "1", 1);
    bench_case("cfg2 enc0 fwd fix-up", 4096, 512, 5000, 0, 0, 3, 0, 0);
    bench_case("cfg2 enc1 fwd fix-up", 4096, 307, 3000, 0, 0, 3, 0, 0);
    bench_case("cfg5 enc fwd fix-up", 4096, 1024, 24000, 0, 0, 3, 0, 0);
    bench_case("cfg3 decoder out fix-up", 4096, 5000, 512, 0, 0, 3, 0, 0);
    unsetenv("FXN_BENCH_FIX");
    bench_case("cfg2 enc0 fwd bn128", 4096, 512, 5000, 0, 0, 3, 128, 1);
    bench_case("cfg2 enc0 fwd splitk2", 4096, 512, 5000, 0, 0, 3, 256, 2);
    bench_case("cfg2 enc0 fwd 1-term", 4096, 512, 5000, 0, 0, 1, 256, 1);
    bench_case("cfg2 enc0 fwd auto", 4096, 512, 5000, 0, 0, 3, 0, 0);
    bench_case("cfg2 enc0 wgrad auto", 512, 5000, 4096, 1, 1, 3, 0, -1);
    bench_case("cfg2 enc1 wgrad auto", 307, 3000, 4096, 1, 1, 3, 0, -1);
    bench_case("cfg2 dgrad h<-latent", 4096, 512, 256, 0, 1, 3, 0, 0);
    bench_case("cfg3 decoder out", 4096, 5000, 512, 0, 0, 3, 0, 0);
    bench_case("cfg4 fc dgrad", 4096, 64000, 128, 0, 1, 3, 0, 0);
    bench_case("cfg5 enc wgrad auto", 1024, 24000, 4096, 1, 1, 3, 0, -1);
    bench_case("cfg2 enc0 wgrad", 512, 5000, 4096, 1, 1, 3, 256, 1);
    bench_case("cfg2 enc0 wgrad bn128", 512, 5000, 4096, 1, 1, 3, 128, 1);
    bench_case("cfg2 enc1 fwd", 4096, 307, 3000, 0, 0, 3, 0, 1);
    bench_case("cfg2 layer_out", 4096, 256, 512, 0, 0, 3, 0, 1);
    bench_case("cfg5 enc fwd", 4096, 1024, 24000, 0, 0, 3, 256, 1);
    bench_case("cfg5 enc wgrad", 1024, 24000, 4096, 1, 1, 3, 256, 1);
    bench_case("square 8192", 8192, 8192, 8192, 0, 0, 3, 256, 1);
    bench_case("square 8192 1-term", 8192, 8192, 8192, 0, 0, 1, 256, 1);
  }
  return fails ? 1 : 0;
}
