// BatchNorm1d (+ activation + dropout) forward and backward as column-parallel elementwise kernels.
//
// Replaces aten::native_batch_norm(_backward), clamp_min/leaky_relu, bernoulli_/mul/div behind
//   MLP:              Linear -> BatchNorm1d -> ReLU -> Dropout(0.1)       flexynesis/modules.py:145-148   (order 0)
//   Encoder/Decoder:  Linear -> LeakyReLU(0.2) -> BatchNorm1d             flexynesis/modules.py:21-36     (order 1;
//                     the LeakyReLU runs in the producing GEMM's epilogue, this file normalises its output)
//   flexGCN:          conv -> BatchNorm1d(B*N rows) -> act -> Dropout(0.2) flexynesis/modules.py:253-257  (order 0)
// Batch statistics arrive as per-row-tile partials (sum, M2 about the tile mean) written by the producer
// (fxn_gemm epilogue or fxn_col_stats) and are merged with Chan's formula, so no extra pass over the
// activations is needed and there is no E[x^2]-E[x]^2 cancellation.
#include "fxn_internal.h"
#include "ptx.cuh"
#include "rng.cuh"
#include <cstdlib>

namespace fxn {

__device__ __forceinline__ float apply_act(float y, int act) {
  switch (act) {
    case 1: return fmaxf(y, 0.f);                          // relu
    case 2: return y > 0.f ? y : 0.01f * y;                // nn.LeakyReLU() default slope (flexGCN 'leakyrelu')
    case 3: return 1.f / (1.f + __expf(-y));               // sigmoid
    case 4: return tanhf(y);
    case 5: return 0.5f * y * (1.f + erff(y * 0.70710678118654752f));   // gelu (erf form)
    default: return y;
  }
}
// derivative of the activation given pre-activation y
__device__ __forceinline__ float act_grad(float y, int act) {
  switch (act) {
    case 1: return y > 0.f ? 1.f : 0.f;
    case 2: return y > 0.f ? 1.f : 0.01f;
    case 3: { const float s = 1.f / (1.f + __expf(-y)); return s * (1.f - s); }
    case 4: { const float t = tanhf(y); return 1.f - t * t; }
    case 5: return 0.5f * (1.f + erff(y * 0.70710678118654752f)) + y * 0.3989422804014327f * __expf(-0.5f * y * y);
    default: return 1.f;
  }
}

// Narrow, contiguous matrices (flexGCN normalises [B*N x 32]) are processed as [rows / fold x 64] so that all 256 threads
// of a block carry data; column c of the folded view belongs to channel c % period. The Philox element index is
// unchanged by the fold (r * cols / 8 + c / 8 is the same number in both views).
__device__ __forceinline__ int chan(int c, int period) { return period ? c % period : c; }

struct BnArgs {
  const float* V; long long ldv;     // input of the norm [rows x cols]
  long long rows; int cols;
  const float* partials; int ntiles; int tile_rows; int pld;   // [ntiles][2][pld], pld >= cols
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* num_batches;
  float momentum, eps;
  int train;
  int act;                // activation applied AFTER the norm (order 0); 0 = none
  float p_drop;           // dropout after the activation; 0 = none
  const uint8_t* mask; long long ldm;  // optional explicit keep mask (test replay)
  unsigned long long seed; const long long* seed_dev;
  float* out; long long ldo;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; long long ldp;
  float* saved;           // [2][cols]: mean, rstd
  int rpb;                // rows per block (multiple of 32)
  long long stat_rows;    // rows the statistics are taken over (= rows * fold)
  int pcols;              // number of real channels: the per-channel vectors have this length
  int period;             // 0, or pcols when `fold` consecutive rows of a contiguous narrow matrix are viewed as one row
  uint8_t* keep_bits;     // optional: keep flags per run of 8 columns, index = the Philox element index
};

constexpr int BN_COLS = 64;     // columns per block (8 per thread x 8 threads)
constexpr int BN_ROWS = 64;     // rows per block (2 per thread): many small blocks keep HBM requests in flight
constexpr int BN_THREADS = 256;
#ifndef BN_MIN_BLOCKS
#define BN_MIN_BLOCKS 4          // 64 registers per thread: 32 warps per SM instead of 24 (the kernels are latency / issue bound)
#endif

// Rows per block: 64 keeps many small blocks (and HBM requests) in flight for the [4096 x 512]-sized activations of the
// MLP encoders; tall inputs (flexGCN normalises B*N = 8 M rows of 32 channels) get fatter blocks so that the grid stays
// within a few waves of 148 SMs and the per-block prologue / column atomics stay negligible.
static inline int bn_rows_per_block(long long rows, int col_blocks) {
  static const int min_rows = [] { const char* e = getenv("FXN_BN_MIN_ROWS"); return e ? atoi(e) : BN_ROWS; }();
  static const int waves = [] { const char* e = getenv("FXN_BN_WAVES"); return e ? atoi(e) : 16; }();
  const long long target_blocks = 148LL * waves;
  long long rpb = (rows * col_blocks + target_blocks - 1) / target_blocks;
  rpb = (rpb + 31) / 32 * 32;
  if (rpb < min_rows) rpb = min_rows;
  return static_cast<int>(rpb);
}

// merge tile partials for one column -> (mean, biased var)
__device__ __forceinline__ void merge_stats(const float* partials, int ntiles, int tile_rows, long long rows, int cols,
                                            int c, float& mean, float& var) {   // `cols` = leading dim of partials
  float total = 0.f;
  for (int t = 0; t < ntiles; ++t) total += partials[(static_cast<long long>(t) * 2) * cols + c];
  mean = total / static_cast<float>(rows);
  float m2 = 0.f;
  for (int t = 0; t < ntiles; ++t) {
    const long long r0 = static_cast<long long>(t) * tile_rows;
    const float n = static_cast<float>(min(static_cast<long long>(tile_rows), rows - r0));
    const float s = partials[(static_cast<long long>(t) * 2) * cols + c];
    const float d = s / n - mean;
    m2 += partials[(static_cast<long long>(t) * 2 + 1) * cols + c] + n * d * d;
  }
  var = m2 / static_cast<float>(rows);
}

__global__ void __launch_bounds__(BN_THREADS, BN_MIN_BLOCKS) bn_fwd_kernel(const BnArgs a) {
  __shared__ float s_scale[BN_COLS], s_shift[BN_COLS];
  __shared__ float s_red[4][BN_COLS];
  __shared__ float s_mean[BN_COLS];
  const int c0 = blockIdx.x * BN_COLS;
  {
    // batch statistics: 4 threads per column stride over the tile partials (Chan merge in two passes)
    const int col = threadIdx.x & (BN_COLS - 1), lane4 = threadIdx.x >> 6;
    const int c = c0 + col;
    const bool ok = c < a.cols;
    const int pc = chan(c, a.period);
    if (a.train) {
      // Chan merge of the tile partials. Each of the 4 threads of a column fetches its share with up to 8 INDEPENDENT loads
      // per batch: one dependent load per iteration cost an L2 round trip each and ~5 us of a 13 us kernel at 32 tiles.
      float part = 0.f;
      float sv[8], mv[8];
      if (ok)
        for (int t0 = lane4; t0 < a.ntiles; t0 += 32) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int t = t0 + 4 * u;
            sv[u] = t < a.ntiles ? a.partials[(static_cast<long long>(t) * 2) * a.pld + pc] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) part += sv[u];
        }
      s_red[lane4][col] = part;
      __syncthreads();
      if (lane4 == 0) s_mean[col] = (s_red[0][col] + s_red[1][col] + s_red[2][col] + s_red[3][col]) / static_cast<float>(a.stat_rows);
      __syncthreads();
      const float mean = s_mean[col];
      float m2 = 0.f;
      if (ok)
        for (int t0 = lane4; t0 < a.ntiles; t0 += 32) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int t = t0 + 4 * u;
            const bool in = t < a.ntiles;
            sv[u] = in ? a.partials[(static_cast<long long>(t) * 2) * a.pld + pc] : 0.f;
            mv[u] = in ? a.partials[(static_cast<long long>(t) * 2 + 1) * a.pld + pc] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int t = t0 + 4 * u;
            if (t < a.ntiles) {
              const long long r0 = static_cast<long long>(t) * a.tile_rows;
              const float n = static_cast<float>(min(static_cast<long long>(a.tile_rows), a.stat_rows - r0));
              const float d = sv[u] / n - mean;
              m2 += mv[u] + n * d * d;
            }
          }
        }
      __syncthreads();
      s_red[lane4][col] = m2;
      __syncthreads();
    }
    if (lane4 == 0) {
      float sc = 0.f, sh = 0.f;
      if (ok) {
        float mean, var;
        if (a.train) {
          mean = s_mean[col];
          var = (s_red[0][col] + s_red[1][col] + s_red[2][col] + s_red[3][col]) / static_cast<float>(a.stat_rows);
        } else {
          mean = a.running_mean[pc];
          var = a.running_var[pc];
        }
        const float rstd = rsqrtf(var + a.eps);
        sc = a.gamma[pc] * rstd;
        sh = a.beta[pc] - mean * sc;
        if (a.train && blockIdx.y == 0 && c == pc) {      // one writer per channel (folded views repeat channels)
          if (a.saved) { a.saved[pc] = mean; a.saved[a.pcols + pc] = rstd; }
          if (a.running_mean) {
            const float n = static_cast<float>(a.stat_rows);
            a.running_mean[pc] = (1.f - a.momentum) * a.running_mean[pc] + a.momentum * mean;
            a.running_var[pc] = (1.f - a.momentum) * a.running_var[pc] + a.momentum * var * n / (n - 1.f);
          }
        }
      }
      s_scale[col] = sc;
      s_shift[col] = sh;
    }
  }
  if (a.train && a.num_batches && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *a.num_batches += 1;
  __syncthreads();

  const int tc = (threadIdx.x & 7) * 8;            // column offset inside the slab
  const int tr = threadIdx.x >> 3;                 // 0..31
  const int c = c0 + tc;
  const int pcols = (a.cols + 7) & ~7;             // planes are zero-filled up to pad8(cols) only
  if (c >= pcols || (c >= a.cols && a.out_hi == nullptr)) return;
  const long long r_begin = static_cast<long long>(blockIdx.y) * a.rpb;
  const long long r_end = min(a.rows, r_begin + a.rpb);
  const bool drop = a.train && a.p_drop > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  const bool vec_in = ((reinterpret_cast<uintptr_t>(a.V) & 15) == 0) && (a.ldv % 4 == 0) && (c + 8 <= a.cols);
  for (long long r = r_begin + tr; r < r_end; r += 32) {
    float x[8];
    const float* src = a.V + r * a.ldv + c;
    if (vec_in) {
      const float4 u = *reinterpret_cast<const float4*>(src);
      const float4 w = *reinterpret_cast<const float4*>(src + 4);
      x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w; x[4] = w.x; x[5] = w.y; x[6] = w.z; x[7] = w.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (c + j < a.cols) ? src[j] : 0.f;
    }
    uint32_t keep = 0xFFu;
    if (drop) {
      if (a.mask) {
        keep = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c + j < a.cols && a.mask[r * a.ldm + c + j]) keep |= 1u << j;
      } else {
        keep = dropout_keep8(step_seed(a.seed, a.seed_dev),
                             static_cast<unsigned long long>(r) * ((a.cols + 7) / 8) + (c >> 3), a.p_drop);
      }
      if (a.keep_bits) a.keep_bits[static_cast<unsigned long long>(r) * ((a.cols + 7) / 8) + (c >> 3)] = static_cast<uint8_t>(keep);
    }
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = fmaf(x[j], s_scale[tc + j], s_shift[tc + j]);
      v = apply_act(v, a.act);
      v = ((keep >> j) & 1u) ? v * keep_scale : 0.f;
      y[j] = (c + j < a.cols) ? v : 0.f;
    }
    if (a.out) {
      float* dst = a.out + r * a.ldo + c;
      if (((reinterpret_cast<uintptr_t>(a.out) & 15) == 0) && (a.ldo % 4 == 0) && (c + 8 <= a.cols)) {
        *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(y[4], y[5], y[6], y[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c + j < a.cols) dst[j] = y[j];
      }
    }
    if (a.out_hi && c < pcols) {
      __align__(16) __nv_bfloat16 h[8];
      __align__(16) __nv_bfloat16 l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_bf16(y[j], h[j], l[j]);
      *reinterpret_cast<uint4*>(a.out_hi + r * a.ldp + c) = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(a.out_lo + r * a.ldp + c) = *reinterpret_cast<const uint4*>(l);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct BnBwdArgs {
  const float* V; long long ldv;        // forward input of the norm
  const float* dOut; long long ldg;     // gradient wrt the block output (after act/dropout)
  long long rows; int cols;
  const float* gamma; const float* beta;
  const float* saved;                   // [2][cols] mean, rstd
  int act; float p_drop;
  const uint8_t* mask; long long ldm; unsigned long long seed; const long long* seed_dev;
  int pre_act;                          // order 1: V = leaky_relu_0.2(Z); multiply dV by (V > 0 ? 1 : 0.2) to get dZ
  float* sums;                          // [3][cols]: sum g, sum g*xhat (pass 1), sum dZ (pass 2, optional)
  float* dgamma; float* dbeta; float* dbias;    // outputs written in pass 2 by blockIdx.y == 0 (dbias may be null)
  float* dV; long long ldd;             // optional fp32 gradient wrt the Linear output
  __nv_bfloat16* dv_hi; __nv_bfloat16* dv_lo; long long ldp;
  float grad_scale;                     // multiplies dOut (1 unless the caller folds a loss weight in)
  int acc_affine;                       // dgamma/dbeta += (module applied several times per step)
  int rpb;                              // rows per block (multiple of 32)
  long long stat_rows; int pcols, period;   // see BnArgs
  const uint8_t* keep_bits;                 // optional: flags stored by the forward pass (replaces mask / Philox)
};

// recompute g = dOut * dropout * act'(y) for 8 columns of row r; also returns xhat
__device__ __forceinline__ void bn_bwd_load(const BnBwdArgs& a, long long r, int c, const float* s_mean,
                                            const float* s_rstd, const float* s_gamma, const float* s_beta, int tc,
                                            float* g, float* xhat) {
  const bool drop = a.p_drop > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - a.p_drop) : 1.f;
  uint32_t keep = 0xFFu;
  if (drop) {
    if (a.keep_bits) {
      keep = a.keep_bits[static_cast<unsigned long long>(r) * ((a.cols + 7) / 8) + (c >> 3)];
    } else if (a.mask) {
      keep = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c + j < a.cols && a.mask[r * a.ldm + c + j]) keep |= 1u << j;
    } else {
      keep = dropout_keep8(step_seed(a.seed, a.seed_dev),
                           static_cast<unsigned long long>(r) * ((a.cols + 7) / 8) + (c >> 3), a.p_drop);
    }
  }
  float vv[8], gg[8];
  const bool vec = (c + 8 <= a.cols) && (a.ldv % 4 == 0) && (a.ldg % 4 == 0) &&
                   (((reinterpret_cast<uintptr_t>(a.V) | reinterpret_cast<uintptr_t>(a.dOut)) & 15) == 0);
  if (vec) {
    const float4 v0 = *reinterpret_cast<const float4*>(a.V + r * a.ldv + c), v1 = *reinterpret_cast<const float4*>(a.V + r * a.ldv + c + 4);
    const float4 g0 = *reinterpret_cast<const float4*>(a.dOut + r * a.ldg + c), g1 = *reinterpret_cast<const float4*>(a.dOut + r * a.ldg + c + 4);
    vv[0] = v0.x; vv[1] = v0.y; vv[2] = v0.z; vv[3] = v0.w; vv[4] = v1.x; vv[5] = v1.y; vv[6] = v1.z; vv[7] = v1.w;
    gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      vv[j] = (c + j < a.cols) ? a.V[r * a.ldv + c + j] : 0.f;
      gg[j] = (c + j < a.cols) ? a.dOut[r * a.ldg + c + j] : 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (c + j < a.cols) {
      const float v = vv[j];
      const float xh = (v - s_mean[tc + j]) * s_rstd[tc + j];
      const float y = fmaf(xh, s_gamma[tc + j], s_beta[tc + j]);
      float d = gg[j] * a.grad_scale;
      d = ((keep >> j) & 1u) ? d * keep_scale : 0.f;
      g[j] = d * act_grad(y, a.act);
      xhat[j] = xh;
    } else {
      g[j] = 0.f;
      xhat[j] = 0.f;
    }
  }
}

__global__ void __launch_bounds__(BN_THREADS, BN_MIN_BLOCKS) bn_bwd_reduce_kernel(const BnBwdArgs a) {
  __shared__ float s_mean[BN_COLS], s_rstd[BN_COLS], s_gamma[BN_COLS], s_beta[BN_COLS];
  __shared__ float s_acc[2][32][BN_COLS + 1];
  const int c0 = blockIdx.x * BN_COLS;
  if (threadIdx.x < BN_COLS) {
    const int c = c0 + threadIdx.x;
    const bool ok = c < a.cols;
    const int pc = chan(c, a.period);
    s_mean[threadIdx.x] = ok ? a.saved[pc] : 0.f;
    s_rstd[threadIdx.x] = ok ? a.saved[a.pcols + pc] : 0.f;
    s_gamma[threadIdx.x] = ok ? a.gamma[pc] : 0.f;
    s_beta[threadIdx.x] = ok ? a.beta[pc] : 0.f;
  }
  __syncthreads();
  const int tc = (threadIdx.x & 7) * 8;
  const int tr = threadIdx.x >> 3;
  const int c = c0 + tc;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  const long long r_begin = static_cast<long long>(blockIdx.y) * a.rpb;
  const long long r_end = min(a.rows, r_begin + a.rpb);
  if (c < a.cols) {
    for (long long r = r_begin + tr; r < r_end; r += 32) {
      float g[8], xh[8];
      bn_bwd_load(a, r, c, s_mean, s_rstd, s_gamma, s_beta, tc, g, xh);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1[j] += g[j]; s2[j] = fmaf(g[j], xh[j], s2[j]); }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { s_acc[0][tr][tc + j] = s1[j]; s_acc[1][tr][tc + j] = s2[j]; }
  __syncthreads();
  if (threadIdx.x < 2 * BN_COLS) {
    const int which = threadIdx.x / BN_COLS, col = threadIdx.x % BN_COLS;
    if (c0 + col < a.cols) {
      float t = 0.f;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) t += s_acc[which][i][col];
      atomicAdd(a.sums + static_cast<long long>(which) * a.pcols + chan(c0 + col, a.period), t);
    }
  }
}

__global__ void __launch_bounds__(BN_THREADS, BN_MIN_BLOCKS) bn_bwd_apply_kernel(const BnBwdArgs a) {
  __shared__ float s_mean[BN_COLS], s_rstd[BN_COLS], s_gamma[BN_COLS], s_beta[BN_COLS], s_m1[BN_COLS], s_m2[BN_COLS];
  __shared__ float s_acc[32][BN_COLS + 1];
  const int c0 = blockIdx.x * BN_COLS;
  const float inv_n = 1.f / static_cast<float>(a.stat_rows);
  if (threadIdx.x < BN_COLS) {
    const int c = c0 + threadIdx.x;
    const bool ok = c < a.cols;
    const int pc = chan(c, a.period);
    s_mean[threadIdx.x] = ok ? a.saved[pc] : 0.f;
    s_rstd[threadIdx.x] = ok ? a.saved[a.pcols + pc] : 0.f;
    s_gamma[threadIdx.x] = ok ? a.gamma[pc] : 0.f;
    s_beta[threadIdx.x] = ok ? a.beta[pc] : 0.f;
    const float sum_g = ok ? a.sums[pc] : 0.f, sum_gx = ok ? a.sums[a.pcols + pc] : 0.f;
    s_m1[threadIdx.x] = sum_g * inv_n;
    s_m2[threadIdx.x] = sum_gx * inv_n;
    if (ok && blockIdx.y == 0 && c == pc) {
      if (a.dbeta) a.dbeta[pc] = (a.acc_affine ? a.dbeta[pc] : 0.f) + sum_g;
      if (a.dgamma) a.dgamma[pc] = (a.acc_affine ? a.dgamma[pc] : 0.f) + sum_gx;
    }
  }
  __syncthreads();
  const int tc = (threadIdx.x & 7) * 8;
  const int tr = threadIdx.x >> 3;
  const int c = c0 + tc;
  float sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sb[j] = 0.f;
  const long long r_begin = static_cast<long long>(blockIdx.y) * a.rpb;
  const long long r_end = min(a.rows, r_begin + a.rpb);
  const int pcols = (a.cols + 7) & ~7;
  if (c < a.cols || (a.dv_hi && c < pcols)) {
    for (long long r = r_begin + tr; r < r_end; r += 32) {
      float g[8], xh[8], dz[8];
      bn_bwd_load(a, r, c, s_mean, s_rstd, s_gamma, s_beta, tc, g, xh);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = s_gamma[tc + j] * s_rstd[tc + j] * (g[j] - s_m1[tc + j] - xh[j] * s_m2[tc + j]);
        if (a.pre_act && c + j < a.cols) d *= (a.V[r * a.ldv + c + j] > 0.f) ? 1.f : 0.2f;
        dz[j] = (c + j < a.cols) ? d : 0.f;
        sb[j] += dz[j];
      }
      if (a.dV) {
        if ((c + 8 <= a.cols) && (a.ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.dV) & 15) == 0)) {
          *reinterpret_cast<float4*>(a.dV + r * a.ldd + c) = make_float4(dz[0], dz[1], dz[2], dz[3]);
          *reinterpret_cast<float4*>(a.dV + r * a.ldd + c + 4) = make_float4(dz[4], dz[5], dz[6], dz[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c + j < a.cols) a.dV[r * a.ldd + c + j] = dz[j];
        }
      }
      if (a.dv_hi && c < pcols) {
        __align__(16) __nv_bfloat16 h[8];
        __align__(16) __nv_bfloat16 l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split_bf16(dz[j], h[j], l[j]);
        *reinterpret_cast<uint4*>(a.dv_hi + r * a.ldp + c) = *reinterpret_cast<const uint4*>(h);
        *reinterpret_cast<uint4*>(a.dv_lo + r * a.ldp + c) = *reinterpret_cast<const uint4*>(l);
      }
    }
  }
  if (a.dbias) {   // column sum of dZ (non-zero only when an activation sits between the bias and the norm)
#pragma unroll
    for (int j = 0; j < 8; ++j) s_acc[tr][tc + j] = sb[j];
    __syncthreads();
    if (threadIdx.x < BN_COLS && c0 + threadIdx.x < a.cols) {
      float t = 0.f;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) t += s_acc[i][threadIdx.x];
      atomicAdd(a.dbias + chan(c0 + threadIdx.x, a.period), t);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// stand-alone column statistics (for inputs that do not come out of an fxn_gemm epilogue)
// partials layout identical to the GEMM's: [ntiles][2][cols] with tile_rows rows per tile
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ V, long long ldv, long long rows,
                                                        int cols, int tile_rows, float* __restrict__ partials) {
  // block = (64 columns) x (one row tile); 256 threads = 4 row-lanes x 64 columns
  __shared__ float s_sum[4][64], s_m2[4][64];
  const int col = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + col;
  const long long r0 = static_cast<long long>(blockIdx.y) * tile_rows;
  const long long r1 = min(rows, r0 + tile_rows);
  float s = 0.f;
  if (c < cols)
    for (long long r = r0 + rl; r < r1; r += 4) s += V[r * ldv + c];
  s_sum[rl][col] = s;
  __syncthreads();
  const float total = s_sum[0][col] + s_sum[1][col] + s_sum[2][col] + s_sum[3][col];
  const float mu = total / static_cast<float>(r1 - r0);
  float m2 = 0.f;
  if (c < cols)
    for (long long r = r0 + rl; r < r1; r += 4) {
      const float d = V[r * ldv + c] - mu;
      m2 = fmaf(d, d, m2);
    }
  s_m2[rl][col] = m2;
  __syncthreads();
  if (rl == 0 && c < cols) {
    partials[(static_cast<long long>(blockIdx.y) * 2) * cols + c] = total;
    partials[(static_cast<long long>(blockIdx.y) * 2 + 1) * cols + c] = s_m2[0][col] + s_m2[1][col] + s_m2[2][col] + s_m2[3][col];
  }
}

}  // namespace fxn

using namespace fxn;

extern "C" int fxn_bn_act_fwd(const fxn_bn_fwd_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d || !d->V || !d->gamma || !d->beta) return set_error(FXN_ERR_ARG, "fxn_bn_act_fwd: null argument");
  if (d->rows <= 0 || d->cols <= 0) return set_error(FXN_ERR_ARG, "fxn_bn_act_fwd: empty input");
  if (d->train && d->rows < 2)
    return set_error(FXN_ERR_ARG, "Expected more than 1 value per channel when training");   // as torch raises
  if (d->train && !d->partials) return set_error(FXN_ERR_ARG, "fxn_bn_act_fwd: train mode needs column partials");
  if (!d->train && (!d->running_mean || !d->running_var))
    return set_error(FXN_ERR_ARG, "fxn_bn_act_fwd: eval mode needs running statistics");
  if (d->out_hi && (!d->out_lo || d->ldp % 8 != 0)) return set_error(FXN_ERR_ARG, "fxn_bn_act_fwd: bad planes");
  BnArgs a;
  a.V = d->V; a.ldv = d->ldv; a.rows = d->rows; a.cols = d->cols;
  a.partials = d->partials; a.ntiles = d->ntiles; a.tile_rows = d->tile_rows;
  a.pld = d->partials_ld > 0 ? d->partials_ld : d->cols;
  a.gamma = d->gamma; a.beta = d->beta;
  a.running_mean = d->running_mean; a.running_var = d->running_var;
  a.num_batches = reinterpret_cast<long long*>(d->num_batches_tracked);
  a.momentum = d->momentum; a.eps = d->eps; a.train = d->train; a.act = d->act; a.p_drop = d->p_drop;
  a.mask = d->mask; a.ldm = d->ldm; a.seed = d->seed; a.seed_dev = static_cast<const long long*>(d->seed_dev);
  a.out = d->out; a.ldo = d->ldo;
  a.out_hi = static_cast<__nv_bfloat16*>(d->out_hi); a.out_lo = static_cast<__nv_bfloat16*>(d->out_lo); a.ldp = d->ldp;
  a.saved = d->saved;
  a.keep_bits = d->keep_bits;
  a.stat_rows = d->stat_rows > 0 ? d->stat_rows : a.rows; a.pcols = a.cols; a.period = 0;
  {
    // fold narrow contiguous matrices into 64-wide rows (see chan())
    const int f = (a.cols == 8 || a.cols == 16 || a.cols == 32) ? 64 / a.cols : 1;
    const bool contiguous = a.ldv == a.cols && (!a.out || a.ldo == a.cols) && (!a.out_hi || a.ldp == a.cols) &&
                            (!a.mask || a.ldm == a.cols);
    if (f > 1 && contiguous && a.rows % f == 0 && a.rows >= 4096) {
      a.period = a.cols; a.rows /= f; a.cols = 64;
      a.ldv = 64; a.ldo = 64; a.ldp = 64; a.ldm = 64;
    }
  }
  const int width = a.out_hi ? ((a.cols + 7) & ~7) : a.cols;
  a.rpb = bn_rows_per_block(a.rows, ceil_div(width, BN_COLS));
  dim3 grid(ceil_div(width, BN_COLS), ceil_div(a.rows, a.rpb));
  bn_fwd_kernel<<<grid, BN_THREADS, 0, stream>>>(a);
  FXN_CHECK_LAUNCH("bn_fwd");
  return 0;
}

extern "C" int fxn_bn_act_bwd(const fxn_bn_bwd_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d || !d->V || !d->dOut || !d->gamma || !d->beta || !d->saved || !d->sums)
    return set_error(FXN_ERR_ARG, "fxn_bn_act_bwd: null argument");
  if (d->dv_hi && (!d->dv_lo || d->ldp % 8 != 0)) return set_error(FXN_ERR_ARG, "fxn_bn_act_bwd: bad planes");
  BnBwdArgs a;
  a.V = d->V; a.ldv = d->ldv; a.dOut = d->dOut; a.ldg = d->ldg; a.rows = d->rows; a.cols = d->cols;
  a.gamma = d->gamma; a.beta = d->beta; a.saved = d->saved; a.act = d->act; a.p_drop = d->p_drop;
  a.mask = d->mask; a.ldm = d->ldm; a.seed = d->seed; a.seed_dev = static_cast<const long long*>(d->seed_dev);
  a.pre_act = d->pre_act;
  a.sums = d->sums; a.dgamma = d->dgamma; a.dbeta = d->dbeta; a.dbias = d->dbias;
  a.dV = d->dV; a.ldd = d->ldd;
  a.dv_hi = static_cast<__nv_bfloat16*>(d->dv_hi); a.dv_lo = static_cast<__nv_bfloat16*>(d->dv_lo); a.ldp = d->ldp;
  a.grad_scale = d->grad_scale == 0.f ? 1.f : d->grad_scale;
  a.acc_affine = d->accumulate_affine;
  a.keep_bits = d->keep_bits;
  a.stat_rows = d->stat_rows > 0 ? d->stat_rows : a.rows; a.pcols = a.cols; a.period = 0;
  if (d->phase < 0 || d->phase > 2) return set_error(FXN_ERR_ARG, "fxn_bn_act_bwd: phase must be 0, 1 or 2");
  {
    const int f = (a.cols == 8 || a.cols == 16 || a.cols == 32) ? 64 / a.cols : 1;
    const bool contiguous = a.ldv == a.cols && a.ldg == a.cols && (!a.dV || a.ldd == a.cols) &&
                            (!a.dv_hi || a.ldp == a.cols) && (!a.mask || a.ldm == a.cols);
    if (f > 1 && contiguous && a.rows % f == 0 && a.rows >= 4096) {
      a.period = a.cols; a.rows /= f; a.cols = 64;
      a.ldv = 64; a.ldg = 64; a.ldd = 64; a.ldp = 64; a.ldm = 64;
    }
  }
  const int width = a.dv_hi ? ((a.cols + 7) & ~7) : a.cols;
  a.rpb = bn_rows_per_block(a.rows, ceil_div(width, BN_COLS));
  if (d->phase != 2) {
    cudaError_t e = cudaSuccess;
    if (!d->prezeroed) {
      e = cudaMemsetAsync(a.sums, 0, sizeof(float) * 2 * a.pcols, stream);
      if (e == cudaSuccess && a.dbias) e = cudaMemsetAsync(a.dbias, 0, sizeof(float) * a.pcols, stream);
    }
    if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "bn_bwd memset: %s", cudaGetErrorString(e));
    dim3 grid(ceil_div(a.cols, BN_COLS), ceil_div(a.rows, a.rpb));
    bn_bwd_reduce_kernel<<<grid, BN_THREADS, 0, stream>>>(a);
    FXN_CHECK_LAUNCH("bn_bwd_reduce");
  }
  if (d->phase != 1) {
    dim3 grid2(ceil_div(width, BN_COLS), ceil_div(a.rows, a.rpb));
    bn_bwd_apply_kernel<<<grid2, BN_THREADS, 0, stream>>>(a);
    FXN_CHECK_LAUNCH("bn_bwd_apply");
  }
  return 0;
}

extern "C" int fxn_col_stats(const float* V, long long ldv, long long rows, int cols, int tile_rows, float* partials,
                             void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!V || !partials || rows <= 0 || cols <= 0 || tile_rows <= 0)
    return set_error(FXN_ERR_ARG, "fxn_col_stats: bad argument");
  dim3 grid(ceil_div(cols, 64), ceil_div(rows, tile_rows));
  col_stats_kernel<<<grid, 256, 0, stream>>>(V, ldv, rows, cols, tile_rows, partials);
  FXN_CHECK_LAUNCH("col_stats");
  return 0;
}
