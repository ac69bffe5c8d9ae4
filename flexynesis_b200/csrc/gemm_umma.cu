// fxn_gemm: C[M,N] = sum over bf16 split terms of A[M,K] * B[N,K]^T  (fp32 accumulate in TMEM)
//
// One kernel serves every dense contraction of the training step (SURVEY.md 2b K1,K4,K5,K9,K11,K16):
//   forward   Z  = X  * W^T          A = X  [B x d]  K-major, B = W  [h x d] K-major
//   dgrad     dD = dO * W            A = dO [B x o]  K-major, B = W  [o x h] MN-major (K = o rows)
//   wgrad     dW = dZ^T * X          A = dZ [B x h]  MN-major (K = B rows), B = X [B x d] MN-major
// so the same row-major bf16 planes are used by all three without any transposed copy.
//
// fp32 fidelity: every fp32 operand is stored as two bf16 planes (hi, lo), x = hi + lo (+2^-17 rel).
// nterms = 3 issues hi*hi + hi*lo + lo*hi per K step (fp32-grade, 4e-6 rms); nterms = 1 issues hi*hi only.
//
// Structure (Blackwell-native): warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle, mbarrier
// complete_tx), warp 1 = single-thread tcgen05.mma issuer with the accumulator in TMEM, warps 2..5 =
// epilogue (tcgen05.ld -> registers -> shared staging tile -> coalesced global stores, optional bias,
// bf16 hi/lo planes of the result and per-tile column statistics for the BatchNorm that follows).
#include "fxn_internal.h"
#include "ptx.cuh"
#include "gemm_common.cuh"
#include <cstdlib>

namespace fxn {

constexpr int BM = 128;       // UMMA M (cta_group::1)
constexpr int BK = 64;        // one 128B swizzle atom of bf16 along K
constexpr int UK = 16;        // UMMA K for 16-bit inputs
constexpr int GEMM_THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr int MAX_STAGES = 8;

struct GemmKernelArgs {
  int M, N, K;
  int bn;                 // tile N (multiple of 16, <= 256)
  int a_mn, b_mn;         // operand majorness (0 = K-major, 1 = MN-major)
  int nterms;             // 1 or 3
  int stages;
  int kb_per_split;       // K blocks per blockIdx.z
  int splitk;
  uint32_t tmem_cols;
  float* C; long long ldc;
  const float* bias;
  __nv_bfloat16* c_hi; __nv_bfloat16* c_lo; long long ldp;
  float* colstats;        // [m_tiles][2][N]: (sum, M2 about the tile mean); stats_mode 1 = sum only
  int stats_mode;
  int vec_c;              // C rows are 16B aligned -> float4 stores
  int epi_act;            // 0 none, 1 relu, 3 sigmoid, 6 leaky_relu(0.2); applied after alpha and bias
  int accumulate;         // C += result
  float alpha; const float* alpha_dev;
  const float* mse_x; long long ldx; float* mse_acc;   // fused sigmoid-MSE: see fxn_gemm_desc
  const float* gauss_ra; const float* gauss_rb; float gauss_inv;   // epi_act 7: exp(-max(ra[m]+rb[n]-2acc,0)*inv)
  float stats_alpha; const float* stats_alpha_dev;     // scale of the stats_mode 3 column sums
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const GemmKernelArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[MAX_STAGES];
  __shared__ uint64_t empty_bar[MAX_STAGES];
  __shared__ uint64_t accum_bar;
  __shared__ uint32_t tmem_base_slot;

  // 1024B alignment for SWIZZLE_128B operand tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * p.bn;
  const int m0 = blockIdx.y * BM;
  const int nplanes = (p.nterms == 3) ? 2 : 1;
  const uint32_t a_bytes = BM * BK * 2;
  const uint32_t b_bytes = static_cast<uint32_t>(p.bn) * BK * 2;
  const uint32_t stage_bytes = nplanes * (a_bytes + b_bytes);

  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > kb_total) kb_end = kb_total;
  const int nkb = kb_end - kb_begin;   // host guarantees >= 1

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (nplanes == 2) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + static_cast<size_t>(stage) * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
        const int k0 = kb * BK;
        for (int pl = 0; pl < nplanes; ++pl) {
          uint8_t* sa = st + pl * a_bytes;
          uint8_t* sb = st + nplanes * a_bytes + pl * b_bytes;
          const CUtensorMap* ta = pl ? &tmA_lo : &tmA_hi;
          const CUtensorMap* tb = pl ? &tmB_lo : &tmB_hi;
          if (p.a_mn == 0) {
            tma_load_2d(sa, ta, &full_bar[stage], k0, m0);                    // box {64 k, 128 rows}
          } else {
            for (int a = 0; a < BM / 64; ++a)                                 // box {64 m, 64 k rows}
              tma_load_2d(sa + a * (BK * 128), ta, &full_bar[stage], m0 + a * 64, k0);
          }
          if (p.b_mn == 0) {
            tma_load_2d(sb, tb, &full_bar[stage], k0, n0);                    // box {64 k, bn rows}
          } else {
            for (int a = 0; a < p.bn / 64; ++a)
              tma_load_2d(sb + a * (BK * 128), tb, &full_bar[stage], n0 + a * 64, k0);
          }
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, p.bn, p.a_mn, p.b_mn);
      // K-major SW128: 8-row groups of 1024B (SBO), LBO unused(1). Advance K by 16 elems = 32B inside the atom.
      // MN-major SW128: 64-element MN atoms, each BK rows x 128B; SBO = 1024 (8 k rows), LBO = BK*128 (next MN atom);
      //                 advance K by 16 rows = 2048B.
      const uint32_t a_lbo = p.a_mn ? BK * 128 : 16, b_lbo = p.b_mn ? BK * 128 : 16;
      const uint32_t a_kstep = p.a_mn ? UK * 128 : UK * 2, b_kstep = p.b_mn ? UK * 128 : UK * 2;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
        const uint32_t sa_hi = st, sa_lo = st + a_bytes;
        const uint32_t sb_hi = st + nplanes * a_bytes, sb_lo = sb_hi + b_bytes;
#pragma unroll
        for (int kk = 0; kk < BK / UK; ++kk) {
          const uint64_t da_hi = umma_smem_desc_sw128(sa_hi + kk * a_kstep, a_lbo, 1024);
          const uint64_t db_hi = umma_smem_desc_sw128(sb_hi + kk * b_kstep, b_lbo, 1024);
          if (nplanes == 2) {
            const uint64_t da_lo = umma_smem_desc_sw128(sa_lo + kk * a_kstep, a_lbo, 1024);
            const uint64_t db_lo = umma_smem_desc_sw128(sb_lo + kk * b_kstep, b_lbo, 1024);
            umma_bf16(tmem_acc, da_lo, db_hi, idesc, acc);   // small terms first
            acc = 1;
            umma_bf16(tmem_acc, da_hi, db_lo, idesc, acc);
          }
          umma_bf16(tmem_acc, da_hi, db_hi, idesc, acc);
          acc = 1;
        }
        umma_commit(&empty_bar[stage]);          // frees this smem stage when the MMAs above retire
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&accum_bar);                   // accumulator complete
    }
  } else {
    // ===================== Epilogue: 4 warps, TMEM lane quarter = warp % 4 =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;               // row inside the tile == TMEM lane
    const int lds = p.bn + 4;                    // staging row stride (floats): conflict-free both ways
    float* stg = reinterpret_cast<float*>(smem);
    const int et = threadIdx.x - 64;             // 0..127

    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const bool add_bias = (p.bias != nullptr) && (blockIdx.z == 0);
    const float alpha = p.alpha * (p.alpha_dev ? __ldg(p.alpha_dev) : 1.f);
    for (int c0 = 0; c0 < p.bn; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0);
      if (p.bn - c0 >= 32) {
        tmem_ld_32x32(taddr, v);
      } else {
        tmem_ld_32x16(taddr, v);
#pragma unroll
        for (int j = 16; j < 32; ++j) v[j] = 0;
      }
      tmem_ld_wait();
      const int ncols = (p.bn - c0 >= 32) ? 32 : 16;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j < ncols) {
          float4 o;
          o.x = __uint_as_float(v[j + 0]) * alpha;
          o.y = __uint_as_float(v[j + 1]) * alpha;
          o.z = __uint_as_float(v[j + 2]) * alpha;
          o.w = __uint_as_float(v[j + 3]) * alpha;
          if (add_bias) {
            const int n = n0 + c0 + j;
            if (n + 0 < p.N) o.x += __ldg(p.bias + n + 0);
            if (n + 1 < p.N) o.y += __ldg(p.bias + n + 1);
            if (n + 2 < p.N) o.z += __ldg(p.bias + n + 2);
            if (n + 3 < p.N) o.w += __ldg(p.bias + n + 3);
          }
          if (p.epi_act == 7) {
            const float ra = __ldg(p.gauss_ra + min(m0 + row, p.M - 1));
            const int n = n0 + c0 + j;
            const float r0 = __ldg(p.gauss_rb + min(n + 0, p.N - 1)), r1 = __ldg(p.gauss_rb + min(n + 1, p.N - 1));
            const float r2 = __ldg(p.gauss_rb + min(n + 2, p.N - 1)), r3 = __ldg(p.gauss_rb + min(n + 3, p.N - 1));
            o.x = __expf(-fmaxf(ra + r0 - 2.f * o.x, 0.f) * p.gauss_inv);
            o.y = __expf(-fmaxf(ra + r1 - 2.f * o.y, 0.f) * p.gauss_inv);
            o.z = __expf(-fmaxf(ra + r2 - 2.f * o.z, 0.f) * p.gauss_inv);
            o.w = __expf(-fmaxf(ra + r3 - 2.f * o.w, 0.f) * p.gauss_inv);
          } else if (p.epi_act) {
            o.x = epi_activation(o.x, p.epi_act);
            o.y = epi_activation(o.y, p.epi_act);
            o.z = epi_activation(o.z, p.epi_act);
            o.w = epi_activation(o.w, p.epi_act);
          }
          *reinterpret_cast<float4*>(stg + row * lds + c0 + j) = o;
        }
      }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");   // epilogue warps only

    const int ew = et >> 5;                      // 0..3
    const int mrows = min(BM, p.M - m0);
    const int ncolsv = min(p.bn, p.N - n0);
    // ---- fp32 C (coalesced rows) ----
    if (p.C != nullptr) {
      for (int r = ew; r < mrows; r += 4) {
        float* crow = p.C + static_cast<long long>(m0 + r) * p.ldc + n0;
        const float* srow = stg + r * lds;
        if (p.splitk > 1) {
          for (int c = lane; c < ncolsv; c += 32) atomicAdd(crow + c, srow[c]);
        } else if (p.accumulate) {
          for (int c = lane; c < ncolsv; c += 32) crow[c] += srow[c];
        } else if (p.vec_c) {
          for (int c = lane * 4; c < ncolsv; c += 128) {
            const float4 o = *reinterpret_cast<const float4*>(srow + c);
            if (c + 3 < ncolsv) {
              *reinterpret_cast<float4*>(crow + c) = o;
            } else {
              crow[c] = o.x;
              if (c + 1 < ncolsv) crow[c + 1] = o.y;
              if (c + 2 < ncolsv) crow[c + 2] = o.z;
            }
          }
        } else {
          for (int c = lane; c < ncolsv; c += 32) crow[c] = srow[c];
        }
      }
    }
    // ---- fused sigmoid-MSE (Decoder output + reconstruction loss): staged value is x_hat; accumulate
    //      sum (x_hat - x)^2 and replace the staged tile by G = (x_hat - x) * x_hat * (1 - x_hat) ----
    if (p.mse_x != nullptr) {
      float sq = 0.f;
      for (int r = ew; r < mrows; r += 4) {
        float* srow = stg + r * lds;
        const float* xrow = p.mse_x + static_cast<long long>(m0 + r) * p.ldx + n0;
        for (int c = lane; c < ncolsv; c += 32) {
          const float xh = srow[c];
          const float df = xh - __ldg(xrow + c);
          sq = fmaf(df, df, sq);
          srow[c] = df * xh * (1.f - xh);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (lane == 0) atomicAdd(p.mse_acc, sq);
      __syncwarp();
    }
    // ---- bf16 hi/lo planes of the result (operand of the next GEMM); ldp % 8 == 0, zero-fills to ldp pad ----
    if (p.c_hi != nullptr) {
      for (int r = ew; r < mrows; r += 4) {
        const float* srow = stg + r * lds;
        const long long off = static_cast<long long>(m0 + r) * p.ldp + n0;
        for (int c = lane * 8; c < p.bn; c += 256) {
          if (n0 + c >= ((p.N + 7) & ~7)) break;      // zero-fill only this GEMM's own pad8(N) columns
          __align__(16) __nv_bfloat16 h[8];
          __align__(16) __nv_bfloat16 l[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = (c + j < ncolsv) ? srow[c + j] : 0.f;
            split_bf16(x, h[j], l[j]);
          }
          *reinterpret_cast<uint4*>(p.c_hi + off + c) = *reinterpret_cast<const uint4*>(h);
          *reinterpret_cast<uint4*>(p.c_lo + off + c) = *reinterpret_cast<const uint4*>(l);
        }
      }
    }
    // ---- per-tile column statistics over the valid rows ----
    if (p.stats_mode != 0 && p.mse_x != nullptr) asm volatile("bar.sync 1, 128;" ::: "memory");   // tile was rewritten
    if (p.stats_mode != 0) {
      const float sscale = p.stats_alpha * (p.stats_alpha_dev ? __ldg(p.stats_alpha_dev) : 1.f);
      for (int c = et; c < ncolsv; c += EPI_THREADS) {
        float s = 0.f;
        for (int r = 0; r < mrows; ++r) s += stg[r * lds + c];
        float m2 = 0.f;
        if (p.stats_mode == 2) {
          const float mu = s / static_cast<float>(mrows);
          for (int r = 0; r < mrows; ++r) {
            const float d = stg[r * lds + c] - mu;
            m2 = fmaf(d, d, m2);
          }
        }
        if (p.stats_mode == 3) {
          atomicAdd(p.colstats + n0 + c, s * sscale);   // plain column sums into a zeroed [N] vector (bias gradients)
        } else {
          float* dst = p.colstats + (static_cast<long long>(blockIdx.y) * 2) * p.N + n0 + c;
          dst[0] = s;
          dst[p.N] = m2;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, p.tmem_cols);
  }
}

// -------------------------------------------------------------------------------------------------
// Host side
// -------------------------------------------------------------------------------------------------
// Tile width / split-K selection by a small analytic cost model (times in microseconds, B200 constants).
// A CTA's k-block costs max(MMA issue time, operand load time); the chip moves at most ~L2_BW bytes/us from L2 to
// the SMs, a single SM at most SM_BW; a launch costs its waves times the CTA time.
struct TileChoice { int bn; int splitk; };

static TileChoice choose_tile(int M, int N, int K, int nterms, int b_mn, bool allow_split) {
  const double SM_CLK = 1.85e3;           // cycles per us under load
  const double L2_BW = 9.0e6;             // bytes per us, whole chip (~9 TB/s L2 -> SM)
  const double SM_BW = 1.6e5;             // bytes per us, one SM
  const int nplanes = nterms == 3 ? 2 : 1;
  const int kb_total = (K + BK - 1) / BK;
  const int mt = (M + BM - 1) / BM;
  const int step = b_mn ? 64 : 16;
  TileChoice best{256, 1};
  double best_t = 1e30;
  for (int bn = step; bn <= 256; bn += step) {
    const int nt = (N + bn - 1) / bn;
    if (nt > 1 && bn < 64) continue;                       // narrow tiles only when one tile covers N
    if (nt > 1 && (nt - 1) * bn >= N) continue;
    const int stage_bytes = nplanes * (BM * BK * 2 + bn * BK * 2);
    int stages = (231424 - 1024) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 1) continue;
    for (int sk = 1; sk <= 32; sk *= 2) {
      if (sk > 1 && (!allow_split || kb_total / sk < 2)) break;
      const int ctas = mt * nt * sk;
      const int kb = (kb_total + sk - 1) / sk;
      const int active = ctas < 148 ? ctas : 148;
      const double t_mma = nterms * 4.0 * (bn / 2.0) / SM_CLK;                 // 4 K-steps x nterms MMAs of 128 x bn x 16
      double bw = L2_BW / active;
      if (bw > SM_BW) bw = SM_BW;
      const double t_load = stage_bytes / bw;
      // with few stages the first loads are exposed; deep pipelines hide everything but the slower of the two rates
      const double lat = 1.6;                                                 // TMA round trip
      double t_loop = kb * (t_mma > t_load ? t_mma : t_load);
      const double exposed = lat * (stages >= kb ? 1.0 : (stages >= 3 ? 1.0 : 1.0 + 0.5 * kb / stages));
      const double t_epi = 1.0 + 3.0 * bn / 256.0 + (sk > 1 ? 1.5 * bn / 256.0 : 0.0);
      const double t_cta = 2.0 + exposed + t_loop + t_epi;
      const int waves = (ctas + 147) / 148;
      double t = waves * t_cta + (sk > 1 ? 2.0 : 0.0);                        // split-K also pays a memset
      if (t < best_t) { best_t = t; best = TileChoice{bn, sk}; }
    }
  }
  return best;
}

}  // namespace fxn

using namespace fxn;

extern "C" int fxn_gemm(const fxn_gemm_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d) return set_error(FXN_ERR_ARG, "null descriptor");
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return set_error(FXN_ERR_ARG, "fxn_gemm: M,N,K must be positive");
  if (d->nterms != 1 && d->nterms != 3) return set_error(FXN_ERR_ARG, "fxn_gemm: nterms must be 1 or 3");
  if (!d->a_hi || !d->b_hi || (d->nterms == 3 && (!d->a_lo || !d->b_lo)))
    return set_error(FXN_ERR_ARG, "fxn_gemm: missing operand plane");
  if (!d->C && !d->c_hi) return set_error(FXN_ERR_ARG, "fxn_gemm: no output");
  {
    // FXN_GEMM_KERNEL=1 selects the one-tile-per-CTA kernel below (kept as the A/B reference); default is the
    // persistent CTA-pair kernel of gemm_umma2.cu
    static const int which = [] { const char* e = getenv("FXN_GEMM_KERNEL"); return e ? atoi(e) : 2; }();
    if (which != 1) return gemm2_dispatch(d, stream);
  }

  GemmKernelArgs p;
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.a_mn = d->a_mn_major ? 1 : 0;
  p.b_mn = d->b_mn_major ? 1 : 0;
  p.nterms = d->nterms;
  const bool split_ok = d->splitk < 0 && d->C && !d->c_hi && !d->colstats && !d->epi_act && !d->accumulate && !d->mse_x;
  const TileChoice tc_auto = choose_tile(d->M, d->N, d->K, d->nterms, p.b_mn, split_ok);
  p.bn = d->block_n > 0 ? d->block_n : tc_auto.bn;
  if (p.bn % 16 != 0 || p.bn > 256 || (p.b_mn && p.bn % 64 != 0))
    return set_error(FXN_ERR_ARG, "fxn_gemm: invalid block_n %d", p.bn);
  const int nplanes = d->nterms == 3 ? 2 : 1;
  const int stage_bytes = nplanes * (BM * BK * 2 + p.bn * BK * 2);
  const int epi_bytes = BM * (p.bn + 4) * 4;
  const int max_smem = 231424 - 1024;   // opt-in dynamic limit (232448 - 1024 static) minus the 1024B alignment slack
  int stages = max_smem / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  const int kb_total = (d->K + BK - 1) / BK;
  int splitk = d->splitk > 1 ? d->splitk : 1;
  if (d->splitk < 0) splitk = (d->block_n > 0) ? 1 : tc_auto.splitk;   // auto
  if (splitk > kb_total) splitk = kb_total;
  int kb_per = (kb_total + splitk - 1) / splitk;
  splitk = (kb_total + kb_per - 1) / kb_per;      // no empty splits
  if (stages > kb_per) stages = kb_per;
  if (stages < 1) return set_error(FXN_ERR_ARG, "fxn_gemm: tile does not fit in shared memory");
  int smem_bytes = stages * stage_bytes;
  if (smem_bytes < epi_bytes) smem_bytes = epi_bytes;
  smem_bytes += 1024;
  p.stages = stages;
  p.kb_per_split = kb_per;
  p.splitk = splitk;
  uint32_t tc = 32;
  while (tc < static_cast<uint32_t>(p.bn)) tc <<= 1;
  p.tmem_cols = tc;
  p.C = d->C; p.ldc = d->ldc;
  p.bias = d->bias;
  p.c_hi = static_cast<__nv_bfloat16*>(d->c_hi);
  p.c_lo = static_cast<__nv_bfloat16*>(d->c_lo);
  p.ldp = d->ldp;
  p.colstats = d->colstats;
  p.stats_mode = d->colstats ? (d->stats_mode ? d->stats_mode : 2) : 0;
  if (splitk > 1 && (p.c_hi || p.stats_mode || !d->C))
    return set_error(FXN_ERR_ARG, "fxn_gemm: split-K supports the fp32 output only");
  if (p.c_hi && (!p.c_lo || d->ldp % 8 != 0 || (reinterpret_cast<uintptr_t>(d->c_hi) & 15) ||
                 (reinterpret_cast<uintptr_t>(d->c_lo) & 15)))
    return set_error(FXN_ERR_ARG, "fxn_gemm: output planes need both pointers, 16B alignment and ldp %% 8 == 0");
  p.vec_c = (d->C && (reinterpret_cast<uintptr_t>(d->C) & 15) == 0 && d->ldc % 4 == 0) ? 1 : 0;
  p.epi_act = d->epi_act;
  p.accumulate = d->accumulate;
  p.alpha = d->alpha == 0.f ? 1.f : d->alpha;
  p.alpha_dev = d->alpha_dev;
  p.mse_x = d->mse_x; p.ldx = d->ldx; p.mse_acc = d->mse_acc;
  if (p.mse_x && (!p.mse_acc || splitk > 1 || (p.stats_mode && p.stats_mode != 3)))
    return set_error(FXN_ERR_ARG, "fxn_gemm: fused MSE needs mse_acc and excludes split-K / per-tile statistics");
  p.gauss_ra = d->gauss_ra; p.gauss_rb = d->gauss_rb; p.gauss_inv = d->gauss_inv;
  if (p.epi_act == 7 && (!p.gauss_ra || !p.gauss_rb)) return set_error(FXN_ERR_ARG, "fxn_gemm: epi_act 7 needs row norms");
  p.stats_alpha = d->stats_alpha == 0.f ? 1.f : d->stats_alpha;
  p.stats_alpha_dev = d->stats_alpha_dev;
  if (splitk > 1 && (p.epi_act || p.accumulate))
    return set_error(FXN_ERR_ARG, "fxn_gemm: split-K excludes epilogue activation / accumulate");

  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  // stored arrays: K-major operand = [MN rows x K cols]; MN-major operand = [K rows x MN cols]
  const long long a_rows = p.a_mn ? d->K : d->M, a_cols = p.a_mn ? d->M : d->K;
  const long long b_rows = p.b_mn ? d->K : d->N, b_cols = p.b_mn ? d->N : d->K;
  const int a_box = p.a_mn ? BK : BM, b_box = p.b_mn ? BK : p.bn;
  if ((rc = make_tensor_map(&ta_hi, d->a_hi, a_rows, a_cols, d->lda, a_box))) return rc;
  if ((rc = make_tensor_map(&tb_hi, d->b_hi, b_rows, b_cols, d->ldb, b_box))) return rc;
  if (nplanes == 2) {
    if ((rc = make_tensor_map(&ta_lo, d->a_lo, a_rows, a_cols, d->lda, a_box))) return rc;
    if ((rc = make_tensor_map(&tb_lo, d->b_lo, b_rows, b_cols, d->ldb, b_box))) return rc;
  } else {
    ta_lo = ta_hi;
    tb_lo = tb_hi;
  }

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 231424);
    if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  if (p.stats_mode == 3 && !d->outputs_prezeroed) {   // plain column sums are accumulated with atomics: zero the target first
    cudaError_t e = cudaMemsetAsync(d->colstats, 0, sizeof(float) * d->N, stream);
    if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "colsum memset: %s", cudaGetErrorString(e));
  }
  if (splitk > 1 && !d->outputs_prezeroed) {
    cudaError_t e = cudaMemset2DAsync(d->C, d->ldc * sizeof(float), 0, d->N * sizeof(float), d->M, stream);
    if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "split-K memset: %s", cudaGetErrorString(e));
  }
  dim3 grid((d->N + p.bn - 1) / p.bn, (d->M + BM - 1) / BM, splitk);
  gemm_umma_kernel<<<grid, GEMM_THREADS, smem_bytes, stream>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

extern "C" int fxn_gemm_stat_tiles(int M) { return (M + BM - 1) / BM; }
