// The supervisor-head section of a training step as three CUDA-core kernels (narrow heads only: every target variable's
// MLP is L -> sh <= 64 -> C <= 16, where a tensor-core tile would be mostly padding and the time goes into launches).
//
// Replaces, for ALL target variables at once (flexynesis/modules.py:135-150 MLP.forward per variable,
// flexynesis/models/direct_pred.py:131-132 the loop over self.MLPs, :146-190 compute_loss and their autograd duals):
//   heads_l1_kernel   Zh = F W1cat^T + b1 for a 16-row block, per-block BatchNorm partials (sum, M2 about the block
//                     mean), valid-label counts;
//   heads_mid_kernel  Chan merge of the partials -> BatchNorm -> ReLU -> Dropout -> layer_out -> MSE / cross-entropy
//                     and, in the same pass, d loss / d logits (the loss weights exp(-s_k) are parameters, the
//                     normalising counts come from heads_l1) -> gradient of layer_out -> gradient w.r.t. the BatchNorm
//                     output G, its column sums for the BatchNorm backward, d layer_out.weight / bias;
//   heads_bwd_kernel  dZh = gamma rstd (G - mean G - xhat mean(G xhat)), its operand planes (for the layer_1 weight
//                     gradient GEMM), dF = dZh W1cat as operand planes (+ fp32), column sums of dF (bias gradient of the
//                     layer that produced F), d gamma / d beta.
// The generic path (fxn_gemm + fxn_bn_act_* + fxn_head_out_*) ran this section as 12 latency-bound launches, ~60 us of a
// 470 us config-2 step (profiles/r01_timeline_v7_cfg2.log); it stays for wide heads (config 5: sh = 256) and Cox heads.
#include "fxn_internal.h"
#include "ptx.cuh"
#include "rng.cuh"
#include <math_constants.h>

namespace fxn {

constexpr int HF_ROWS = 16;          // rows per CTA (256 CTAs at B = 4096: two per SM, so four warps per scheduler hide latency)
constexpr int HF_RPW = 2;            // rows per warp
constexpr int HF_THREADS = 256;      // 8 warps x 2 rows
constexpr int HF_MAXC = 16;

__device__ __forceinline__ float hf_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float hf_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// layer_1 of every head on a 32-row block + BatchNorm partials + label counts
// ---------------------------------------------------------------------------------------------------------------
template <int NCOL>
__global__ void __launch_bounds__(HF_THREADS) heads_l1_kernel(const fxn_heads_desc d) {
  extern __shared__ __align__(16) float hf_smem[];
  const int L = d.L, shp = (d.sh + 7) & ~7, width = d.nv * shp, wst = width + 1;
  float* sF = hf_smem;                               // [32][L]
  float* sW = sF + HF_ROWS * L;                      // [L][wst]  (W1cat transposed; padded columns are zero)
  float* sRed = sW + static_cast<size_t>(L) * wst;   // [8][2][64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * HF_ROWS;
  const int nvalid = min(HF_ROWS, d.B - r0);
  // stage F rows (zero beyond B) and the transposed layer_1 weights. Loads are issued in batches of 8 independent 16-byte
  // requests per thread: a one-load-per-iteration loop pays a full L2 round trip per element and made this 2 us kernel
  // take 17 us (profiles/r02_timeline_heads_v1.log).
  {
    const int L4 = L >> 2, n4 = HF_ROWS * L4;
    for (int base = tid; base < n4; base += HF_THREADS * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * HF_THREADS;
        const int r = i / L4, k4 = i - r * L4;
        v[u] = (i < n4 && r < nvalid) ? __ldg(reinterpret_cast<const float4*>(d.F + static_cast<long long>(r0 + r) * d.ldf) + k4)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * HF_THREADS;
        if (i < n4) reinterpret_cast<float4*>(sF)[i] = v[u];
      }
    }
    const int m4 = width * L4;          // (column of Zh, 4 consecutive k)
    for (int base = tid; base < m4; base += HF_THREADS * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * HF_THREADS;
        const int c = i / L4, k4 = i - c * L4;
        const int vv = c / shp, j = c - vv * shp;
        v[u] = (i < m4 && j < d.sh) ? __ldg(reinterpret_cast<const float4*>(d.var[vv].W1 + static_cast<long long>(j) * L) + k4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * HF_THREADS;
        if (i < m4) {
          const int c = i / L4, k = (i - c * L4) * 4;
          sW[(k + 0) * wst + c] = v[u].x; sW[(k + 1) * wst + c] = v[u].y;
          sW[(k + 2) * wst + c] = v[u].z; sW[(k + 3) * wst + c] = v[u].w;
        }
      }
    }
  }
  __syncthreads();
  constexpr int ncol = NCOL;                         // column slots per lane (width <= 32: 1, else 2)
  float acc[HF_RPW][2];
#pragma unroll
  for (int r = 0; r < HF_RPW; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
  const float* f0 = sF + (warp * HF_RPW) * L;
  const int c1 = lane + 32 < width ? lane + 32 : lane;     // second column slot (reads stay in range when unused)
  const int c0 = lane < width ? lane : 0;
#pragma unroll 2
  for (int k = 0; k < L; k += 4) {
    float4 fr[HF_RPW];
#pragma unroll
    for (int r = 0; r < HF_RPW; ++r) fr[r] = *reinterpret_cast<const float4*>(f0 + r * L + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float w0 = sW[(k + kk) * wst + c0];
      const float w1 = NCOL > 1 ? sW[(k + kk) * wst + c1] : 0.f;
#pragma unroll
      for (int r = 0; r < HF_RPW; ++r) {
        const float f = kk == 0 ? fr[r].x : kk == 1 ? fr[r].y : kk == 2 ? fr[r].z : fr[r].w;
        acc[r][0] = fmaf(f, w0, acc[r][0]);
        if (NCOL > 1) acc[r][1] = fmaf(f, w1, acc[r][1]);
      }
    }
  }
  // bias, store, per-block column statistics over the valid rows
  float s[2] = {0.f, 0.f};
#pragma unroll
  for (int cs = 0; cs < 2; ++cs) {
    const int c = lane + 32 * cs;
    if (cs < ncol && c < width) {
      const int v = c / shp, j = c - v * shp;
      const float b = j < d.sh ? __ldg(d.var[v].b1 + j) : 0.f;
#pragma unroll
      for (int r = 0; r < HF_RPW; ++r) {
        const int row = warp * HF_RPW + r;
        acc[r][cs] += b;
        if (row < nvalid) {
          d.Zh[static_cast<long long>(r0 + row) * d.ldz + c] = acc[r][cs];
          s[cs] += acc[r][cs];
        }
      }
      sRed[(warp * 2 + 0) * 64 + c] = s[cs];
    }
  }
  __syncthreads();
#pragma unroll
  for (int cs = 0; cs < 2; ++cs) {
    const int c = lane + 32 * cs;
    if (cs < ncol && c < width) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += sRed[(w * 2 + 0) * 64 + c];
      const float mu = tot / static_cast<float>(nvalid);
      float m2 = 0.f;
#pragma unroll
      for (int r = 0; r < HF_RPW; ++r)
        if (warp * HF_RPW + r < nvalid) { const float q = acc[r][cs] - mu; m2 = fmaf(q, q, m2); }
      sRed[(warp * 2 + 1) * 64 + c] = m2;
      s[cs] = tot;
    }
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int cs = 0; cs < 2; ++cs) {
      const int c = lane + 32 * cs;
      if (cs < ncol && c < width) {
        float m2 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) m2 += sRed[(w * 2 + 1) * 64 + c];
        // layout [2][width][nblk rounded up to 32]: heads_mid reads a column's partials as consecutive 16-byte vectors
        const int nbp = (static_cast<int>(gridDim.x) + 31) & ~31;
        d.partials[static_cast<long long>(c) * nbp + blockIdx.x] = s[cs];
        d.partials[static_cast<long long>(width + c) * nbp + blockIdx.x] = m2;
      }
    }
  }
  // number of valid labels per variable (the denominators of the mean losses)
  if (warp == 1) {
    for (int v = 0; v < d.nv; ++v) {
      const float* y = d.var[v].y;
      if (y == nullptr) continue;
      float c = 0.f;
      if (lane < nvalid) {
        const float yv = y[r0 + lane];
        c = (!isnan(yv) && (d.var[v].kind == 1 || yv != -1.f)) ? 1.f : 0.f;
      }
      c = hf_warp_sum(c);
      if (lane == 0 && c > 0.f) atomicAdd(d.acc + 2 * d.var[v].slot + 1, c);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// BatchNorm -> ReLU -> Dropout -> layer_out -> loss (+ backward to G)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HF_THREADS) heads_mid_kernel(const fxn_heads_desc d) {
  __shared__ float s_scale[64], s_shift[64], s_mean[64], s_rstd[64];
  __shared__ float s_red[HF_THREADS];
  __shared__ float s_wout[HF_MAXC * 64];   // [c][column of Zh]: layer_out.weight of the variable that owns the column
  __shared__ float s_dw[HF_MAXC * 64];     // block partial of d layer_out.weight, same indexing
  __shared__ float s_db[FXN_HEADS_MAX_VARS * HF_MAXC];
  __shared__ float s_sum[2][64];
  __shared__ float s_loss[FXN_HEADS_MAX_VARS];
  __shared__ float s_lossacc[HF_THREADS / 32][FXN_HEADS_MAX_VARS];
  const int shp = (d.sh + 7) & ~7, width = d.nv * shp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * HF_ROWS;
  const int nblk = (d.B + HF_ROWS - 1) / HF_ROWS;
  // ---- batch statistics: Chan merge of the per-block partials. Thread (column, part) owns 32-block spans of the column's
  // partials and fetches each span with 8 independent 16-byte loads (serial scalar loads cost one L2 round trip each and
  // made this kernel 35 us) ----
  {
    const int c = tid & 63, part = tid >> 6;       // 4 parts per column slot
    const bool ok = c < width;
    const int nbp = (nblk + 31) & ~31;
    if (d.train) {
      float t = 0.f;
      if (ok)
        for (int b0 = part * 32; b0 < nblk; b0 += 128) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(d.partials + static_cast<long long>(c) * nbp + b0 + 4 * u);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int b = b0 + 4 * u;
            t += (b < nblk ? v[u].x : 0.f) + (b + 1 < nblk ? v[u].y : 0.f) + (b + 2 < nblk ? v[u].z : 0.f) + (b + 3 < nblk ? v[u].w : 0.f);
          }
        }
      s_red[tid] = t;
      __syncthreads();
      const float mean = (s_red[c] + s_red[64 + c] + s_red[128 + c] + s_red[192 + c]) / static_cast<float>(d.B);
      __syncthreads();
      float m2 = 0.f;
      if (ok)
        for (int b0 = part * 32; b0 < nblk; b0 += 128) {
          float4 v[8], w[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            v[u] = *reinterpret_cast<const float4*>(d.partials + static_cast<long long>(c) * nbp + b0 + 4 * u);
            w[u] = *reinterpret_cast<const float4*>(d.partials + static_cast<long long>(width + c) * nbp + b0 + 4 * u);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float sv[4] = {v[u].x, v[u].y, v[u].z, v[u].w}, mv[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int b = b0 + 4 * u + e;
              if (b < nblk) {
                const float n = static_cast<float>(min(HF_ROWS, d.B - b * HF_ROWS));
                const float q = sv[e] / n - mean;
                m2 += mv[e] + n * q * q;
              }
            }
          }
        }
      s_red[tid] = m2;
      __syncthreads();
      if (part == 0 && ok) {
        const float var = (s_red[c] + s_red[64 + c] + s_red[128 + c] + s_red[192 + c]) / static_cast<float>(d.B);
        const int v = c / shp, j = c - v * shp;
        const float rstd = rsqrtf(var + d.eps);
        s_mean[c] = mean; s_rstd[c] = rstd;
        if (j < d.sh) {
          const float g = __ldg(d.var[v].gamma + j), be = __ldg(d.var[v].beta + j);
          s_scale[c] = g * rstd; s_shift[c] = be - mean * g * rstd;
          if (blockIdx.x == 0) {
            d.saved[c] = mean; d.saved[width + c] = rstd;
            float* rm = d.var[v].running_mean; float* rv = d.var[v].running_var;
            if (rm) {
              const float n = static_cast<float>(d.B);
              rm[j] = (1.f - d.momentum) * rm[j] + d.momentum * mean;
              rv[j] = (1.f - d.momentum) * rv[j] + d.momentum * var * n / (n - 1.f);
            }
            if (j == 0 && d.var[v].num_batches_tracked) *reinterpret_cast<long long*>(d.var[v].num_batches_tracked) += 1;
          }
        } else { s_scale[c] = 0.f; s_shift[c] = 0.f; }
      }
    } else if (part == 0 && ok) {
      const int v = c / shp, j = c - v * shp;
      if (j < d.sh) {
        const float mean = d.var[v].running_mean[j], rstd = rsqrtf(d.var[v].running_var[j] + d.eps);
        const float g = __ldg(d.var[v].gamma + j), be = __ldg(d.var[v].beta + j);
        s_mean[c] = mean; s_rstd[c] = rstd; s_scale[c] = g * rstd; s_shift[c] = be - mean * g * rstd;
      } else { s_mean[c] = 0.f; s_rstd[c] = 0.f; s_scale[c] = 0.f; s_shift[c] = 0.f; }
    }
  }
  for (int i = tid; i < HF_MAXC * 64; i += HF_THREADS) {
    const int c = i >> 6, col = i & 63;
    float w = 0.f;
    if (col < width) {
      const int v = col / shp, j = col - v * shp;
      if (c < d.var[v].C && j < d.sh) w = __ldg(d.var[v].Wout + c * d.sh + j);
    }
    s_wout[i] = w;
    s_dw[i] = 0.f;
  }
  if (tid < FXN_HEADS_MAX_VARS * HF_MAXC) s_db[tid] = 0.f;
  if (tid < 128) s_sum[tid >> 6][tid & 63] = 0.f;
  if (tid < FXN_HEADS_MAX_VARS) s_loss[tid] = 0.f;
  if (tid < (HF_THREADS / 32) * FXN_HEADS_MAX_VARS) s_lossacc[tid / FXN_HEADS_MAX_VARS][tid % FXN_HEADS_MAX_VARS] = 0.f;
  __syncthreads();

  const bool drop = d.train && d.p_drop > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - d.p_drop) : 1.f;
  const int ncol = (width + 31) >> 5;
  float g1[2] = {0.f, 0.f}, g2[2] = {0.f, 0.f};       // column sums of G and G * xhat over this warp's rows
  // this warp's rows of Zh, fetched together (independent loads)
  float zrow[HF_RPW][2];
#pragma unroll
  for (int r = 0; r < HF_RPW; ++r)
#pragma unroll
    for (int cs = 0; cs < 2; ++cs) {
      const int row = r0 + warp * HF_RPW + r, c = lane + 32 * cs;
      zrow[r][cs] = (row < d.B && cs < ncol && c < width) ? d.Zh[static_cast<long long>(row) * d.ldz + c] : 0.f;
    }
#pragma unroll
  for (int r = 0; r < HF_RPW; ++r) {
    const int row = r0 + warp * HF_RPW + r;
    if (row >= d.B) break;                             // warp-uniform
    float act[2], xh[2], yv_pre[2];
    bool keep[2];
#pragma unroll
    for (int cs = 0; cs < 2; ++cs) {
      const int c = lane + 32 * cs;
      act[cs] = 0.f; xh[cs] = 0.f; yv_pre[cs] = 0.f; keep[cs] = false;
      if (cs < ncol && c < width) {
        const int v = c / shp, j = c - v * shp;
        if (j < d.sh) {
          const float z = zrow[r][cs];
          xh[cs] = (z - s_mean[c]) * s_rstd[c];
          yv_pre[cs] = fmaf(z, s_scale[c], s_shift[c]);
          bool k = true;
          if (drop) {
            if (d.var[v].mask) k = d.var[v].mask[static_cast<long long>(row) * d.var[v].ldm + j] != 0;
            else
              k = (dropout_keep8(step_seed(d.var[v].seed, static_cast<const long long*>(d.seed_dev)),
                                 static_cast<unsigned long long>(row) * ((d.sh + 7) / 8) + (j >> 3), d.p_drop) >> (j & 7)) & 1u;
          }
          keep[cs] = k;
          act[cs] = k ? fmaxf(yv_pre[cs], 0.f) * keep_scale : 0.f;
        }
      }
    }
    // ---- per variable: logits, loss, d logits, gradient of layer_out. Lane c (< C) holds class c. ----
    float gcol[2] = {0.f, 0.f};
    for (int v = 0; v < d.nv; ++v) {
      const int C = d.var[v].C, kind = d.var[v].kind;
      const int cbeg = v * shp;
      const bool mine0 = lane >= cbeg && lane < cbeg + shp, mine1 = lane + 32 >= cbeg && lane + 32 < cbeg + shp;
      float mylogit = -CUDART_INF_F;
      for (int c = 0; c < C; ++c) {
        float p = 0.f;
        if (mine0) p = act[0] * s_wout[c * 64 + lane];
        if (mine1) p = fmaf(act[1], s_wout[c * 64 + lane + 32], p);
        p = hf_warp_sum(p);
        if (lane == c) mylogit = p + (d.var[v].bout ? __ldg(d.var[v].bout + c) : 0.f);
      }
      if (lane < C) d.var[v].logits[static_cast<long long>(row) * C + lane] = mylogit;
      const float* y = d.var[v].y;
      if (y == nullptr) continue;
      const float yv = y[row];
      const float cnt = d.acc[2 * d.var[v].slot + 1];
      const float w = d.var[v].log_var ? __expf(-__ldg(d.var[v].log_var)) : 1.f;
      float mydl = 0.f;
      if (kind == 1) {
        if (!isnan(yv)) {
          const float e = __shfl_sync(0xffffffffu, mylogit, 0) - yv;
          if (lane == 0) { s_lossacc[warp][v] += e * e; if (cnt > 0.f) mydl = 2.f * e / cnt * w; }
        }
      } else if (!isnan(yv) && yv != -1.f) {
        const float mx = hf_warp_max(mylogit);
        const float ex = lane < C ? expf(mylogit - mx) : 0.f;
        const float e = hf_warp_sum(ex);
        const int yi = static_cast<int>(static_cast<long long>(yv));   // y.long(): truncation
        const float picked = (yi >= 0 && yi < C) ? __shfl_sync(0xffffffffu, mylogit, yi & 31) : CUDART_NAN_F;
        if (lane == 0) s_lossacc[warp][v] += (mx + logf(e)) - picked;
        if (cnt > 0.f && lane < C) mydl = (ex / e - (lane == yi ? 1.f : 0.f)) / cnt * w;
      }
      if (!d.backward) continue;
      for (int c = 0; c < C; ++c) {
        const float dlc = __shfl_sync(0xffffffffu, mydl, c);
        if (dlc == 0.f) continue;                       // warp-uniform
        if (mine0) { gcol[0] = fmaf(dlc, s_wout[c * 64 + lane], gcol[0]); if (act[0] != 0.f) atomicAdd(&s_dw[c * 64 + lane], dlc * act[0]); }
        if (mine1) { gcol[1] = fmaf(dlc, s_wout[c * 64 + lane + 32], gcol[1]); if (act[1] != 0.f) atomicAdd(&s_dw[c * 64 + lane + 32], dlc * act[1]); }
      }
      if (lane < C && d.var[v].dbout && mydl != 0.f) atomicAdd(&s_db[v * HF_MAXC + lane], mydl);
    }
    if (d.backward) {
#pragma unroll
      for (int cs = 0; cs < 2; ++cs) {
        const int c = lane + 32 * cs;
        if (cs < ncol && c < width) {
          // through Dropout and ReLU: G = d loss / d (BatchNorm output)
          const float g = (keep[cs] && yv_pre[cs] > 0.f) ? gcol[cs] * keep_scale : 0.f;
          d.G[static_cast<long long>(row) * d.ldg + c] = g;
          g1[cs] += g;
          g2[cs] = fmaf(g, xh[cs], g2[cs]);
        }
      }
    }
  }
  // ---- block reductions -> global accumulators ----
  if (lane == 0)
    for (int v = 0; v < d.nv; ++v) if (s_lossacc[warp][v] != 0.f) atomicAdd(&s_loss[v], s_lossacc[warp][v]);
  if (d.backward) {
#pragma unroll
    for (int cs = 0; cs < 2; ++cs) {
      const int c = lane + 32 * cs;
      if (cs < ncol && c < width) { atomicAdd(&s_sum[0][c], g1[cs]); atomicAdd(&s_sum[1][c], g2[cs]); }
    }
  }
  __syncthreads();
  if (tid < d.nv && d.var[tid].y != nullptr && s_loss[tid] != 0.f) atomicAdd(d.acc + 2 * d.var[tid].slot, s_loss[tid]);
  if (d.backward) {
    if (tid < 128) {
      const int c = tid & 63;
      if (c < width) atomicAdd(d.sums + (tid >> 6) * width + c, s_sum[tid >> 6][c]);
    }
    for (int i = tid; i < HF_MAXC * 64; i += HF_THREADS) {
      const int c = i >> 6, col = i & 63;
      if (col < width && s_dw[i] != 0.f) {
        const int v = col / shp, j = col - v * shp;
        if (c < d.var[v].C && j < d.sh) atomicAdd(d.var[v].dWout + c * d.sh + j, s_dw[i]);
      }
    }
    if (tid < d.nv * HF_MAXC) {
      const int v = tid / HF_MAXC, c = tid % HF_MAXC;
      if (c < d.var[v].C && d.var[v].dbout && s_db[tid] != 0.f) atomicAdd(d.var[v].dbout + c, s_db[tid]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// BatchNorm backward apply + dF = dZh W1cat (+ its column sums) + planes
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HF_THREADS) heads_bwd_kernel(const fxn_heads_desc d) {
  extern __shared__ __align__(16) float hf_smem[];
  const int L = d.L, shp = (d.sh + 7) & ~7, width = d.nv * shp;
  float* sW = hf_smem;                               // [width][L]  (rows of padded columns are zero)
  float* sDz = sW + static_cast<size_t>(width) * L;  // [32][width + 1]
  float* sCol = sDz + HF_ROWS * (width + 1);         // [8][L] column sums of dF per warp (only with dbias)
  __shared__ float s_coef[64], s_m1[64], s_m2[64], s_mean[64], s_rstd[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * HF_ROWS;
  const int nvalid = min(HF_ROWS, d.B - r0);
  {
    const int L4 = L >> 2, m4 = width * L4;
    for (int base = tid; base < m4; base += HF_THREADS * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * HF_THREADS;
        const int c = i / L4, k4 = i - c * L4;
        const int vv = c / shp, j = c - vv * shp;
        v[u] = (i < m4 && j < d.sh) ? __ldg(reinterpret_cast<const float4*>(d.var[vv].W1 + static_cast<long long>(j) * L) + k4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * HF_THREADS;
        if (i < m4) reinterpret_cast<float4*>(sW)[i] = v[u];
      }
    }
  }
  if (tid < 64) {
    const int c = tid;
    float coef = 0.f, m1 = 0.f, m2 = 0.f, mean = 0.f, rstd = 0.f;
    if (c < width) {
      const int v = c / shp, j = c - v * shp;
      if (j < d.sh) {
        const float sg = d.sums[c], sgx = d.sums[width + c];
        mean = d.saved[c]; rstd = d.saved[width + c];
        coef = __ldg(d.var[v].gamma + j) * rstd;
        m1 = sg / static_cast<float>(d.B); m2 = sgx / static_cast<float>(d.B);
        if (blockIdx.x == 0) {
          if (d.var[v].dbeta) d.var[v].dbeta[j] = sg;
          if (d.var[v].dgamma) d.var[v].dgamma[j] = sgx;
        }
      }
    }
    s_coef[c] = coef; s_m1[c] = m1; s_m2[c] = m2; s_mean[c] = mean; s_rstd[c] = rstd;
  }
  __syncthreads();
  // dZh for this block's rows -> shared tile + operand planes (loads batched: 2 x 8 independent requests per thread)
  for (int base = tid; base < HF_ROWS * width; base += HF_THREADS * 8) {
    float g[8], z[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * HF_THREADS;
      const int r = i / width, c = i - r * width;
      const bool ok = i < HF_ROWS * width && r < nvalid;
      g[u] = ok ? d.G[static_cast<long long>(r0 + r) * d.ldg + c] : 0.f;
      z[u] = ok ? d.Zh[static_cast<long long>(r0 + r) * d.ldz + c] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * HF_THREADS;
      if (i >= HF_ROWS * width) continue;
      const int r = i / width, c = i - r * width;
      float dz = 0.f;
      if (r < nvalid) {
        const long long row = r0 + r;
        const float xh = (z[u] - s_mean[c]) * s_rstd[c];
        dz = s_coef[c] * (g[u] - s_m1[c] - xh * s_m2[c]);
        __nv_bfloat16 h, l;
        split_bf16(dz, h, l);
        static_cast<__nv_bfloat16*>(d.dz_hi)[row * d.ldzp + c] = h;
        static_cast<__nv_bfloat16*>(d.dz_lo)[row * d.ldzp + c] = l;
      }
      sDz[r * (width + 1) + c] = dz;
    }
  }
  __syncthreads();
  // dF[r][k] = sum_c dZh[r][c] W1cat[c][k]: warp = 4 rows, lane = columns k = lane + 32 i
  const int Lp = (L + 7) & ~7;
  for (int kb = 0; kb < L; kb += 256) {
    float acc[HF_RPW][8];
#pragma unroll
    for (int r = 0; r < HF_RPW; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[r][i] = 0.f;
#pragma unroll 4
    for (int c = 0; c < width; ++c) {
      float dzr[HF_RPW];
#pragma unroll
      for (int r = 0; r < HF_RPW; ++r) dzr[r] = sDz[(warp * HF_RPW + r) * (width + 1) + c];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kb + lane + 32 * i;
        const float w = k < L ? sW[c * L + k] : 0.f;
#pragma unroll
        for (int r = 0; r < HF_RPW; ++r) acc[r][i] = fmaf(dzr[r], w, acc[r][i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kb + lane + 32 * i;
      float cs = 0.f;
#pragma unroll
      for (int r = 0; r < HF_RPW; ++r) {
        const int rr = warp * HF_RPW + r;
        if (rr < nvalid && k < Lp) {
          const long long row = r0 + rr;
          const float v = k < L ? acc[r][i] : 0.f;
          if (d.df_hi) {
            __nv_bfloat16 h, l;
            split_bf16(v, h, l);
            static_cast<__nv_bfloat16*>(d.df_hi)[row * d.ldfp + k] = h;
            static_cast<__nv_bfloat16*>(d.df_lo)[row * d.ldfp + k] = l;
          }
          if (d.dF && k < L) d.dF[row * d.lddf + k] = v;
          cs += v;
        }
      }
      if (d.dbias && k < L) sCol[warp * L + k] = cs;
    }
  }
  if (d.dbias) {
    __syncthreads();
    for (int k = tid; k < L; k += HF_THREADS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sCol[w * L + k];
      atomicAdd(d.dbias + k, t);
    }
  }
}

static int heads_check(const fxn_heads_desc* d, const char* who) {
  if (!d || !d->F || !d->Zh || d->B <= 0 || d->L <= 0 || d->sh <= 0 || d->nv <= 0 || d->nv > FXN_HEADS_MAX_VARS)
    return set_error(FXN_ERR_ARG, "%s: bad argument", who);
  int maxc = 0;
  for (int v = 0; v < d->nv; ++v) maxc = d->var[v].C > maxc ? d->var[v].C : maxc;
  if (!fxn_heads_fused_ok(d->L, d->sh, d->nv, maxc)) return set_error(FXN_ERR_UNSUPPORTED, "%s: heads too wide for the fused path", who);
  for (int v = 0; v < d->nv; ++v)
    if (d->var[v].kind != 1 && d->var[v].kind != 2 && d->var[v].y) return set_error(FXN_ERR_UNSUPPORTED, "%s: MSE / cross-entropy heads only", who);
  return 0;
}

}  // namespace fxn

using namespace fxn;

extern "C" int fxn_heads_fused_ok(int L, int sh, int nv, int maxC) {
  const int shp = (sh + 7) & ~7, width = nv * shp;
  if (nv < 1 || nv > FXN_HEADS_MAX_VARS || width > 64 || maxC > HF_MAXC || L % 4 != 0) return 0;
  const long long smem_l1 = (static_cast<long long>(HF_ROWS) * L + static_cast<long long>(L) * (width + 1) + 8 * 2 * 64) * 4;
  const long long smem_bw = (static_cast<long long>(width) * L + HF_ROWS * (width + 1) + 8LL * L) * 4;
  return smem_l1 <= 200 * 1024 && smem_bw <= 200 * 1024;
}

extern "C" int fxn_heads_fwd(const fxn_heads_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = heads_check(d, "fxn_heads_fwd");
  if (rc) return rc;
  if (!d->partials || !d->saved || !d->acc || (d->backward && (!d->G || !d->sums)))
    return set_error(FXN_ERR_ARG, "fxn_heads_fwd: missing workspace");
  if (d->train && d->B < 2) return set_error(FXN_ERR_ARG, "Expected more than 1 value per channel when training");
  const int shp = (d->sh + 7) & ~7, width = d->nv * shp;
  const int blocks = ceil_div(d->B, HF_ROWS);
  const size_t smem = (static_cast<size_t>(HF_ROWS) * d->L + static_cast<size_t>(d->L) * (width + 1) + 8 * 2 * 64) * sizeof(float);
  static size_t attr_l1[2] = {48 * 1024, 48 * 1024};
  const int wide = width > 32 ? 1 : 0;
  if (smem > attr_l1[wide]) {
    cudaError_t e = wide ? cudaFuncSetAttribute(heads_l1_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
                         : cudaFuncSetAttribute(heads_l1_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "heads_l1 attr: %s", cudaGetErrorString(e));
    attr_l1[wide] = smem;
  }
  if (wide) heads_l1_kernel<2><<<blocks, HF_THREADS, smem, stream>>>(*d);
  else heads_l1_kernel<1><<<blocks, HF_THREADS, smem, stream>>>(*d);
  FXN_CHECK_LAUNCH("heads_l1");
  heads_mid_kernel<<<blocks, HF_THREADS, 0, stream>>>(*d);
  FXN_CHECK_LAUNCH("heads_mid");
  return 0;
}

extern "C" int fxn_heads_bwd(const fxn_heads_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = heads_check(d, "fxn_heads_bwd");
  if (rc) return rc;
  if (!d->G || !d->sums || !d->saved || !d->dz_hi || !d->dz_lo || (!d->df_hi && !d->dF) || (d->df_hi && !d->df_lo))
    return set_error(FXN_ERR_ARG, "fxn_heads_bwd: missing buffer");
  const int shp = (d->sh + 7) & ~7, width = d->nv * shp;
  const int blocks = ceil_div(d->B, HF_ROWS);
  const size_t smem = (static_cast<size_t>(width) * d->L + HF_ROWS * (width + 1) + 8 * static_cast<size_t>(d->L)) * sizeof(float);
  static size_t attr_bw = 48 * 1024;
  if (smem > attr_bw) {
    cudaError_t e = cudaFuncSetAttribute(heads_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "heads_bwd attr: %s", cudaGetErrorString(e));
    attr_bw = smem;
  }
  if (d->dbias && d->zero_dbias) {
    cudaError_t e = cudaMemsetAsync(d->dbias, 0, sizeof(float) * d->L, stream);
    if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "heads_bwd memset: %s", cudaGetErrorString(e));
  }
  heads_bwd_kernel<<<blocks, HF_THREADS, smem, stream>>>(*d);
  FXN_CHECK_LAUNCH("heads_bwd");
  return 0;
}
