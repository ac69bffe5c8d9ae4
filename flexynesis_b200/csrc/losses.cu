// Supervisor-head output layer + losses, forward and backward, on CUDA cores (C <= a few dozen columns: a
// tensor-core tile would be > 90 % padding).
//
// Replaces, per target variable:
//   MLP.layer_out                         flexynesis/modules.py:149        logits = D * W2^T (+ b2 iff C > 1, :126-130)
//   compute_loss (MSE / cross-entropy)    flexynesis/models/direct_pred.py:146-190
//   cox_ph_loss                           flexynesis/modules.py:265-305
//   compute_total_loss                    flexynesis/models/direct_pred.py:192-223
//   triplet_loss                          flexynesis/models/triplet_encoder.py:178-194
// Data-dependent Python branches of the reference (no valid label -> loss 0; non-finite Cox -> 0) are
// in-kernel predicates here, so a step never synchronises with the host.
#include "fxn_internal.h"
#include "ptx.cuh"
#include <math_constants.h>

namespace fxn {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int HEAD_WARPS = 8;
constexpr int MAX_CLASSES = 128;

// ------------------------------------------------------------------------------------------------
// forward: logits + (sum of per-row losses, number of valid rows) for MSE / CE
// kind: 0 = no loss here (Cox risk score, loss by cox_kernel), 1 = MSE, 2 = cross-entropy
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HEAD_WARPS * 32)
head_out_fwd_kernel(const float* __restrict__ D, long long ldd, int rows, int sh, const float* __restrict__ W,
                    const float* __restrict__ bias, int C, float* __restrict__ logits, long long ldl, int kind,
                    const float* __restrict__ y, float* __restrict__ acc) {
  __shared__ float s_logit[HEAD_WARPS][MAX_CLASSES];
  __shared__ float s_loss[HEAD_WARPS], s_cnt[HEAD_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float loss = 0.f, cnt = 0.f;
  for (int b = blockIdx.x * HEAD_WARPS + warp; b < rows; b += gridDim.x * HEAD_WARPS) {
    const float* drow = D + static_cast<long long>(b) * ldd;
    for (int c = 0; c < C; ++c) {
      float p = 0.f;
      for (int k = lane; k < sh; k += 32) p = fmaf(drow[k], __ldg(W + static_cast<long long>(c) * sh + k), p);
      p = warp_sum(p);
      if (bias) p += __ldg(bias + c);
      if (lane == 0) {
        logits[static_cast<long long>(b) * ldl + c] = p;
        s_logit[warp][c] = p;
      }
    }
    __syncwarp();
    if (kind == 1) {
      const float yv = y[b];
      if (!isnan(yv)) {
        const float d = s_logit[warp][0] - yv;
        loss += d * d;
        cnt += 1.f;
      }
    } else if (kind == 2) {
      const float yv = y[b];
      if (!isnan(yv) && yv != -1.f) {
        float mx = -CUDART_INF_F;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, s_logit[warp][c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float e = 0.f;
        for (int c = lane; c < C; c += 32) e += expf(s_logit[warp][c] - mx);
        e = warp_sum(e);
        const int yi = static_cast<int>(static_cast<long long>(yv));   // y.long(): truncation
        const float picked = (yi >= 0 && yi < C) ? s_logit[warp][yi] : CUDART_NAN_F;
        loss += (mx + logf(e)) - picked;
        cnt += 1.f;
      }
    }
    __syncwarp();
  }
  if (kind == 0) return;
  if (lane == 0) { s_loss[warp] = loss; s_cnt[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, n = 0.f;
    for (int w = 0; w < HEAD_WARPS; ++w) { l += s_loss[w]; n += s_cnt[w]; }
    if (n > 0.f) { atomicAdd(acc, l); atomicAdd(acc + 1, n); }
  }
}

// ------------------------------------------------------------------------------------------------
// backward: dlogits -> dD (stored), dW2 / db2 (accumulated with atomics; caller zeroes them)
// kind 3 = Cox: dlogit = coef[b] (from cox_kernel)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HEAD_WARPS * 32)
head_out_bwd_kernel(const float* __restrict__ D, long long ldd, int rows, int sh, const float* __restrict__ W, int C,
                    const float* __restrict__ logits, long long ldl, int kind, const float* __restrict__ y,
                    const float* __restrict__ acc, const float* __restrict__ coef, const float* __restrict__ weight,
                    float* __restrict__ dD, long long ldg, float* __restrict__ dW, float* __restrict__ dbias) {
  extern __shared__ float s_dyn[];
  float* s_dW = s_dyn;                              // [C][sh]
  float* s_db = s_dyn + static_cast<size_t>(C) * sh;  // [C]
  float* s_dl = s_db + C;                           // [HEAD_WARPS][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < C * sh + C; i += blockDim.x) s_dyn[i] = 0.f;
  __syncthreads();
  const float w = weight ? *weight : 1.f;
  const float cnt = (kind == 1 || kind == 2) ? acc[1] : 1.f;
  float* dl = s_dl + warp * C;
  for (int b = blockIdx.x * HEAD_WARPS + warp; b < rows; b += gridDim.x * HEAD_WARPS) {
    const float* lrow = logits + static_cast<long long>(b) * ldl;
    // ---- dlogits for this row ----
    if (kind == 1) {
      const float yv = y[b];
      if (lane == 0) dl[0] = (!isnan(yv) && cnt > 0.f) ? 2.f * (lrow[0] - yv) / cnt * w : 0.f;
    } else if (kind == 2) {
      const float yv = y[b];
      const bool valid = !isnan(yv) && yv != -1.f && cnt > 0.f;
      float mx = -CUDART_INF_F;
      for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lrow[c]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float e = 0.f;
      for (int c = lane; c < C; c += 32) e += expf(lrow[c] - mx);
      e = warp_sum(e);
      const int yi = static_cast<int>(static_cast<long long>(yv));
      for (int c = lane; c < C; c += 32) {
        const float sm = expf(lrow[c] - mx) / e;
        dl[c] = valid ? (sm - (c == yi ? 1.f : 0.f)) / cnt * w : 0.f;
      }
    } else {  // Cox
      if (lane == 0) dl[0] = coef[b] * w;
    }
    __syncwarp();
    // ---- dD row, dW2, db2 ----
    const float* drow = D + static_cast<long long>(b) * ldd;
    float* grow = dD + static_cast<long long>(b) * ldg;
    for (int k = lane; k < sh; k += 32) {
      const float dk = drow[k];
      float g = 0.f;
      for (int c = 0; c < C; ++c) {
        const float d = dl[c];
        g = fmaf(d, __ldg(W + static_cast<long long>(c) * sh + k), g);
        if (d != 0.f) atomicAdd(&s_dW[c * sh + k], d * dk);
      }
      grow[k] = g;
    }
    if (dbias)
      for (int c = lane; c < C; c += 32)
        if (dl[c] != 0.f) atomicAdd(&s_db[c], dl[c]);
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * sh; i += blockDim.x)
    if (s_dW[i] != 0.f) atomicAdd(dW + i, s_dW[i]);
  if (dbias)
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      if (s_db[c] != 0.f) atomicAdd(dbias + c, s_db[c]);
}

// ------------------------------------------------------------------------------------------------
// Cox partial likelihood: one CTA sorts the batch by duration (descending, bitonic in shared memory),
// scans exp(o) forward (risk-set sums) and e/S backward (gradient), and writes loss + d loss / d o.
// ------------------------------------------------------------------------------------------------
constexpr int COX_THREADS = 1024;

__device__ __forceinline__ float block_scan_inclusive(float v, float* s_warp, float& total) {
  // inclusive scan over the block in thread order; s_warp has 32 floats
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  total = s_warp[31];
  return v + (warp > 0 ? s_warp[warp - 1] : 0.f);
}

__global__ void __launch_bounds__(COX_THREADS)
cox_kernel(const float* __restrict__ o, long long ldo, const float* __restrict__ dur, const float* __restrict__ evt,
           int n, int npow2, float* __restrict__ coef, float* __restrict__ acc) {
  extern __shared__ float s_cox[];
  float* s_key = s_cox;                                   // [npow2]
  int* s_idx = reinterpret_cast<int*>(s_cox + npow2);     // [npow2]
  float* s_val = s_cox + 2 * static_cast<size_t>(npow2);  // [npow2] scan workspace
  __shared__ float s_warp[32];
  __shared__ double s_red[32];
  __shared__ int s_nvalid;
  __shared__ float s_events;
  const int tid = threadIdx.x;
  if (tid == 0) { s_nvalid = 0; s_events = 0.f; }
  __syncthreads();
  int myvalid = 0;
  float myev = 0.f;
  for (int i = tid; i < npow2; i += COX_THREADS) {
    float k = -CUDART_INF_F;
    int id = -1;
    if (i < n) {
      const float t = dur[i], e = evt[i];
      if (!isnan(t) && !isnan(e)) { k = t; id = i; ++myvalid; myev += e; }
    }
    s_key[i] = k;
    s_idx[i] = id;
  }
  atomicAdd(&s_nvalid, myvalid);
  atomicAdd(&s_events, myev);
  __syncthreads();
  // bitonic sort, descending by key; invalid rows (idx < 0) last; ties by ascending row index
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npow2; i += COX_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float ka = s_key[i], kb = s_key[ixj];
          const int ia = s_idx[i], ib = s_idx[ixj];
          // "a before b" in the final order
          const bool a_first = (ib < 0) ? true : (ia < 0) ? false : (ka > kb) || (ka == kb && ia < ib);
          const bool up = ((i & k) == 0);
          if (a_first != up) { s_key[i] = kb; s_key[ixj] = ka; s_idx[i] = ib; s_idx[ixj] = ia; }
        }
      }
      __syncthreads();
    }
  }
  const int nv = s_nvalid;
  const float events = s_events;
  // forward scan of hazards in sorted order; each thread owns a contiguous chunk
  const int chunk = (npow2 + COX_THREADS - 1) / COX_THREADS;
  const int beg = tid * chunk, end = min(beg + chunk, nv);
  float local = 0.f;
  for (int i = beg; i < end; ++i) {
    const float h = expf(o[static_cast<long long>(s_idx[i]) * ldo]);
    local += h;
    s_val[i] = local;          // chunk-local inclusive sum
  }
  float tot;
  const float incl = block_scan_inclusive(local, s_warp, tot);
  const float offset = incl - local;
  double num = 0.0;            // sum over events of (o_i - log S_i)
  float rlocal = 0.f;
  for (int i = beg; i < end; ++i) {
    const float S = s_val[i] + offset;
    const int id = s_idx[i];
    const float e = evt[id];
    s_val[i] = S;
    if (e == 1.f) {
      num += static_cast<double>(o[static_cast<long long>(id) * ldo]) - static_cast<double>(logf(S));
      rlocal += 1.f / S;
    }
  }
  // reverse scan of e/S: suffix sum R_i = sum_{j >= i, e_j = 1} 1/S_j
  float rtot;
  const float rincl = block_scan_inclusive(rlocal, s_warp, rtot);
  float suffix = rtot - rincl;  // contribution of all later chunks
  // block reduce num
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) num += __shfl_xor_sync(0xffffffffu, num, off);
  if (lane == 0) s_red[warp] = num;
  __syncthreads();
  double total_num = 0.0;
  for (int w = 0; w < COX_THREADS / 32; ++w) total_num += s_red[w];
  float loss = static_cast<float>(-total_num / static_cast<double>(events));
  const bool ok = (nv > 0) && isfinite(loss);
  if (!ok) loss = 0.f;
  // gradient: d loss / d o_k = -(1/E) * ([e_k = 1] - h_k * R_k)
  for (int i = end - 1; i >= beg; --i) {
    const int id = s_idx[i];
    const float e = evt[id];
    if (e == 1.f) suffix += 1.f / s_val[i];
    const float h = expf(o[static_cast<long long>(id) * ldo]);
    coef[id] = ok ? -((e == 1.f ? 1.f : 0.f) - h * suffix) / events : 0.f;
  }
  for (int i = tid; i < n; i += COX_THREADS) {   // rows excluded from the likelihood get no gradient
    const float t = dur[i], e = evt[i];
    if (isnan(t) || isnan(e)) coef[i] = 0.f;
  }
  if (tid == 0) { acc[0] = loss; acc[1] = 1.f; }
}

// ------------------------------------------------------------------------------------------------
// Cox partial likelihood on the whole chip: the same sums as cox_kernel without the sort. Row j is in the risk set of
// row i iff it comes before-or-at i in the (duration descending, row index ascending) order, i.e.
// before(j, i) = t_j > t_i || (t_j == t_i && j <= i); so S_i = sum_j before(j, i) exp(o_j) and, for the gradient,
// R_b = sum over events i with before(b, i) of 1 / S_i. Both are [n x n] pairwise passes: a CTA owns 256 rows x 512
// columns (columns staged in shared memory as (t, value) pairs and broadcast), partial row sums go out as fp32 atomics.
// 16.7 M pairs at n = 4096 are a few microseconds on 128 CTAs, against ~100 us for the single-CTA bitonic sort + scans,
// and there is no row limit. Rows with NaN duration / event are outside every sum.
// ------------------------------------------------------------------------------------------------
constexpr int CP_ROWS = 256, CP_COLS = 512;

// ws: [S n][R n][events, nvalid][double num]
template <int PASS>
__global__ void __launch_bounds__(CP_ROWS)
cox_pair_kernel(const float* __restrict__ o, long long ldo, const float* __restrict__ dur, const float* __restrict__ evt, int n,
                float* __restrict__ ws) {
  __shared__ float2 s_col[CP_COLS];
  __shared__ double s_red[CP_ROWS / 32];
  __shared__ float s_cnt[2][CP_ROWS / 32];
  float* S = ws;
  float* R = ws + n;
  float* counts = ws + 2 * static_cast<size_t>(n);
  double* num = reinterpret_cast<double*>(ws + 2 * static_cast<size_t>(n) + 2);
  const int tid = threadIdx.x, j0 = blockIdx.y * CP_COLS;
  for (int jj = tid; jj < CP_COLS; jj += CP_ROWS) {
    const int j = j0 + jj;
    float t = 0.f, v = 0.f;
    if (j < n) {
      const float tj = dur[j], ej = evt[j];
      if (!isnan(tj) && !isnan(ej)) {
        t = tj;
        if (PASS == 1) v = expf(o[static_cast<long long>(j) * ldo]);
        else v = ej == 1.f ? 1.f / S[j] : 0.f;
      }
    }
    s_col[jj] = make_float2(t, v);          // rows outside the likelihood carry value 0
  }
  __syncthreads();
  const int i = blockIdx.x * CP_ROWS + tid;
  float ti = 0.f, ei = 0.f;
  bool valid = false;
  if (i < n) {
    ti = dur[i]; ei = evt[i];
    valid = !isnan(ti) && !isnan(ei);
  }
  float acc = 0.f;
  if (valid) {
    const int rel = i - j0;                 // tie rule on indices, relative to this column block
#pragma unroll 8
    for (int jj = 0; jj < CP_COLS; ++jj) {
      const float2 c = s_col[jj];
      bool in;
      if (PASS == 1) in = (c.x > ti) || (c.x == ti && jj <= rel);      // before(j, i)
      else in = (ti > c.x) || (ti == c.x && rel <= jj);                // before(i, j)
      acc += in ? c.y : 0.f;
    }
    if (acc != 0.f) atomicAdd((PASS == 1 ? S : R) + i, acc);
  }
  if (blockIdx.y == 0) {                    // once per row block: the scalar sums
    const int lane = tid & 31, warp = tid >> 5;
    if (PASS == 1) {
      float ev = valid ? ei : 0.f, nv = valid ? 1.f : 0.f;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) { ev += __shfl_xor_sync(0xffffffffu, ev, off); nv += __shfl_xor_sync(0xffffffffu, nv, off); }
      if (lane == 0) { s_cnt[0][warp] = ev; s_cnt[1][warp] = nv; }
      __syncthreads();
      if (tid == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < CP_ROWS / 32; ++w) { a += s_cnt[0][w]; b += s_cnt[1][w]; }
        atomicAdd(counts, a);
        atomicAdd(counts + 1, b);
      }
    } else {
      double t = 0.0;
      if (valid && ei == 1.f) t = static_cast<double>(o[static_cast<long long>(i) * ldo]) - static_cast<double>(logf(S[i]));
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
      if (lane == 0) s_red[warp] = t;
      __syncthreads();
      if (tid == 0) {
        double a = 0.0;
        for (int w = 0; w < CP_ROWS / 32; ++w) a += s_red[w];
        atomicAdd(num, a);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
cox_finish_kernel(const float* __restrict__ o, long long ldo, const float* __restrict__ dur, const float* __restrict__ evt,
                  int n, const float* __restrict__ ws, float* __restrict__ coef, float* __restrict__ acc) {
  const float* R = ws + n;
  const float events = ws[2 * static_cast<size_t>(n)], nv = ws[2 * static_cast<size_t>(n) + 1];
  const double num = *reinterpret_cast<const double*>(ws + 2 * static_cast<size_t>(n) + 2);
  float loss = static_cast<float>(-num / static_cast<double>(events));
  const bool ok = (nv > 0.f) && isfinite(loss);
  if (!ok) loss = 0.f;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float t = dur[i], e = evt[i];
    float c = 0.f;
    if (ok && !isnan(t) && !isnan(e)) c = -((e == 1.f ? 1.f : 0.f) - expf(o[static_cast<long long>(i) * ldo]) * R[i]) / events;
    coef[i] = c;
  }
  if (i == 0) { acc[0] = loss; acc[1] = 1.f; }
}

// ------------------------------------------------------------------------------------------------
// total loss (Kendall uncertainty weighting) + upstream weights for the backward pass
// kinds[k]: 1 = mean of (sum, count) in acc[k]; 3 = value already in acc[k][0]
// ------------------------------------------------------------------------------------------------
__global__ void total_loss_kernel(int n, const float* __restrict__ acc, const int* __restrict__ kinds,
                                  const float* const* __restrict__ log_vars, float* const* __restrict__ dlog_vars,
                                  int weighting, float* __restrict__ out) {
  // out: [0..n) losses, [n] total, [n+1] val_total (unweighted sum), [n+2 .. 2n+2) weights
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float total = 0.f, val = 0.f;
  const bool wgt = weighting && n > 1;
  for (int k = 0; k < n; ++k) {
    float l;
    if (kinds[k] == 3) l = acc[2 * k];
    else l = acc[2 * k + 1] > 0.f ? acc[2 * k] / acc[2 * k + 1] : 0.f;
    out[k] = l;
    val += l;
    float w = 1.f;
    if (wgt) {
      const float s = *log_vars[k];
      w = expf(-s);
      total += w * l + s;
      if (dlog_vars && dlog_vars[k]) *dlog_vars[k] = 1.f - w * l;
    } else {
      total += l;
    }
    out[n + 2 + k] = w;
  }
  out[n] = total;
  out[n + 1] = val;
}

// ------------------------------------------------------------------------------------------------
// triplet margin loss on fused embeddings, forward + gradient planes in one pass
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
triplet_fwd_kernel(const float* __restrict__ A, const float* __restrict__ P, const float* __restrict__ N,
                   long long ld, int rows, int L, float margin, float* __restrict__ rowloss, float* __restrict__ acc) {
  __shared__ float s_part[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float loss = 0.f;
  for (int b = blockIdx.x * 8 + warp; b < rows; b += gridDim.x * 8) {
    const float* a = A + static_cast<long long>(b) * ld;
    const float* p = P + static_cast<long long>(b) * ld;
    const float* q = N + static_cast<long long>(b) * ld;
    float dp = 0.f, dn = 0.f;
    for (int k = lane; k < L; k += 32) {
      const float u = a[k] - p[k], v = a[k] - q[k];
      dp = fmaf(u, u, dp);
      dn = fmaf(v, v, dn);
    }
    dp = warp_sum(dp);
    dn = warp_sum(dn);
    const float l = fmaxf(dp - dn + margin, 0.f);
    if (lane == 0) rowloss[b] = l;
    loss += l;
  }
  if (lane == 0) s_part[warp] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(acc, t);
    if (blockIdx.x == 0) atomicAdd(acc + 1, static_cast<float>(rows));
  }
}

// dA = 2(n - p)/B, dP = -2(a - p)/B, dN = 2(a - n)/B on active rows, times the loss weight; added to `add*` if given
__global__ void __launch_bounds__(256)
triplet_bwd_kernel(const float* __restrict__ A, const float* __restrict__ P, const float* __restrict__ N,
                   long long ld, int rows, int L, const float* __restrict__ rowloss, const float* __restrict__ weight,
                   float* __restrict__ dA, float* __restrict__ dP, float* __restrict__ dN, long long ldg,
                   int accumulate_a) {
  const float w = (weight ? *weight : 1.f) * 2.f / static_cast<float>(rows);
  const long long total = static_cast<long long>(rows) * L;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / L), k = static_cast<int>(i - static_cast<long long>(b) * L);
    const bool active = rowloss[b] > 0.f;
    const float a = A[b * ld + k], p = P[b * ld + k], q = N[b * ld + k];
    const float ga = active ? w * (q - p) : 0.f;
    if (accumulate_a) dA[b * ldg + k] += ga; else dA[b * ldg + k] = ga;
    dP[b * ldg + k] = active ? -w * (a - p) : 0.f;
    dN[b * ldg + k] = active ? w * (a - q) : 0.f;
  }
}

}  // namespace fxn

using namespace fxn;

extern "C" int fxn_head_out_fwd(const float* D, long long ldd, int rows, int sh, const float* W, const float* bias,
                                int C, float* logits, long long ldl, int kind, const float* y, float* acc,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!D || !W || !logits || rows <= 0 || sh <= 0 || C <= 0) return set_error(FXN_ERR_ARG, "fxn_head_out_fwd: bad argument");
  if (C > MAX_CLASSES) return set_error(FXN_ERR_UNSUPPORTED, "fxn_head_out_fwd: more than %d classes", MAX_CLASSES);
  if ((kind == 1 || kind == 2) && (!y || !acc)) return set_error(FXN_ERR_ARG, "fxn_head_out_fwd: loss needs y and acc");
  if (kind == 1 && C != 1) return set_error(FXN_ERR_ARG, "fxn_head_out_fwd: MSE head must have one output");
  int blocks = ceil_div(rows, HEAD_WARPS);
  if (blocks > 148 * 4) blocks = 148 * 4;
  head_out_fwd_kernel<<<blocks, HEAD_WARPS * 32, 0, stream>>>(D, ldd, rows, sh, W, bias, C, logits, ldl, kind, y, acc);
  FXN_CHECK_LAUNCH("head_out_fwd");
  return 0;
}

extern "C" int fxn_head_out_bwd(const float* D, long long ldd, int rows, int sh, const float* W, int C,
                                const float* logits, long long ldl, int kind, const float* y, const float* acc,
                                const float* coef, const float* weight, float* dD, long long ldg, float* dW,
                                float* dbias, int prezeroed, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!D || !W || !logits || !dD || !dW) return set_error(FXN_ERR_ARG, "fxn_head_out_bwd: null argument");
  if (C > MAX_CLASSES) return set_error(FXN_ERR_UNSUPPORTED, "fxn_head_out_bwd: more than %d classes", MAX_CLASSES);
  if (kind == 3 && !coef) return set_error(FXN_ERR_ARG, "fxn_head_out_bwd: Cox needs coef");
  const size_t smem = (static_cast<size_t>(C) * sh + C + static_cast<size_t>(HEAD_WARPS) * C) * sizeof(float);
  if (smem > 200 * 1024) return set_error(FXN_ERR_UNSUPPORTED, "fxn_head_out_bwd: C*sh too large");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(head_out_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "head_out_bwd attr: %s", cudaGetErrorString(e));
  }
  if (!prezeroed) {
    cudaError_t e = cudaMemsetAsync(dW, 0, sizeof(float) * C * sh, stream);
    if (e == cudaSuccess && dbias) e = cudaMemsetAsync(dbias, 0, sizeof(float) * C, stream);
    if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "head_out_bwd memset: %s", cudaGetErrorString(e));
  }
  int blocks = ceil_div(rows, HEAD_WARPS * 2);
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  head_out_bwd_kernel<<<blocks, HEAD_WARPS * 32, smem, stream>>>(D, ldd, rows, sh, W, C, logits, ldl, kind, y, acc, coef,
                                                                weight, dD, ldg, dW, dbias);
  FXN_CHECK_LAUNCH("head_out_bwd");
  return 0;
}

extern "C" int fxn_cox_max_rows(void) { return 16384; }

extern "C" int fxn_cox_fwd(const float* o, long long ldo, const float* durations, const float* events, int n,
                           float* coef, float* acc, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!o || !durations || !events || !coef || !acc || n <= 0) return set_error(FXN_ERR_ARG, "fxn_cox_fwd: bad argument");
  if (n > 16384) return set_error(FXN_ERR_UNSUPPORTED, "fxn_cox_fwd: batch of %d rows exceeds the single-CTA limit 16384", n);
  int npow2 = 32;
  while (npow2 < n) npow2 <<= 1;
  const size_t smem = 3 * static_cast<size_t>(npow2) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(cox_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16384 * 4);
    if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "cox attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  cox_kernel<<<1, COX_THREADS, smem, stream>>>(o, ldo, durations, events, n, npow2, coef, acc);
  FXN_CHECK_LAUNCH("cox");
  return 0;
}

extern "C" long long fxn_cox_ws_floats(int n) { return 2LL * n + 4; }

extern "C" int fxn_cox_fwd_ws(const float* o, long long ldo, const float* durations, const float* events, int n, float* coef,
                              float* acc, float* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!o || !durations || !events || !coef || !acc || !workspace || n <= 0)
    return set_error(FXN_ERR_ARG, "fxn_cox_fwd_ws: bad argument");
  if (reinterpret_cast<uintptr_t>(workspace) & 7) return set_error(FXN_ERR_ARG, "fxn_cox_fwd_ws: workspace must be 8-byte aligned");
  cudaError_t e = cudaMemsetAsync(workspace, 0, sizeof(float) * static_cast<size_t>(fxn_cox_ws_floats(n)), stream);
  if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "cox workspace memset: %s", cudaGetErrorString(e));
  dim3 grid((n + CP_ROWS - 1) / CP_ROWS, (n + CP_COLS - 1) / CP_COLS);
  cox_pair_kernel<1><<<grid, CP_ROWS, 0, stream>>>(o, ldo, durations, events, n, workspace);
  FXN_CHECK_LAUNCH("cox_pair_1");
  cox_pair_kernel<2><<<grid, CP_ROWS, 0, stream>>>(o, ldo, durations, events, n, workspace);
  FXN_CHECK_LAUNCH("cox_pair_2");
  cox_finish_kernel<<<(n + 255) / 256, 256, 0, stream>>>(o, ldo, durations, events, n, workspace, coef, acc);
  FXN_CHECK_LAUNCH("cox_finish");
  return 0;
}

extern "C" int fxn_total_loss(int n, const float* acc, const int* kinds, const float* const* log_vars,
                              float* const* dlog_vars, int weighting, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0 || !acc || !kinds || !out) return set_error(FXN_ERR_ARG, "fxn_total_loss: bad argument");
  if (weighting && n > 1 && !log_vars) return set_error(FXN_ERR_ARG, "fxn_total_loss: weighting needs log_vars");
  total_loss_kernel<<<1, 32, 0, stream>>>(n, acc, kinds, log_vars, dlog_vars, weighting, out);
  FXN_CHECK_LAUNCH("total_loss");
  return 0;
}

extern "C" int fxn_triplet_fwd(const float* A, const float* P, const float* N, long long ld, int rows, int L,
                               float margin, float* rowloss, float* acc, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!A || !P || !N || !rowloss || !acc || rows <= 0 || L <= 0) return set_error(FXN_ERR_ARG, "fxn_triplet_fwd: bad argument");
  int blocks = ceil_div(rows, 8);
  if (blocks > 148 * 4) blocks = 148 * 4;
  triplet_fwd_kernel<<<blocks, 256, 0, stream>>>(A, P, N, ld, rows, L, margin, rowloss, acc);
  FXN_CHECK_LAUNCH("triplet_fwd");
  return 0;
}

extern "C" int fxn_triplet_bwd(const float* A, const float* P, const float* N, long long ld, int rows, int L,
                               const float* rowloss, const float* weight, float* dA, float* dP, float* dN,
                               long long ldg, int accumulate_a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!A || !P || !N || !rowloss || !dA || !dP || !dN) return set_error(FXN_ERR_ARG, "fxn_triplet_bwd: null argument");
  int blocks = ceil_div(static_cast<long long>(rows) * L, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  triplet_bwd_kernel<<<blocks, 256, 0, stream>>>(A, P, N, ld, rows, L, rowloss, weight, dA, dP, dN, ldg, accumulate_a);
  FXN_CHECK_LAUNCH("triplet_bwd");
  return 0;
}
