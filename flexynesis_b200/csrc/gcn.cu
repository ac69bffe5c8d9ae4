// GCN layer of flexGCN (flexynesis/modules.py:252-257) on batched dense node features [B, N, F] with one shared graph.
//
// The reference calls torch_geometric.nn.GCNConv (modules.py:221-226, :254); its published algorithm is
//   H = X W^T ;  O[b, v, :] = sum_{(u -> v) in E^} w_uv * H[b, u, :] + bias ,  w_uv = deg_u^-1/2 deg_v^-1/2
// with E^ = E + the missing self loops and deg = in-degree on the directed list. Aggregation and the
// linear map commute, and F_in <= emb for every layer flexGCN builds, so the kernels aggregate the NARROW tensor and
// transform afterwards: O[b, v, :] = W * (sum_u w_uv X[b, u, :]) + bias. No [B, N, emb] message tensor and no
// scatter-add: the graph is a CSR by destination (forward, weight gradient) and a CSR by source (input gradient), both
// built once from edge_index; each warp owns one node, lane = channel, so every global access is a coalesced row.
// HBM-bound: per layer the forward reads X once (neighbour re-reads of a 256 KB sample slice hit L1/L2) and writes O
// once; the per-channel BatchNorm statistics of O are produced here as per-sample partials (sum, M2 about the sample
// mean) so the norm needs no extra pass.
#include "fxn_internal.h"
#include "ptx.cuh"

namespace fxn {

constexpr int GCN_THREADS = 256;
constexpr int GCN_WARPS = GCN_THREADS / 32;
constexpr int GCN_MAXC = 32;     // channels per node handled by one warp (node_embedding_dim <= 32 in the reference's space)

// sum over the in-edges of node v of w_e * X[b, src_e, lane]   (lane < F)
__device__ __forceinline__ float gather_row(const float* __restrict__ Xb, int F, const int* __restrict__ rowptr,
                                            const int* __restrict__ col, const float* __restrict__ w, int v, int lane) {
  const int e0 = __ldg(rowptr + v), e1 = __ldg(rowptr + v + 1);
  float acc = 0.f;
  int e = e0;
  for (; e + 4 <= e1; e += 4) {          // 4 independent row loads in flight
    const int u0 = __ldg(col + e), u1 = __ldg(col + e + 1), u2 = __ldg(col + e + 2), u3 = __ldg(col + e + 3);
    const float w0 = __ldg(w + e), w1 = __ldg(w + e + 1), w2 = __ldg(w + e + 2), w3 = __ldg(w + e + 3);
    float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
    if (lane < F) {
      x0 = Xb[static_cast<long long>(u0) * F + lane];
      x1 = Xb[static_cast<long long>(u1) * F + lane];
      x2 = Xb[static_cast<long long>(u2) * F + lane];
      x3 = Xb[static_cast<long long>(u3) * F + lane];
    }
    acc = fmaf(w0, x0, acc); acc = fmaf(w1, x1, acc); acc = fmaf(w2, x2, acc); acc = fmaf(w3, x3, acc);
  }
  for (; e < e1; ++e) {
    const int u = __ldg(col + e);
    const float we = __ldg(w + e);
    const float x = (lane < F) ? Xb[static_cast<long long>(u) * F + lane] : 0.f;
    acc = fmaf(we, x, acc);
  }
  return acc;
}

// One CTA per sample (grid-stride over samples), one warp per node (stride over nodes), lane = channel.
__global__ void __launch_bounds__(GCN_THREADS)
gcn_fwd_kernel(const float* __restrict__ X, int B, int N, int Fin, const int* __restrict__ rowptr,
               const int* __restrict__ col, const float* __restrict__ w, const float* __restrict__ W,
               const float* __restrict__ bias, int emb, float* __restrict__ O, float* __restrict__ partials) {
  __shared__ float s_n[GCN_WARPS], s_mean[GCN_WARPS][GCN_MAXC], s_m2[GCN_WARPS][GCN_MAXC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // lane c' keeps row c' of W (emb x Fin) and its bias in registers for the whole kernel
  float wrow[GCN_MAXC];
#pragma unroll
  for (int c = 0; c < GCN_MAXC; ++c) wrow[c] = (lane < emb && c < Fin) ? __ldg(W + lane * Fin + c) : 0.f;
  const float bl = (lane < emb) ? __ldg(bias + lane) : 0.f;

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* Xb = X + static_cast<long long>(b) * N * Fin;
    float* Ob = O + static_cast<long long>(b) * N * emb;
    float shift = 0.f, s1 = 0.f, s2 = 0.f;
    int cnt = 0;
    for (int v = warp; v < N; v += GCN_WARPS) {
      const float agg = gather_row(Xb, Fin, rowptr, col, w, v, lane);
      float o = bl;
#pragma unroll
      for (int c = 0; c < GCN_MAXC; ++c)
        if (c < Fin) o = fmaf(wrow[c], __shfl_sync(0xffffffffu, agg, c), o);
      if (lane < emb) Ob[static_cast<long long>(v) * emb + lane] = o;
      if (cnt == 0) shift = o;            // shifted sums: no cancellation when |mean| >> std
      const float d = o - shift;
      s1 += d;
      s2 = fmaf(d, d, s2);
      ++cnt;
    }
    if (partials != nullptr) {
      const float n = static_cast<float>(cnt);
      const float mean_w = cnt ? shift + s1 / n : 0.f;
      const float m2_w = cnt ? fmaxf(s2 - s1 * s1 / n, 0.f) : 0.f;
      __syncthreads();                    // previous sample's readers are done with the shared arrays
      if (lane == 0) s_n[warp] = n;
      s_mean[warp][lane] = mean_w;
      s_m2[warp][lane] = m2_w;
      __syncthreads();
      if (warp == 0 && lane < emb) {      // Chan merge of the 8 warp partitions -> (sum, M2 about the sample mean)
        float tn = 0.f, tm = 0.f, tm2 = 0.f;
#pragma unroll
        for (int k = 0; k < GCN_WARPS; ++k) {
          const float nk = s_n[k];
          if (nk > 0.f) {
            const float delta = s_mean[k][lane] - tm;
            const float nn = tn + nk;
            tm += delta * nk / nn;
            tm2 += s_m2[k][lane] + delta * delta * tn * nk / nn;
            tn = nn;
          }
        }
        partials[(static_cast<long long>(b) * 2) * emb + lane] = tm * tn;
        partials[(static_cast<long long>(b) * 2 + 1) * emb + lane] = tm2;
      }
    }
  }
}

// Backward of one layer for a sample: weight gradient (recomputes the forward aggregate) and, optionally, the input
// gradient through the transposed graph. dW / dbias accumulate over all (b, v) -> registers -> shared -> one atomic
// per CTA and element.
__global__ void __launch_bounds__(GCN_THREADS)
gcn_bwd_kernel(const float* __restrict__ X, const float* __restrict__ dO, int B, int N, int Fin, int emb,
               const int* __restrict__ rowptr_in, const int* __restrict__ col_in, const float* __restrict__ w_in,
               const int* __restrict__ rowptr_out, const int* __restrict__ col_out, const float* __restrict__ w_out,
               const float* __restrict__ W, float* __restrict__ dW, float* __restrict__ dbias, float* __restrict__ dX) {
  __shared__ float s_dw[GCN_MAXC][GCN_MAXC + 1];
  __shared__ float s_db[GCN_MAXC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < GCN_MAXC * (GCN_MAXC + 1); i += GCN_THREADS) (&s_dw[0][0])[i] = 0.f;
  if (threadIdx.x < GCN_MAXC) s_db[threadIdx.x] = 0.f;
  __syncthreads();
  // lane c keeps column c of W (for dX[c] = sum_c' W[c', c] g[c']); lane c' accumulates row c' of dW
  float wcol[GCN_MAXC], dwacc[GCN_MAXC];
#pragma unroll
  for (int k = 0; k < GCN_MAXC; ++k) {
    wcol[k] = (dX != nullptr && lane < Fin && k < emb) ? __ldg(W + k * Fin + lane) : 0.f;
    dwacc[k] = 0.f;
  }
  float dbacc = 0.f;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* Xb = X + static_cast<long long>(b) * N * Fin;
    const float* dOb = dO + static_cast<long long>(b) * N * emb;
    for (int v = warp; v < N; v += GCN_WARPS) {
      // ---- weight gradient: dW[c', c] += dO[b, v, c'] * Agg[b, v, c] ----
      const float agg = gather_row(Xb, Fin, rowptr_in, col_in, w_in, v, lane);
      const float g = (lane < emb) ? dOb[static_cast<long long>(v) * emb + lane] : 0.f;
      dbacc += g;
#pragma unroll
      for (int c = 0; c < GCN_MAXC; ++c)
        if (c < Fin) dwacc[c] = fmaf(g, __shfl_sync(0xffffffffu, agg, c), dwacc[c]);
      // ---- input gradient: dX[b, v, c] = sum_c' W[c', c] * (sum_{(v -> t)} w dO[b, t, c']) ----
      if (dX != nullptr) {
        const float gg = gather_row(dOb, emb, rowptr_out, col_out, w_out, v, lane);
        float dx = 0.f;
#pragma unroll
        for (int k = 0; k < GCN_MAXC; ++k)
          if (k < emb) dx = fmaf(wcol[k], __shfl_sync(0xffffffffu, gg, k), dx);
        if (lane < Fin) dX[(static_cast<long long>(b) * N + v) * Fin + lane] = dx;
      }
    }
  }
  if (lane < emb) {
#pragma unroll
    for (int c = 0; c < GCN_MAXC; ++c)
      if (c < Fin) atomicAdd(&s_dw[lane][c], dwacc[c]);
    atomicAdd(&s_db[lane], dbacc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < emb * Fin; i += GCN_THREADS) atomicAdd(dW + i, s_dw[i / Fin][i % Fin]);
  if (threadIdx.x < emb && dbias != nullptr) atomicAdd(dbias + threadIdx.x, s_db[threadIdx.x]);
}

// Chan merge of [ntiles][2][pld] column partials into one (sum, M2) record [2][cols]: lets the BatchNorm kernels run
// with ntiles = 1 when the producer emitted thousands of small tiles (one per sample).
__global__ void __launch_bounds__(256)
merge_col_stats_kernel(const float* __restrict__ partials, int ntiles, int tile_rows, long long rows, int cols, int pld,
                       float* __restrict__ merged) {
  __shared__ double s_a[256];
  const int c = blockIdx.x;
  double sum = 0.0;
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) sum += partials[(static_cast<long long>(t) * 2) * pld + c];
  s_a[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_a[threadIdx.x] += s_a[threadIdx.x + o];
    __syncthreads();
  }
  const double total = s_a[0];
  const double mean = total / static_cast<double>(rows);
  __syncthreads();
  double m2 = 0.0;
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
    const long long r0 = static_cast<long long>(t) * tile_rows;
    const double n = static_cast<double>(min(static_cast<long long>(tile_rows), rows - r0));
    const double d = partials[(static_cast<long long>(t) * 2) * pld + c] / n - mean;
    m2 += partials[(static_cast<long long>(t) * 2 + 1) * pld + c] + n * d * d;
  }
  s_a[threadIdx.x] = m2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_a[threadIdx.x] += s_a[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    merged[c] = static_cast<float>(total);
    merged[cols + c] = static_cast<float>(s_a[0]);
  }
}

}  // namespace fxn

using namespace fxn;

static int gcn_grid(int B) {
  int blocks = 148 * 4;       // 4 CTAs of 256 threads per SM keep ~32 warps of gathers in flight
  return blocks < B ? blocks : B;
}

extern "C" int fxn_gcn_fwd(const float* X, int B, int N, int Fin, const int* rowptr, const int* col, const float* w,
                           const float* W, const float* bias, int emb, float* O, float* partials, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!X || !rowptr || !col || !w || !W || !bias || !O) return set_error(FXN_ERR_ARG, "fxn_gcn_fwd: null argument");
  if (B <= 0 || N <= 0 || Fin <= 0 || emb <= 0) return set_error(FXN_ERR_ARG, "fxn_gcn_fwd: empty input");
  if (Fin > GCN_MAXC || emb > GCN_MAXC)
    return set_error(FXN_ERR_UNSUPPORTED, "fxn_gcn_fwd: at most %d channels per node (got in=%d, out=%d)", GCN_MAXC, Fin, emb);
  gcn_fwd_kernel<<<gcn_grid(B), GCN_THREADS, 0, stream>>>(X, B, N, Fin, rowptr, col, w, W, bias, emb, O, partials);
  FXN_CHECK_LAUNCH("gcn_fwd");
  return 0;
}

extern "C" int fxn_gcn_bwd(const float* X, const float* dO, int B, int N, int Fin, int emb, const int* rowptr_in,
                           const int* col_in, const float* w_in, const int* rowptr_out, const int* col_out,
                           const float* w_out, const float* W, float* dW, float* dbias, float* dX, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!X || !dO || !rowptr_in || !col_in || !w_in || !W || !dW) return set_error(FXN_ERR_ARG, "fxn_gcn_bwd: null argument");
  if (dX && (!rowptr_out || !col_out || !w_out)) return set_error(FXN_ERR_ARG, "fxn_gcn_bwd: dX needs the CSR by source");
  if (B <= 0 || N <= 0 || Fin <= 0 || emb <= 0) return set_error(FXN_ERR_ARG, "fxn_gcn_bwd: empty input");
  if (Fin > GCN_MAXC || emb > GCN_MAXC)
    return set_error(FXN_ERR_UNSUPPORTED, "fxn_gcn_bwd: at most %d channels per node (got in=%d, out=%d)", GCN_MAXC, Fin, emb);
  cudaError_t e = cudaMemsetAsync(dW, 0, sizeof(float) * emb * Fin, stream);
  if (e == cudaSuccess && dbias) e = cudaMemsetAsync(dbias, 0, sizeof(float) * emb, stream);
  if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "gcn_bwd memset: %s", cudaGetErrorString(e));
  gcn_bwd_kernel<<<gcn_grid(B), GCN_THREADS, 0, stream>>>(X, dO, B, N, Fin, emb, rowptr_in, col_in, w_in, rowptr_out,
                                                         col_out, w_out, W, dW, dbias, dX);
  FXN_CHECK_LAUNCH("gcn_bwd");
  return 0;
}

extern "C" int fxn_merge_col_stats(const float* partials, int ntiles, int tile_rows, long long rows, int cols, int pld,
                                   float* merged, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!partials || !merged || ntiles <= 0 || tile_rows <= 0 || rows <= 0 || cols <= 0)
    return set_error(FXN_ERR_ARG, "fxn_merge_col_stats: bad argument");
  merge_col_stats_kernel<<<cols, 256, 0, stream>>>(partials, ntiles, tile_rows, rows, cols, pld > 0 ? pld : cols, merged);
  FXN_CHECK_LAUNCH("merge_col_stats");
  return 0;
}
