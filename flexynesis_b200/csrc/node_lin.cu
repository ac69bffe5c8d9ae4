// Root-weight term of the GraphConv / SAGEConv layers of flexGCN (flexynesis/modules.py:221-226: conv in {GC, SAGE}; the
// CLI default is GC, flexynesis/__main__.py:1048-1057). torch_geometric's published algorithms:
//   GraphConv: O[b, v] = lin_rel( sum_{u -> v} X[b, u] ) + lin_root( X[b, v] )          (lin_rel has the bias)
//   SAGEConv : O[b, v] = lin_l ( mean_{u -> v} X[b, u] ) + lin_r  ( X[b, v] )           (lin_l has the bias)
// The neighbour term is the same aggregate-then-transform pass as GCNConv with other edge weights (1, or 1/in-degree; no
// self loops added) and runs through fxn_gcn_fwd / fxn_gcn_bwd. This file adds the per-node root term on top:
//   forward : O[b, v, :] += Wr X[b, v, :]   and the per-sample BatchNorm partials of the finished O
//   backward: dWr = sum_{b, v} dO[b, v]^T X[b, v] ,  dX[b, v, :] += Wr^T dO[b, v, :]
// One CTA per sample (grid-stride), one warp per node, lane = channel: every global access is a coalesced row.
// HBM-bound: forward reads X and O once and writes O once; backward reads dO and X once and updates dX once.
#include "fxn_internal.h"
#include "ptx.cuh"

namespace fxn {

constexpr int NL_THREADS = 256;
constexpr int NL_WARPS = NL_THREADS / 32;
constexpr int NL_MAXC = 32;

__global__ void __launch_bounds__(NL_THREADS)
node_lin_fwd_kernel(const float* __restrict__ X, int B, int N, int Fin, const float* __restrict__ Wr, int emb,
                    float* __restrict__ O, float* __restrict__ partials) {
  __shared__ float s_n[NL_WARPS], s_mean[NL_WARPS][NL_MAXC], s_m2[NL_WARPS][NL_MAXC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float wrow[NL_MAXC];                       // lane c keeps row c of Wr [emb x Fin]
#pragma unroll
  for (int f = 0; f < NL_MAXC; ++f) wrow[f] = (lane < emb && f < Fin) ? __ldg(Wr + lane * Fin + f) : 0.f;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* Xb = X + static_cast<long long>(b) * N * Fin;
    float* Ob = O + static_cast<long long>(b) * N * emb;
    float shift = 0.f, s1 = 0.f, s2 = 0.f;
    int cnt = 0;
    for (int v = warp; v < N; v += NL_WARPS) {
      const float xv = (lane < Fin) ? Xb[static_cast<long long>(v) * Fin + lane] : 0.f;
      float o = (lane < emb) ? Ob[static_cast<long long>(v) * emb + lane] : 0.f;
#pragma unroll
      for (int f = 0; f < NL_MAXC; ++f)
        if (f < Fin) o = fmaf(wrow[f], __shfl_sync(0xffffffffu, xv, f), o);
      if (lane < emb) Ob[static_cast<long long>(v) * emb + lane] = o;
      if (cnt == 0) shift = o;               // shifted sums: no cancellation when |mean| >> std
      const float d = o - shift;
      s1 += d;
      s2 = fmaf(d, d, s2);
      ++cnt;
    }
    if (partials != nullptr) {
      const float n = static_cast<float>(cnt);
      const float mean_w = cnt ? shift + s1 / n : 0.f;
      const float m2_w = cnt ? fmaxf(s2 - s1 * s1 / n, 0.f) : 0.f;
      __syncthreads();
      if (lane == 0) s_n[warp] = n;
      s_mean[warp][lane] = mean_w;
      s_m2[warp][lane] = m2_w;
      __syncthreads();
      if (warp == 0 && lane < emb) {         // Chan merge of the warp partitions -> (sum, M2 about the sample mean)
        float tn = 0.f, tm = 0.f, tm2 = 0.f;
#pragma unroll
        for (int k = 0; k < NL_WARPS; ++k) {
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

__global__ void __launch_bounds__(NL_THREADS)
node_lin_bwd_kernel(const float* __restrict__ X, const float* __restrict__ dO, int B, int N, int Fin, int emb,
                    const float* __restrict__ Wr, float* __restrict__ dWr, float* __restrict__ dX) {
  __shared__ float s_dw[NL_MAXC * NL_MAXC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float wcol[NL_MAXC];                       // lane f keeps column f of Wr: wcol[c] = Wr[c][f]
  float acc[NL_MAXC];                        // lane c accumulates dWr[c][0..Fin)
#pragma unroll
  for (int c = 0; c < NL_MAXC; ++c) {
    wcol[c] = (lane < Fin && c < emb) ? __ldg(Wr + c * Fin + lane) : 0.f;
    acc[c] = 0.f;
  }
  for (int i = threadIdx.x; i < NL_MAXC * NL_MAXC; i += NL_THREADS) s_dw[i] = 0.f;
  __syncthreads();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* Xb = X + static_cast<long long>(b) * N * Fin;
    const float* Gb = dO + static_cast<long long>(b) * N * emb;
    float* dXb = dX ? dX + static_cast<long long>(b) * N * Fin : nullptr;
    for (int v = warp; v < N; v += NL_WARPS) {
      const float g = (lane < emb) ? Gb[static_cast<long long>(v) * emb + lane] : 0.f;
      const float xv = (lane < Fin) ? Xb[static_cast<long long>(v) * Fin + lane] : 0.f;
#pragma unroll
      for (int f = 0; f < NL_MAXC; ++f)
        if (f < Fin) acc[f] = fmaf(g, __shfl_sync(0xffffffffu, xv, f), acc[f]);
      if (dXb != nullptr) {
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < NL_MAXC; ++c)
          if (c < emb) d = fmaf(__shfl_sync(0xffffffffu, g, c), wcol[c], d);
        if (lane < Fin) dXb[static_cast<long long>(v) * Fin + lane] += d;
      }
    }
  }
  if (lane < emb) {
#pragma unroll
    for (int f = 0; f < NL_MAXC; ++f)
      if (f < Fin) atomicAdd(&s_dw[lane * NL_MAXC + f], acc[f]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < emb * Fin; i += NL_THREADS) atomicAdd(dWr + i, s_dw[(i / Fin) * NL_MAXC + (i % Fin)]);
}

}  // namespace fxn

using namespace fxn;

static int node_lin_grid(int B) {
  const int blocks = 148 * 4;
  return blocks < B ? blocks : B;
}

extern "C" int fxn_node_lin_fwd(const float* X, int B, int N, int Fin, const float* Wr, int emb, float* O, float* partials,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!X || !Wr || !O) return set_error(FXN_ERR_ARG, "fxn_node_lin_fwd: null argument");
  if (B <= 0 || N <= 0 || Fin <= 0 || emb <= 0) return set_error(FXN_ERR_ARG, "fxn_node_lin_fwd: empty input");
  if (Fin > NL_MAXC || emb > NL_MAXC)
    return set_error(FXN_ERR_UNSUPPORTED, "fxn_node_lin_fwd: at most %d channels per node (got in=%d, out=%d)", NL_MAXC, Fin, emb);
  node_lin_fwd_kernel<<<node_lin_grid(B), NL_THREADS, 0, stream>>>(X, B, N, Fin, Wr, emb, O, partials);
  FXN_CHECK_LAUNCH("node_lin_fwd");
  return 0;
}

extern "C" int fxn_node_lin_bwd(const float* X, const float* dO, int B, int N, int Fin, int emb, const float* Wr, float* dWr,
                                float* dX, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!X || !dO || !Wr || !dWr) return set_error(FXN_ERR_ARG, "fxn_node_lin_bwd: null argument");
  if (B <= 0 || N <= 0 || Fin <= 0 || emb <= 0) return set_error(FXN_ERR_ARG, "fxn_node_lin_bwd: empty input");
  if (Fin > NL_MAXC || emb > NL_MAXC)
    return set_error(FXN_ERR_UNSUPPORTED, "fxn_node_lin_bwd: at most %d channels per node (got in=%d, out=%d)", NL_MAXC, Fin, emb);
  cudaError_t e = cudaMemsetAsync(dWr, 0, sizeof(float) * emb * Fin, stream);
  if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "node_lin_bwd memset: %s", cudaGetErrorString(e));
  node_lin_bwd_kernel<<<node_lin_grid(B), NL_THREADS, 0, stream>>>(X, dO, B, N, Fin, emb, Wr, dWr, dX);
  FXN_CHECK_LAUNCH("node_lin_bwd");
  return 0;
}
