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
#include <cstdlib>

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


// ================================================================================================================
// Shared-memory staged variants (the default): the per-sample node-feature slab is staged ONCE in shared memory and
// every neighbour gather is a shared-memory read, so HBM sees each activation exactly once per pass and L2 is not
// asked to serve 11 re-reads of 128-byte rows at DRAM-like latency (the L2-gather kernels above ran at 3 % of the HBM
// roofline, profiles/r01_breakdown_v3_cfg4.log).
//  * narrow inputs (F_in <= 4, the first layer: one feature per omics layer): stage X[b] and its aggregate, expand to emb
//    channels with lane = channel.
//  * wide inputs: aggregation and the linear map commute and both are separable over OUTPUT channels, so a CTA owns
//    (sample, group of 8 output channels): it computes H[:, group] = X[b] W[group]^T into shared memory (64 KB for 2000
//    nodes -> 3 CTAs per SM) and gathers from there. X[b] is re-read once per group from L2.
// ================================================================================================================
constexpr int GCN_CG = 8;                       // channels per group in the wide kernels

template <int C>
__device__ __forceinline__ void gather_accumulate(const float* __restrict__ s_src, int u, float we, float* acc) {
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int j = 0; j < C; j += 4) {
      const float4 x = *reinterpret_cast<const float4*>(s_src + u * C + j);
      acc[j] = fmaf(we, x.x, acc[j]); acc[j + 1] = fmaf(we, x.y, acc[j + 1]);
      acc[j + 2] = fmaf(we, x.z, acc[j + 2]); acc[j + 3] = fmaf(we, x.w, acc[j + 3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < C; ++j) acc[j] = fmaf(we, s_src[u * C + j], acc[j]);
  }
}

// acc[0..C) = sum over the CSR row of node v of w_e * s_src[col_e][0..C). The edge records come from L2 (the graph is
// shared by every sample but does not fit beside the slab in shared memory), so they are fetched eight at a time
// before the shared-memory gathers: one exposed L2 latency per 8 edges instead of one per edge.
template <int C>
__device__ __forceinline__ void gather_smem(const float* __restrict__ s_src, const int* __restrict__ rowptr,
                                            const int* __restrict__ col, const float* __restrict__ w, int v, float* acc) {
#pragma unroll
  for (int j = 0; j < C; ++j) acc[j] = 0.f;
  const int e0 = __ldg(rowptr + v), e1 = __ldg(rowptr + v + 1);
  for (int e = e0; e < e1; e += 8) {
    int u[8];
    float we[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const bool ok = e + t < e1;
      u[t] = ok ? __ldg(col + e + t) : 0;
      we[t] = ok ? __ldg(w + e + t) : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 8; ++t)
      if (e + t < e1) gather_accumulate<C>(s_src, u[t], we[t], acc);
  }
}

// block-wide sum of `n` per-thread values (n <= 16) -> s_out[n]; s_tmp has GCN_WARPS * 16 floats
__device__ __forceinline__ void block_sum_vec(float* vals, int n, float* s_tmp, float* s_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = 0; j < n; ++j) {
    float v = vals[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_tmp[warp * 16 + j] = v;
  }
  __syncthreads();
  if (threadIdx.x < n) {
    float t = 0.f;
    for (int k = 0; k < GCN_WARPS; ++k) t += s_tmp[k * 16 + threadIdx.x];
    s_out[threadIdx.x] = t;
  }
  __syncthreads();
}

// sH[u][0..8) = sum_c sWg[j][c] * rows[u][c] for every node u, with COALESCED row loads: 8 lanes share a node, lane k loads
// the float4 of channels 4k..4k+3 (a warp instruction covers 4 rows = 4 full 128-byte lines; a thread-per-node load of
// the same rows costs 32 L1 wavefronts per instruction and made the LSU pipe the bottleneck). Each lane forms the 8
// partial dot products of its 4 channels, then a recursive-halving butterfly over the 8 lanes (4 + 2 + 1 shuffles)
// leaves output channel k on lane k. Requires C % 4 == 0 and C <= 32.
__device__ __forceinline__ void transform_rows_coalesced(const float* __restrict__ rows, int N, int C,
                                                         const float* __restrict__ sWg, float* __restrict__ sH) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nl = lane >> 3, k = lane & 7;
  const bool have = 4 * k < C;
  float4 wv[GCN_CG];
#pragma unroll
  for (int j = 0; j < GCN_CG; ++j)
    wv[j] = have ? *reinterpret_cast<const float4*>(sWg + j * GCN_MAXC + 4 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
  // 4 passes (16 nodes) per iteration with all four row loads issued before the first use: the loop is otherwise one
  // exposed DRAM/L2 latency per pass (30 % of all stall samples in profiles/r01_ncu_gcn_fwd_wide_v1.txt)
  for (int base = warp * 16; base < N; base += GCN_WARPS * 16) {
    float4 xs[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int u = base + 4 * t + nl;
      xs[t] = (u < N && have) ? __ldg(reinterpret_cast<const float4*>(rows + static_cast<long long>(u) * C + 4 * k))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int u = base + 4 * t + nl;
      const float4 x = xs[t];
      float p[GCN_CG];
#pragma unroll
      for (int j = 0; j < GCN_CG; ++j) p[j] = wv[j].x * x.x + wv[j].y * x.y + wv[j].z * x.z + wv[j].w * x.w;
      float q[4], r2[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float send = (k & 4) ? p[i] : p[i + 4];
        const float keep = (k & 4) ? p[i + 4] : p[i];
        q[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float send = (k & 2) ? q[i] : q[i + 2];
        const float keep = (k & 2) ? q[i + 2] : q[i];
        r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      const float send = (k & 1) ? r2[0] : r2[1];
      const float keep = (k & 1) ? r2[1] : r2[0];
      const float h = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      if (u < N) sH[u * GCN_CG + k] = h;
    }
  }
}

// Write the 8-channel results of the 32 consecutive nodes owned by a warp (vals[0..8) on lane = node) to
// dst[node][c0 .. c0+8) through a 1 KB per-warp staging tile: each store instruction covers 4 nodes x 32 bytes.
__device__ __forceinline__ void store_group_rows(float* __restrict__ s_stage, const float* vals, float* __restrict__ dst, int ld,
                                                 int node0, int N, int c0, int cmax) {
  const int lane = threadIdx.x & 31;
  const int nl = lane >> 3, k = lane & 7;
  __syncwarp();
  *reinterpret_cast<float4*>(s_stage + lane * GCN_CG) = make_float4(vals[0], vals[1], vals[2], vals[3]);
  *reinterpret_cast<float4*>(s_stage + lane * GCN_CG + 4) = make_float4(vals[4], vals[5], vals[6], vals[7]);
  __syncwarp();
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    const int n = pass * 4 + nl;
    if (node0 + n < N && c0 + k < cmax) dst[static_cast<long long>(node0 + n) * ld + c0 + k] = s_stage[n * GCN_CG + k];
  }
}

// ---- narrow forward: one CTA per sample ----
template <int F>
__global__ void __launch_bounds__(GCN_THREADS)
gcn_fwd_narrow_kernel(const float* __restrict__ X, int B, int N, const int* __restrict__ rowptr, const int* __restrict__ col,
                      const float* __restrict__ w, const float* __restrict__ W, const float* __restrict__ bias, int emb,
                      float* __restrict__ O, float* __restrict__ partials) {
  extern __shared__ __align__(16) float s_dyn[];
  float* sX = s_dyn;                 // [N][F]
  float* sA = s_dyn + N * F;         // [N][F] aggregate
  __shared__ float s_tmp[GCN_WARPS * 16], s_out[16];
  __shared__ float s_s1[GCN_WARPS][GCN_MAXC], s_s2[GCN_WARPS][GCN_MAXC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float wrow[F];
#pragma unroll
  for (int f = 0; f < F; ++f) wrow[f] = lane < emb ? __ldg(W + lane * F + f) : 0.f;
  const float bl = lane < emb ? __ldg(bias + lane) : 0.f;
  (void)s_tmp; (void)s_out;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* Xb = X + static_cast<long long>(b) * N * F;
    __syncthreads();
    for (int i = threadIdx.x; i < N * F; i += GCN_THREADS) sX[i] = Xb[i];
    __syncthreads();
    for (int v = threadIdx.x; v < N; v += GCN_THREADS) {
      float acc[F];
      gather_smem<F>(sX, rowptr, col, w, v, acc);
#pragma unroll
      for (int f = 0; f < F; ++f) sA[v * F + f] = acc[f];
    }
    __syncthreads();
    float* Ob = O + static_cast<long long>(b) * N * emb;
    float s1 = 0.f, s2 = 0.f;                       // sums of (o - bias): the shift removes the common offset
    for (int v = warp; v < N; v += GCN_WARPS) {
      float d = 0.f;
#pragma unroll
      for (int f = 0; f < F; ++f) d = fmaf(wrow[f], sA[v * F + f], d);
      if (lane < emb) Ob[static_cast<long long>(v) * emb + lane] = d + bl;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
    if (partials != nullptr) {
      s_s1[warp][lane] = s1; s_s2[warp][lane] = s2;
      __syncthreads();
      if (warp == 0 && lane < emb) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 0; k < GCN_WARPS; ++k) { t1 += s_s1[k][lane]; t2 += s_s2[k][lane]; }
        const float n = static_cast<float>(N);
        partials[(static_cast<long long>(b) * 2) * emb + lane] = t1 + n * bl;
        partials[(static_cast<long long>(b) * 2 + 1) * emb + lane] = fmaxf(t2 - t1 * t1 / n, 0.f);
      }
    }
  }
}

// ---- narrow backward (first layer: no input gradient): one CTA per sample ----
template <int F>
__global__ void __launch_bounds__(GCN_THREADS)
gcn_bwd_narrow_kernel(const float* __restrict__ X, const float* __restrict__ dO, int B, int N, int emb,
                      const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ w,
                      float* __restrict__ dW, float* __restrict__ dbias) {
  extern __shared__ __align__(16) float s_dyn[];
  float* sX = s_dyn;
  float* sA = s_dyn + N * F;
  __shared__ float s_dw[GCN_MAXC][F + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < GCN_MAXC)
    for (int f = 0; f <= F; ++f) s_dw[threadIdx.x][f] = 0.f;
  float dwacc[F], dbacc = 0.f;
#pragma unroll
  for (int f = 0; f < F; ++f) dwacc[f] = 0.f;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* Xb = X + static_cast<long long>(b) * N * F;
    __syncthreads();
    for (int i = threadIdx.x; i < N * F; i += GCN_THREADS) sX[i] = Xb[i];
    __syncthreads();
    for (int v = threadIdx.x; v < N; v += GCN_THREADS) {
      float acc[F];
      gather_smem<F>(sX, rowptr, col, w, v, acc);
#pragma unroll
      for (int f = 0; f < F; ++f) sA[v * F + f] = acc[f];
    }
    __syncthreads();
    const float* dOb = dO + static_cast<long long>(b) * N * emb;
    for (int v = warp; v < N; v += GCN_WARPS) {
      const float g = lane < emb ? dOb[static_cast<long long>(v) * emb + lane] : 0.f;
      dbacc += g;
#pragma unroll
      for (int f = 0; f < F; ++f) dwacc[f] = fmaf(g, sA[v * F + f], dwacc[f]);
    }
  }
  __syncthreads();
  if (lane < emb) {
#pragma unroll
    for (int f = 0; f < F; ++f) atomicAdd(&s_dw[lane][f], dwacc[f]);
    atomicAdd(&s_dw[lane][F], dbacc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < emb * F; i += GCN_THREADS) atomicAdd(dW + i, s_dw[i / F][i % F]);
  if (threadIdx.x < emb && dbias != nullptr) atomicAdd(dbias + threadIdx.x, s_dw[threadIdx.x][F]);
}

// ---- wide forward: one CTA per (sample, group of GCN_CG output channels) ----
__global__ void __launch_bounds__(GCN_THREADS, 3)
gcn_fwd_wide_kernel(const float* __restrict__ X, int B, int N, int Fin, const int* __restrict__ rowptr,
                    const int* __restrict__ col, const float* __restrict__ w, const float* __restrict__ W,
                    const float* __restrict__ bias, int emb, float* __restrict__ O, float* __restrict__ partials) {
  extern __shared__ __align__(16) float s_dyn[];
  float* sH = s_dyn;                                   // [N][GCN_CG]
  __shared__ __align__(16) float sW[GCN_CG * GCN_MAXC];              // rows c0 .. c0+CG of W (zero beyond emb)
  __shared__ __align__(16) float s_stage[GCN_WARPS * 32 * GCN_CG];
  __shared__ float s_tmp[GCN_WARPS * 16], s_out[16];
  // the grid is a multiple of the group count: a CTA keeps one channel group for its whole life (weights staged once)
  const int groups = (emb + GCN_CG - 1) / GCN_CG;
  const int c0 = static_cast<int>(blockIdx.x % groups) * GCN_CG;
  for (int i = threadIdx.x; i < GCN_CG * Fin; i += GCN_THREADS) {
    const int j = i / Fin, c = i - j * Fin;
    sW[j * GCN_MAXC + c] = (c0 + j < emb) ? __ldg(W + (c0 + j) * Fin + c) : 0.f;
  }
  for (int b = blockIdx.x / groups; b < B; b += gridDim.x / groups) {
    __syncthreads();
    const float* Xb = X + static_cast<long long>(b) * N * Fin;
    const bool vec = (Fin % 4 == 0) && ((reinterpret_cast<uintptr_t>(Xb) & 15) == 0);
    if (vec) {
      transform_rows_coalesced(Xb, N, Fin, sW, sH);
    } else {
      for (int u = threadIdx.x; u < N; u += GCN_THREADS) {
        float h[GCN_CG];
#pragma unroll
        for (int j = 0; j < GCN_CG; ++j) h[j] = 0.f;
        const float* xr = Xb + static_cast<long long>(u) * Fin;
        for (int c = 0; c < Fin; ++c) {
          const float x = xr[c];
#pragma unroll
          for (int j = 0; j < GCN_CG; ++j) h[j] = fmaf(sW[j * GCN_MAXC + c], x, h[j]);
        }
        *reinterpret_cast<float4*>(sH + u * GCN_CG) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(sH + u * GCN_CG + 4) = make_float4(h[4], h[5], h[6], h[7]);
      }
    }
    __syncthreads();
    float bj[GCN_CG], st[2 * GCN_CG];
#pragma unroll
    for (int j = 0; j < GCN_CG; ++j) { bj[j] = (c0 + j < emb) ? __ldg(bias + c0 + j) : 0.f; st[j] = 0.f; st[GCN_CG + j] = 0.f; }
    for (int v0 = (threadIdx.x >> 5) * 32; v0 < N; v0 += GCN_THREADS) {      // a warp owns 32 consecutive nodes
      const int v = v0 + (threadIdx.x & 31);
      float acc[GCN_CG];
      if (v < N) {
        gather_smem<GCN_CG>(sH, rowptr, col, w, v, acc);
#pragma unroll
        for (int j = 0; j < GCN_CG; ++j) { st[j] += acc[j]; st[GCN_CG + j] = fmaf(acc[j], acc[j], st[GCN_CG + j]); acc[j] += bj[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < GCN_CG; ++j) acc[j] = 0.f;
      }
      store_group_rows(s_stage + (threadIdx.x >> 5) * (32 * GCN_CG), acc, O + static_cast<long long>(b) * N * emb, emb, v0, N, c0, emb);
    }
    if (partials != nullptr) {
      block_sum_vec(st, 2 * GCN_CG, s_tmp, s_out);
      if (threadIdx.x < GCN_CG && c0 + threadIdx.x < emb) {
        const float n = static_cast<float>(N);
        const float t1 = s_out[threadIdx.x], t2 = s_out[GCN_CG + threadIdx.x];
        const float bb = __ldg(bias + c0 + threadIdx.x);
        partials[(static_cast<long long>(b) * 2) * emb + c0 + threadIdx.x] = t1 + n * bb;
        partials[(static_cast<long long>(b) * 2 + 1) * emb + c0 + threadIdx.x] = fmaxf(t2 - t1 * t1 / n, 0.f);
      }
    }
  }
}

// ---- wide backward: one CTA per (sample, group of GCN_CG INPUT channels) ----
// phase A: dW[:, group] += dO[b]^T (A^ X[b])[:, group]   (aggregate recomputed from the staged slab)
// phase B: dX[b][:, group] = A^T (dO[b] W)[:, group]     (transform first, then gather over the out-edges)
__global__ void __launch_bounds__(GCN_THREADS, 3)
gcn_bwd_wide_kernel(const float* __restrict__ X, const float* __restrict__ dO, int B, int N, int Fin, int emb,
                    const int* __restrict__ rowptr_in, const int* __restrict__ col_in, const float* __restrict__ w_in,
                    const int* __restrict__ rowptr_out, const int* __restrict__ col_out, const float* __restrict__ w_out,
                    const float* __restrict__ W, float* __restrict__ dW, float* __restrict__ dbias, float* __restrict__ dX) {
  extern __shared__ __align__(16) float s_dyn[];
  float* sS = s_dyn;                                   // [N][GCN_CG]: X slab in phase A, transformed dO in phase B
  __shared__ __align__(16) float sWt[GCN_CG * GCN_MAXC];             // sWt[j][c'] = W[c'][c0 + j]
  __shared__ float s_dw[GCN_MAXC][GCN_CG + 1];
  __shared__ __align__(16) float s_stage[GCN_WARPS * 32 * GCN_CG];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = (Fin + GCN_CG - 1) / GCN_CG;
  const int c0 = static_cast<int>(blockIdx.x % groups) * GCN_CG;     // fixed input-channel group per CTA
  for (int i = threadIdx.x; i < GCN_MAXC * (GCN_CG + 1); i += GCN_THREADS) (&s_dw[0][0])[i] = 0.f;
  for (int i = threadIdx.x; i < GCN_CG * emb; i += GCN_THREADS) {
    const int j = i / emb, cc = i - j * emb;
    sWt[j * GCN_MAXC + cc] = (c0 + j < Fin) ? __ldg(W + cc * Fin + c0 + j) : 0.f;
  }
  float dwacc[GCN_CG], dbacc = 0.f;
#pragma unroll
  for (int j = 0; j < GCN_CG; ++j) dwacc[j] = 0.f;
  for (int b = blockIdx.x / groups; b < B; b += gridDim.x / groups) {
    __syncthreads();
    const float* Xb = X + static_cast<long long>(b) * N * Fin;
    const float* dOb = dO + static_cast<long long>(b) * N * emb;
    for (int i = threadIdx.x; i < N * GCN_CG; i += GCN_THREADS) {
      const int u = i / GCN_CG, j = i - u * GCN_CG;
      sS[i] = (c0 + j < Fin) ? Xb[static_cast<long long>(u) * Fin + c0 + j] : 0.f;
    }
    __syncthreads();
    // ---- phase A: lane c' accumulates dW[c', c0 + j] += dO[b, v, c'] * Agg[v, j]; Agg[v, j] is broadcast from lane j ----
    for (int base = warp * 32; base < N; base += GCN_WARPS * 32) {
      const int v = base + lane;
      float agg[GCN_CG];
      if (v < N) {
        gather_smem<GCN_CG>(sS, rowptr_in, col_in, w_in, v, agg);
      } else {
#pragma unroll
        for (int j = 0; j < GCN_CG; ++j) agg[j] = 0.f;
      }
      const int cnt = min(32, N - base);
#pragma unroll 8
      for (int t = 0; t < cnt; ++t) {
        const float g = lane < emb ? dOb[static_cast<long long>(base + t) * emb + lane] : 0.f;
        dbacc += g;
#pragma unroll
        for (int j = 0; j < GCN_CG; ++j) dwacc[j] = fmaf(g, __shfl_sync(0xffffffffu, agg[j], t), dwacc[j]);
      }
    }
    __syncthreads();                                   // all gathers of phase A are done: the slab can be overwritten
    // ---- phase B ----
    if (dX != nullptr) {
      const bool vec = (emb % 4 == 0) && ((reinterpret_cast<uintptr_t>(dOb) & 15) == 0);
      if (vec) {
        transform_rows_coalesced(dOb, N, emb, sWt, sS);
      } else {
        for (int t = threadIdx.x; t < N; t += GCN_THREADS) {
          float h[GCN_CG];
#pragma unroll
          for (int j = 0; j < GCN_CG; ++j) h[j] = 0.f;
          const float* gr = dOb + static_cast<long long>(t) * emb;
          for (int cc = 0; cc < emb; ++cc) {
            const float g = gr[cc];
#pragma unroll
            for (int j = 0; j < GCN_CG; ++j) h[j] = fmaf(sWt[j * GCN_MAXC + cc], g, h[j]);
          }
          *reinterpret_cast<float4*>(sS + t * GCN_CG) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(sS + t * GCN_CG + 4) = make_float4(h[4], h[5], h[6], h[7]);
        }
      }
      __syncthreads();
      for (int u0 = warp * 32; u0 < N; u0 += GCN_THREADS) {
        const int u = u0 + lane;
        float acc[GCN_CG];
        if (u < N) {
          gather_smem<GCN_CG>(sS, rowptr_out, col_out, w_out, u, acc);
        } else {
#pragma unroll
          for (int j = 0; j < GCN_CG; ++j) acc[j] = 0.f;
        }
        store_group_rows(s_stage + warp * (32 * GCN_CG), acc, dX + static_cast<long long>(b) * N * Fin, Fin, u0, N, c0, Fin);
      }
    }
  }
  // one flush per CTA: registers -> shared -> global
  __syncthreads();
  if (lane < emb) {
#pragma unroll
    for (int j = 0; j < GCN_CG; ++j) atomicAdd(&s_dw[lane][j], dwacc[j]);
    if (c0 == 0) atomicAdd(&s_dw[lane][GCN_CG], dbacc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < emb * GCN_CG; i += GCN_THREADS) {
    const int cc = i / GCN_CG, j = i - cc * GCN_CG;
    if (c0 + j < Fin) atomicAdd(dW + cc * Fin + c0 + j, s_dw[cc][j]);
  }
  if (c0 == 0 && threadIdx.x < emb && dbias != nullptr) atomicAdd(dbias + threadIdx.x, s_dw[threadIdx.x][GCN_CG]);
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

// 0: L2-gather kernels (slab too large for shared memory), 1: narrow staged (F_in <= 4, no input gradient needed),
// 2: channel-group staged. FXN_GCN_VARIANT=0 forces the L2-gather kernels (A/B reference).
static int gcn_variant(int N, int Fin, bool narrow_ok) {
  static const int forced = [] { const char* e = getenv("FXN_GCN_VARIANT"); return e ? atoi(e) : -1; }();
  if (forced == 0) return 0;
  const size_t limit = 200 * 1024;
  if (narrow_ok && Fin <= 4 && sizeof(float) * 2 * N * Fin <= limit) return 1;
  if (sizeof(float) * N * GCN_CG <= limit) return 2;
  return 0;
}
static bool gcn_smem_attr(const void* fn, size_t smem) {
  if (smem <= 48 * 1024) return true;
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) == cudaSuccess;
}

// persistent grid of the channel-group kernels: as many CTAs as fit (<= 6 per SM), a multiple of the group count
static int gcn_wide_grid(int B, int groups, size_t smem_per_cta) {
  int per_sm = static_cast<int>(220 * 1024 / smem_per_cta);
  per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
  long long cap = 148LL * per_sm / groups * groups;
  if (cap < groups) cap = groups;
  const long long items = static_cast<long long>(B) * groups;
  return static_cast<int>(items < cap ? items : cap);
}

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
  const int variant = gcn_variant(N, Fin, true);
  if (variant == 1) {
    const size_t smem = sizeof(float) * 2 * N * Fin;
    const int grid = B < 148 * 4 ? B : 148 * 4;
#define FXN_NARROW_FWD(F)                                                                                              \
  do {                                                                                                                 \
    if (!gcn_smem_attr(reinterpret_cast<const void*>(gcn_fwd_narrow_kernel<F>), smem)) return set_error(FXN_ERR_CUDA, "gcn smem"); \
    gcn_fwd_narrow_kernel<F><<<grid, GCN_THREADS, smem, stream>>>(X, B, N, rowptr, col, w, W, bias, emb, O, partials);   \
  } while (0)
    switch (Fin) { case 1: FXN_NARROW_FWD(1); break; case 2: FXN_NARROW_FWD(2); break; case 3: FXN_NARROW_FWD(3); break;
                   default: FXN_NARROW_FWD(4); break; }
#undef FXN_NARROW_FWD
    FXN_CHECK_LAUNCH("gcn_fwd_narrow");
    return 0;
  }
  if (variant == 2) {
    const size_t smem = sizeof(float) * N * GCN_CG;
    if (!gcn_smem_attr(reinterpret_cast<const void*>(gcn_fwd_wide_kernel), smem)) return set_error(FXN_ERR_CUDA, "gcn smem");
    const int groups = (emb + GCN_CG - 1) / GCN_CG;
    gcn_fwd_wide_kernel<<<gcn_wide_grid(B, groups, smem + 2048), GCN_THREADS, smem, stream>>>(X, B, N, Fin, rowptr, col, w, W, bias,
                                                                                            emb, O, partials);
    FXN_CHECK_LAUNCH("gcn_fwd_wide");
    return 0;
  }
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
  const int variant = gcn_variant(N, Fin, dX == nullptr);
  if (variant == 1) {
    const size_t smem = sizeof(float) * 2 * N * Fin;
    const int grid = B < 148 * 4 ? B : 148 * 4;
#define FXN_NARROW_BWD(F)                                                                                              \
  do {                                                                                                                 \
    if (!gcn_smem_attr(reinterpret_cast<const void*>(gcn_bwd_narrow_kernel<F>), smem)) return set_error(FXN_ERR_CUDA, "gcn smem"); \
    gcn_bwd_narrow_kernel<F><<<grid, GCN_THREADS, smem, stream>>>(X, dO, B, N, emb, rowptr_in, col_in, w_in, dW, dbias); \
  } while (0)
    switch (Fin) { case 1: FXN_NARROW_BWD(1); break; case 2: FXN_NARROW_BWD(2); break; case 3: FXN_NARROW_BWD(3); break;
                   default: FXN_NARROW_BWD(4); break; }
#undef FXN_NARROW_BWD
    FXN_CHECK_LAUNCH("gcn_bwd_narrow");
    return 0;
  }
  if (variant != 0) {
    const size_t smem = sizeof(float) * N * GCN_CG;
    if (!gcn_smem_attr(reinterpret_cast<const void*>(gcn_bwd_wide_kernel), smem)) return set_error(FXN_ERR_CUDA, "gcn smem");
    const int groups = (Fin + GCN_CG - 1) / GCN_CG;
    gcn_bwd_wide_kernel<<<gcn_wide_grid(B, groups, smem + 4096), GCN_THREADS, smem, stream>>>(
        X, dO, B, N, Fin, emb, rowptr_in, col_in, w_in, rowptr_out, col_out, w_out, W, dW, dbias, dX);
    FXN_CHECK_LAUNCH("gcn_bwd_wide");
    return 0;
  }
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
