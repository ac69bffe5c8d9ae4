// Neighbour aggregation of flexGCN's wide layers as a PURE gather, so that the per-node linear map can run on the tensor
// cores (fxn_gemm over [B * N x C] rows) instead of on CUDA cores inside the aggregate kernel.
//
//   out[b, v, :] = sum over the CSR row of node v of w_e * in[b, col_e, :]            (C channels, C % 16 == 0)
//
// Forward of a GCNConv layer (torch_geometric, flexynesis/modules.py:221-226, :254):
//   G = A^ X (this kernel, CSR by destination, output as bf16 operand planes) ; O = G W^T + bias (fxn_gemm, BatchNorm
//   partials in its epilogue).
// Backward: T = dO W (fxn_gemm) ; dX = A^T T (this kernel over the CSR by source, fp32 output) ; dW = dO^T G (fxn_gemm,
//   stream-K: G is kept from the forward pass, nothing is re-aggregated).
// The fused CUDA-core kernels this replaces (gcn_fwd_wide / gcn_bwd_wide, csrc/gcn.cu) ran at 10-12 % of the HBM roofline:
// their pace was set by the 32 x 8 per-node transforms (about half of their instructions) and shared-memory gathers of
// 8-channel rows (profiles/r01_timeline_v7_cfg4.log). Here a CTA owns (sample, group of 16 channels): the [N x 16] slab is
// staged once in shared memory (64 bytes per node, XOR-swizzled 16-byte slots so that random rows spread over the banks),
// lane = node, 4 x LDS.128 + 16 FMA per edge. HBM sees every activation once per pass: read C * 4 B, write C * 4 B per node.
#include "fxn_internal.h"
#include "ptx.cuh"
#include <cstdlib>

namespace fxn {

constexpr int GG_C = 16;                 // channel granularity the entry point requires (C % 16 == 0)

// float offset of 16-byte slot s of node u in a slab with GC channels per node: the slots of a 128-byte line are
// XOR-permuted with the line index, so lanes that read the same slot of random rows spread over all 32 banks
template <int GC>
__device__ __forceinline__ int gg_slot(int u, int s) {
  if constexpr (GC == 16) return u * 16 + 4 * (s ^ ((u >> 1) & 3));
  else return u * 8 + 4 * (s ^ ((u >> 2) & 1));
}

template <bool PLANES, int GC, int GG_THREADS>
__global__ void __launch_bounds__(GG_THREADS)
graph_gather_kernel(const float* __restrict__ in, int B, int N, int C, const int* __restrict__ rowptr,
                    const int* __restrict__ col, const float* __restrict__ w, float* __restrict__ out,
                    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  extern __shared__ __align__(16) float gg_slab[];          // [N][GC], swizzled
  constexpr int SL = GC / 4;                                // 16-byte slots per node
  const int groups = C / GC;
  const int g = static_cast<int>(blockIdx.x % groups);      // a CTA keeps its channel group
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = blockIdx.x / groups; b < B; b += gridDim.x / groups) {
    const float* src = in + (static_cast<long long>(b) * N) * C + g * GC;
    __syncthreads();                                        // previous sample's gathers are done with the slab
    // stage: 4 threads per node (one 16-byte slot each), 8 independent loads in flight per thread
    for (int base = threadIdx.x; base < N * SL; base += GG_THREADS * 8) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + k * GG_THREADS;
        v[k] = i < N * SL ? __ldg(reinterpret_cast<const float4*>(src + static_cast<long long>(i / SL) * C) + (i % SL))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + k * GG_THREADS;
        if (i < N * SL) *reinterpret_cast<float4*>(gg_slab + gg_slot<GC>(i / SL, i % SL)) = v[k];
      }
    }
    __syncthreads();
    for (int v0 = warp * 32; v0 < N; v0 += GG_THREADS) {    // a warp owns 32 consecutive nodes, lane = node
      const int v = v0 + lane;
      float acc[GC];
#pragma unroll
      for (int j = 0; j < GC; ++j) acc[j] = 0.f;
      if (v < N) {
        const int e0 = __ldg(rowptr + v), e1 = __ldg(rowptr + v + 1);
        for (int e = e0; e < e1; e += 8) {                  // edge records come from L2: eight at a time
          int u[8];
          float we[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const bool ok = e + t < e1;
            u[t] = ok ? __ldg(col + e + t) : 0;
            we[t] = ok ? __ldg(w + e + t) : 0.f;
          }
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            if (e + t < e1) {
#pragma unroll
              for (int s = 0; s < SL; ++s) {
                const float4 x = *reinterpret_cast<const float4*>(gg_slab + gg_slot<GC>(u[t], s));
                acc[4 * s] = fmaf(we[t], x.x, acc[4 * s]);
                acc[4 * s + 1] = fmaf(we[t], x.y, acc[4 * s + 1]);
                acc[4 * s + 2] = fmaf(we[t], x.z, acc[4 * s + 2]);
                acc[4 * s + 3] = fmaf(we[t], x.w, acc[4 * s + 3]);
              }
            }
          }
        }
        const long long row = static_cast<long long>(b) * N + v;
        if constexpr (PLANES) {
          __align__(16) __nv_bfloat16 h[GC];
          __align__(16) __nv_bfloat16 l[GC];
#pragma unroll
          for (int j = 0; j < GC; ++j) split_bf16(acc[j], h[j], l[j]);
          uint4* ph = reinterpret_cast<uint4*>(out_hi + row * C + g * GC);
          uint4* pl = reinterpret_cast<uint4*>(out_lo + row * C + g * GC);
#pragma unroll
          for (int q = 0; q < GC / 8; ++q) {
            ph[q] = reinterpret_cast<const uint4*>(h)[q];
            pl[q] = reinterpret_cast<const uint4*>(l)[q];
          }
        } else {
          float4* po = reinterpret_cast<float4*>(out + row * C + g * GC);
#pragma unroll
          for (int s = 0; s < SL; ++s) po[s] = make_float4(acc[4 * s], acc[4 * s + 1], acc[4 * s + 2], acc[4 * s + 3]);
        }
      }
    }
  }
}

// ---- variant without a staged slab: the sample's [N x C] rows are gathered through L1 / L2 ----
// A group of LPN = C / 8 lanes owns a node (eight channels = two float4 per lane), a warp advances 32 / LPN nodes at once and
// a CTA walks the nodes of ONE sample, so that the rows it gathers (N * C * 4 bytes, 256 KB at N = 2000, C = 32) are served by
// its SM's L1 or by L2 for their ~deg uses and by HBM once. Edge records are read LPN at a time (lane q of the group takes
// edge e + q, coalesced) and handed round by shuffles; the 2 * LPN row loads of a batch are independent. No shared memory,
// no barriers. The first version of this kernel (one float4 per lane, predicated loads, 64-bit index arithmetic) was bound by
// instruction ISSUE, not by memory (ncu: 1.2 G warp instructions, issue slots 77 % busy at 1.5 ms): everything per edge
// except the loads and FMAs is overhead, so a lane now covers eight channels, the padded slots of a batch re-read the node's
// last edge with weight 0 instead of branching, offsets are 32-bit, and `order` (nodes sorted by degree) makes the 32 / LPN
// nodes a warp advances together finish together.
struct F8 { float v[8]; };
// one 256-bit load (sm_100: LDG.E.256): a node's 128-byte row is covered by four lanes of ONE instruction, so L1 looks the line
// up once per edge (two float4 loads per lane touched every sector twice: l1tex 86 % busy, profiles/r02_gather_rows_v2.txt)
__device__ __forceinline__ F8 ldg_f8(const float* p) {
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}

__device__ __forceinline__ int ldg_stream_s32(const int* p) {       // edge records: read once per pass, kept out of L1
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// C = row stride (all channels); one pass covers the LPN * 8 channels from channel (item % cgroups) * LPN * 8 of one sample
template <bool PLANES, int LPN, bool HINTS>
__global__ void __launch_bounds__(1024, 1)
graph_gather_rows_kernel(const float* __restrict__ in, int B, int N, int C, const int* __restrict__ rowptr,
                         const int* __restrict__ col, const float* __restrict__ w, const int* __restrict__ order,
                         float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  constexpr int NPW = 32 / LPN;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / LPN, q = lane % LPN;
  const int ngroups = (N + NPW - 1) / NPW;
  const int cgroups = C / (LPN * 8);
  for (int item = blockIdx.x; item < B * cgroups; item += gridDim.x) {
    const int b = item / cgroups, ch0 = (item % cgroups) * LPN * 8 + q * 8;
    const float* base = in + static_cast<long long>(b) * N * C + ch0;
    for (int grp = warp; grp < ngroups; grp += nwarps) {
      const int slot = grp * NPW + sub;
      const bool valid = slot < N;
      const int v = valid ? (order ? __ldg(order + slot) : slot) : 0;
      const int e0 = valid ? __ldg(rowptr + v) : 0, e1 = valid ? __ldg(rowptr + v + 1) : 0;
      const int elast = e1 - 1;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int e = e0; __any_sync(0xffffffffu, e < e1); e += LPN) {
        // lane q fetches edge e + q; past the end of the row it re-reads the row's last edge (own row if the node has no
        // edge at all) with weight 0, so the loads below need no predicate
        const int idx = min(e + q, elast);
        int off;
        float wu;
        if constexpr (HINTS) {
          off = (e1 > e0 ? ldg_stream_s32(col + idx) : v) * C;
          wu = e + q < e1 ? __int_as_float(ldg_stream_s32(reinterpret_cast<const int*>(w) + idx)) : 0.f;
        } else {
          off = (e1 > e0 ? __ldg(col + idx) : v) * C;
          wu = e + q < e1 ? __ldg(w + idx) : 0.f;
        }
        constexpr int TB = LPN < 4 ? LPN : 4;                // row loads kept in flight per lane: 2 * TB float4
#pragma unroll
        for (int t0 = 0; t0 < LPN; t0 += TB) {
          F8 x[TB];
          float wt[TB];
#pragma unroll
          for (int t = 0; t < TB; ++t) {
            const float* rp = base + __shfl_sync(0xffffffffu, off, t0 + t, LPN);
            wt[t] = __shfl_sync(0xffffffffu, wu, t0 + t, LPN);
            x[t] = ldg_f8(rp);
          }
#pragma unroll
          for (int t = 0; t < TB; ++t) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt[t], x[t].v[j], acc[j]);
          }
        }
      }
      if (valid) {
        const long long o = (static_cast<long long>(b) * N + v) * C + ch0;
        if constexpr (PLANES) {
          __align__(16) __nv_bfloat16 h[8];
          __align__(16) __nv_bfloat16 l[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_bf16(acc[j], h[j], l[j]);
          *reinterpret_cast<uint4*>(out_hi + o) = *reinterpret_cast<const uint4*>(h);
          *reinterpret_cast<uint4*>(out_lo + o) = *reinterpret_cast<const uint4*>(l);
        } else {
          *reinterpret_cast<float4*>(out + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
          *reinterpret_cast<float4*>(out + o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
      }
    }
  }
}

// ---- Chan merge of MANY row-tile partials (one per 128 rows of a [B * N x C] GEMM output) into one (sum, M2) record ----
// pass 1: column sums (double atomics) ; pass 2: sum over tiles of M2_t + n_t (mean_t - mean)^2.
__global__ void __launch_bounds__(256)
merge_big_kernel(const float* __restrict__ partials, int ntiles, int tile_rows, long long rows, int cols, int pld, int fold,
                 double* __restrict__ scratch, int pass) {
  // fold > 1: the producer viewed the [rows * fold x cols] matrix as [rows x fold * cols] (fold consecutive rows per GEMM
  // row): folded column q * cols + c carries channel c, so every tile contributes `fold` records per channel
  // thread = (tile lane 0..7, column 0..31) x column blocks along blockIdx.y; blockIdx.x strides over chunks of 64 tiles
  const int c = blockIdx.y * 32 + (threadIdx.x & 31);
  const int tl = threadIdx.x >> 5;
  __shared__ double s_red[8][32];
  double acc = 0.0;
  const double mean = pass == 2 && c < cols ? scratch[c] / (static_cast<double>(rows) * fold) : 0.0;
  if (c < cols) {
   for (int q = 0; q < fold; ++q)
    for (int t0 = blockIdx.x * 64 + tl; t0 < ntiles; t0 += gridDim.x * 64) {
      float sv[8], mv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int t = t0 + 8 * k;
        const bool ok = t < ntiles;
        sv[k] = ok ? partials[(static_cast<long long>(t) * 2) * pld + q * cols + c] : 0.f;
        mv[k] = (ok && pass == 2) ? partials[(static_cast<long long>(t) * 2 + 1) * pld + q * cols + c] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int t = t0 + 8 * k;
        if (t >= ntiles) continue;
        if (pass == 1) {
          acc += static_cast<double>(sv[k]);
        } else {
          const long long r0 = static_cast<long long>(t) * tile_rows;
          const double n = static_cast<double>(min(static_cast<long long>(tile_rows), rows - r0));
          const double d = static_cast<double>(sv[k]) / n - mean;
          acc += static_cast<double>(mv[k]) + n * d * d;
        }
      }
    }
  }
  s_red[tl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (tl == 0 && c < cols) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s_red[k][threadIdx.x & 31];
    atomicAdd(scratch + (pass == 1 ? 0 : cols) + c, t);
  }
}
__global__ void merge_big_finish_kernel(const double* __restrict__ scratch, int cols, float* __restrict__ merged) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) { merged[c] = static_cast<float>(scratch[c]); merged[cols + c] = static_cast<float>(scratch[cols + c]); }
}

}  // namespace fxn

using namespace fxn;

extern "C" int fxn_graph_gather_ok(int N, int C) {
  return C > 0 && C % GG_C == 0 && static_cast<long long>(N) * GG_C * 4 <= 200 * 1024;
}

extern "C" int fxn_graph_gather(const float* in, int B, int N, int C, const int* rowptr, const int* col, const float* w,
                                const int* order, float* out, void* out_hi, void* out_lo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!in || !rowptr || !col || !w || (!out && (!out_hi || !out_lo))) return set_error(FXN_ERR_ARG, "fxn_graph_gather: null argument");
  if (B <= 0 || N <= 0 || !fxn_graph_gather_ok(N, C))
    return set_error(FXN_ERR_UNSUPPORTED, "fxn_graph_gather: needs C %% 16 == 0 and a [N x 16] fp32 slab within shared memory");
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (out && (reinterpret_cast<uintptr_t>(out) & 15)) ||
      (out_hi && ((reinterpret_cast<uintptr_t>(out_hi) | reinterpret_cast<uintptr_t>(out_lo)) & 15)))
    return set_error(FXN_ERR_ARG, "fxn_graph_gather: buffers must be 16-byte aligned");
  // variant (FXN_GG_MODE): 0 = 16-channel groups, 256 threads; 1 = 16-channel groups, 512 threads; 2 = 8-channel groups, 256
  // threads (64 KB slab at N = 2000: three CTAs per SM overlap each other's staging and gather phases)
  static const int mode = [] { const char* e = getenv("FXN_GG_MODE"); return e ? atoi(e) : 3; }();
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(out_hi);
  __nv_bfloat16* lo = static_cast<__nv_bfloat16*>(out_lo);
  if (mode == 3 && C % 16 == 0 && static_cast<long long>(N) * C < (1LL << 31)) {
    // default: rows gathered through L1 / L2, no slab
    static const int th = [] { const char* e = getenv("FXN_GG_THREADS"); return e ? atoi(e) : 1024; }();
    // one CTA (= one sample in flight) per SM: with two, the 2 x 148 x 256 KB of rows being gathered plus the output stream
    // no longer stay in L2 and HBM reads triple (ncu: 3.6 GB instead of 1.1 GB, profiles/r02_gather_rows_v3.txt)
    static const int per = [] { const char* e = getenv("FXN_GG_PER_SM"); return e ? atoi(e) : 1; }();
    static const int lpn_env = [] { const char* e = getenv("FXN_GG_LPN"); return e ? atoi(e) : 0; }();
    static const int hints = [] { const char* e = getenv("FXN_GG_HINTS"); return e ? atoi(e) : 0; }();
    int lpn = lpn_env ? lpn_env : (C % 32 == 0 ? 4 : 2);             // channels per pass = 8 * lpn
    if (C % (8 * lpn)) lpn = 2;
    const long long items_r = static_cast<long long>(B) * (C / (8 * lpn));
    const int grid_r = static_cast<int>(items_r < 148LL * per ? items_r : 148LL * per);
#define FXN_GGR(PL, LPN, H)                                                                                                \
  graph_gather_rows_kernel<PL, LPN, H><<<grid_r, th, 0, stream>>>(in, B, N, C, rowptr, col, w, order, out, hi, lo)
#define FXN_GGR2(PL, LPN) do { if (hints) FXN_GGR(PL, LPN, true); else FXN_GGR(PL, LPN, false); } while (0)
    if (lpn != 2 && lpn != 4) return set_error(FXN_ERR_ARG, "FXN_GG_LPN must be 2 or 4");
    if (out_hi) { if (lpn == 2) FXN_GGR2(true, 2); else FXN_GGR2(true, 4); }
    else { if (lpn == 2) FXN_GGR2(false, 2); else FXN_GGR2(false, 4); }
#undef FXN_GGR2
#undef FXN_GGR
    FXN_CHECK_LAUNCH("graph_gather_rows");
    return 0;
  }
  const int gc = mode == 2 ? 8 : 16, threads = mode == 1 ? 512 : 256;
  const size_t smem = static_cast<size_t>(N) * gc * sizeof(float);
  const int groups = C / gc;
  int per_sm = static_cast<int>(220 * 1024 / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
  if (threads == 512 && per_sm > 2) per_sm = 2;
  long long cap = 148LL * per_sm / groups * groups;
  if (cap < groups) cap = groups;
  const long long items = static_cast<long long>(B) * groups;
  const int grid = static_cast<int>(items < cap ? items : cap);
#define FXN_GG_LAUNCH(PL, GCV, TH)                                                                                         \
  do {                                                                                                                     \
    static size_t attr = 48 * 1024;                                                                                        \
    if (smem > attr) {                                                                                                     \
      cudaError_t e = cudaFuncSetAttribute(graph_gather_kernel<PL, GCV, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           static_cast<int>(smem));                                                        \
      if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "graph_gather attr: %s", cudaGetErrorString(e));                \
      attr = smem;                                                                                                         \
    }                                                                                                                      \
    graph_gather_kernel<PL, GCV, TH><<<grid, TH, smem, stream>>>(in, B, N, C, rowptr, col, w, out, hi, lo);                \
  } while (0)
  if (out_hi) {
    if (mode == 2) FXN_GG_LAUNCH(true, 8, 256); else if (mode == 1) FXN_GG_LAUNCH(true, 16, 512); else FXN_GG_LAUNCH(true, 16, 256);
  } else {
    if (mode == 2) FXN_GG_LAUNCH(false, 8, 256); else if (mode == 1) FXN_GG_LAUNCH(false, 16, 512); else FXN_GG_LAUNCH(false, 16, 256);
  }
#undef FXN_GG_LAUNCH
  FXN_CHECK_LAUNCH("graph_gather");
  return 0;
}

extern "C" int fxn_merge_col_stats_big(const float* partials, int ntiles, int tile_rows, long long rows, int cols, int pld,
                                       int fold, float* merged, double* scratch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!partials || !merged || !scratch || ntiles <= 0 || tile_rows <= 0 || rows <= 0 || cols <= 0 || fold < 1)
    return set_error(FXN_ERR_ARG, "fxn_merge_col_stats_big: bad argument");
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * cols, stream);
  if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "merge_big memset: %s", cudaGetErrorString(e));
  int bx = (ntiles + 63) / 64;
  if (bx > 148 * 4) bx = 148 * 4;
  dim3 grid(bx, (cols + 31) / 32);
  const int ld = pld > 0 ? pld : cols * fold;
  merge_big_kernel<<<grid, 256, 0, stream>>>(partials, ntiles, tile_rows, rows, cols, ld, fold, scratch, 1);
  FXN_CHECK_LAUNCH("merge_big_1");
  merge_big_kernel<<<grid, 256, 0, stream>>>(partials, ntiles, tile_rows, rows, cols, ld, fold, scratch, 2);
  FXN_CHECK_LAUNCH("merge_big_2");
  merge_big_finish_kernel<<<(cols + 63) / 64, 64, 0, stream>>>(scratch, cols, merged);
  FXN_CHECK_LAUNCH("merge_big_finish");
  return 0;
}
