// Library plumbing: version, thread-local error string, launch counter, fp32 -> bf16 hi/lo planes.
#include "fxn_internal.h"
#include "ptx.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace fxn {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// 8 columns per thread: two float4 loads when the source row is 16B aligned, one 16B store per plane.
__global__ void split_planes_kernel(const float* __restrict__ src, long long ld_src, long long rows, long long cols,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ldp,
                                    int vec_src) {
  const long long chunks_per_row = (cols + 7) / 8;   // zero-fill to pad8(cols) only: dst may be a column window
  const long long total = rows * chunks_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / chunks_per_row;
    const long long c = (i - r * chunks_per_row) * 8;
    float x[8];
    const float* s = src + r * ld_src + c;
    if (vec_src && c + 8 <= cols) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(s));
      const float4 b = __ldg(reinterpret_cast<const float4*>(s) + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
      x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (c + j < cols) ? __ldg(s + j) : 0.f;
    }
    __align__(16) __nv_bfloat16 h[8];
    __align__(16) __nv_bfloat16 l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_bf16(x[j], h[j], l[j]);
    *reinterpret_cast<uint4*>(hi + r * ldp + c) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + r * ldp + c) = *reinterpret_cast<const uint4*>(l);
  }
}

// Batch feeder: planes[b, :] = split(src[idx[b], :]) -- a device-side permutation gather over the HBM-resident
// dataset (replaces per-sample __getitem__ + default_collate, flexynesis/data.py:980-995), fused with the split.
__global__ void gather_rows_kernel(const float* __restrict__ src, long long ld_src, const long long* __restrict__ idx,
                                   long long nrows, long long cols, float* __restrict__ out, long long ldo,
                                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ldp,
                                   int vec_src) {
  const long long chunks_per_row = (cols + 7) / 8;
  const long long total = nrows * chunks_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / chunks_per_row;
    const long long c = (i - r * chunks_per_row) * 8;
    const long long sr = idx ? idx[r] : r;
    const float* s = src + sr * ld_src + c;
    float x[8];
    if (vec_src && c + 8 <= cols) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(s));
      const float4 b = __ldg(reinterpret_cast<const float4*>(s) + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
      x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (c + j < cols) ? __ldg(s + j) : 0.f;
    }
    if (out) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c + j < cols) out[r * ldo + c + j] = x[j];
    }
    if (hi) {
      __align__(16) __nv_bfloat16 h[8];
      __align__(16) __nv_bfloat16 l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_bf16(x[j], h[j], l[j]);
      *reinterpret_cast<uint4*>(hi + r * ldp + c) = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(lo + r * ldp + c) = *reinterpret_cast<const uint4*>(l);
    }
  }
}

}  // namespace fxn

using namespace fxn;

extern "C" int fxn_gather_rows(const float* src, long long ld_src, const long long* idx, long long nrows,
                               long long cols, float* out, long long ldo, void* hi, void* lo, long long ldp,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!src || (!out && !hi)) return set_error(FXN_ERR_ARG, "fxn_gather_rows: null pointer");
  if (nrows <= 0 || cols <= 0) return 0;
  if (hi && (!lo || ldp % 8 != 0 || ldp < cols)) return set_error(FXN_ERR_ARG, "fxn_gather_rows: bad planes");
  const int vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && ld_src % 4 == 0) ? 1 : 0;
  const long long total = nrows * ((cols + 7) / 8);
  int blocks = ceil_div(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_rows_kernel<<<blocks, 256, 0, stream>>>(src, ld_src, idx, nrows, cols, out, ldo,
                                                 static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), ldp, vec);
  FXN_CHECK_LAUNCH("gather_rows");
  return 0;
}

extern "C" int fxn_version(void) { return 100; }
extern "C" const char* fxn_last_error(void) { return g_err; }
extern "C" long long fxn_launch_count(void) { return g_launches.load(); }
extern "C" void fxn_reset_launch_count(void) { g_launches.store(0); }

extern "C" int fxn_split_planes(const float* src, long long ld_src, long long rows, long long cols, void* hi,
                                void* lo, long long ldp, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!src || !hi || !lo) return set_error(FXN_ERR_ARG, "fxn_split_planes: null pointer");
  if (rows <= 0 || cols <= 0) return 0;
  if (ldp % 8 != 0 || ldp < cols) return set_error(FXN_ERR_ARG, "fxn_split_planes: ld_planes must be >= cols and %% 8 == 0");
  if ((reinterpret_cast<uintptr_t>(hi) & 15) || (reinterpret_cast<uintptr_t>(lo) & 15))
    return set_error(FXN_ERR_ARG, "fxn_split_planes: planes must be 16B aligned");
  const int vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && ld_src % 4 == 0) ? 1 : 0;
  const long long total = rows * ((cols + 7) / 8);
  int blocks = ceil_div(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  split_planes_kernel<<<blocks, 256, 0, stream>>>(src, ld_src, rows, cols, static_cast<__nv_bfloat16*>(hi),
                                                  static_cast<__nv_bfloat16*>(lo), ldp, vec);
  FXN_CHECK_LAUNCH("split_planes");
  return 0;
}
