// Step policy on a flat parameter arena: gradient-norm clipping + Adam in two launches, and the refresh of the
// bf16 operand planes of every weight matrix in one launch.
//
// Replaces what Lightning does between training_step and the next batch as configured by the reference:
//   clip_grad_norm_(params, 1.0) ('norm' algorithm)     flexynesis/main.py:216-217
//   torch.optim.Adam(lr=config['lr']) defaults          flexynesis/models/direct_pred.py:135-144
// The engine keeps all trainable tensors of a model as views into ONE flat fp32 buffer (same for grads, m, v),
// so the whole update is a single grid-stride pass and the DDP all-reduce is a single contiguous message.
#include "fxn_internal.h"
#include "ptx.cuh"

namespace fxn {

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, long long n, float grad_scale,
                                                         double* __restrict__ sumsq, long long* __restrict__ step) {
  __shared__ double s_part[8];
  double acc = 0.0;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = g4[i];
    const float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    acc += static_cast<double>(s);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n - n4 * 4)) {
    const float v = g[n4 * 4 + threadIdx.x];
    acc += static_cast<double>(v) * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(sumsq, t * static_cast<double>(grad_scale) * grad_scale);
    if (blockIdx.x == 0 && step) *step += 1;
  }
}

__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n, float lr, float beta1, float beta2, float eps, float max_norm, float grad_scale,
                 double* __restrict__ sumsq, const long long* __restrict__ step, float* __restrict__ norm_out) {
  const float total_norm = static_cast<float>(sqrt(*reinterpret_cast<const volatile double*>(sumsq)));
  float coef = max_norm > 0.f ? max_norm / (total_norm + 1e-6f) : 1.f;
  coef = fminf(coef, 1.f) * grad_scale;
  const double t = static_cast<double>(*step);
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), t));
  const float bc2 = static_cast<float>(1.0 - pow(static_cast<double>(beta2), t));
  const float step_size = lr / bc1;
  const float bc2_sqrt = sqrtf(bc2);
  if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total_norm;
  // arenas are 16B aligned and n % 8 == 0 (every parameter slot is padded to 8 floats): float4 all the way
  const long long n4 = n / 4;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gi = ga[j] * coef;
      const float mi = ma[j] + (gi - ma[j]) * omb1;           // lerp_
      const float vi = va[j] * beta2 + omb2 * gi * gi;         // mul_ + addcmul_
      ma[j] = mi;
      va[j] = vi;
      const float denom = sqrtf(vi) / bc2_sqrt + eps;
      pa[j] = pa[j] - step_size * (mi / denom);
    }
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  // hand the norm scratch back ZEROED for the next step: every block read it on entry, the block that leaves last clears it
  // (replaces a memset node in front of grad_sumsq_kernel: ~5 us of serialisation per step in a captured graph)
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* count = reinterpret_cast<unsigned*>(sumsq + 1);
    __threadfence();
    if (atomicAdd(count, 1u) + 1u == gridDim.x) {
      *sumsq = 0.0;
      *count = 0u;
    }
  }
}

struct SplitSeg {
  long long src_off, rows, cols, ld_src, dst_off, ldp;
};

__global__ void __launch_bounds__(256)
split_multi_kernel(const float* __restrict__ src, const SplitSeg* __restrict__ segs, __nv_bfloat16* __restrict__ hi,
                   __nv_bfloat16* __restrict__ lo) {
  const SplitSeg s = segs[blockIdx.y];
  const long long chunks = (s.cols + 7) / 8;        // only this segment's own pad8(cols) columns (rows may be shared)
  const long long total = s.rows * chunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / chunks, c = (i - r * chunks) * 8;
    const float* q = src + s.src_off + r * s.ld_src + c;
    __align__(16) __nv_bfloat16 h[8];
    __align__(16) __nv_bfloat16 l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_bf16((c + j < s.cols) ? q[j] : 0.f, h[j], l[j]);
    *reinterpret_cast<uint4*>(hi + s.dst_off + r * s.ldp + c) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + s.dst_off + r * s.ldp + c) = *reinterpret_cast<const uint4*>(l);
  }
}

}  // namespace fxn

using namespace fxn;

extern "C" int fxn_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                                  float lr, float beta1, float beta2, float eps, float max_norm, float grad_scale,
                                  double* sumsq_scratch, long long* step_counter, float* norm_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!params || !grads || !exp_avg || !exp_avg_sq || !sumsq_scratch || !step_counter || n <= 0)
    return set_error(FXN_ERR_ARG, "fxn_clip_adam_step: bad argument");
  if ((reinterpret_cast<uintptr_t>(grads) & 15) || (reinterpret_cast<uintptr_t>(params) & 15) ||
      (reinterpret_cast<uintptr_t>(exp_avg) & 15) || (reinterpret_cast<uintptr_t>(exp_avg_sq) & 15) || (n % 4) != 0)
    return set_error(FXN_ERR_ARG, "fxn_clip_adam_step: arenas must be 16B aligned with n %% 4 == 0");
  if (reinterpret_cast<uintptr_t>(sumsq_scratch) & 7) return set_error(FXN_ERR_ARG, "fxn_clip_adam_step: scratch must be 8B aligned");
  int blocks = ceil_div(n, 256 * 4 * 2);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  const float gs = grad_scale == 0.f ? 1.f : grad_scale;
  grad_sumsq_kernel<<<blocks, 256, 0, stream>>>(grads, n, gs, sumsq_scratch, step_counter);
  FXN_CHECK_LAUNCH("grad_sumsq");
  clip_adam_kernel<<<blocks, 256, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, max_norm, gs,
                                              sumsq_scratch, step_counter, norm_out);
  FXN_CHECK_LAUNCH("clip_adam");
  return 0;
}

extern "C" int fxn_split_planes_multi(const float* src, const void* segments_dev, int nseg, long long max_seg_elems,
                                      void* hi, void* lo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!src || !segments_dev || !hi || !lo || nseg <= 0) return set_error(FXN_ERR_ARG, "fxn_split_planes_multi: bad argument");
  int bx = ceil_div(max_seg_elems / 8 + 1, 256);
  if (bx > 148 * 2) bx = 148 * 2;
  if (bx < 1) bx = 1;
  dim3 grid(bx, nseg);
  split_multi_kernel<<<grid, 256, 0, stream>>>(src, static_cast<const SplitSeg*>(segments_dev),
                                              static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo));
  FXN_CHECK_LAUNCH("split_planes_multi");
  return 0;
}
