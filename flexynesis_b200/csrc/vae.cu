// supervised_vae specific elementwise / reduction kernels: reparameterisation, squared row norms, the assembly of
// the MMD loss from the Gaussian-kernel column sums, and the MMD gradient combine. The O(B^2 L) work itself (Gram
// matrices, K*Z products) runs on the tcgen05 GEMM (gemm_umma.cu, epi_act 7) -- the reference's [B, B, L] broadcast
// tensors (flexynesis/models/supervised_vae.py:494-513, > 90 % of its step time) are never materialised.
#include "fxn_internal.h"
#include "ptx.cuh"
#include "rng.cuh"

namespace fxn {

__global__ void __launch_bounds__(256)
reparam_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ s, const float* __restrict__ eps,
                   long long ld, long long rows, int cols, float* __restrict__ z, __nv_bfloat16* __restrict__ hi,
                   __nv_bfloat16* __restrict__ lo, long long ldp) {
  const int chunks = (cols + 7) / 8;
  const long long total = rows * chunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / chunks;
    const int c = static_cast<int>(i - r * chunks) * 8;
    __align__(16) __nv_bfloat16 h[8];
    __align__(16) __nv_bfloat16 l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = 0.f;
      if (c + j < cols) {
        const long long o = r * ld + c + j;
        v = fmaf(s[o], eps[o], mean[o]);
        z[o] = v;
      }
      split_bf16(v, h[j], l[j]);
    }
    *reinterpret_cast<uint4*>(hi + r * ldp + c) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo + r * ldp + c) = *reinterpret_cast<const uint4*>(l);
  }
}

// block = 64 columns x 4 row lanes, rows strided by gridDim.y * 4
__global__ void __launch_bounds__(256)
reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ eps, long long ld, long long rows, int cols,
                   __nv_bfloat16* __restrict__ dm_hi, __nv_bfloat16* __restrict__ dm_lo,
                   __nv_bfloat16* __restrict__ ds_hi, __nv_bfloat16* __restrict__ ds_lo, long long ldp,
                   float* __restrict__ dbias_mean, float* __restrict__ dbias_s) {
  __shared__ float s_a[4][64], s_b[4][64];
  const int col = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + col;
  const int pcols = (cols + 7) & ~7;
  float sa = 0.f, sb = 0.f;
  if (c < pcols) {
    for (long long r = static_cast<long long>(blockIdx.y) * 4 + rl; r < rows; r += static_cast<long long>(gridDim.y) * 4) {
      float g = 0.f, ge = 0.f;
      if (c < cols) {
        g = dz[r * ld + c];
        ge = g * eps[r * ld + c];
      }
      __nv_bfloat16 h, l;
      split_bf16(g, h, l);
      dm_hi[r * ldp + c] = h; dm_lo[r * ldp + c] = l;
      split_bf16(ge, h, l);
      ds_hi[r * ldp + c] = h; ds_lo[r * ldp + c] = l;
      sa += g; sb += ge;
    }
  }
  s_a[rl][col] = sa; s_b[rl][col] = sb;
  __syncthreads();
  if (rl == 0 && c < cols) {
    atomicAdd(dbias_mean + c, s_a[0][col] + s_a[1][col] + s_a[2][col] + s_a[3][col]);
    atomicAdd(dbias_s + c, s_b[0][col] + s_b[1][col] + s_b[2][col] + s_b[3][col]);
  }
}

__global__ void __launch_bounds__(256)
row_sqnorm_kernel(const float* __restrict__ X, long long ld, long long rows, int cols, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long r = static_cast<long long>(blockIdx.x) * 8 + warp; r < rows; r += static_cast<long long>(gridDim.x) * 8) {
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float v = X[r * ld + c];
      s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[r] = s;
  }
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += s_red[w];
  return t;
}

__global__ void __launch_bounds__(1024)
mmd_finish_kernel(const float* __restrict__ cs_zz, const float* __restrict__ cs_tt, const float* __restrict__ cs_tz,
                  const float* __restrict__ mse_acc, const int* __restrict__ dims, int nlayers, int B, int P,
                  float* __restrict__ acc) {
  __shared__ double s_red[32];
  double zz = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) zz += cs_zz[i];
  zz = block_sum(zz, s_red) / (static_cast<double>(B) * B);
  double total = 0.0;
  for (int l = 0; l < nlayers; ++l) {
    double tt = 0.0, tz = 0.0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) tt += cs_tt[static_cast<long long>(l) * P + i];
    for (int i = threadIdx.x; i < B; i += blockDim.x) tz += cs_tz[static_cast<long long>(l) * B + i];
    tt = block_sum(tt, s_red) / (static_cast<double>(P) * P);
    tz = block_sum(tz, s_red) / (static_cast<double>(P) * B);
    total += tt + zz - 2.0 * tz + static_cast<double>(mse_acc[l]) / (static_cast<double>(B) * dims[l]);
  }
  if (threadIdx.x == 0) {
    acc[0] = static_cast<float>(total / nlayers);
    acc[1] = 1.f;
  }
}

__global__ void __launch_bounds__(256)
mmd_grad_kernel(const float* __restrict__ z, long long ldz, const float* __restrict__ cs_zz,
                const float* __restrict__ KZ, const float* __restrict__ cs_tz, const float* __restrict__ KT,
                long long ldk, int nlayers, int B, int L, int P, const float* __restrict__ weight,
                float* __restrict__ dz, long long ldd) {
  const float w = weight ? *weight : 1.f;
  const float l2 = static_cast<float>(L) * static_cast<float>(L);
  const float c1 = -4.f / (static_cast<float>(B) * static_cast<float>(B) * l2) * w;
  const float c2 = -4.f / (static_cast<float>(P) * static_cast<float>(B) * l2) * w / static_cast<float>(nlayers);
  const long long total = static_cast<long long>(B) * L;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int a = static_cast<int>(i / L), k = static_cast<int>(i - static_cast<long long>(a) * L);
    const float zv = z[a * ldz + k];
    float g = c1 * (zv * cs_zz[a] - KZ[a * ldk + k]);
    for (int l = 0; l < nlayers; ++l)
      g += c2 * (KT[(static_cast<long long>(l) * B + a) * ldk + k] - zv * cs_tz[static_cast<long long>(l) * B + a]);
    dz[a * ldd + k] += g;
  }
}

// out[r, c] ~ N(0, 1) for c < cols (pad columns up to ld are left untouched); 4 values per Philox call
__global__ void __launch_bounds__(256)
randn_kernel(float* __restrict__ out, long long ld, long long rows, int cols, unsigned long long seed,
             const long long* __restrict__ seed_dev) {
  const unsigned long long s = step_seed(seed, seed_dev);
  const int quads = (cols + 3) / 4;
  const long long total = rows * quads;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / quads;
    const int c = static_cast<int>(i - r * quads) * 4;
    const uint4 u = philox4x32(static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), static_cast<uint32_t>(s),
                               static_cast<uint32_t>(s >> 32));
    const float2 a = box_muller(u.x, u.y), b = box_muller(u.z, u.w);
    float* o = out + r * ld + c;
    o[0] = a.x;
    if (c + 1 < cols) o[1] = a.y;
    if (c + 2 < cols) o[2] = b.x;
    if (c + 3 < cols) o[3] = b.y;
  }
}

__global__ void loss_weights_kernel(int n, const float* const* __restrict__ log_vars, int weighting, float* __restrict__ wts) {
  const int k = threadIdx.x;
  if (k < n) wts[k] = (weighting && n > 1) ? expf(-*log_vars[k]) : 1.f;
}

}  // namespace fxn

using namespace fxn;

extern "C" int fxn_reparam_fwd(const float* mean, const float* s, const float* eps, long long ld, long long rows,
                               int cols, float* z, void* z_hi, void* z_lo, long long ldp, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!mean || !s || !eps || !z || !z_hi || !z_lo) return set_error(FXN_ERR_ARG, "fxn_reparam_fwd: null argument");
  if (ldp % 8 != 0) return set_error(FXN_ERR_ARG, "fxn_reparam_fwd: bad planes");
  int blocks = ceil_div(rows * ((cols + 7) / 8), 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  reparam_fwd_kernel<<<blocks, 256, 0, stream>>>(mean, s, eps, ld, rows, cols, z, static_cast<__nv_bfloat16*>(z_hi),
                                                 static_cast<__nv_bfloat16*>(z_lo), ldp);
  FXN_CHECK_LAUNCH("reparam_fwd");
  return 0;
}

extern "C" int fxn_reparam_bwd(const float* dz, const float* eps, long long ld, long long rows, int cols, void* dm_hi,
                               void* dm_lo, void* ds_hi, void* ds_lo, long long ldp, float* dbias_mean, float* dbias_s,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!dz || !eps || !dm_hi || !dm_lo || !ds_hi || !ds_lo || !dbias_mean || !dbias_s)
    return set_error(FXN_ERR_ARG, "fxn_reparam_bwd: null argument");
  cudaError_t e = cudaMemsetAsync(dbias_mean, 0, sizeof(float) * cols, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbias_s, 0, sizeof(float) * cols, stream);
  if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "reparam_bwd memset: %s", cudaGetErrorString(e));
  int by = ceil_div(rows, 4 * 8);
  if (by > 148 * 2) by = 148 * 2;
  dim3 grid(ceil_div(((cols + 7) & ~7), 64), by);
  reparam_bwd_kernel<<<grid, 256, 0, stream>>>(dz, eps, ld, rows, cols, static_cast<__nv_bfloat16*>(dm_hi),
                                               static_cast<__nv_bfloat16*>(dm_lo), static_cast<__nv_bfloat16*>(ds_hi),
                                               static_cast<__nv_bfloat16*>(ds_lo), ldp, dbias_mean, dbias_s);
  FXN_CHECK_LAUNCH("reparam_bwd");
  return 0;
}

extern "C" int fxn_row_sqnorm(const float* X, long long ld, long long rows, int cols, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!X || !out || rows <= 0 || cols <= 0) return set_error(FXN_ERR_ARG, "fxn_row_sqnorm: bad argument");
  int blocks = ceil_div(rows, 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  row_sqnorm_kernel<<<blocks, 256, 0, stream>>>(X, ld, rows, cols, out);
  FXN_CHECK_LAUNCH("row_sqnorm");
  return 0;
}

extern "C" int fxn_mmd_finish(const float* cs_zz, const float* cs_tt, const float* cs_tz, const float* mse_acc,
                              const int* dims, int nlayers, int B, int P, float* acc, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!cs_zz || !cs_tt || !cs_tz || !mse_acc || !dims || !acc || nlayers <= 0)
    return set_error(FXN_ERR_ARG, "fxn_mmd_finish: bad argument");
  mmd_finish_kernel<<<1, 1024, 0, stream>>>(cs_zz, cs_tt, cs_tz, mse_acc, dims, nlayers, B, P, acc);
  FXN_CHECK_LAUNCH("mmd_finish");
  return 0;
}

extern "C" int fxn_mmd_grad(const float* z, long long ldz, const float* cs_zz, const float* KZ, const float* cs_tz,
                            const float* KT, long long ldk, int nlayers, int B, int L, int P, const float* weight,
                            float* dz, long long ldd, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!z || !cs_zz || !KZ || !cs_tz || !KT || !dz) return set_error(FXN_ERR_ARG, "fxn_mmd_grad: null argument");
  int blocks = ceil_div(static_cast<long long>(B) * L, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  mmd_grad_kernel<<<blocks, 256, 0, stream>>>(z, ldz, cs_zz, KZ, cs_tz, KT, ldk, nlayers, B, L, P, weight, dz, ldd);
  FXN_CHECK_LAUNCH("mmd_grad");
  return 0;
}

extern "C" int fxn_loss_weights(int n, const float* const* log_vars, int weighting, float* wts, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0 || n > 1024 || !wts) return set_error(FXN_ERR_ARG, "fxn_loss_weights: bad argument");
  if (weighting && n > 1 && !log_vars) return set_error(FXN_ERR_ARG, "fxn_loss_weights: weighting needs log_vars");
  loss_weights_kernel<<<1, 1024, 0, stream>>>(n, log_vars, weighting, wts);
  FXN_CHECK_LAUNCH("loss_weights");
  return 0;
}

extern "C" int fxn_randn(float* out, long long ld, long long rows, int cols, unsigned long long seed,
                         const void* seed_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!out || rows <= 0 || cols <= 0 || ld < cols) return set_error(FXN_ERR_ARG, "fxn_randn: bad argument");
  int blocks = ceil_div(rows * ((cols + 3) / 4), 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  randn_kernel<<<blocks, 256, 0, stream>>>(out, ld, rows, cols, seed, static_cast<const long long*>(seed_dev));
  FXN_CHECK_LAUNCH("randn");
  return 0;
}
