// Data-parallel step fused with the NVSwitch collectives (NVLink SHARP / "NVLS" multicast objects).
//
// The reference has no multi-GPU path (pl.Trainer(devices=1), flexynesis/main.py:223). The plain port of DDP is
// "all-reduce the gradient arena with NCCL, then run clip + Adam on every rank" -- W ranks each read and write the whole
// model. Here the parameter and gradient arenas of every rank are mapped behind ONE multicast address (torch symmetric
// memory), the arena is cut into W slices, and rank r
//   1. pulls slice r of the SUMMED gradient with multimem.ld_reduce (the switch adds the W copies in flight: one load
//      instruction replaces a reduce-scatter), scales it by 1/W, keeps it, and accumulates its squared norm;
//   2. publishes that partial norm into slot r of every rank with one multimem.st;
//   -- system-wide barrier --
//   3. clips with the global norm, runs Adam on slice r only (moments exist for the slice only), and stores the new
//      parameters with multimem.st, which lands in every rank's parameter arena (one store replaces an all-gather).
// Per step and GPU: P/W elements in through the switch, P/W out, instead of 2 P (W-1)/W for a ring all-reduce, and the
// optimizer touches P/W instead of P elements.
#include "fxn_internal.h"
#include "ptx.cuh"

namespace fxn {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc_addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc_addr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_st_scalar(float* mc_addr, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc_addr), "f"(v) : "memory");
}

// ---- barrier folded into the kernel that needs it ----
// The stand-alone barrier kernel (below) costs a launch, a drain and a second launch around ~1 us of actual signalling; at
// two ranks the three of them were 32 us of a 94 us optimizer tail (profiles/r02_dp_timeline_cfg2_n2_v0.log). Folded in:
// thread 0 of block 0 signals every rank (stream order makes this rank's earlier kernels complete), thread 0 of EVERY block
// waits for the local copy of the flag to show `world` arrivals for the next epoch, and the block that finishes last bumps
// the epoch (all blocks have read it by then). `sync` = {mc_flags, local_flags, epoch, done_counter}; mc_flags == nullptr
// skips it (host-side barriers).
struct DpSync {
  unsigned* mc_flags;
  const unsigned* local_flags;
  unsigned* epoch;
  unsigned* done;
  unsigned* go;
  int slot, world;
};
struct DpFused {
  unsigned lo[8], hi[8];      // element ranges of the arena whose reduce-scatter happened inside the weight-gradient GEMMs
  int n;
  float* inbox;               // this rank's inbox arena (peers' contributions to the ranges above)
};
struct DpPeers {
  const float* p[8];          // every rank's gradient arena (peer-mapped), in rank order
  int n;                      // 0: use the multicast address (in-switch reduction)
};
__device__ __forceinline__ void dp_sync_enter(const DpSync& sy) {
  if (sy.mc_flags == nullptr) return;
  if (threadIdx.x == 0) {
    const unsigned next = sy.epoch[sy.slot] + 1u;
    if (blockIdx.x == 0) {
      // block 0 talks to the other ranks (system scope); the rest of the grid watches a local word at device scope -- a
      // thousand blocks polling the system-scope flag slowed the multicast traffic they were waiting for by 2x
      const unsigned target = next * static_cast<unsigned>(sy.world);
      __threadfence_system();
      asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(sy.mc_flags + sy.slot), "r"(1u) : "memory");
      const long long t0 = clock64();
      unsigned seen;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(sy.local_flags + sy.slot) : "memory");
        if (clock64() - t0 > 20000000000LL) {
          printf("fxn: data-parallel barrier timeout (slot %d, have %u, want %u)\n", sy.slot, seen, target);
          __trap();
        }
      } while (static_cast<int>(seen - target) < 0);
      __threadfence_system();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(sy.go + sy.slot), "r"(next) : "memory");
    } else {
      unsigned seen;
      const long long t0 = clock64();
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(sy.go + sy.slot) : "memory");
        if (static_cast<int>(seen - next) < 0) __nanosleep(200);
        if (clock64() - t0 > 24000000000LL) __trap();
      } while (static_cast<int>(seen - next) < 0);
    }
  }
  __syncthreads();
}
// call from ONE thread per block after the block's work
__device__ __forceinline__ void dp_sync_leave(const DpSync& sy) {
  if (sy.mc_flags == nullptr) return;
  __threadfence();
  if (atomicAdd(sy.done + sy.slot, 1u) + 1u == gridDim.x) {
    sy.epoch[sy.slot] += 1u;
    sy.done[sy.slot] = 0u;
  }
}

// grad_local[i] = scale * sum_ranks grad[i] for i in [begin, end) (multiples of 4); slot `rank` of the symmetric partial
// array receives the squared norm of that slice on every rank. scratch: {double sum, unsigned arrivals}, zero on entry and
// left zero on exit.
__global__ void __launch_bounds__(256)
dp_reduce_kernel(const float* __restrict__ mc_grad, float* __restrict__ grad_local, long long begin, long long end,
                 float scale, float* __restrict__ mc_partials, int rank, double* __restrict__ scratch_sum,
                 unsigned* __restrict__ scratch_count, long long* __restrict__ step, const DpSync sy, const DpPeers peers,
                 const DpFused fused) {
  __shared__ double s_part[8];
  dp_sync_enter(sy);                                    // every rank's gradients are complete and visible
  double acc = 0.0;
  const long long n4 = (end - begin) / 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  // a load through the switch takes microseconds: four independent ones per thread per round
  for (long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < n4; i0 += 4 * stride) {
    float4 g[4];
    bool pulled[4];
    // ranges already reduce-scattered by the GEMMs (fxn_gemm_desc.rs_*): own contribution + inbox, both local; clear the inbox
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      pulled[k] = false;
      if (i0 + k * stride < n4 && fused.n > 0) {
        const unsigned off = static_cast<unsigned>(begin + 4 * (i0 + k * stride));
        bool in = false;
        for (int r = 0; r < fused.n; ++r) in = in || (off >= fused.lo[r] && off < fused.hi[r]);
        if (in) {
          const float4 a = *reinterpret_cast<const float4*>(grad_local + off);
          const float4 b = __ldcg(reinterpret_cast<const float4*>(fused.inbox + off));
          *reinterpret_cast<float4*>(fused.inbox + off) = make_float4(0.f, 0.f, 0.f, 0.f);
          g[k] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
          pulled[k] = true;
        }
      }
    }
    if (peers.n > 0) {
      // the W copies are read over NVLink as plain peer loads (all in flight at once) and added in rank order
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (i0 + k * stride < n4 && !pulled[k]) {
          const long long off = begin + 4 * (i0 + k * stride);
          float4 v[8];
#pragma unroll
          for (int r = 0; r < 8; ++r)
            if (r < peers.n) v[r] = __ldcg(reinterpret_cast<const float4*>(peers.p[r] + off));
          float4 t = v[0];
#pragma unroll
          for (int r = 1; r < 8; ++r)
            if (r < peers.n) { t.x += v[r].x; t.y += v[r].y; t.z += v[r].z; t.w += v[r].w; }
          g[k] = t;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i0 + k * stride < n4 && !pulled[k]) g[k] = multimem_ld_reduce_add(mc_grad + begin + 4 * (i0 + k * stride));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k * stride < n4) {
        g[k].x *= scale; g[k].y *= scale; g[k].z *= scale; g[k].w *= scale;
        *reinterpret_cast<float4*>(grad_local + begin + 4 * (i0 + k * stride)) = g[k];
        acc += static_cast<double>(g[k].x * g[k].x + g[k].y * g[k].y + g[k].z * g[k].z + g[k].w * g[k].w);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(scratch_sum, t);
    __threadfence();
    const unsigned arrived = atomicAdd(scratch_count, 1u) + 1u;
    if (arrived == gridDim.x) {                       // last block: publish this rank's partial to every rank
      __threadfence();
      const double total = *reinterpret_cast<volatile double*>(scratch_sum);
      multimem_st_scalar(mc_partials + rank, static_cast<float>(total));
      *scratch_sum = 0.0;
      *scratch_count = 0u;
      if (step) *step += 1;
      __threadfence_system();
    }
    dp_sync_leave(sy);
  }
}

// clip_grad_norm_(max_norm) with the global norm assembled from the W partials, Adam on [begin, end), new parameters
// multicast into every rank's arena. exp_avg / exp_avg_sq are indexed like the arena (only the slice is touched).
__global__ void __launch_bounds__(256)
dp_adam_bcast_kernel(float* __restrict__ mc_param, const float* __restrict__ param_local, const float* __restrict__ grad_local,
                     float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, long long begin, long long end,
                     const float* __restrict__ partials, int world, float lr, float beta1, float beta2, float eps,
                     float max_norm, const long long* __restrict__ step, float* __restrict__ norm_out, const DpSync sy) {
  dp_sync_enter(sy);                                    // all partial norms have landed everywhere
  float total = 0.f;
  for (int r = 0; r < world; ++r) total += __ldcg(partials + r);      // written by peers through the multicast address
  const float norm = sqrtf(total);
  const float coef = max_norm > 0.f ? fminf(max_norm / (norm + 1e-6f), 1.f) : 1.f;
  if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
  // same arithmetic as clip_adam_kernel (optim.cu); *step was incremented by fxn_dp_reduce_sumsq of this step
  const double t = static_cast<double>(*step);
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), t));
  const float bc2 = static_cast<float>(1.0 - pow(static_cast<double>(beta2), t));
  const float step_size = lr / bc1;
  const float bc2_sqrt = sqrtf(bc2);
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  const long long n4 = (end - begin) / 4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long off = begin + 4 * i;
    const float4 g4 = *reinterpret_cast<const float4*>(grad_local + off);
    float4 m4 = *reinterpret_cast<const float4*>(exp_avg + off);
    float4 v4 = *reinterpret_cast<const float4*>(exp_avg_sq + off);
    float4 p4 = *reinterpret_cast<const float4*>(param_local + off);
    const float g[4] = {g4.x * coef, g4.y * coef, g4.z * coef, g4.w * coef};
    float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w}, p[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      m[j] = m[j] + (g[j] - m[j]) * omb1;
      v[j] = v[j] * beta2 + omb2 * g[j] * g[j];
      p[j] = p[j] - step_size * (m[j] / (sqrtf(v[j]) / bc2_sqrt + eps));
    }
    *reinterpret_cast<float4*>(exp_avg + off) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(exp_avg_sq + off) = make_float4(v[0], v[1], v[2], v[3]);
    multimem_st(mc_param + off, make_float4(p[0], p[1], p[2], p[3]));
  }
  if (sy.mc_flags != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) dp_sync_leave(sy);
  }
}

__global__ void dp_step_inc_kernel(long long* step) { *step += 1; }

// System-wide barrier INSIDE the stream (one thread per rank): add 1 to flag `slot` of EVERY rank through the multicast
// address (one multimem.red), then wait until the local copy has collected `world` arrivals for this rank's epoch. The
// epoch counter lives on the device, so a captured graph replays it correctly; release / acquire at system scope orders
// the gradient / parameter traffic of the kernels before and after it. Replaces three host-enqueued symmetric-memory
// barriers per step (they kept the step out of a CUDA graph and cost ~30 us each in the round-1 scaling runs).
__global__ void dp_barrier_kernel(unsigned* __restrict__ mc_flags, const unsigned* __restrict__ local_flags,
                                  unsigned* __restrict__ epoch, int slot, int world) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned e = epoch[slot] + 1u;
  epoch[slot] = e;
  __threadfence_system();
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_flags + slot), "r"(1u) : "memory");
  const unsigned target = e * static_cast<unsigned>(world);
  const long long t0 = clock64();
  unsigned seen;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(local_flags + slot) : "memory");
    if (clock64() - t0 > 20000000000LL) {               // ~10 s: a rank is missing -- fail the launch instead of hanging the GPU
      printf("fxn: data-parallel barrier timeout (slot %d, have %u, want %u)\n", slot, seen, target);
      __trap();
    }
  } while (static_cast<int>(seen - target) < 0);
  __threadfence_system();
}

}  // namespace fxn

using namespace fxn;

static DpSync make_sync(void* mc_flags, const void* local_flags, void* epoch, int slot, int world) {
  DpSync sy;
  sy.mc_flags = static_cast<unsigned*>(mc_flags);
  sy.local_flags = static_cast<const unsigned*>(local_flags);
  sy.epoch = static_cast<unsigned*>(epoch);
  sy.done = sy.epoch ? sy.epoch + 16 : nullptr;           // epoch[0..15] epochs, [16..31] finished-block counters,
  sy.go = sy.epoch ? sy.epoch + 32 : nullptr;             // [32..47] device-scope release words
  sy.slot = slot;
  sy.world = world;
  return sy;
}

extern "C" int fxn_dp_reduce_sumsq(const void* mc_grad, float* grad_local, long long begin, long long end, float scale,
                                   void* mc_partials, int rank, void* scratch16, long long* step_counter, void* mc_flags,
                                   const void* local_flags, void* epoch32, int slot, int world,
                                   const float* const* peer_grads, int npeers, float* inbox, const long long* fused_ranges,
                                   int nranges, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!mc_grad || !grad_local || !mc_partials || !scratch16) return set_error(FXN_ERR_ARG, "fxn_dp_reduce_sumsq: null argument");
  if (begin % 4 || end % 4 || end < begin || (reinterpret_cast<uintptr_t>(mc_grad) & 15) || (reinterpret_cast<uintptr_t>(grad_local) & 15))
    return set_error(FXN_ERR_ARG, "fxn_dp_reduce_sumsq: slice bounds and bases must be 16-byte aligned");
  if (mc_flags && (!local_flags || !epoch32 || slot < 0 || slot >= 16 || world < 1))
    return set_error(FXN_ERR_ARG, "fxn_dp_reduce_sumsq: bad barrier arguments");
  long long n4 = (end - begin) / 4;
  int blocks = static_cast<int>((n4 + 255) / 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  DpPeers peers;
  peers.n = 0;
  if (peer_grads) {
    if (npeers < 1 || npeers > 8) return set_error(FXN_ERR_ARG, "fxn_dp_reduce_sumsq: 1..8 peer arenas");
    for (int r = 0; r < npeers; ++r) {
      if (!peer_grads[r] || (reinterpret_cast<uintptr_t>(peer_grads[r]) & 15))
        return set_error(FXN_ERR_ARG, "fxn_dp_reduce_sumsq: peer arenas must be 16-byte aligned");
      peers.p[r] = peer_grads[r];
    }
    peers.n = npeers;
  }
  DpFused fused;
  fused.n = 0;
  fused.inbox = inbox;
  if (nranges > 0) {
    if (!inbox || !fused_ranges || nranges > 8 || (reinterpret_cast<uintptr_t>(inbox) & 15))
      return set_error(FXN_ERR_ARG, "fxn_dp_reduce_sumsq: <= 8 fused ranges and a 16-byte aligned inbox");
    for (int r = 0; r < nranges; ++r) {
      const long long lo = fused_ranges[2 * r], hi = fused_ranges[2 * r + 1];
      if (lo < 0 || hi < lo || lo % 4 || hi % 4 || hi >= (1LL << 31))
        return set_error(FXN_ERR_ARG, "fxn_dp_reduce_sumsq: fused ranges must be 4-aligned and below 2^31");
      fused.lo[r] = static_cast<unsigned>(lo);
      fused.hi[r] = static_cast<unsigned>(hi);
    }
    fused.n = nranges;
  }
  dp_reduce_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float*>(mc_grad), grad_local, begin, end, scale,
                                               static_cast<float*>(mc_partials), rank, static_cast<double*>(scratch16),
                                               reinterpret_cast<unsigned*>(static_cast<char*>(scratch16) + 8), step_counter,
                                               make_sync(mc_flags, local_flags, epoch32, slot, world), peers, fused);
  FXN_CHECK_LAUNCH("dp_reduce");
  return 0;
}

extern "C" int fxn_dp_adam_bcast(void* mc_param, const float* param_local, const float* grad_local, float* exp_avg,
                                 float* exp_avg_sq, long long begin, long long end, const float* partials, int world, float lr,
                                 float beta1, float beta2, float eps, float max_norm, const long long* step_counter,
                                 float* norm_out, void* mc_flags, const void* local_flags, void* epoch32, int slot,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!mc_param || !param_local || !grad_local || !exp_avg || !exp_avg_sq || !partials || !step_counter)
    return set_error(FXN_ERR_ARG, "fxn_dp_adam_bcast: null argument");
  if (begin % 4 || end % 4 || end < begin) return set_error(FXN_ERR_ARG, "fxn_dp_adam_bcast: slice bounds must be multiples of 4");
  if (mc_flags && (!local_flags || !epoch32 || slot < 0 || slot >= 16))
    return set_error(FXN_ERR_ARG, "fxn_dp_adam_bcast: bad barrier arguments");
  long long n4 = (end - begin) / 4;
  int blocks = static_cast<int>((n4 + 255) / 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  dp_adam_bcast_kernel<<<blocks, 256, 0, stream>>>(static_cast<float*>(mc_param), param_local, grad_local, exp_avg, exp_avg_sq,
                                                   begin, end, partials, world, lr, beta1, beta2, eps, max_norm, step_counter,
                                                   norm_out, make_sync(mc_flags, local_flags, epoch32, slot, world));
  FXN_CHECK_LAUNCH("dp_adam_bcast");
  return 0;
}

extern "C" int fxn_dp_barrier(void* mc_flags, const void* local_flags, void* epoch, int slot, int world, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!mc_flags || !local_flags || !epoch || slot < 0 || slot >= 16 || world < 1)
    return set_error(FXN_ERR_ARG, "fxn_dp_barrier: bad argument");
  dp_barrier_kernel<<<1, 32, 0, stream>>>(static_cast<unsigned*>(mc_flags), static_cast<const unsigned*>(local_flags),
                                          static_cast<unsigned*>(epoch), slot, world);
  FXN_CHECK_LAUNCH("dp_barrier");
  return 0;
}
