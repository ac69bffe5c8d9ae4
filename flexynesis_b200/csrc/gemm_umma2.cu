// gemm2: persistent tcgen05 GEMM, C[M,N] = sum over bf16 split terms of A[M,K] * B[N,K]^T, one kernel for every dense
// contraction of the training step.
//
// Design points and why (profiles/r01_*, profiles/r02_*):
//  * CTA PAIRS (tcgen05 cta_group::2, cluster 2x1x1): a pair computes a 256 x bn tile; each CTA stages its own 128 rows
//    of A and only HALF of the B tile, the MMA reads both halves through the pair's shared-memory window. The mid-size
//    GEMMs of the step (M = 4096, K = 5000, N = 512) are bound by L2 -> SM operand traffic, not by the tensor pipe:
//    128x128 tiles move 650 MB per launch, pair tiles of 256x256 move 324 MB.
//  * PERSISTENT CTAs with a DOUBLE-BUFFERED TMEM accumulator (2 x bn columns): the epilogue of tile i (TMEM -> registers
//    -> global) overlaps the TMA/MMA main loop of tile i+1; the small-K GEMMs (dgrad K = 128..512, decoder, Gram) were
//    epilogue-serialised at one tile per CTA launch.
//  * CHUNKED EPILOGUE: 32-column chunks are transposed through a 4.5 KB per-warp staging buffer instead of a full
//    128 x bn fp32 tile, which frees the shared memory for pipeline stages.
//  * STREAM-K for plain fp32 outputs (the wgrads, K = batch): the (tile, k-block) space is cut into equal contiguous
//    ranges, one per CTA pair, partial tiles are combined with 16-byte vector reductions into the zeroed C.
//  * STREAM-K WITH FIX-UP for every other epilogue (streamk = 2): the group that holds the LAST k-block of a tile owns
//    its epilogue; groups holding earlier k-blocks store their raw partial accumulators into their own workspace slot
//    (register order, fully coalesced, no shared-memory staging, no atomics) and raise a flag; the owner adds the slots
//    of its contributors to its own accumulator, in a fixed order, before the fused epilogue. Each group runs its trailing (contributor)
//    segment FIRST, so an owner never waits on work that is queued behind another owner's wait. This is what lets a
//    32-tile problem (config 2's [4096 x 512] layer) run MMA-paced 256-wide tiles on a caller-chosen number of SMs.
//
// Roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA of the pair only) + TMEM
// allocation, warps 2..5 = epilogue (TMEM lane quarter = warp % 4).
#include "fxn_internal.h"
#include "ptx.cuh"
#include "gemm_common.cuh"
#include <cstdlib>
#include <cstdio>
#include <cmath>

namespace fxn {

constexpr int G2_BM = 128;          // rows of A per CTA
constexpr int G2_BK = 64;           // one 128-byte swizzle atom of bf16 along K
constexpr int G2_UK = 16;
constexpr int G2_MAX_STAGES = 8;
constexpr int G2_CHUNK = 32;        // epilogue column chunk
constexpr int G2_STG_LD = 36;       // staging row stride in floats (16-byte aligned rows, spreads banks)
constexpr int G2_STG_BYTES = 4 * 32 * G2_STG_LD * 4;
constexpr int G2_MAX_DYN_SMEM = 229376;   // 227 KB per CTA minus the static shared memory (barriers, statistics exchange)

struct Gemm2Args {
  int M, N, K;
  int bn;                 // tile width of the CTA group (columns of C)
  int a_mn, b_mn;
  int nterms;
  int stages;
  int streamk;            // 1: equal (tile, k-block) ranges per group, reductions into zeroed C
  int tiles_m, tiles_n, kb_total;
  uint32_t tmem_cols;
  float* C; long long ldc;
  const float* bias;
  __nv_bfloat16* c_hi; __nv_bfloat16* c_lo; long long ldp;
  float* colstats; int stats_mode;
  int epi_act; int accumulate;
  float alpha; const float* alpha_dev;
  const float* mse_x; long long ldx; float* mse_acc;
  const float* gauss_ra; const float* gauss_rb; float gauss_inv;
  float stats_alpha; const float* stats_alpha_dev;
  int trace;
  int epi_variant;        // index of the specialised epilogue loop for interior chunks, -1 = generic only
  float* fix_ws;          // streamk == 2: [groups][CG][4 quarters][bn / 32][8][32 lanes][4] fp32 partial accumulators
  unsigned* fix_flags;    // streamk == 2: [2][groups]: contributor arrivals, owner departures (zero between launches)
  // Data-parallel reduce-scatter fused into the stream-K reductions (rs_world > 1): C lies in this rank's gradient arena
  // (base rs_base), the arena is cut into rs_world slices of rs_per elements (the last one takes the remainder), and a
  // reduction whose target element belongs to another rank's slice goes to THAT rank's inbox arena over NVLink instead
  // of into the local C. The exchange rides under the GEMM's main loop; csrc/dp.cu adds own + inbox afterwards.
  float* rs_inbox[8];
  const float* rs_base;
  unsigned rs_per;
  int rs_world, rs_rank;
};

// where the reduction for local element address `cp` has to go (rs_world > 1): the local C for this rank's own slice, the
// owner's inbox otherwise. per is a multiple of 4 and every arena offset below 2^31, so a 16-byte aligned float4 never
// straddles two slices.
__device__ __forceinline__ float* rs_target(const Gemm2Args& p, float* cp) {
  const unsigned off = static_cast<unsigned>(cp - p.rs_base);
  int owner = static_cast<int>(off / p.rs_per);
  owner = owner < p.rs_world ? owner : p.rs_world - 1;
  return owner == p.rs_rank ? cp : p.rs_inbox[owner] + off;
}

// ---- PTX pieces that differ between cta_group::1 and ::2 ----
template <int CG> __device__ __forceinline__ void g2_tmem_alloc(uint32_t* dst, uint32_t ncols) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG> __device__ __forceinline__ void g2_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int CG>
__device__ __forceinline__ void g2_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 1) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
                 "r"(idesc), "r"(acc) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
                 "r"(idesc), "r"(acc) : "memory");
  }
}
// arrive on the barrier at this shared-memory offset in every CTA of the group once the issued MMAs have retired
template <int CG> __device__ __forceinline__ void g2_commit(uint64_t* bar) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  } else {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)), "h"(mask) : "memory");
  }
}
// TMA tile load whose byte count is credited to the LEADER CTA's barrier (peer bit of the address cleared)
template <int CG>
__device__ __forceinline__ void g2_tma_load(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  if constexpr (CG == 1) {
    tma_load_2d(smem_dst, m, bar, c0, c1);
  } else {
    const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c0), "r"(c1) : "memory");
  }
}
// one lane of a converged warp (elect.sync): the issuing lane of the TMA / MMA roles
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier with this local offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// ---- optional timeline of CTA 0 / CTA 1 (clock64 stamps), enabled with FXN_GEMM_TRACE=1; read by fxn_debug_gemm_trace ----
__device__ long long g2_trace[2][16];
__device__ unsigned long long g2_cta_time[2][512];     // %globaltimer at start / end of every CTA (trace mode)
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define G2_STAMP(slot)                                                        \
  do {                                                                        \
    if (p.trace && blockIdx.x < 2) g2_trace[blockIdx.x][slot] = clock64();    \
  } while (0)

// ---- work decomposition shared by the three roles ----
struct Segment { int mt, nt, kb0, kb1; long long tile; };
struct WorkIter {
  long long u, end;     // stream-K: position / end in the (tile, k-block) space; tile mode: tile index / tile count
  long long tail0;      // stream-K: start of the trailing partial segment, which is processed FIRST (-1: none / done)
  long long units, tail_end;
  int step;             // tile mode: stride
  int kb_total, tiles_n, streamk, group, ngroups;
  __device__ __forceinline__ long long bound(int g) const { return units * g / ngroups; }
  __device__ __forceinline__ WorkIter(const Gemm2Args& p, int group_, int ngroups_) {
    kb_total = p.kb_total; tiles_n = p.tiles_n; streamk = p.streamk; group = group_; ngroups = ngroups_;
    const long long tiles = static_cast<long long>(p.tiles_m) * p.tiles_n;
    units = tiles * kb_total;
    tail0 = -1; tail_end = 0;
    if (streamk) {
      u = bound(group);
      end = bound(group + 1);
      step = 0;
      // A range that stops inside a tile ends with a CONTRIBUTOR segment (it does not hold the tile's last k-block). Run it
      // first: the owner of that tile is the next group, and it must not have to wait until this group has worked through
      // everything in front of the segment (or, transitively, through this group's own wait for ITS contributors).
      const long long last_tile_start = end / kb_total * kb_total;
      if (end != last_tile_start && last_tile_start > u) { tail0 = last_tile_start; tail_end = end; end = last_tile_start; }
    } else {
      u = group; end = tiles; step = ngroups;
    }
  }
  __device__ __forceinline__ void fill(Segment& s, long long a, long long b) const {
    const long long tile = a / kb_total;
    s.tile = tile;
    s.kb0 = static_cast<int>(a - tile * kb_total);
    s.kb1 = static_cast<int>(b - tile * kb_total);
    s.mt = static_cast<int>(tile / tiles_n); s.nt = static_cast<int>(tile - static_cast<long long>(s.mt) * tiles_n);
  }
  __device__ __forceinline__ bool next(Segment& s) {
    if (streamk) {
      if (tail0 >= 0) { fill(s, tail0, tail_end); tail0 = -1; return true; }
      if (u >= end) return false;
      const long long tile_end = (u / kb_total + 1) * kb_total;
      const long long b = tile_end < end ? tile_end : end;
      fill(s, u, b);
      u = b;
      return true;
    }
    if (u >= end) return false;
    s.tile = u;
    s.mt = static_cast<int>(u / tiles_n); s.nt = static_cast<int>(u - static_cast<long long>(s.mt) * tiles_n);
    s.kb0 = 0; s.kb1 = kb_total;
    u += step;
    return true;
  }
  // group that holds unit x (largest g with bound(g) <= x)
  __device__ __forceinline__ int group_of(long long x) const {
    int g = static_cast<int>(x * ngroups / units);
    while (g + 1 < ngroups && bound(g + 1) <= x) ++g;
    while (g > 0 && bound(g) > x) --g;
    return g;
  }
};


// explicit shared-window accesses: a staging pointer that travels through a struct loses its address space and the
// compiler falls back to generic LD/ST (measured: 4x slower epilogue)
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f2(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts_f4(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
// 16-byte vector reduction into global memory (REDG.E.ADD.F32x4): one L2 atomic request per four accumulator values
__device__ __forceinline__ void red_add_f4(float* p, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

__device__ __forceinline__ float4 ldcg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Stream-K partial tile -> C for an interior chunk whose rows are 16-byte aligned: lane = (row % 4, four consecutive
// columns), 8 passes over the staged 32 x 32 chunk. The scalar form issued 32 REDs per lane and chunk and made the
// weight-gradient epilogue as long as its main loop (profiles/r02_wgrad_red.log).
__device__ __forceinline__ void epi_red4(const Gemm2Args& p, uint32_t stg_s, int lane, int mbase, int col0, float alpha,
                                         bool add_bias, int rows_valid) {
  const int r4 = lane >> 3, c4 = (lane & 7) * 4;
  float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
  if (add_bias) {
    b0 = __ldg(p.bias + col0 + c4); b1 = __ldg(p.bias + col0 + c4 + 1);
    b2 = __ldg(p.bias + col0 + c4 + 2); b3 = __ldg(p.bias + col0 + c4 + 3);
  }
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = lds_f4(stg_s + ((4 * i + r4) * G2_STG_LD + c4) * 4);
  float* cp = p.C + static_cast<long long>(mbase + r4) * p.ldc + col0 + c4;
  if (p.rs_world > 1) {
    // one owner for all eight rows of this lane (the usual case): one address translation; otherwise row by row
    float* first = rs_target(p, cp);
    float* last = rs_target(p, cp + static_cast<long long>(28) * p.ldc);
    const bool uniform = (first - cp) == (last - (cp + static_cast<long long>(28) * p.ldc));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (4 * i + r4 < rows_valid) {
        float* row = cp + static_cast<long long>(4 * i) * p.ldc;
        float* dst = uniform ? row + (first - cp) : rs_target(p, row);
        red_add_f4(dst, fmaf(v[i].x, alpha, b0), fmaf(v[i].y, alpha, b1), fmaf(v[i].z, alpha, b2), fmaf(v[i].w, alpha, b3));
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (4 * i + r4 < rows_valid)
      red_add_f4(cp + static_cast<long long>(4 * i) * p.ldc, fmaf(v[i].x, alpha, b0), fmaf(v[i].y, alpha, b1),
                 fmaf(v[i].z, alpha, b2), fmaf(v[i].w, alpha, b3));
}

// ---- epilogue inner loop: 16 passes over a staged 32 x 32 chunk, lane = (row parity, column pair) ----
struct EpiRowCtx {
  uint32_t stg;           // shared-window address of this lane's first staged element (row rr, columns 2cp, 2cp + 1)
  int mbase, rr, col, rows_valid, padN;
  bool ok0, ok1, c_vec2, x_vec2, atomic_out;
  float alpha, b0, b1, rb0, rb1;
};

// RT = true: every mode is read from the descriptor at run time and every access is bounds-checked (edge chunks and
// rare mode combinations). RT = false: modes are template constants and the chunk is known to be interior (32 valid
// rows, 32 valid columns, vector-aligned), so the loop body is a handful of instructions.
// CMODE: 0 no fp32 output, 1 store, 2 accumulate, 3 reduction (stream-K). STATS: 0 none, 1 sums, 2 sums + M2, 3 column sums.
template <bool RT, int ACT, int CMODE, bool PLANES, int STATS, bool MSE>
__device__ __forceinline__ void epi_rows(const Gemm2Args& p, const EpiRowCtx& cx, float& s0, float& s1, float& sq_acc) {
  const int act = RT ? p.epi_act : ACT;
  const int cmode = RT ? (p.C == nullptr ? 0 : (cx.atomic_out ? 3 : (p.accumulate ? 2 : 1))) : CMODE;
  const bool planes = RT ? (p.c_hi != nullptr) : PLANES;
  const int stats = RT ? p.stats_mode : STATS;
  const bool mse = RT ? (p.mse_x != nullptr) : MSE;
  const long long row0 = cx.mbase + cx.rr;
  float* cptr = cmode ? p.C + row0 * p.ldc + cx.col : nullptr;
  const float* xptr = mse ? p.mse_x + row0 * p.ldx + cx.col : nullptr;
  __nv_bfloat16* hptr = planes ? p.c_hi + row0 * p.ldp + cx.col : nullptr;
  __nv_bfloat16* lptr = planes ? p.c_lo + row0 * p.ldp + cx.col : nullptr;
  // Gaussian-kernel epilogue: the row norms of this warp's 32 rows are fetched ONCE per chunk (lane = row) and handed out
  // by shuffle; a load per loop iteration exposed 16 L2 latencies per chunk (the [4096 x 4096] Gram GEMM ran 6x slower
  // than a plain-store GEMM of its size).
  float ra_lane = 0.f;
  if (act == 7) ra_lane = __ldg(p.gauss_ra + min(cx.mbase + static_cast<int>(threadIdx.x & 31), p.M - 1));
  // Fused reconstruction error: the 16 target values this lane needs come from HBM (x is read exactly once per step).
  // Issued one per loop iteration they cost 16 exposed memory latencies per chunk (the Decoder output GEMM ran at 1/6 of
  // the encoder GEMM's rate, profiles/r01_timeline_cfg3_before_mse_prefetch.log); interior chunks fetch all 16 up front.
  float2 xt[16];
  if constexpr (!RT && MSE) {
#pragma unroll
    for (int i = 0; i < 16; ++i) xt[i] = __ldg(reinterpret_cast<const float2*>(xptr + static_cast<long long>(2 * i) * p.ldx));
  }
  if constexpr (!RT && CMODE == 2 && !MSE) {      // C += : the old values, same reasoning
#pragma unroll
    for (int i = 0; i < 16; ++i) xt[i] = *reinterpret_cast<const float2*>(cptr + static_cast<long long>(2 * i) * p.ldc);
  }
  // Interior chunks read their 16 staged values up front: the epilogue runs ONE warp per scheduler, so every
  // shared-memory load left inside the loop is an exposed latency (the loads are volatile asm and are not hoisted by
  // the compiler); with the values in registers the 16 iterations are independent instruction streams.
  float2 xin[16];
  if constexpr (!RT) {
#pragma unroll
    for (int i = 0; i < 16; ++i) xin[i] = lds_f2(cx.stg + 2 * i * G2_STG_LD * 4);
  }
#pragma unroll (RT ? 4 : 16)
  for (int i = 0; i < 16; ++i) {
    const int r = 2 * i + cx.rr;
    const bool rok = r < cx.rows_valid;      // specialised variants too: a ragged last row quarter (M % 32 != 0) stays on them
    const bool ok0 = RT ? cx.ok0 : true, ok1 = RT ? cx.ok1 : true;
    float2 x;
    if constexpr (RT) x = lds_f2(cx.stg + 2 * i * G2_STG_LD * 4); else x = xin[i];
    x.x = fmaf(x.x, cx.alpha, cx.b0);
    x.y = fmaf(x.y, cx.alpha, cx.b1);
    if (act == 7) {
      const float ra = __shfl_sync(0xffffffffu, ra_lane, r);
      x.x = __expf(-fmaxf(ra + cx.rb0 - 2.f * x.x, 0.f) * p.gauss_inv);
      x.y = __expf(-fmaxf(ra + cx.rb1 - 2.f * x.y, 0.f) * p.gauss_inv);
    } else if (act == 1) {
      x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f);
    } else if (act == 6) {
      x.x = x.x > 0.f ? x.x : 0.2f * x.x; x.y = x.y > 0.f ? x.y : 0.2f * x.y;
    } else if (act == 3) {
      // approximate reciprocal (MUFU.RCP, ~1 ulp): an IEEE division costs ~10 dependent instructions per element and the
      // epilogue runs one warp per scheduler, so dependent-instruction count is what its time is made of
      x.x = __fdividef(1.f, 1.f + __expf(-x.x)); x.y = __fdividef(1.f, 1.f + __expf(-x.y));
    }
    if (RT) { if (!ok0) x.x = 0.f; if (!ok1) x.y = 0.f; }
    if (cmode != 0 && rok) {
      float* cp_ = cptr + static_cast<long long>(2 * i) * p.ldc;
      if (cmode == 3) {
        if (p.rs_world > 1) {
          if (ok0) atomicAdd(rs_target(p, cp_), x.x);
          if (ok1) atomicAdd(rs_target(p, cp_ + 1), x.y);
        } else {
          if (ok0) atomicAdd(cp_, x.x);
          if (ok1) atomicAdd(cp_ + 1, x.y);
        }
      } else if (!RT || (ok1 && cx.c_vec2)) {
        float2 o = x;
        if constexpr (!RT && CMODE == 2 && !MSE) { o.x += xt[i].x; o.y += xt[i].y; }
        else if (cmode == 2) { const float2 old = *reinterpret_cast<const float2*>(cp_); o.x += old.x; o.y += old.y; }
        *reinterpret_cast<float2*>(cp_) = o;
      } else {
        if (ok0) cp_[0] = cmode == 2 ? cp_[0] + x.x : x.x;
        if (ok1) cp_[1] = cmode == 2 ? cp_[1] + x.y : x.y;
      }
    }
    if (mse) {
      // fused Decoder output: x = x_hat; accumulate (x_hat - x)^2, continue with G = (x_hat - x) x_hat (1 - x_hat)
      float t0 = 0.f, t1 = 0.f;
      if (rok) {
        const float* xp = xptr + static_cast<long long>(2 * i) * p.ldx;
        if constexpr (!RT) { t0 = xt[i].x; t1 = xt[i].y; (void)xp; }
        else if (ok1 && cx.x_vec2) { const float2 t = __ldg(reinterpret_cast<const float2*>(xp)); t0 = t.x; t1 = t.y; }
        else { if (ok0) t0 = __ldg(xp); if (ok1) t1 = __ldg(xp + 1); }
      }
      const float d0 = (rok && ok0) ? x.x - t0 : 0.f, d1 = (rok && ok1) ? x.y - t1 : 0.f;
      sq_acc = fmaf(d0, d0, sq_acc);
      sq_acc = fmaf(d1, d1, sq_acc);
      x.x = d0 * x.x * (1.f - x.x);
      x.y = d1 * x.y * (1.f - x.y);
    }
    if (planes && rok && (!RT || cx.col < cx.padN)) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(x.x, h0, l0);
      split_bf16(x.y, h1, l1);
      const long long off = static_cast<long long>(2 * i) * p.ldp;
      *reinterpret_cast<__nv_bfloat162*>(hptr + off) = __halves2bfloat162(h0, h1);
      *reinterpret_cast<__nv_bfloat162*>(lptr + off) = __halves2bfloat162(l0, l1);
    }
    if (stats != 0) {
      if (rok) { s0 += x.x; s1 += x.y; }
      if (stats == 2) sts_f2(cx.stg + 2 * i * G2_STG_LD * 4, x);   // the M2 pass reads it back
    }
  }
}

// EW = number of epilogue warps: 4 (one per TMEM lane quarter) or 8 (two per quarter, alternating 32-column chunks: two
// warps per scheduler hide each other's instruction latency; used for short-K, epilogue-bound problems where the second
// set of staging buffers does not cost a pipeline stage that matters).
template <int CG, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
             const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const Gemm2Args p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[G2_MAX_STAGES];
  __shared__ uint64_t empty_bar[G2_MAX_STAGES];
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_stat_all[EW / 4][4][2][G2_CHUNK];     // cross-warp merge of column statistics (per chunk parity)

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int group = blockIdx.x / CG, ngroups = gridDim.x / CG;
  const int nplanes = (p.nterms == 3) ? 2 : 1;
  const int bnl = p.bn / CG;                                   // rows of the B tile staged by this CTA
  const uint32_t a_bytes = G2_BM * G2_BK * 2;
  const uint32_t b_bytes = static_cast<uint32_t>(bnl) * G2_BK * 2;
  const uint32_t stage_bytes = nplanes * (a_bytes + b_bytes);
  float* stg_all = reinterpret_cast<float*>(smem + static_cast<size_t>(p.stages) * stage_bytes);

  if (threadIdx.x == 0) G2_STAMP(0);
  if (p.trace && threadIdx.x == 0 && blockIdx.x < 512) g2_cta_time[0][blockIdx.x] = globaltimer_ns();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (nplanes == 2) { tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_lo); }
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], EW * CG); }
    fence_mbar_init();
  }
  if (warp == 1) g2_tmem_alloc<CG>(&tmem_base_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();      // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (threadIdx.x == 0) G2_STAMP(1);

  // The two single-thread roles run with the WHOLE warp in convergent control flow and one elected lane issuing: every
  // address / descriptor is then warp-uniform for the compiler (uniform registers feed UTMALDG / UTCHMMA directly).
  // Written as `if (lane == 0)` the same code compiled to a per-instruction "which lane holds the operand" loop around each
  // TMA and MMA issue, ~1060 cycles of issue work per k-block: every tile narrower than 224 columns ran at that floor
  // instead of at its MMA time (profiles/r02_bn_sweep_before.log).
  if (warp == 0) {
    // ===================== TMA producer (every CTA of the group) =====================
    WorkIter it(p, group, ngroups);
    Segment sg;
    int stage = 0;
    uint32_t phase = 0;
    while (it.next(sg)) {
      const int m0 = (sg.mt * CG + static_cast<int>(rank)) * G2_BM;
      const int n0 = sg.nt * p.bn + static_cast<int>(rank) * bnl;
      for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* st = smem + static_cast<size_t>(stage) * stage_bytes;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], stage_bytes * CG);
          const int k0 = kb * G2_BK;
          for (int pl = 0; pl < nplanes; ++pl) {
            uint8_t* sa = st + pl * a_bytes;
            uint8_t* sb = st + nplanes * a_bytes + pl * b_bytes;
            const CUtensorMap* ta = pl ? &tmA_lo : &tmA_hi;
            const CUtensorMap* tb = pl ? &tmB_lo : &tmB_hi;
            if (p.a_mn == 0) {
              g2_tma_load<CG>(sa, ta, &full_bar[stage], k0, m0);                       // box {64 k, 128 rows}
            } else {
              for (int a = 0; a < G2_BM / 64; ++a)                                     // box {64 m, 64 k rows}
                g2_tma_load<CG>(sa + a * (G2_BK * 128), ta, &full_bar[stage], m0 + a * 64, k0);
            }
            if (p.b_mn == 0) {
              g2_tma_load<CG>(sb, tb, &full_bar[stage], k0, n0);                       // box {64 k, bnl rows}
            } else {
              for (int a = 0; a < bnl / 64; ++a)
                g2_tma_load<CG>(sb + a * (G2_BK * 128), tb, &full_bar[stage], n0 + a * 64, k0);
            }
          }
          if (kb == sg.kb0) G2_STAMP(2);        // first load of the (last) segment issued
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    if (lane == 0) G2_STAMP(3);                 // all loads issued
  } else if (warp == 1) {
    // ===================== MMA issuer: the leader CTA's warp 1, one elected lane =====================
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(G2_BM * CG, p.bn, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? G2_BK * 128 : 16, b_lbo = p.b_mn ? G2_BK * 128 : 16;
      const uint32_t a_kstep = p.a_mn ? G2_UK * 128 : G2_UK * 2, b_kstep = p.b_mn ? G2_UK * 128 : G2_UK * 2;
      // descriptors of one stage differ from those of stage 0 / k-step 0 only in the 14-bit start-address field
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t da0 = umma_smem_desc_sw128(smem0, a_lbo, 1024);
      const uint64_t db0 = umma_smem_desc_sw128(smem0 + nplanes * a_bytes, b_lbo, 1024);
      WorkIter it(p, group, ngroups);
      Segment sg;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_iter = 0;
      while (it.next(sg)) {
        const uint32_t as = tile_iter & 1, aphase = (tile_iter >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);              // epilogues of both CTAs have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * static_cast<uint32_t>(p.bn);
        uint32_t acc = 0;
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            if (kb == sg.kb0 && tile_iter == 0) G2_STAMP(4);   // first stage landed
            const uint32_t soff = static_cast<uint32_t>(stage) * stage_bytes;
#pragma unroll
            for (int kk = 0; kk < G2_BK / G2_UK; ++kk) {
              const uint64_t da_hi = da0 + ((soff + kk * a_kstep) >> 4);
              const uint64_t db_hi = db0 + ((soff + kk * b_kstep) >> 4);
              if (nplanes == 2) {
                const uint64_t da_lo = da_hi + (a_bytes >> 4);
                const uint64_t db_lo = db_hi + (b_bytes >> 4);
                g2_mma<CG>(tmem_acc, da_lo, db_hi, idesc, acc);   // small terms first
                g2_mma<CG>(tmem_acc, da_hi, db_lo, idesc, 1u);
                g2_mma<CG>(tmem_acc, da_hi, db_hi, idesc, 1u);
              } else {
                g2_mma<CG>(tmem_acc, da_hi, db_hi, idesc, acc);
              }
              acc = 1;
            }
            g2_commit<CG>(&empty_bar[stage]);                    // frees this stage in every CTA of the group
            if (kb + 1 == sg.kb1) {
              g2_commit<CG>(&tfull_bar[as]);                     // accumulator complete -> both epilogues
              if (tile_iter == 0) G2_STAMP(5);                 // first tile fully issued
            }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        ++tile_iter;
      }
      if (lane == 0) G2_STAMP(6);                              // all MMAs issued
    }
  } else {
    // ===================== Epilogue: 4 warps, TMEM lane quarter = warp % 4 =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;                          // 0 (EW = 4); EW = 8: which chunk parity this warp takes
    constexpr int CH_STEP = EW / 4;
    float (*s_stat)[2][G2_CHUNK] = s_stat_all[half];
    const uint32_t stg_s = smem_u32(stg_all + (warp - 2) * (32 * G2_STG_LD));
    const int rr = lane >> 4, cp = lane & 15;                 // transposed layout: 2 rows x 16 column pairs per pass
    const float alpha = p.alpha * (p.alpha_dev ? __ldg(p.alpha_dev) : 1.f);
    const float sscale = p.stats_alpha * (p.stats_alpha_dev ? __ldg(p.stats_alpha_dev) : 1.f);
    const int padN = (p.N + 7) & ~7;
    const bool c_vec2 = p.C && ((reinterpret_cast<uintptr_t>(p.C) & 7) == 0) && (p.ldc % 2 == 0);
    const bool x_vec2 = p.mse_x && ((reinterpret_cast<uintptr_t>(p.mse_x) & 7) == 0) && (p.ldx % 2 == 0);
    const bool c_vec4 = p.C && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (p.ldc % 4 == 0);
    WorkIter it(p, group, ngroups);
    Segment sg;
    uint32_t tile_iter = 0;
    while (it.next(sg)) {
      const uint32_t as = tile_iter & 1, aphase = (tile_iter >> 1) & 1;
      const int mbase = (sg.mt * CG + static_cast<int>(rank)) * G2_BM + q * 32;   // first row of this warp's quarter
      const int n0 = sg.nt * p.bn;
      const int rows_valid = max(0, min(32, p.M - mbase));
      // stream-K with fix-up (streamk == 2): the segment holding the tile's last k-block owns the epilogue
      const bool fix = p.streamk == 2;
      const bool owner = sg.kb1 == p.kb_total;
      const bool contributor = fix && !owner;
      const bool has_contrib = fix && owner && sg.kb0 > 0;
      const bool add_bias = p.bias != nullptr && (fix ? owner : sg.kb0 == 0);
      // a stream-K segment that covers the whole K range of its tile is the only contributor: plain stores into the zeroed C
      // (with the fused reduce-scatter every tile goes through the reductions: a plain store could not reach a peer's inbox)
      const bool whole_tile = sg.kb0 == 0 && sg.kb1 == p.kb_total && p.rs_world <= 1;
      const bool atomic_out = p.streamk == 1 && !whole_tile;
      // Workspace slot g = the partial accumulator of group g's (single) contributor segment, this warp's part being
      // [rank][quarter][chunk][8][32 lanes][4] floats in register order; flag word g counts the epilogue warps of group g
      // whose slot stores are visible, flag word ngroups + g the owner warps of group g that are done with their slots.
      const size_t nch_ws = static_cast<size_t>((p.bn + G2_CHUNK - 1) / G2_CHUNK);
      const size_t slot_floats = static_cast<size_t>(CG) * 4 * nch_ws * 1024;
      const size_t warp_off = ((static_cast<size_t>(rank) * 4 + q) * nch_ws) * 1024 + lane * 4;
      float* my_slot = contributor ? p.fix_ws + static_cast<size_t>(group) * slot_floats + warp_off : nullptr;
      int ncontrib = 0;
      mbar_wait(&tfull_bar[as], aphase);
      if (warp == 2 && lane == 0 && tile_iter == 0) G2_STAMP(7);   // first accumulator ready
      tc_fence_after();
      if (has_contrib) {
        // contributors = the groups directly in front of this one whose (non-empty) ranges reach into this tile
        const long long tile_start = sg.tile * p.kb_total;
        for (int g = group - 1; g >= 0 && it.bound(g + 1) > tile_start; --g) ++ncontrib;
        if (lane == 0) {
          const long long t0 = clock64();
          for (int c = 1; c <= ncontrib; ++c)
            while (ld_acquire_u32(p.fix_flags + (group - c)) < 4u * CG) {
              if (clock64() - t0 > 4000000000LL) { printf("fxn: gemm fix-up wait timeout (block %d)\n", blockIdx.x); __trap(); }
            }
        }
        __syncwarp();
      }
      float sq_acc = 0.f;
      const int ncols_tile = min(p.bn, p.N - n0);                        // valid columns of this tile (may be <= 0 never)
      int nchunks = (min(p.bn, padN - n0) + G2_CHUNK - 1) / G2_CHUNK;
      // a row quarter that lies entirely below the matrix (ragged M: 307 rows in 256-row pair tiles) has nothing to write;
      // per-tile statistics keep all four warps in step (they meet at a named barrier per chunk), fix-up slots are read whole
      if (rows_valid == 0 && p.stats_mode != 2 && !fix) nchunks = 0;
      if (half >= nchunks) {
        // this warp has no chunk in this tile: release its share of the accumulator right away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&tempty_bar[as], 0);
      }
      for (int ch = half; ch < nchunks; ch += CH_STEP) {
        const int c0 = ch * G2_CHUNK;
        uint32_t v[32];
        const uint32_t taddr = tmem_base + as * static_cast<uint32_t>(p.bn) + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(c0);
        if (p.bn - c0 >= 32) {
          tmem_ld_32x32(taddr, v);
        } else {
          tmem_ld_32x16(taddr, v);
#pragma unroll
          for (int j = 16; j < 32; ++j) v[j] = 0;
        }
        tmem_ld_wait();
        if (ch + CH_STEP >= nchunks) {
          // every accumulator value this warp needs is in registers: hand its share of the TMEM stage back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tempty_bar[as], 0);
        }
        if (contributor) {
          // raw partial accumulators -> this group's slot, in register order: one warp-level 16-byte store instruction
          // covers 512 contiguous bytes (no shared-memory staging, no atomics: every slot has exactly one writer)
          float* wp_ = my_slot + static_cast<size_t>(ch) * 1024;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcg(reinterpret_cast<float4*>(wp_ + j * 128),
                   make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3])));
          continue;
        }
        for (int c = 1; c <= ncontrib; ++c) {      // owner of a split tile: add the contributors' partials (fixed order)
          const float* wp_ = p.fix_ws + static_cast<size_t>(group - c) * slot_floats + warp_off + static_cast<size_t>(ch) * 1024;
          float4 pv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pv[j] = ldcg_f4(wp_ + j * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + pv[j].x);
            v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + pv[j].y);
            v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + pv[j].z);
            v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + pv[j].w);
          }
        }
        __syncwarp();
        // lane = row: 32 consecutive columns -> staging (row stride 36 floats)
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts_f4(stg_s + (lane * G2_STG_LD + j) * 4, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                 __uint_as_float(v[j + 3]));
        __syncwarp();
        // transposed passes: lane handles columns (col, col + 1) of rows 2i + rr
        const int col = n0 + c0 + 2 * cp;
        const bool ok0 = col < p.N, ok1 = col + 1 < p.N;
        float b0 = 0.f, b1 = 0.f, rb0 = 0.f, rb1 = 0.f;
        if (add_bias) { if (ok0) b0 = __ldg(p.bias + col); if (ok1) b1 = __ldg(p.bias + col + 1); }
        if (p.epi_act == 7) { rb0 = __ldg(p.gauss_rb + min(col, p.N - 1)); rb1 = __ldg(p.gauss_rb + min(col + 1, p.N - 1)); }
        float s0 = 0.f, s1 = 0.f;
        EpiRowCtx cx;
        cx.stg = stg_s + (rr * G2_STG_LD + 2 * cp) * 4;
        cx.mbase = mbase; cx.rr = rr; cx.col = col; cx.rows_valid = rows_valid; cx.ok0 = ok0; cx.ok1 = ok1;
        cx.alpha = alpha; cx.b0 = b0; cx.b1 = b1; cx.rb0 = rb0; cx.rb1 = rb1; cx.padN = padN;
        cx.c_vec2 = c_vec2; cx.x_vec2 = x_vec2; cx.atomic_out = atomic_out;
        // specialised loops need 32 valid, vector-aligned columns; rows may be ragged except where the loop fetches its
        // old C / reconstruction targets up front (variants 7 and 8)
        const bool full_rows = rows_valid == 32 || (rows_valid > 0 && p.epi_variant != 7 && p.epi_variant != 8);
        const bool interior = full_rows && n0 + c0 + G2_CHUNK <= p.N && (c_vec2 || p.C == nullptr) &&
                              (x_vec2 || p.mse_x == nullptr);
        int variant = interior ? p.epi_variant : -1;
        if (variant == 6) variant = whole_tile ? 0 : (c_vec4 ? 12 : 6);
        switch (variant) {
          //                         ACT CMODE PLANES STATS MSE
          case 0:  epi_rows<false, 0, 1, false, 0, false>(p, cx, s0, s1, sq_acc); break;   // plain fp32 output
          case 1:  epi_rows<false, 0, 1, false, 2, false>(p, cx, s0, s1, sq_acc); break;   // MLP layer_1 forward (+ BN partials)
          case 2:  epi_rows<false, 6, 1, false, 2, false>(p, cx, s0, s1, sq_acc); break;   // Encoder/Decoder hidden forward
          case 3:  epi_rows<false, 0, 1, true, 0, false>(p, cx, s0, s1, sq_acc); break;    // fp32 + planes
          case 4:  epi_rows<false, 0, 0, true, 0, false>(p, cx, s0, s1, sq_acc); break;    // planes only
          case 5:  epi_rows<false, 0, 0, true, 3, false>(p, cx, s0, s1, sq_acc); break;    // planes + column sums
          case 6:  epi_rows<false, 0, 3, false, 0, false>(p, cx, s0, s1, sq_acc); break;   // stream-K reductions
          case 7:  epi_rows<false, 0, 2, false, 0, false>(p, cx, s0, s1, sq_acc); break;   // C +=
          case 8:  epi_rows<false, 3, 0, true, 3, true>(p, cx, s0, s1, sq_acc); break;     // Decoder output + MSE
          case 9:  epi_rows<false, 7, 0, true, 3, false>(p, cx, s0, s1, sq_acc); break;    // Gaussian-kernel Gram
          case 10: epi_rows<false, 0, 1, true, 3, false>(p, cx, s0, s1, sq_acc); break;    // fp32 + planes + column sums
          case 11: epi_rows<false, 6, 1, false, 0, false>(p, cx, s0, s1, sq_acc); break;   // hidden forward, eval mode
          case 12: epi_red4(p, stg_s, lane, mbase, n0 + c0, alpha, add_bias, rows_valid); break;   // stream-K, vector reductions
          default: epi_rows<true, 0, 0, false, 0, false>(p, cx, s0, s1, sq_acc); break;    // edges / anything else
        }
        if (p.stats_mode != 0) {
          // column sums over this warp's 32 rows (combine the two row-parity halves)
          s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
          if (p.stats_mode == 3) {
            if (rr == 0) {
              if (ok0) atomicAdd(p.colstats + col, s0 * sscale);
              if (ok1) atomicAdd(p.colstats + col + 1, s1 * sscale);
            }
          } else {
            float m20 = 0.f, m21 = 0.f;
            if (p.stats_mode == 2 && rows_valid > 0) {
              const float mu0 = s0 / static_cast<float>(rows_valid), mu1 = s1 / static_cast<float>(rows_valid);
              __syncwarp();
              // all 16 staged values first (independent loads), then the arithmetic: issued one per iteration the loads cost
              // 16 exposed shared-memory latencies per chunk, which sets the pace of short-K launches ([B * N x 256] GCN layers)
              float2 xs[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) xs[i] = lds_f2(stg_s + ((2 * i + rr) * G2_STG_LD + 2 * cp) * 4);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (2 * i + rr < rows_valid) {
                  m20 = fmaf(xs[i].x - mu0, xs[i].x - mu0, m20);
                  m21 = fmaf(xs[i].y - mu1, xs[i].y - mu1, m21);
                }
              }
              m20 += __shfl_xor_sync(0xffffffffu, m20, 16);
              m21 += __shfl_xor_sync(0xffffffffu, m21, 16);
            }
            // merge the four 32-row quarters of this CTA's 128 rows (Chan) -> one partial per 128-row tile
            if (rr == 0) {
              s_stat[q][0][2 * cp] = s0; s_stat[q][0][2 * cp + 1] = s1;
              s_stat[q][1][2 * cp] = m20; s_stat[q][1][2 * cp + 1] = m21;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
            if (q == 0 && lane < 32) {
              const int c = n0 + c0 + lane;
              const int tile_row0 = (sg.mt * CG + static_cast<int>(rank)) * G2_BM;
              if (c < p.N && tile_row0 < p.M) {        // the second CTA of the last pair may own no valid row at all
                float tn = 0.f, tm = 0.f, tm2 = 0.f, tsum = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float nk = static_cast<float>(max(0, min(32, p.M - (tile_row0 + k * 32))));
                  if (nk > 0.f) {
                    const float sk = s_stat[k][0][lane];
                    const float delta = sk / nk - tm;
                    const float nn = tn + nk;
                    tm += delta * nk / nn;
                    tm2 += s_stat[k][1][lane] + delta * delta * tn * nk / nn;
                    tn = nn;
                    tsum += sk;
                  }
                }
                const long long t128 = tile_row0 / G2_BM;
                float* dst = p.colstats + (t128 * 2) * p.N + c;
                dst[0] = tsum;
                dst[p.N] = tm2;
              }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
          }
        }
        __syncwarp();
      }
      if (contributor) {
        __threadfence();                       // this lane's slot stores are visible before the flag is raised
        __syncwarp();
        if (lane == 0) atomicAdd(p.fix_flags + group, 1u);
      } else if (ncontrib > 0) {
        __syncwarp();
        if (lane == 0) {
          // the last owner warp to leave resets the flags it consumed for the next launch (every owner warp of this tile is
          // past its wait by then; a contributor group feeds exactly one owner)
          const unsigned old = atomicAdd(p.fix_flags + ngroups + group, 1u);
          if (old == 4u * CG - 1u) {
            for (int c = 1; c <= ncontrib; ++c) p.fix_flags[group - c] = 0u;
            p.fix_flags[ngroups + group] = 0u;
          }
        }
      }
      if (p.mse_x != nullptr && !contributor) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq_acc += __shfl_xor_sync(0xffffffffu, sq_acc, o);
        if (lane == 0) atomicAdd(p.mse_acc, sq_acc);
      }
      (void)ncols_tile;
      if (warp == 2 && lane == 0 && tile_iter == 0) G2_STAMP(8);   // first tile stored
      ++tile_iter;
    }
    if (warp == 2 && lane == 0) G2_STAMP(9);                       // all tiles stored
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();      // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == 1) {
    tc_fence_after();
    g2_tmem_dealloc<CG>(tmem_base, p.tmem_cols);
  }
  if (threadIdx.x == 0) G2_STAMP(10);
  if (p.trace && threadIdx.x == 0 && blockIdx.x < 512) g2_cta_time[1][blockIdx.x] = globaltimer_ns();
}

}  // namespace fxn

using namespace fxn;

namespace {

struct Plan { int cg, bn, stages, streamk, groups, tiles_m, tiles_n; };

// Pick CTA-group size, tile width and scheduling mode with a small analytic cost model (SM cycles). Constants from the
// pipeline traces (profiles/r02_bn_sweep_after.log): a k-block of a pair tile costs max(MMA time = nterms * 4
// instructions of bn / 2 cycles + ~90, operand ingest at ~58 B/clk per SM) -- 691 / 857 / 1057 / 1244 / 1604 cycles at
// bn = 64 / 128 / 160 / 192 / 256, i.e. MMA-paced from 128 columns up; the epilogue costs ~1450 cycles per 32-column
// chunk and only the last one of a CTA is exposed.
//   mode 0  whole tiles (any epilogue)
//   mode 1  stream-K, vector reductions into a zeroed plain fp32 C
//   mode 2  stream-K with fix-up through a workspace (any epilogue; needs fxn_gemm_desc.fix_ws)
// group_limit > 0 caps the number of CTA groups (the engine splits the SMs between GEMMs that run concurrently).
Plan make_plan(int M, int N, int K, int nterms, int b_mn, bool plain_c, int force_bn, bool fix_ok = false, int group_limit = 0,
               bool only_reductions = false) {
  const int nplanes = nterms == 3 ? 2 : 1;
  const int kb_total = (K + G2_BK - 1) / G2_BK;
  const int cg = M > G2_BM ? 2 : 1;
  const int unit = b_mn ? 64 * cg : 16 * cg;                   // per-CTA B rows: multiple of 64 (MN-major) or 16
  int max_groups = 148 / cg;
  if (group_limit > 0 && group_limit < max_groups) max_groups = group_limit;
  const int tiles_m = (M + G2_BM * cg - 1) / (G2_BM * cg);
  const double SETUP = 3500.0, EPI_CHUNK = 1450.0, SM_BW = 58.0, L2_BW = 11000.0, MEMSET = 4000.0, FIX = 2500.0;
  static const int force_sk = [] { const char* e = getenv("FXN_GEMM_FORCE_STREAMK"); return e ? atoi(e) : 0; }();
  static const int no_fix = [] { const char* e = getenv("FXN_GEMM_NO_FIXUP"); return e ? atoi(e) : 0; }();
  Plan best{};
  double best_t = 1e300;
  for (int bn = unit; bn <= 256; bn += unit) {
    if (force_bn > 0 && bn != (force_bn + unit - 1) / unit * unit && !(bn == 256 && force_bn > 256)) continue;
    const int tiles_n = (N + bn - 1) / bn;
    if (tiles_n > 1 && (tiles_n - 1) * bn >= N) continue;
    if (force_bn <= 0 && tiles_n > 1 && bn < 64) continue;     // narrow tiles only when one tile covers N
    const int stage_bytes = nplanes * (G2_BM * G2_BK * 2 + (bn / cg) * G2_BK * 2);
    const int budget = G2_MAX_DYN_SMEM - 1024 - G2_STG_BYTES;
    int stages = budget / stage_bytes;
    if (stages > G2_MAX_STAGES) stages = G2_MAX_STAGES;
    if (stages < 2) continue;
    const long long tiles = static_cast<long long>(tiles_m) * tiles_n;
    const double t_mma = nterms * 4.0 * (bn / 2.0) + 90.0;
    const double t_epi = EPI_CHUNK * ((bn + 31) / 32);
    for (int mode = 0; mode < 3; ++mode) {
      if (only_reductions && mode != 1) continue;              // fused reduce-scatter: stream-K reductions, whatever K
      if (mode == 1 && (!plain_c || (kb_total < 4 && !only_reductions))) continue;
      if (mode == 2 && (plain_c || !fix_ok || no_fix || kb_total < 8 || bn % 32 != 0)) continue;
      if (mode == 0 && force_sk && plain_c && kb_total >= 4) continue;
      int groups;
      double t;
      if (mode == 0) {
        groups = static_cast<int>(tiles < max_groups ? tiles : max_groups);
        const double t_load = fmax(stage_bytes / SM_BW, static_cast<double>(stage_bytes) * groups * cg / L2_BW);
        const double t_main = kb_total * fmax(t_mma, t_load) + (stages < 3 ? 0.25 * kb_total * t_load : 0.0);
        const long long tpg = (tiles + groups - 1) / groups;
        t = SETUP + t_main + (tpg - 1) * fmax(t_main, t_epi) + t_epi;
      } else {
        const long long units = tiles * kb_total;
        const long long g = units / (mode == 2 ? 4 : 2);     // every group gets several k-blocks
        groups = static_cast<int>(g < max_groups ? g : max_groups);
        if (groups < 1) continue;
        const double kb_per = static_cast<double>(units) / groups;
        const double t_load = fmax(stage_bytes / SM_BW, static_cast<double>(stage_bytes) * groups * cg / L2_BW);
        if (mode == 1) {
          const double segs = 1.0 + (kb_per < kb_total ? 1.0 : kb_per / kb_total);
          t = SETUP + MEMSET + kb_per * fmax(t_mma, t_load) + segs * 0.7 * t_epi;
        } else {
          // one exposed owner epilogue (reads its slot), the contributor's register-order reductions are short
          const double split = (tiles % groups == 0) ? 0.0 : 1.0;
          t = SETUP + kb_per * fmax(t_mma, t_load) + t_epi * (1.0 + 0.35 * split) + FIX * split;
        }
      }
      if (t < best_t) {
        best_t = t;
        best.cg = cg; best.bn = bn; best.stages = stages; best.streamk = mode; best.groups = groups;
        best.tiles_m = tiles_m; best.tiles_n = tiles_n;
      }
    }
  }
  return best;
}

template <int CG, int EW>
cudaError_t launch_gemm2(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi,
                         const CUtensorMap& tb_lo, const Gemm2Args& p, int groups, int smem_bytes, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<CG, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_MAX_DYN_SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(groups * CG, 1, 1);
  cfg.blockDim = dim3(64 + 32 * EW, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (p.trace) {
    int nclusters = -1;
    cudaOccupancyMaxActiveClusters(&nclusters, gemm2_kernel<CG, EW>, &cfg);
    fprintf(stderr, "[gemm2] max co-resident clusters of %d CTAs at %d B smem: %d (launching %d)\n", CG, smem_bytes, nclusters,
            groups);
  }
  return cudaLaunchKernelEx(&cfg, gemm2_kernel<CG, EW>, ta_hi, ta_lo, tb_hi, tb_lo, p);
}

}  // namespace

namespace fxn {

int gemm2_dispatch(const fxn_gemm_desc* d, cudaStream_t stream) {
  Gemm2Args p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.a_mn = d->a_mn_major ? 1 : 0;
  p.b_mn = d->b_mn_major ? 1 : 0;
  p.nterms = d->nterms;
  const bool plain_c = (d->splitk < 0 || d->splitk > 1) && d->C && !d->c_hi && !d->colstats && !d->epi_act && !d->accumulate && !d->mse_x;
  const bool fix_ok = d->fix_ws != nullptr && d->fix_flags != nullptr;
  const bool rs = d->rs_world > 1;
  if (rs) {
    if (!plain_c) return set_error(FXN_ERR_UNSUPPORTED, "fxn_gemm: the fused reduce-scatter needs a plain fp32 output with splitk < 0");
    if (d->rs_world > 8 || d->rs_rank < 0 || d->rs_rank >= d->rs_world || !d->rs_base || !d->rs_inbox || d->rs_per <= 0 ||
        d->rs_per % 4 || d->rs_per >= (1LL << 31))
      return set_error(FXN_ERR_ARG, "fxn_gemm: bad reduce-scatter arguments (2..8 ranks, slice a multiple of 4 elements)");
    const long long off0 = d->C - d->rs_base;
    if (off0 < 0 || off0 + static_cast<long long>(d->M - 1) * d->ldc + d->N >= (1LL << 31))
      return set_error(FXN_ERR_ARG, "fxn_gemm: C must lie in the first 2^31 elements of the reduce-scatter arena");
    for (int r = 0; r < d->rs_world; ++r) {
      if (r != d->rs_rank && !d->rs_inbox[r]) return set_error(FXN_ERR_ARG, "fxn_gemm: missing peer inbox");
      p.rs_inbox[r] = d->rs_inbox[r];
    }
    p.rs_base = d->rs_base;
    p.rs_per = static_cast<unsigned>(d->rs_per);
    p.rs_world = d->rs_world;
    p.rs_rank = d->rs_rank;
  }
  Plan pl = make_plan(d->M, d->N, d->K, d->nterms, p.b_mn, plain_c, d->block_n, fix_ok, d->max_groups, rs);
  if (pl.stages < 1) return set_error(FXN_ERR_ARG, "fxn_gemm: tile does not fit in shared memory");
  if (pl.streamk == 2) {
    // one slot per group: [cg][4][ceil(bn / 32)][1024] floats; two flag words per group
    const long long need = static_cast<long long>(pl.groups) * pl.cg * 4 * ((pl.bn + 31) / 32) * 1024 * sizeof(float);
    if (d->fix_ws_bytes < need || d->fix_flags_count < 2LL * pl.groups)
      pl = make_plan(d->M, d->N, d->K, d->nterms, p.b_mn, plain_c, d->block_n, false, d->max_groups);   // workspace too small
  }
  p.fix_ws = d->fix_ws;
  p.fix_flags = static_cast<unsigned*>(d->fix_flags);
  p.bn = pl.bn;
  p.stages = pl.stages;
  p.streamk = pl.streamk;
  p.tiles_m = pl.tiles_m; p.tiles_n = pl.tiles_n;
  p.kb_total = (d->K + G2_BK - 1) / G2_BK;
  static const int force_stages = [] { const char* e = getenv("FXN_GEMM_STAGES"); return e ? atoi(e) : 0; }();
  if (force_stages > 0 && force_stages < p.stages) p.stages = force_stages;      // experiments: pipeline depth
  if (p.stages > p.kb_total && !p.streamk && static_cast<long long>(pl.tiles_m) * pl.tiles_n <= pl.groups)
    p.stages = p.kb_total < 1 ? 1 : p.kb_total;                 // one tile per group: no need for more stages than k-blocks
  uint32_t tc = 32;
  while (tc < static_cast<uint32_t>(2 * p.bn)) tc <<= 1;
  p.tmem_cols = tc;
  p.C = d->C; p.ldc = d->ldc;
  p.bias = d->bias;
  p.c_hi = static_cast<__nv_bfloat16*>(d->c_hi);
  p.c_lo = static_cast<__nv_bfloat16*>(d->c_lo);
  p.ldp = d->ldp;
  p.colstats = d->colstats;
  p.stats_mode = d->colstats ? (d->stats_mode ? d->stats_mode : 2) : 0;
  if (p.c_hi && (!p.c_lo || d->ldp % 8 != 0 || (reinterpret_cast<uintptr_t>(d->c_hi) & 15) ||
                 (reinterpret_cast<uintptr_t>(d->c_lo) & 15)))
    return set_error(FXN_ERR_ARG, "fxn_gemm: output planes need both pointers, 16B alignment and ldp %% 8 == 0");
  p.epi_act = d->epi_act;
  p.accumulate = d->accumulate;
  p.alpha = d->alpha == 0.f ? 1.f : d->alpha;
  p.alpha_dev = d->alpha_dev;
  p.mse_x = d->mse_x; p.ldx = d->ldx; p.mse_acc = d->mse_acc;
  if (p.mse_x && (!p.mse_acc || (p.stats_mode && p.stats_mode != 3)))
    return set_error(FXN_ERR_ARG, "fxn_gemm: fused MSE needs mse_acc and excludes per-tile statistics");
  p.gauss_ra = d->gauss_ra; p.gauss_rb = d->gauss_rb; p.gauss_inv = d->gauss_inv;
  if (p.epi_act == 7 && (!p.gauss_ra || !p.gauss_rb)) return set_error(FXN_ERR_ARG, "fxn_gemm: epi_act 7 needs row norms");
  p.stats_alpha = d->stats_alpha == 0.f ? 1.f : d->stats_alpha;
  p.stats_alpha_dev = d->stats_alpha_dev;
  {
    const int cmode = !d->C ? 0 : (p.streamk == 1 ? 3 : (d->accumulate ? 2 : 1));
    const int act = p.epi_act, st = p.stats_mode;
    const bool pl_ = p.c_hi != nullptr, ms = p.mse_x != nullptr;
    struct V { int act, cmode, planes, stats, mse; };
    static const V table[12] = {{0, 1, 0, 0, 0}, {0, 1, 0, 2, 0}, {6, 1, 0, 2, 0}, {0, 1, 1, 0, 0}, {0, 0, 1, 0, 0}, {0, 0, 1, 3, 0},
                                {0, 3, 0, 0, 0}, {0, 2, 0, 0, 0}, {3, 0, 1, 3, 1}, {7, 0, 1, 3, 0}, {0, 1, 1, 3, 0}, {6, 1, 0, 0, 0}};
    p.epi_variant = -1;
    for (int i = 0; i < 12; ++i)
      if (table[i].act == act && table[i].cmode == cmode && table[i].planes == (pl_ ? 1 : 0) && table[i].stats == st &&
          table[i].mse == (ms ? 1 : 0))
        p.epi_variant = i;
    if (const char* e = getenv("FXN_GEMM_GENERIC_EPILOGUE")) if (atoi(e)) p.epi_variant = -1;
  }
  static const int trace = [] { const char* e = getenv("FXN_GEMM_TRACE"); return e ? atoi(e) : 0; }();
  p.trace = trace;
  if (trace)
    fprintf(stderr, "[gemm2] M=%d N=%d K=%d cg=%d bn=%d stages=%d groups=%d streamk=%d tiles=%dx%d\n", d->M, d->N, d->K, pl.cg,
            pl.bn, p.stages, pl.groups, pl.streamk, pl.tiles_m, pl.tiles_n);

  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  const long long a_rows = p.a_mn ? d->K : d->M, a_cols = p.a_mn ? d->M : d->K;
  const long long b_rows = p.b_mn ? d->K : d->N, b_cols = p.b_mn ? d->N : d->K;
  const int a_box = p.a_mn ? G2_BK : G2_BM, b_box = p.b_mn ? G2_BK : p.bn / pl.cg;
  if ((rc = make_tensor_map(&ta_hi, d->a_hi, a_rows, a_cols, d->lda, a_box))) return rc;
  if ((rc = make_tensor_map(&tb_hi, d->b_hi, b_rows, b_cols, d->ldb, b_box))) return rc;
  if (d->nterms == 3) {
    if ((rc = make_tensor_map(&ta_lo, d->a_lo, a_rows, a_cols, d->lda, a_box))) return rc;
    if ((rc = make_tensor_map(&tb_lo, d->b_lo, b_rows, b_cols, d->ldb, b_box))) return rc;
  } else {
    ta_lo = ta_hi;
    tb_lo = tb_hi;
  }
  if (p.stats_mode == 3 && !d->outputs_prezeroed) {
    cudaError_t e = cudaMemsetAsync(d->colstats, 0, sizeof(float) * d->N, stream);
    if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "colsum memset: %s", cudaGetErrorString(e));
  }
  if (p.streamk == 1 && !d->outputs_prezeroed) {
    cudaError_t e = cudaMemset2DAsync(d->C, d->ldc * sizeof(float), 0, d->N * sizeof(float), d->M, stream);
    if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "stream-K memset: %s", cudaGetErrorString(e));
  }
  const int nplanes = d->nterms == 3 ? 2 : 1;
  const int stage_bytes = nplanes * (G2_BM * G2_BK * 2 + (p.bn / pl.cg) * G2_BK * 2);
  // Epilogue width. FXN_GEMM_EPI_WARPS=8 (4 = never) gives short-K, epilogue-bound tile-mode problems eight
  // epilogue warps, provided the second set of staging buffers still leaves as many stages as there are k-blocks to
  // prefetch (or at least 3).
  // Default: eight warps when K <= 4 k-blocks -- such launches ARE their epilogue (GCN layers folded to [B * N / 8 x 256]:
  // 1389 -> 964 us, the flatten -> fc input gradient [4096 x 64000], K = 128: 577 -> 421 us; profiles/r02_timeline_cfg4_*).
  static const int epi_env = [] { const char* e = getenv("FXN_GEMM_EPI_WARPS"); return e ? atoi(e) : 0; }();
  int ew = 4;
  if (epi_env != 4 && !p.streamk && p.kb_total <= (epi_env == 8 ? 16 : 4)) {   // (the fix-up protocol counts four epilogue warps per CTA)
    int st8 = (G2_MAX_DYN_SMEM - 1024 - 2 * G2_STG_BYTES) / stage_bytes;
    if (st8 > p.stages) st8 = p.stages;
    if (st8 >= 3 || (st8 >= 2 && st8 >= p.kb_total)) { ew = 8; p.stages = st8; }
  }
  const int smem_bytes = p.stages * stage_bytes + (ew / 4) * G2_STG_BYTES + 1024;
  cudaError_t e;
  if (ew == 8)
    e = pl.cg == 2 ? launch_gemm2<2, 8>(ta_hi, ta_lo, tb_hi, tb_lo, p, pl.groups, smem_bytes, stream)
                   : launch_gemm2<1, 8>(ta_hi, ta_lo, tb_hi, tb_lo, p, pl.groups, smem_bytes, stream);
  else
    e = pl.cg == 2 ? launch_gemm2<2, 4>(ta_hi, ta_lo, tb_hi, tb_lo, p, pl.groups, smem_bytes, stream)
                   : launch_gemm2<1, 4>(ta_hi, ta_lo, tb_hi, tb_lo, p, pl.groups, smem_bytes, stream);
  if (e != cudaSuccess) return set_error(FXN_ERR_CUDA, "gemm2 launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

}  // namespace fxn

// Debug: clock64 stamps of CTA 0 and CTA 1 of the last traced launch (FXN_GEMM_TRACE=1), relative to each CTA's start.
// Host-side planner exposed for tests and tooling (no device needed): the plan fxn_gemm would use for a problem.
// out[0..7] = {cta_group, block_n, stages, streamk, groups (CTA pairs / CTAs), tiles_m, tiles_n, dynamic smem bytes}.
extern "C" int fxn_gemm_plan(int M, int N, int K, int nterms, int b_mn_major, int plain_c, int block_n, int* out8) {
  if (!out8 || M <= 0 || N <= 0 || K <= 0 || (nterms != 1 && nterms != 3))
    return fxn::set_error(FXN_ERR_ARG, "fxn_gemm_plan: bad argument");
  // plain_c: 0 = fused epilogue, whole tiles only; 1 = plain fp32 C (stream-K eligible); 2 = fused epilogue with a fix-up
  // workspace (stream-K with fix-up eligible). block_n may carry a group limit in its upper half: block_n | limit << 16.
  const int limit = block_n >> 16;
  block_n &= 0xFFFF;
  const Plan pl = make_plan(M, N, K, nterms, b_mn_major ? 1 : 0, plain_c == 1, block_n, plain_c == 2, limit);
  if (pl.stages < 1) return fxn::set_error(FXN_ERR_ARG, "fxn_gemm_plan: tile does not fit in shared memory");
  const int nplanes = nterms == 3 ? 2 : 1;
  const int stage_bytes = nplanes * (G2_BM * G2_BK * 2 + (pl.bn / pl.cg) * G2_BK * 2);
  out8[0] = pl.cg; out8[1] = pl.bn; out8[2] = pl.stages; out8[3] = pl.streamk; out8[4] = pl.groups;
  out8[5] = pl.tiles_m; out8[6] = pl.tiles_n; out8[7] = pl.stages * stage_bytes + G2_STG_BYTES + 1024;
  return 0;
}

// Debug: %globaltimer (ns) at start and end of CTA i of the last traced launch, relative to the earliest start
// (-1 = CTA index not launched).
extern "C" int fxn_debug_gemm_cta_times(long long* start_ns, long long* end_ns, int n) {
  static unsigned long long h[2][512];
  cudaError_t e = cudaMemcpyFromSymbol(h, fxn::g2_cta_time, sizeof(h));
  if (e != cudaSuccess) return fxn::set_error(FXN_ERR_CUDA, "trace copy: %s", cudaGetErrorString(e));
  if (n > 512) n = 512;
  unsigned long long t0 = ~0ull;
  for (int i = 0; i < n; ++i) if (h[0][i] != 0 && h[0][i] < t0) t0 = h[0][i];
  for (int i = 0; i < n; ++i) {
    start_ns[i] = h[0][i] ? static_cast<long long>(h[0][i] - t0) : -1;
    end_ns[i] = h[0][i] ? static_cast<long long>(h[1][i] - t0) : -1;
  }
  return 0;
}

extern "C" int fxn_debug_gemm_trace(long long* out32) {
  long long h[2][16];
  cudaError_t e = cudaMemcpyFromSymbol(h, fxn::g2_trace, sizeof(h));
  if (e != cudaSuccess) return fxn::set_error(FXN_ERR_CUDA, "trace copy: %s", cudaGetErrorString(e));
  for (int c = 0; c < 2; ++c)
    for (int i = 0; i < 16; ++i) out32[c * 16 + i] = h[c][i] ? h[c][i] - h[c][0] : -1;
  return 0;
}

extern "C" int fxn_gemm(const fxn_gemm_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d) return fxn::set_error(FXN_ERR_ARG, "null descriptor");
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return fxn::set_error(FXN_ERR_ARG, "fxn_gemm: M,N,K must be positive");
  if (d->nterms != 1 && d->nterms != 3) return fxn::set_error(FXN_ERR_ARG, "fxn_gemm: nterms must be 1 or 3");
  if (!d->a_hi || !d->b_hi || (d->nterms == 3 && (!d->a_lo || !d->b_lo)))
    return fxn::set_error(FXN_ERR_ARG, "fxn_gemm: missing operand plane");
  if (!d->C && !d->c_hi) return fxn::set_error(FXN_ERR_ARG, "fxn_gemm: no output");
  return fxn::gemm2_dispatch(d, stream);
}

extern "C" int fxn_gemm_stat_tiles(int M) { return (M + fxn::G2_BM - 1) / fxn::G2_BM; }

// Bytes of fix-up workspace / number of 32-bit flag words that cover every plan fxn_gemm can choose.
extern "C" long long fxn_gemm_fix_ws_bytes(void) { return 74LL * 2 * 4 * 8 * 1024 * static_cast<long long>(sizeof(float)); }
extern "C" int fxn_gemm_fix_flag_words(void) { return 2 * 148; }
