// Micro-benchmark: how fast can one SM (and the whole chip) ingest TMA tiles into shared memory?
// Each CTA runs a producer thread issuing `boxes` TMA loads of {64 bf16, rows} per stage into an S-stage ring and a consumer
// thread that frees a stage `delay` cycles after it landed (standing in for the MMAs that read it). Prints bytes/clk/SM,
// GB/s per SM and chip-wide for a sweep of stage sizes, depths, CTA counts and source footprints (L2-resident / HBM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_ingest_bench tools/tma_ingest_bench.cu -lcuda
#include "../flexynesis_b200/csrc/ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace fxn;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(64, 1)
ingest_kernel(const __grid_constant__ CUtensorMap tm, int rows_per_box, int boxes, int stages, int kblocks, int delay,
              int row_tiles, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[16], empty_bar[16];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t box_bytes = rows_per_box * 128u, stage_bytes = box_bytes * boxes;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  // each CTA walks its own row tile (like an A operand), k-blocks along the columns
  const int tile = blockIdx.x % row_tiles;
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
      for (int b = 0; b < boxes; ++b)
        tma_load_2d(smem + stage * stage_bytes + b * box_bytes, &tm, &full_bar[stage], kb * 64,
                    (tile * boxes + b) * rows_per_box);
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      if (delay > 0) { const long long t = clock64(); while (clock64() - t < delay) {} }
      mbar_arrive(&empty_bar[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
}

int main(int argc, char** argv) {
  PFN_encodeTiled enc = nullptr;
  {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    enc = reinterpret_cast<PFN_encodeTiled>(p);
  }
  cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  long long* d_cycles; cudaMalloc(&d_cycles, 4096 * 8);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device clock attr %d kHz\n", clk_khz);
  // source matrices: big (HBM: 16384 x 8192 bf16 = 256 MB) and small (L2: 2048 x 8192 = 32 MB)
  struct Src { const char* name; long long rows, cols; } srcs[2] = {{"HBM 1.2GB", 75776, 8192}, {"L2 32MB", 2048, 8192}};
  for (const Src& src : srcs) {
    __nv_bfloat16* base; cudaMalloc(&base, src.rows * src.cols * 2); cudaMemset(base, 0, src.rows * src.cols * 2);
    struct Cfg { int rows_per_box, boxes, stages, grid, delay; };
    std::vector<Cfg> cfgs;
    const int grids[4] = {1, 32, 64, 148};
    for (int g : grids) {
      cfgs.push_back({128, 4, 3, g, 0});    // 64 KB stages x3 (bn=256 pair tile)
      cfgs.push_back({128, 3, 4, g, 0});    // 48 KB x4 (bn=128)
      cfgs.push_back({128, 2, 6, g, 0});    // 32 KB x6
      cfgs.push_back({128, 1, 12, g, 0});   // 16 KB x12
      cfgs.push_back({64, 1, 16, g, 0});    // 8 KB x16 (only 128 KB in flight)
      cfgs.push_back({128, 4, 3, g, 1536}); // with the MMA time of a 256-wide tile holding each stage
      cfgs.push_back({128, 2, 6, g, 768});
    }
    for (const Cfg& c : cfgs) {
      CUtensorMap tm;
      cuuint64_t gdim[2] = {(cuuint64_t)src.cols, (cuuint64_t)src.rows};
      cuuint64_t gstr[1] = {(cuuint64_t)src.cols * 2};
      cuuint32_t box[2] = {64u, (cuuint32_t)c.rows_per_box};
      cuuint32_t estr[2] = {1u, 1u};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const int stage_bytes = c.rows_per_box * 128 * c.boxes;
      const int smem = stage_bytes * c.stages + 1024;
      const int kblocks = 128;
      const int row_tiles = (int)(src.rows / (c.rows_per_box * c.boxes));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int it = 0; it < 2; ++it) ingest_kernel<<<c.grid, 64, smem>>>(tm, c.rows_per_box, c.boxes, c.stages, kblocks, c.delay, row_tiles, d_cycles);
      cudaEventRecord(e0);
      ingest_kernel<<<c.grid, 64, smem>>>(tm, c.rows_per_box, c.boxes, c.stages, kblocks, c.delay, row_tiles, d_cycles);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> cyc(c.grid);
      cudaMemcpy(cyc.data(), d_cycles, c.grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (long long v : cyc) mx = v > mx ? v : mx;
      const double bytes = (double)stage_bytes * kblocks;
      printf("%-10s stage %3d KB x%2d grid %3d delay %4d : %7.0f cyc/kb  %5.1f B/clk/SM  %6.1f GB/s/SM  chip %6.2f TB/s  (%.1f us, %.2f GHz)\n",
             src.name, stage_bytes / 1024, c.stages, c.grid, c.delay, (double)mx / kblocks, bytes / mx,
             bytes / (ms * 1e-3) / 1e9, bytes * c.grid / (ms * 1e-3) / 1e12, ms * 1e3, mx / (ms * 1e-3) / 1e9);
    }
    cudaFree(base);
  }
  return 0;
}
