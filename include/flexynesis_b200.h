/* flexynesis_b200 -- C ABI of the B200-native engine for flexynesis's training hot path.
 *
 * The reference (BIMSBbioinfo/flexynesis) has no FFI: its hot path is Python on top of torch.nn
 * (SURVEY.md section 8b). This header is therefore the *new* boundary that the Python drop-in classes
 * (flexynesis_b200.DirectPred / supervised_vae / MultiTripletNetwork / GNN) bind with ctypes. Each entry
 * point names the reference arithmetic it replaces (file:line under the reference tree).
 *
 * Conventions
 *  - plain C symbols, plain pointers and sizes, no C++/torch types;
 *  - every pointer is a DEVICE pointer unless its name ends in _host; buffers are caller-allocated;
 *  - every call is asynchronous on the caller-supplied CUDA stream (void* = cudaStream_t);
 *  - return 0 on success, <0 on error; fxn_last_error() gives the message (thread-local);
 *  - matrices are row-major; "ld" is the row stride in elements;
 *  - "planes" are the engine's operand format for tensor-core GEMMs: an fp32 matrix x stored as two
 *    bf16 matrices (hi, lo), x = hi + lo up to 2^-17 relative, same ld for both, ld % 8 == 0,
 *    16-byte aligned bases.
 */
#ifndef FLEXYNESIS_B200_H
#define FLEXYNESIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FXN_OK 0
#define FXN_ERR_ARG (-1)
#define FXN_ERR_CUDA (-2)
#define FXN_ERR_UNSUPPORTED (-3)

/* ---- library ---- */
int fxn_version(void);
const char* fxn_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long fxn_launch_count(void);
void fxn_reset_launch_count(void);

/* ---- operand planes ---- */
/* planes(hi,lo)[r, c] = split(src[r, c]); columns [cols, ld_planes) are zero-filled. */
int fxn_split_planes(const float* src, long long ld_src, long long rows, long long cols, void* hi, void* lo,
                     long long ld_planes, void* stream);

/* ---- dense contraction on tcgen05 tensor cores ----
 * C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]).
 * Replaces aten::addmm / aten::mm behind nn.Linear forward and backward:
 *   flexynesis/modules.py:145,149 (MLP.layer_1 / layer_out), :54-56 (Encoder), :101-102 (Decoder),
 *   :261 (flexGCN.fc), flexynesis/models/direct_pred.py:124-128 (fusion_block), and their autograd duals.
 * Operand storage: K-major operand = row-major [MN x K]; MN-major operand = row-major [K x MN].
 */
typedef struct fxn_gemm_desc {
  int M, N, K;
  const void* a_hi; const void* a_lo; long long lda; int a_mn_major;
  const void* b_hi; const void* b_lo; long long ldb; int b_mn_major;
  int nterms;              /* 3: hi*hi + hi*lo + lo*hi (fp32-grade); 1: hi*hi only (bf16-grade) */
  float* C; long long ldc; /* fp32 result, may be NULL if only planes are wanted */
  const float* bias;       /* [N] or NULL */
  void* c_hi; void* c_lo; long long ldp; /* optional planes of the result */
  float* colstats;         /* optional [fxn_gemm_stat_tiles(M)][2][N]: per 128-row tile (sum, M2 about tile mean) */
  int stats_mode;          /* 0/2: sum + M2, 1: sum only (M2 slot written as 0) */
  int splitk;              /* >1: split K over blockIdx.z, fp32 atomics into C (C is zeroed by the call) */
  int block_n;             /* 0 = auto */
} fxn_gemm_desc;
int fxn_gemm(const fxn_gemm_desc* d, void* stream);
int fxn_gemm_stat_tiles(int M);

#ifdef __cplusplus
}
#endif
#endif /* FLEXYNESIS_B200_H */
