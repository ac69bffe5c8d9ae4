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
/* planes(hi,lo)[r, c] = split(src[r, c]); columns [cols, pad8(cols)) are zero-filled. */
int fxn_split_planes(const float* src, long long ld_src, long long rows, long long cols, void* hi, void* lo,
                     long long ld_planes, void* stream);

/* Batch feeder: row gather out[b,:] = src[idx[b],:] (idx int64 device array, NULL = identity) written as fp32
 * and/or planes. Replaces per-sample MultiOmicDataset.__getitem__ + default_collate (flexynesis/data.py:980-995)
 * for a dataset resident in HBM. */
int fxn_gather_rows(const float* src, long long ld_src, const long long* idx, long long nrows, long long cols,
                    float* out, long long ldo, void* hi, void* lo, long long ldp, void* stream);

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
  int stats_mode;          /* 0/2: sum + M2; 1: sum only; 3: colstats is an [N] vector (zeroed by the call) receiving
                              plain column sums by atomics (bias gradients) */
  int splitk;              /* >1: split K over blockIdx.z, fp32 atomics into C (C is zeroed by the call); <0: auto */
  int block_n;             /* 0 = auto */
  int epi_act;             /* applied after alpha and bias: 0 none, 1 relu, 3 sigmoid, 6 leaky_relu(0.2), 7 Gaussian kernel */
  int accumulate;          /* C += result instead of C = result */
  float alpha;             /* result scale (0 means 1) ... */
  const float* alpha_dev;  /* ... times *alpha_dev when not NULL (device scalar, e.g. a loss weight) */
  /* fused Decoder output + reconstruction loss (flexynesis/modules.py:101-102 + supervised_vae.py:549):
   * with epi_act = 3 the staged tile is x_hat = sigmoid(.); if mse_x != NULL the epilogue adds sum (x_hat - x)^2
   * to *mse_acc and the planes (c_hi, c_lo) receive G = (x_hat - x) * x_hat * (1 - x_hat) -- the gradient of the
   * squared error w.r.t. the pre-sigmoid output up to the constant 2/(M*N). C (if given) still receives x_hat. */
  const float* mse_x; long long ldx; float* mse_acc;
  /* epi_act = 7: Gaussian kernel of compute_kernel (supervised_vae.py:494-513) from the Gram GEMM:
   * out[m,n] = exp(-max(ra[m] + rb[n] - 2 * acc[m,n], 0) * gauss_inv), ra / rb = squared row norms of A / B. */
  const float* gauss_ra; const float* gauss_rb; float gauss_inv;
  /* stats_mode 3 column sums are multiplied by stats_alpha (0 means 1) * *stats_alpha_dev (if not NULL) */
  float stats_alpha; const float* stats_alpha_dev;
  /* the caller has already zeroed every accumulation target of this call (a stream-K / split-K C and the stats_mode 3
   * column sums): the library then queues no memset in front of the kernel (the engine zeroes its whole gradient arena
   * once per step, beside the forward pass) */
  int outputs_prezeroed;
  /* Scheduling. max_groups > 0 caps the number of CTA groups (pairs of SMs) the launch may occupy: the engine splits the
   * chip between GEMMs it runs concurrently. fix_ws / fix_flags (optional; fxn_gemm_fix_ws_bytes() bytes and
   * fxn_gemm_fix_flag_words() 32-bit words; the flags zeroed ONCE by the caller; both private to launches that cannot
   * overlap in time) allow stream-K for problems with a fused epilogue: partial accumulators of a split tile travel through
   * the workspace and the group holding the tile's last k-block runs the epilogue; the library leaves the flags zeroed. */
  int max_groups;
  float* fix_ws; long long fix_ws_bytes;
  void* fix_flags; long long fix_flags_count;
  /* Data-parallel weight gradients (rs_world = 2..8; 0 or 1 = off; no reference counterpart, pl.Trainer(devices=1),
   * flexynesis/main.py:223): the reduce-scatter of the gradient is fused into the GEMM. C must be a plain fp32 stream-K
   * output (splitk < 0) inside this rank's gradient arena, which starts at rs_base and is cut into rs_world slices of
   * rs_per elements (a multiple of 4; the last slice takes the remainder). Reductions for elements of this rank's own
   * slice go to C as usual; those for another rank's slice are issued over NVLink into rs_inbox[owner] + (element offset
   * from rs_base), the owner's peer-mapped inbox arena (HOST array of rs_world device pointers; [rs_rank] is ignored). The
   * owner's total for its slice is C + inbox once every rank's GEMM has finished (fxn_dp_reduce_sumsq adds them). */
  int rs_world, rs_rank;
  long long rs_per;
  const float* rs_base;
  float* const* rs_inbox;
} fxn_gemm_desc;
int fxn_gemm(const fxn_gemm_desc* d, void* stream);
/* Debug aid: with FXN_GEMM_TRACE=1 in the environment the persistent kernel stamps clock64 at its pipeline milestones
 * for CTA 0 and CTA 1; this copies the 2 x 16 stamps (cycles since CTA start, -1 = not reached) of the last launch. */
int fxn_debug_gemm_trace(long long* out32);
/* Same mode: %globaltimer (ns, relative to the earliest CTA start) at entry and exit of CTAs 0..n-1 (n <= 512) of the
 * last traced launch -- shows scheduling waves and stragglers. */
int fxn_debug_gemm_cta_times(long long* start_ns, long long* end_ns, int n);
int fxn_gemm_stat_tiles(int M);
long long fxn_gemm_fix_ws_bytes(void);
int fxn_gemm_fix_flag_words(void);
/* The launch plan fxn_gemm would choose (host-side cost model, no device needed): out8 = {cta_group, block_n, stages,
 * streamk, groups, tiles_m, tiles_n, dynamic shared memory bytes}. plain_c = 1: the output is a plain fp32 C (stream-K
 * eligible); plain_c = 2: fused epilogue with a fix-up workspace (stream-K with fix-up eligible, streamk = 2 in the plan);
 * block_n & 0xFFFF > 0 forces the tile width as fxn_gemm_desc.block_n does, block_n >> 16 > 0 caps the groups as
 * fxn_gemm_desc.max_groups does. */
int fxn_gemm_plan(int M, int N, int K, int nterms, int b_mn_major, int plain_c, int block_n, int* out8);

/* ---- BatchNorm1d (+ activation + dropout) ----
 * Forward of  y = dropout(act(BN(V)))  over the rows of V [rows x cols].
 * Replaces aten::native_batch_norm + relu + dropout of MLP.forward (flexynesis/modules.py:146-148), the
 * BatchNorm1d of Encoder/Decoder.hidden_layers (modules.py:28, :78; their LeakyReLU runs in the producing
 * fxn_gemm via epi_act = 6) and flexGCN's bn -> act -> dropout (modules.py:255-257).
 * train != 0: batch statistics from `partials` ([ntiles][2][cols], tile_rows rows per tile, written by
 *   fxn_gemm(colstats) or fxn_col_stats), running stats updated with `momentum` (unbiased variance),
 *   *num_batches_tracked += 1, (mean, rstd) stored in `saved` [2][cols] for the backward pass.
 *   rows < 2 is an error, as in torch ("Expected more than 1 value per channel when training").
 * train == 0: running statistics, no dropout.
 * act: 0 none, 1 relu, 2 leaky_relu(0.01), 3 sigmoid, 4 tanh, 5 gelu.
 * Dropout keep-mask: `mask` (uint8 [rows x cols], ld = ldm) when given, else Philox(seed, element index).
 */
typedef struct fxn_bn_fwd_desc {
  const float* V; long long ldv; long long rows; int cols;
  const float* partials; int ntiles; int tile_rows;
  int partials_ld;                           /* columns of the partials array (0 = cols); lets V be a column window */
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; void* num_batches_tracked; /* int64 */
  float momentum; float eps;
  int train; int act; float p_drop;
  const uint8_t* mask; long long ldm; unsigned long long seed;
  const void* seed_dev;                      /* optional device int64 mixed into the seed (per-step counter) */
  float* out; long long ldo;                 /* optional fp32 output */
  void* out_hi; void* out_lo; long long ldp; /* optional planes of the output */
  float* saved;
  long long stat_rows;                       /* rows the batch statistics cover (0 = rows). Larger than `rows` when the
                                                partials were gathered from several ranks (global-batch BatchNorm): the
                                                tiles then describe stat_rows rows, this call normalises its own `rows` */
  uint8_t* keep_bits;                        /* optional [rows * ceil(cols / 8)] bytes: the dropout keep flags of each run of 8
                                                columns as drawn, so that the backward pass (fxn_bn_bwd_desc.keep_bits) reads
                                                one byte instead of re-evaluating Philox twice; 1 bit per activation */
} fxn_bn_fwd_desc;
int fxn_bn_act_fwd(const fxn_bn_fwd_desc* d, void* stream);

/* Backward of the same block: given dOut (gradient w.r.t. dropout(act(BN(V)))) produce dgamma, dbeta and dV
 * (fp32 and/or planes). pre_act != 0: V is leaky_relu_0.2(Z) of the Linear output Z (Encoder/Decoder order); the
 * result is multiplied by the LeakyReLU derivative so it is dZ, and dbias [cols] receives its column sum.
 * `sums` is scratch [2][cols]. Autograd dual of the forward ops listed above. */
typedef struct fxn_bn_bwd_desc {
  const float* V; long long ldv; const float* dOut; long long ldg; long long rows; int cols;
  const float* gamma; const float* beta; const float* saved;
  int act; float p_drop; const uint8_t* mask; long long ldm; unsigned long long seed;
  const void* seed_dev;
  int pre_act;
  float* sums; float* dgamma; float* dbeta; float* dbias;
  float* dV; long long ldd;
  void* dv_hi; void* dv_lo; long long ldp;
  float grad_scale;      /* multiplies dOut; 0 means 1 */
  int accumulate_affine; /* dgamma/dbeta += instead of = (a module applied several times per step) */
  long long stat_rows;   /* rows the forward statistics covered (0 = rows); see fxn_bn_fwd_desc */
  int prezeroed;         /* `sums` and `dbias` are already zero: queue no memsets */
  int phase;             /* 0: whole backward. 1: only the column reductions into `sums` (sum g, sum g*xhat over this
                            call's rows). 2: only the apply pass, reading `sums` as given -- a data-parallel caller
                            sum-all-reduces `sums` between phase 1 and phase 2 (SyncBN backward) */
  const uint8_t* keep_bits; /* optional: the keep flags the forward call stored (replaces mask / Philox) */
} fxn_bn_bwd_desc;
int fxn_bn_act_bwd(const fxn_bn_bwd_desc* d, void* stream);

/* Column statistics partials of V for inputs that do not come from an fxn_gemm epilogue. */
int fxn_col_stats(const float* V, long long ldv, long long rows, int cols, int tile_rows, float* partials, void* stream);

/* ---- supervisor heads and losses ----
 * logits[rows x C] = D[rows x sh] * W[C x sh]^T (+ bias). kind: 0 none (Cox risk score), 1 MSE over rows with
 * non-NaN y, 2 cross-entropy over rows with y != -1 and non-NaN. acc[0] += sum of row losses, acc[1] += valid rows
 * (caller zeroes acc). Replaces MLP.layer_out (modules.py:149) + compute_loss (direct_pred.py:146-190). */
int fxn_head_out_fwd(const float* D, long long ldd, int rows, int sh, const float* W, const float* bias, int C,
                     float* logits, long long ldl, int kind, const float* y, float* acc, void* stream);
/* Backward: dlogits from (kind, y, acc[1], *weight) or from Cox coefficients (kind 3), then dD[rows x sh] (stored),
 * dW[C x sh] and dbias[C] (accumulated with atomics; zeroed by the call unless `prezeroed`: the caller cleared them). */
int fxn_head_out_bwd(const float* D, long long ldd, int rows, int sh, const float* W, int C, const float* logits,
                     long long ldl, int kind, const float* y, const float* acc, const float* coef,
                     const float* weight, float* dD, long long ldg, float* dW, float* dbias, int prezeroed, void* stream);
/* Cox partial-likelihood loss of risk scores o[n] (stride ldo) -- cox_ph_loss, modules.py:265-305: rows with NaN
 * duration/event dropped, sorted by duration descending, loss = -(sum_{e=1} o_i - log cumsum exp(o))/sum e, 0 when
 * empty or non-finite. acc[0] = loss, acc[1] = 1; coef[n] = d loss / d o. */
int fxn_cox_fwd(const float* o, long long ldo, const float* durations, const float* events, int n, float* coef,
                float* acc, void* stream);
int fxn_cox_max_rows(void);
/* The same loss and coefficients on the whole chip and without a row limit: pairwise passes instead of the single-CTA
 * sort + scans (risk set of row i = rows j with t_j > t_i, or t_j == t_i and j <= i, which is the sorted order the
 * reference's descending argsort + cumsum sees with ties kept in row order). workspace: fxn_cox_ws_floats(n) floats,
 * 8-byte aligned, zeroed by the call. Sums are fp32 atomics (order varies between runs at the 1e-7 level). */
long long fxn_cox_ws_floats(int n);
int fxn_cox_fwd_ws(const float* o, long long ldo, const float* durations, const float* events, int n, float* coef,
                   float* acc, float* workspace, void* stream);
/* compute_total_loss (direct_pred.py:192-223). acc [n][2]; kinds[n] (device): 1 = mean of (sum, count), 3 = value in
 * acc[k][0]. out: [0,n) losses, [n] total, [n+1] unweighted sum (validation objective, :290), [n+2, 2n+2) weights
 * d total / d loss_k. With weighting and n > 1, *dlog_vars[k] = 1 - exp(-s_k) * loss_k. Pointer tables are device arrays. */
int fxn_total_loss(int n, const float* acc, const int* kinds, const float* const* log_vars, float* const* dlog_vars,
                   int weighting, float* out, void* stream);
/* ---- the whole supervisor-head section in three launches (narrow heads: nv * pad8(sh) <= 64, C <= 16, MSE / CE) ----
 * All target variables' MLPs (flexynesis/modules.py:135-150; the loop flexynesis/models/direct_pred.py:131-132) with
 * their losses (direct_pred.py:146-190) and the autograd dual, on CUDA cores.
 * fxn_heads_fwd: Zh = F W1cat^T + b1 -> BatchNorm1d (batch statistics merged from 16-row partials; running statistics
 *   updated) -> ReLU -> Dropout(p_drop) -> layer_out -> logits; with labels: acc[slot] += (sum of row losses, valid
 *   rows); with `backward`: G = d(weighted total loss)/d(BatchNorm output), sums += column sums of (G, G * xhat),
 *   d layer_out.weight / bias accumulated (caller zeroes sums and those gradients).
 * fxn_heads_bwd: dZh = gamma rstd (G - mean G - xhat mean(G xhat)) -> planes (dz_hi, dz_lo) for the layer_1 weight
 *   gradient; dF = dZh W1cat -> planes (df_hi, df_lo), optional fp32 dF, dbias[L] += column sums of dF (caller zeroes);
 *   d gamma / d beta written. */
#define FXN_HEADS_MAX_VARS 8
typedef struct fxn_heads_var {
  int kind;                 /* 1 MSE over rows with non-NaN y, 2 cross-entropy over rows with y != -1 and non-NaN */
  int C;                    /* outputs of layer_out */
  int slot;                 /* row of the loss table */
  const float* W1; const float* b1;                     /* layer_1 [sh x L], [sh] */
  const float* gamma; const float* beta;                /* batchnorm affine [sh] */
  float* running_mean; float* running_var; void* num_batches_tracked;
  const float* Wout; const float* bout;                 /* layer_out [C x sh], [C] or NULL */
  const float* y;                                       /* labels [B] or NULL (no loss for this variable) */
  float* logits;                                        /* [B x C] */
  const uint8_t* mask; long long ldm;                   /* optional explicit dropout keep mask [B x sh] */
  unsigned long long seed;                              /* Philox seed of this head's dropout */
  const float* log_var;                                 /* device scalar s_k: loss weight exp(-s_k); NULL = 1 */
  float* dWout; float* dbout; float* dgamma; float* dbeta;
} fxn_heads_var;
typedef struct fxn_heads_desc {
  int B, L, sh, nv;
  const float* F; long long ldf;            /* input of the heads (fused embedding), fp32 [B x L] */
  float* Zh; long long ldz;                 /* [B x nv * pad8(sh)] layer_1 outputs, kept for the backward pass */
  float* G; long long ldg;                  /* same shape */
  float* partials;                          /* [2][nv * pad8(sh)][ceil(B / 16) rounded up to 32] */
  float* saved;                             /* [2][nv * pad8(sh)] mean, rstd */
  float* sums;                              /* [2][nv * pad8(sh)] */
  float* acc;                               /* loss table [n][2] (caller zeroes) */
  int train; float p_drop; float momentum; float eps;
  const void* seed_dev;                     /* optional device int64 mixed into the dropout seeds (per-step counter) */
  int backward;
  fxn_heads_var var[FXN_HEADS_MAX_VARS];
  void* dz_hi; void* dz_lo; long long ldzp;
  void* df_hi; void* df_lo; long long ldfp;
  float* dF; long long lddf;
  float* dbias; int zero_dbias;             /* zero_dbias != 0: the call clears dbias[L] first */
} fxn_heads_desc;
int fxn_heads_fused_ok(int L, int sh, int nv, int maxC);
int fxn_heads_fwd(const fxn_heads_desc* d, void* stream);
int fxn_heads_bwd(const fxn_heads_desc* d, void* stream);

/* triplet_loss (triplet_encoder.py:178-194): mean_b relu(|a-p|^2 - |a-n|^2 + margin); acc as above. */
int fxn_triplet_fwd(const float* A, const float* P, const float* N, long long ld, int rows, int L, float margin,
                    float* rowloss, float* acc, void* stream);
int fxn_triplet_bwd(const float* A, const float* P, const float* N, long long ld, int rows, int L,
                    const float* rowloss, const float* weight, float* dA, float* dP, float* dN, long long ldg,
                    int accumulate_a, void* stream);

/* ---- supervised_vae latent + MMD (flexynesis/models/supervised_vae.py) ----
 * reparameterization (:187-200): z = mean + s * eps  (s is the raw FC_log_var output, no exp). z as fp32 + planes. */
int fxn_reparam_fwd(const float* mean, const float* s, const float* eps, long long ld, long long rows, int cols,
                    float* z, void* z_hi, void* z_lo, long long ldp, void* stream);
/* dmean = dz, ds = dz * eps as planes; their column sums (bias gradients of FC_mean / FC_log_var) into dbias_mean /
 * dbias_s (zeroed by the call). */
int fxn_reparam_bwd(const float* dz, const float* eps, long long ld, long long rows, int cols, void* dm_hi, void* dm_lo,
                    void* ds_hi, void* ds_lo, long long ldp, float* dbias_mean, float* dbias_s, void* stream);
/* Standard normal draws out[r, c], c < cols (epsilon of reparameterization, torch.randn_like, :198; the MMD prior
 * torch.randn(200, latent), :545): Philox4x32-10 + Box-Muller keyed by (seed, *seed_dev, element). */
int fxn_randn(float* out, long long ld, long long rows, int cols, unsigned long long seed, const void* seed_dev,
              void* stream);
/* out[r] = sum_c X[r,c]^2 */
int fxn_row_sqnorm(const float* X, long long ld, long long rows, int cols, float* out, void* stream);
/* MMD_loss (:532-550) assembled from column sums of the three Gaussian kernel matrices and the fused reconstruction
 * error: loss_i = sum(cs_tt_i)/P^2 + sum(cs_zz)/B^2 - 2 sum(cs_tz_i)/(P B) + mse_acc[i]/(B d_i); acc[0] = mean_i loss_i,
 * acc[1] = 1. cs_tt [n][P], cs_tz [n][B], dims [n] (device int32). */
int fxn_mmd_finish(const float* cs_zz, const float* cs_tt, const float* cs_tz, const float* mse_acc, const int* dims,
                   int nlayers, int B, int P, float* acc, void* stream);
/* dz += w * ( -(4/(B^2 L^2)) (z * cs_zz - KZ) + (1/n) sum_i -(4/(P B L^2)) (KT_i - z * cs_tz_i) ),  w = *weight.
 * KZ [B x L] = K(z,z) Z, KT [n][B x L] = K(t_i,z)^T T_i. */
int fxn_mmd_grad(const float* z, long long ldz, const float* cs_zz, const float* KZ, const float* cs_tz,
                 const float* KT, long long ldk, int nlayers, int B, int L, int P, const float* weight, float* dz,
                 long long ldd, void* stream);
/* wts[k] = weighting && n > 1 ? exp(-*log_vars[k]) : 1 -- the d total / d loss_k factors, available before the
 * forward pass (they depend on parameters only). */
int fxn_loss_weights(int n, const float* const* log_vars, int weighting, float* wts, void* stream);

/* ---- GCN layer of flexGCN (flexynesis/modules.py:252-257; torch_geometric.nn.GCNConv as called at :221-226, :254) ----
 * Batched dense node features X [B, N, Fin] (contiguous fp32), one shared graph given as CSR by DESTINATION node:
 * rowptr[N+1], col[nnz] = source node of each in-edge, w[nnz] = symmetric-normalised weight deg_u^-1/2 deg_v^-1/2 with the
 * missing self loops added (gcn_norm). O[b, v, :] = W * (sum_e w_e X[b, col_e, :]) + bias, W [emb x Fin] = lin.weight.
 * partials (optional) [B][2][emb]: per-sample column statistics (sum, M2 about the sample mean) of O for the
 * BatchNorm1d over B*N rows that follows. Fin, emb <= 32. */
int fxn_gcn_fwd(const float* X, int B, int N, int Fin, const int* rowptr, const int* col, const float* w, const float* W,
                const float* bias, int emb, float* O, float* partials, void* stream);
/* Backward: dW [emb x Fin] and dbias [emb] (zeroed by the call), and, when dX != NULL, dX [B, N, Fin] through the
 * transposed graph (CSR by SOURCE node: rowptr_out, col_out = destination of each out-edge, w_out). */
int fxn_gcn_bwd(const float* X, const float* dO, int B, int N, int Fin, int emb, const int* rowptr_in, const int* col_in,
                const float* w_in, const int* rowptr_out, const int* col_out, const float* w_out, const float* W,
                float* dW, float* dbias, float* dX, void* stream);
/* Root-weight term of torch_geometric's GraphConv / SAGEConv as flexGCN calls them (flexynesis/modules.py:221-226, :254):
 *   GraphConv: lin_rel(sum_{u->v} x_u) + lin_root(x_v) ;  SAGEConv: lin_l(mean_{u->v} x_u) + lin_r(x_v).
 * The neighbour term runs through fxn_gcn_fwd / fxn_gcn_bwd with edge weights 1 (GraphConv) or 1/in-degree (SAGEConv) and
 * no added self loops; these two calls add the per-node term with Wr = lin_root.weight / lin_r.weight [emb x Fin]:
 *   fwd: O[b, v, :] += Wr X[b, v, :], then `partials` (optional, [B][2][emb]) = per-sample (sum, M2) of the finished O;
 *   bwd: dWr = sum_{b,v} dO[b,v]^T X[b,v] (zeroed by the call) and, when dX != NULL, dX[b, v, :] += Wr^T dO[b, v, :]. */
int fxn_node_lin_fwd(const float* X, int B, int N, int Fin, const float* Wr, int emb, float* O, float* partials, void* stream);
int fxn_node_lin_bwd(const float* X, const float* dO, int B, int N, int Fin, int emb, const float* Wr, float* dWr, float* dX,
                     void* stream);
/* Chan-merge [ntiles][2][pld] column partials (tile_rows rows per tile) into one record merged[2][cols] = (sum, M2),
 * which fxn_bn_act_fwd accepts as partials with ntiles = 1, tile_rows = rows. */
int fxn_merge_col_stats(const float* partials, int ntiles, int tile_rows, long long rows, int cols, int pld,
                        float* merged, void* stream);

/* Pure neighbour aggregation out[b, v, :] = sum_{e in CSR row v} w_e * in[b, col_e, :] over [B, N, C] fp32 node features
 * (C % 16 == 0, N * 64 bytes within shared memory: fxn_graph_gather_ok). Output as fp32 (`out`) or as bf16 operand planes
 * (out_hi / out_lo, ld = C). With it a GCNConv layer (torch_geometric; flexynesis/modules.py:221-226, :254) runs as
 * G = A^ X (gather) ; O = G W^T + b (fxn_gemm) and its backward as T = dO W (fxn_gemm) ; dX = A^T T (gather over the CSR by
 * source) ; dW = dO^T G (fxn_gemm): the per-node linear maps are tensor-core GEMMs over B * N rows.
 * `order` (optional, may be NULL): a permutation of the N nodes, e.g. by decreasing CSR row length; nodes are processed in
 * that order so that the nodes a warp advances together have similar degrees. The result does not depend on it. */
int fxn_graph_gather_ok(int N, int C);
int fxn_graph_gather(const float* in, int B, int N, int C, const int* rowptr, const int* col, const float* w,
                     const int* order, float* out, void* out_hi, void* out_lo, void* stream);
/* fxn_merge_col_stats for tens of thousands of tiles (the 128-row tile partials of a [B * N x C] fxn_gemm output):
 * two parallel passes with double-precision atomics; scratch = 2 * cols doubles (zeroed by the call). fold > 1: the GEMM
 * ran on the matrix viewed as [rows x fold * cols] (fold consecutive nodes per GEMM row, block-diagonal weights), so
 * partials has fold * cols columns and column q * cols + c belongs to channel c; `rows` counts the folded rows. */
int fxn_merge_col_stats_big(const float* partials, int ntiles, int tile_rows, long long rows, int cols, int pld, int fold,
                            float* merged, double* scratch, void* stream);

/* ---- step policy ----
 * clip_grad_norm_(params, max_norm) + Adam on flat arenas (flexynesis/main.py:216-217, direct_pred.py:135-144).
 * grads are multiplied by grad_scale (1/world_size after a sum all-reduce) before the norm. *step_counter (int64)
 * is incremented by the call; norm_out (optional) receives the pre-clip norm. sumsq_scratch: 16 bytes (a double and a
 * 32-bit counter), zeroed ONCE by the caller; the call hands it back zeroed (no memset is queued per step). */
int fxn_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                       float beta1, float beta2, float eps, float max_norm, float grad_scale, double* sumsq_scratch,
                       long long* step_counter, float* norm_out, void* stream);
/* Data-parallel step over NVSwitch multicast (no reference counterpart: pl.Trainer(devices=1), flexynesis/main.py:223).
 * mc_* are MULTICAST addresses mapping the same-offset arenas of all ranks (torch symmetric memory). Rank `rank` owns the
 * arena slice [begin, end) (multiples of 4 elements).
 * fxn_dp_reduce_sumsq: grad_local[slice] = scale * sum over ranks (multimem.ld_reduce), its squared norm is stored into
 *   slot `rank` of the symmetric partials array on every rank (multimem.st); *step_counter += 1. scratch16: 16 zeroed bytes.
 * -- a system-wide barrier belongs here --
 * fxn_dp_adam_bcast: global-norm clip + Adam on the slice, new parameters multicast into every rank's arena.
 * Barriers: with mc_flags != NULL each call starts with a barrier between the ranks INSIDE its kernel (before touching any
 * peer data): mc_flags / local_flags as for fxn_dp_barrier, epoch48 a local device uint32[48] zeroed once ([0..15] epochs,
 * [16..47] scratch), `slot` a flag slot no other call site uses. With mc_flags == NULL the caller provides the barriers.
 * peer_grads (HOST array of npeers <= 8 device pointers, rank order, may be NULL): every rank's gradient arena as mapped
 * into this process; when given the W copies are fetched with plain peer loads and added in rank order instead of
 * multimem.ld_reduce on mc_grad.
 * inbox / fused_ranges (HOST array of nranges <= 8 pairs [lo, hi) of element offsets, 4-aligned; may be NULL / 0): ranges of
 * the arena whose reduce-scatter already happened inside the weight-gradient GEMMs (fxn_gemm_desc.rs_*): there the slice's
 * sum is grad_local + inbox (both local), and the inbox is cleared for the next step. */
int fxn_dp_reduce_sumsq(const void* mc_grad, float* grad_local, long long begin, long long end, float scale, void* mc_partials,
                        int rank, void* scratch16, long long* step_counter, void* mc_flags, const void* local_flags,
                        void* epoch48, int slot, int world, const float* const* peer_grads, int npeers, float* inbox,
                        const long long* fused_ranges, int nranges, void* stream);
int fxn_dp_adam_bcast(void* mc_param, const float* param_local, const float* grad_local, float* exp_avg, float* exp_avg_sq,
                      long long begin, long long end, const float* partials, int world, float lr, float beta1, float beta2,
                      float eps, float max_norm, const long long* step_counter, float* norm_out, void* mc_flags,
                      const void* local_flags, void* epoch48, int slot, void* stream);
/* Barrier between the ranks of a data-parallel job, executed in stream order by one device thread per rank: mc_flags is
 * the multicast address of a symmetric uint32[16] array (zeroed once), local_flags this rank's copy, epoch a local device
 * uint32[16] (zeroed once). Every rank must issue the same sequence of barriers. Capturable in a CUDA graph. */
int fxn_dp_barrier(void* mc_flags, const void* local_flags, void* epoch, int slot, int world, void* stream);

/* Refresh the operand planes of many weight matrices in one launch. segments_dev: device array of nseg records
 * {int64 src_off, rows, cols, ld_src, dst_off, ldp} (element offsets into src / the plane arenas). */
int fxn_split_planes_multi(const float* src, const void* segments_dev, int nseg, long long max_seg_elems, void* hi,
                           void* lo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLEXYNESIS_B200_H */
