/*
 * mellon_b200 — C ABI of the B200-native sparse-GP density hot path.
 *
 * The reference (settylab/Mellon v1.7.1) is pure Python/JAX and has no FFI of its own;
 * its boundary for this path is a set of Python call sites.  Every entry point below
 * names the reference call site(s) it replaces (paths relative to
 * /root/reference/mellon).  INTEGRATION.md shows the ctypes stub a Mellon maintainer
 * would add at each of those sites.
 *
 * Conventions
 *   - plain C: pointers, sizes, doubles.  No torch / CUDA types in any signature.
 *   - all matrices are float64, row-major, C-contiguous (ld == cols); a vector is an
 *     (n, 1) matrix.
 *   - `mb_mat` is an opaque handle to a device-resident matrix owned by a context.
 *   - host pointers are caller-owned; the library copies (pinned staging inside).
 *   - return code: 0 ok; <0 error (text via mb_last_error()); >0 only from the
 *     Cholesky entry points = LAPACK-style `info` (1-based index of the first
 *     non-positive pivot) — the host raises the reference's ValueError
 *     (decomposition.py:116-122) on it.
 *   - one context drives ONE GPU; one process per GPU.  With a communicator attached
 *     (mb_comm_init) the cell axis is sharded across ranks: each rank holds the block of
 *     rows of x / K_NM / L that mb_row_block assigns to it and marks those matrices with
 *     mb_mat_set_shard.  The entry points marked [cell sum] add over the cells of ALL
 *     ranks when their operand is so marked (and over the local rows only when it is
 *     not: a replicated matrix is never summed twice), and every rank sees identical bits.
 *   - [cell sum] results do not depend on the number of ranks: the global cell axis is cut
 *     into 32 chunks, a chunk is summed in an order that depends on the chunk alone, and
 *     the 32 chunk sums are combined by one fixed pairwise tree (within a rank and across
 *     ranks alike).  mb_row_block puts rank boundaries on chunk boundaries.
 */
#ifndef MELLON_B200_H
#define MELLON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mb_ctx mb_ctx;
typedef struct mb_mat mb_mat;

/* ---- covariance program ------------------------------------------------------------
 * A covariance expression (cov.py kernels combined with base_cov.py Add/Mul/Pow and
 * nested active_dims) is flattened on the host into a postfix program.  Leaves carry
 * ABSOLUTE column indices into the input matrix (nested active_dims already composed,
 * base_cov.py:310-315, util.py:150-171). */
enum mb_kernel_kind {          /* leaf kinds: cov.py                                  */
  MB_K_MATERN32 = 0,           /* cov.py:62-66    r=sqrt(3) d/ls ; (r+1) e^-r          */
  MB_K_MATERN52 = 1,           /* cov.py:157-161  r=sqrt(5) d/ls ; (r+r^2/3+1) e^-r    */
  MB_K_EXPQUAD = 2,            /* cov.py:255-259  exp(-(d/ls)^2/2)                     */
  MB_K_EXPONENTIAL = 3,        /* cov.py:352-356  exp(-(d/ls)/2)                       */
  MB_K_RATQUAD = 4,            /* cov.py:453-457  ((d/ls)^2/(2a)+1)^-a                 */
  MB_K_LINEAR = 5,             /* cov.py:551-556  x.y/ls                               */
  MB_K_DISTANCE = 6            /* util.py:351-366 the bare distance (ls unused)        */
};
enum mb_kop_code {
  MB_OP_LEAF = 0,              /* push leaf value                                      */
  MB_OP_CONST = 1,             /* push scalar `value`   (scalar right operand)         */
  MB_OP_ADD = 2,               /* base_cov.py:309-315                                  */
  MB_OP_MUL = 3,               /* base_cov.py:375-381                                  */
  MB_OP_POW = 4                /* base_cov.py:449-453  top ** value                    */
};
typedef struct mb_kop {
  int32_t op;                  /* mb_kop_code                                          */
  int32_t kind;                /* mb_kernel_kind (LEAF only)                           */
  double ls;                   /* length scale (LEAF)                                  */
  double alpha;                /* RatQuad alpha (LEAF)                                 */
  double value;                /* CONST value / POW exponent                           */
  int32_t dim_off;             /* LEAF: offset into mb_kprog.dims                      */
  int32_t dim_cnt;             /* LEAF: number of active columns; -1 = all columns     */
} mb_kop;
typedef struct mb_kprog {
  int32_t n_ops;
  int32_t n_dims;
  const mb_kop* ops;
  const int32_t* dims;
} mb_kprog;

#define MB_MAX_LEAVES 4
#define MB_MAX_OPS 16
#define MB_STACK_DEPTH 4

/* ---- context / errors ---------------------------------------------------------------- */
const char* mb_last_error(void);
int mb_version(void);
/* first 16 hex digits of the sha256 over csrc/ and this header the library was compiled from */
const char* mb_source_hash(void);
int mb_device_count(int* n);
int mb_ctx_create(int device, mb_ctx** out);
int mb_ctx_destroy(mb_ctx* ctx);
int mb_ctx_sync(mb_ctx* ctx);
int mb_ctx_info(mb_ctx* ctx, int* device, int* n_sm, int64_t* free_bytes, int64_t* total_bytes);
/* number of this library's kernels launched through the context since creation */
int64_t mb_ctx_launch_count(mb_ctx* ctx);
/* device-side stopwatch on the context's stream (CUDA events): start / stop -> ms.
 * `slot` in [0, 16) lets several intervals be open at once. */
int mb_timer_start(mb_ctx* ctx, int slot);
int mb_timer_stop(mb_ctx* ctx, int slot, double* ms);
/* per-kernel-class stopwatch: while enabled, every launch of a class is bracketed by CUDA
 * events on the context's stream; mb_prof_read returns the launch count and summed device
 * time since the last reset, plus the ALGORITHMIC work of those launches (`work`: bytes for the
 * HBM-bound classes, flops for the GEMM class).  Classes: 0 = K1 covariance build (x != y),
 * 1 = K7 covariance mat-vec, 2 = FP64 GEMM tiles (K3/K4/K2 updates), 3 = K5/K6 fused objective
 * pass, 4 = everything else that is timed (the symmetric landmark covariance K_MM), 5 = int8 digit-slice GEMMs
 * on tcgen05 (work: float64-equivalent flops), 6 = the cuSOLVER Dsyevd library call of the Nystroem path. */
int mb_prof_enable(mb_ctx* ctx, int on);
int mb_prof_reset(mb_ctx* ctx);
int mb_prof_read(mb_ctx* ctx, int cls, int64_t* count, double* ms, double* work);
/* NVTX ranges (nvtx3, no cost unless a profiler is attached): every entry point of this library opens a range named
 * "mellon_b200: <stage>"; these two let the host side bracket its own stages (the L-BFGS-B loop of
 * inference.py:272-288, prepare_inference) so that one timeline shows Lp / L / Gram / L-BFGS-B / transform. */
int mb_range_push(const char* name);
int mb_range_pop(void);
/* overwrite a scratch buffer larger than L2 (cache flush between timed iterations) */
int mb_flush_l2(mb_ctx* ctx);
/* select kernel variants for A/B measurements and parity tests (0 is always the default path):
 *   "cov"      1 = DFMA register-tile covariance kernel, 2 = general (multi-leaf) kernel
 *   "gemm"     1 = DFMA reference GEMM, 2 = 8-warp DMMA tiles, 3 = no split-k, 4 = generic operand loaders
 *   "trsm"     1 = 32-wide substitution leaves for TRSM / Cholesky (no inverted 128-blocks, no blocked TRSV)
 *   "lossgrad" 1 = two-pass objective, 2 = register-fused single pass (0 = bulk-TMA ring)
 *   "cov_i8"   0 = K1 always on the FP64 DMMA kernel; 1 (default) = one exponential-family leaf with D <= 64, >= 4096
 *              cells and >= 256 landmarks on the tcgen05 kind::i8 digit-slice kernel; 2 = at every size (tests)
 *   "i8_issuers" MMA-issuing warps of the int8 GEMM kernels: 1, 2 or 4 (default 4)
 *   "i8_overlap" 1 = the digit pack of the next slab / chunk runs on a side stream under the int8 GEMM of the current
 *              one (same kernels, same bits; measured gain 0-2 %: both go through the L1 / shared-memory port);
 *              0 (default) = in sequence on the library stream
 *   "graph"    0 = launch the Cholesky on the stream instead of replaying its CUDA graph
 *   "i8"       0 = every FP64 product on the DMMA tiles; 1 (default) = the large products on tcgen05 kind::i8 digit
 *              slices: Gram matrices with chunks >= 2048 cells and r >= 512, TRSM updates / tall GEMMs with >= 8192
 *              cells, k >= 256 and >= 128 output columns; 2 = int8 slices at every size (tests, sanitizer runs) */
int mb_set_option(mb_ctx* ctx, const char* key, int value);

/* pinned (page-locked) host buffers, so uploads / the streaming predictor overlap with compute */
int mb_host_alloc(int64_t bytes, void** out);
int mb_host_free(void* p);

/* ---- communicator (NCCL over NVLink, one rank per process/GPU) ----------------------- */
int mb_comm_unique_id(unsigned char* out128);
int mb_comm_init(mb_ctx* ctx, const unsigned char* id128, int rank, int world);
int mb_comm_destroy(mb_ctx* ctx);
int mb_comm_info(mb_ctx* ctx, int* rank, int* world);
/* on != 0: until switched off this rank works ALONE although a communicator is attached — matrices marked as rows
 * [0, n) of n are summed locally, nothing is exchanged.  Reproduces the one-GPU computation inside a multi-GPU job
 * (bench.py and tools/check_multi_gpu.py compare its bits with the sharded result). */
int mb_comm_solo(mb_ctx* ctx, int on);
/* Row block [*row_lo, *row_hi) of rank `rank` of `world` for a cell axis of `global_rows` rows: whole chunks of
 * *chunk_rows = ceil(global_rows / 32) rows, chunks [32 rank / world, 32 (rank + 1) / world).  Pure function. */
int mb_row_block(int64_t global_rows, int rank, int world, int64_t* row_lo, int64_t* row_hi, int64_t* chunk_rows);
/* Mark `m` as rows [row_lo, row_lo + rows) of a matrix of `global_rows` rows whose cell axis is sharded over the
 * ranks (global_rows < 0 clears the mark).  With a communicator attached the block must be the one mb_row_block
 * gives this rank. */
int mb_mat_set_shard(mb_mat* m, int64_t global_rows, int64_t row_lo);
/* sum a device matrix over ranks in place (no-op without a communicator); plain NCCL all-reduce, rank-count
 * dependent rounding: not used on the fit path any more (the [cell sum] entry points reduce through the fixed tree) */
int mb_comm_allreduce(mb_ctx* ctx, mb_mat* a);
/* gather equally-sized row blocks: out(world*rows, cols) <- a(rows, cols) of each rank */
int mb_comm_allgather(mb_ctx* ctx, const mb_mat* a, mb_mat* out);

/* ---- device matrices ------------------------------------------------------------------ */
int mb_mat_alloc(mb_ctx* ctx, int64_t rows, int64_t cols, mb_mat** out);
int mb_mat_free(mb_ctx* ctx, mb_mat* m);
int mb_mat_shape(const mb_mat* m, int64_t* rows, int64_t* cols);
/* copy `nrows` full rows starting at device row `row0` from / to a host buffer */
int mb_mat_upload(mb_ctx* ctx, mb_mat* m, const double* host, int64_t row0, int64_t nrows);
int mb_mat_download(mb_ctx* ctx, const mb_mat* m, double* host, int64_t row0, int64_t nrows);
int mb_mat_copy(mb_ctx* ctx, const mb_mat* src, mb_mat* dst);
int mb_mat_fill(mb_ctx* ctx, mb_mat* m, double v);
/* dst(cols, rows) <- src(rows, cols)^T */
int mb_mat_transpose(mb_ctx* ctx, const mb_mat* src, mb_mat* dst);
/* A += v * I                      util.py:269-293 (stabilize / add_diagonal) */
int mb_mat_add_diag(mb_ctx* ctx, mb_mat* a, double v);
/* A(i, i) += v(i)                 conditional.py:245,320,347 (`K + sigma_g**2 * eye(n)` with one noise level per
 * observation: the (n, p) sigma form of FunctionEstimator) */
int mb_mat_add_diag_vec(mb_ctx* ctx, mb_mat* a, const mb_mat* v);
/* a(i, j) *= s(j)                 decomposition.py:265 (`* sqrt(S)`), inference.py:372 */
int mb_mat_scale_cols(mb_ctx* ctx, mb_mat* a, const mb_mat* s);
/* A(i, :) *= s(i)  (local rows)   conditional.py:531-533 (`A / sigma2` with one sigma per observation, on the
 * transposed layout this library keeps) */
int mb_mat_scale_rows(mb_ctx* ctx, mb_mat* a, const mb_mat* s);
/* dst <- src[:, c0:c0+ncols]      decomposition.py:75-76 (`v[:, -p:]`) */
int mb_mat_copy_cols(mb_ctx* ctx, const mb_mat* src, int64_t c0, int64_t ncols, mb_mat* dst);
/* dst <- src[r0:r0+nrows, :] */
int mb_mat_copy_rows(mb_ctx* ctx, const mb_mat* src, int64_t r0, int64_t nrows, mb_mat* dst);
/* copy the lower triangle onto the upper one (symmetrise a lower-only result) */
int mb_mat_symmetrize(mb_ctx* ctx, mb_mat* a);
/* a *= s                          conditional.py:139-181 (`A / sigma2`), :296-300 */
int mb_mat_scale(mb_ctx* ctx, mb_mat* a, double s);
/* a <- a + b | a * b (op = MB_OP_ADD | MB_OP_MUL; b == NULL: the scalar `value` instead of b) or a ** value
 * (MB_OP_POW), elementwise: base_cov.py:309-315, 375-381, 449-453 for expressions too large for ONE covariance
 * program (more than MB_MAX_LEAVES leaves): the sub-expressions are built by mb_cov_build and combined here, on
 * the device. */
int mb_mat_combine(mb_ctx* ctx, int op, mb_mat* a, const mb_mat* b, double value);
/* out(i) = sum_j a(i, j)^2        conditional.py:417,436,712,731,941,959 (`arraysum(square(A), axis=0)`
 * on the transposed layout this library keeps) */
int mb_mat_row_sumsq(mb_ctx* ctx, const mb_mat* a, mb_mat* out);

/* ---- K1: fused pairwise distance + covariance kernel -----------------------------------
 * K(i, j) = prog(x_i, y_j); replaces util.py:351-366 (distance) + cov.py `k` methods +
 * base_cov.py Add/Mul/Pow.k, i.e. every `cov_func(x, y)` call on the path
 * (decomposition.py:114,199,255-256; conditional.py:237,514,370,655,903). */
int mb_cov_build(mb_ctx* ctx, const mb_kprog* prog, const mb_mat* x, const mb_mat* y, mb_mat* K);
/* diag(i) = prog(x_i, x_i)        base_cov.py:71-93 */
int mb_cov_diag(mb_ctx* ctx, const mb_kprog* prog, const mb_mat* x, mb_mat* out);

/* ---- exact nearest neighbour (the step before the path) ------------------
 * dist(i) = min_{j != i + self_offset} |x_i - all_j| and its index: brute force over the same
 * distance tiles as K1 with a running-minimum epilogue, then the selected pair's distance is
 * recomputed as sqrt(sum (x - y)^2).  Replaces parameters.py:352-433 (pynndescent k=1). */
int mb_nn_distances(mb_ctx* ctx, const mb_mat* x, const mb_mat* all, int64_t self_offset, mb_mat* dist,
                    int64_t* idx_host);

/* ---- k-means landmarks (the step before the path) --------------------------------------------
 * One seeding step of k-means++ as scikit-learn's k_means runs it (sklearn/cluster/_kmeans.py:_kmeans_plusplus, reached
 * from parameters.py:291): for the T <= 16 candidate rows `cand` (T x d) and every row of x,
 *   out(t, i) = min(closest(i), max(|x_i|^2 - 2 x_i . c_t + |c_t|^2, 0))     (closest == NULL: no minimum)
 *   pot(t)    = sum_i out(t, i)                                               (fixed order, returned on the host)
 * xnorm holds |x_i|^2.  The Lloyd assignment step is mb_nn_distances(x, centres, self_offset < -rows). */
int mb_sqdist_min(mb_ctx* ctx, const mb_mat* x, const mb_mat* xnorm, const mb_mat* cand, const mb_mat* closest,
                  mb_mat* out, double* pot_host);

/* ---- K7: fused covariance + mat-vec (never materialises K) -----------------------------
 * out = mu + prog(xq, base) @ w ; w is (m, p), out is (nq, p).
 * Replaces `_mean` of the three conditionals: conditional.py:366-373, 651-658, 899-906. */
int mb_cov_matvec(mb_ctx* ctx, const mb_kprog* prog, const mb_mat* xq, const mb_mat* base,
                  const mb_mat* w, double mu, mb_mat* out);
/* same with HOST queries / results, streamed through the device in double-buffered
 * chunks (base_predictor.py:180-257 `Predictor.mean` body). */
int mb_predict_mean(mb_ctx* ctx, const mb_kprog* prog, const double* xq_host, int64_t nq,
                    int64_t d, const mb_mat* base, const mb_mat* w, double mu, double* out_host);

/* ---- K2 / K3: Cholesky and triangular solves -------------------------------------------- */
/* in-place lower Cholesky of a symmetric matrix (lower triangle read; strict upper
 * zeroed).  Returns info > 0 on a non-positive pivot.  decomposition.py:115,
 * conditional.py:63,73. */
int mb_potrf(mb_ctx* ctx, mb_mat* a);
/* Lp = chol(prog(xu, xu) + diag_add * I)      decomposition.py:79-123 (`_full_rank`) */
int mb_cov_chol(mb_ctx* ctx, const mb_kprog* prog, const mb_mat* xu, double diag_add, mb_mat* Lp);
/* X <- X Lp^-T  (X is n x m, rows independent)  decomposition.py:209
 * `solve_triangular(Lp, C.T, lower=True).T`, conditional.py:519 */
int mb_trsm_right_lt(mb_ctx* ctx, const mb_mat* Lp, mb_mat* X);
/* B <- Lp^-1 B (trans=0) or Lp^-T B (trans=1), B is (m, nrhs)
 * conditional.py:64-65,264,818 ; sklearn Ridge posv (parameters.py:896). */
int mb_tri_solve(mb_ctx* ctx, const mb_mat* Lp, int trans, mb_mat* B);
/* L = prog(x, xu) Lp^-T in one call   decomposition.py:174-210 (`_standard_low_rank`) */
int mb_lowrank_standard(mb_ctx* ctx, const mb_kprog* prog, const mb_mat* x, const mb_mat* xu,
                        const mb_mat* Lp, mb_mat* L);

/* ---- K4: Gram contraction over the cell axis ------------------------------------------- */
/* G = L^T L (r x r, full symmetric)  [cell sum]
 * parameters.py:896 (Ridge normal equations), conditional.py:61 (A A^T),
 * decomposition.py:259 (QR of K_NM via its Gram). */
int mb_gram(mb_ctx* ctx, const mb_mat* L, mb_mat* G);
/* b = L^T t (r x 1)  [cell sum]   parameters.py:896, conditional.py:64 (`dot(A, r_l)`) */
int mb_gemv_t(mb_ctx* ctx, const mb_mat* L, const mb_mat* t, mb_mat* b);
/* z0 = (L^T L + I)^-1 L^T t  [cell sums inside]   parameters.py:877-896 */
int mb_ridge_init(mb_ctx* ctx, const mb_mat* L, const mb_mat* t, double* z0_host);
/* C = alpha op(A) op(B) + beta C ; op = transpose when the flag is 1.  With trans_a = 1, trans_b = 0 and A, B marked
 * as row blocks of sharded matrices the product contracts over the cells: [cell sum] (alpha = 1, beta = 0 only;
 * conditional.py:61-64 `dot(A, y)` on the transposed layout).  Otherwise local. */
int mb_gemm(mb_ctx* ctx, int trans_a, int trans_b, double alpha, const mb_mat* A,
            const mb_mat* B, double beta, mb_mat* C);

/* ---- K5 / K6: MAP objective ---------------------------------------------------------------
 * One fused pass over the local rows of L:
 *   f = L z + mu ; A = exp(f + V)
 *   loss = 1/2 |z|^2 + (k/2) log 2pi - (sum_i (f_i - A_i) + sum_vdr)
 *   grad = z + L^T (A - 1)                                     [cell sum of r+1 doubles]
 * inference.py:35-48, 51-69, 72-92, 167-192 + jax.value_and_grad (inference.py:285).
 * `V` is the per-cell vector of inference.py:84; `sum_vdr` the global sum of :85. */
int mb_loss_grad(mb_ctx* ctx, const mb_mat* L, const mb_mat* V, double sum_vdr, double mu,
                 double k, const double* z_host, double* loss, double* grad_host);
/* f = L z + mu for the local rows    inference.py:66-67, 341-354 */
int mb_transform(mb_ctx* ctx, const mb_mat* L, const double* z_host, double mu, double* f_host);
/* diag(I + L^T diag(A) L)  [cell sum]   inference.py:311-317 in closed form */
int mb_hess_diag(mb_ctx* ctx, const mb_mat* L, const mb_mat* V, double mu, const double* z_host,
                 double* diag_host);

/* ---- symmetric eigen-decomposition (Nystroem paths) -------------------------------------
 * a <- eigenvectors (columns, ascending eigenvalues in w).  decomposition.py:50
 * (`eigh`).  LIBRARY CALL: cuSOLVER Dsyevd (dlopen'ed), the one device routine of this
 * library that is not hand-written; see DESIGN.md. */
int mb_syevd(mb_ctx* ctx, mb_mat* a, mb_mat* w);

#ifdef __cplusplus
}
#endif
#endif
