// K2 / K3: recursive blocked Cholesky, right-side triangular solve (X <- X Lp^-T) and
// vector triangular solves, all float64.  The O(n^3) work runs in the DMMA GEMM
// (mb_gemm.cu); the kernels here are the 32-wide leaves.
#include "mb_common.cuh"

namespace {

constexpr int NB = 32;

// ---- leaf: Cholesky of one (w <= 32) diagonal block, one warp ------------------------------
__global__ void potrf_leaf_kernel(double* __restrict__ A, int64_t lda, int w, int64_t global_off, int* info) {
  __shared__ double S[NB][NB + 1];
  const int lane = threadIdx.x;
  for (int r = 0; r < w; r++)
    if (lane < w) S[r][lane] = A[r * lda + lane];
  __syncwarp();
  for (int j = 0; j < w; j++) {
    double d = S[j][j];
    if (!(d > 0.0)) {
      if (lane == 0) atomicCAS(info, 0, (int)(global_off + j + 1));
      d = nan("");
    }
    d = sqrt(d);
    __syncwarp();
    if (lane == j) S[j][j] = d;
    if (lane > j && lane < w) S[lane][j] = S[lane][j] / d;
    __syncwarp();
    if (lane > j && lane < w) {
      const double lij = S[lane][j];
      for (int k = j + 1; k <= lane; k++) S[lane][k] = fma(-lij, S[k][j], S[lane][k]);
    }
    __syncwarp();
  }
  for (int r = 0; r < w; r++)
    if (lane < w) A[r * lda + lane] = (lane <= r) ? S[r][lane] : 0.0;
}

// ---- leaf: X[:, 0:w] <- X[:, 0:w] T^-T for a (w <= 32) lower-triangular block T --------------
// 128 rows per CTA, one thread per row, row in registers, T broadcast from shared memory.
__global__ void __launch_bounds__(128)
trsm_leaf_kernel(const double* __restrict__ T, int64_t ldt, int w, double* __restrict__ X, int64_t ldx,
                 int64_t nrows) {
  __shared__ double Ts[NB][NB + 1];
  __shared__ double Xs[128][NB + 1];
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += 128) {
    int r = e / NB, c = e % NB;
    Ts[r][c] = (r < w && c < w) ? T[r * ldt + c] : (r == c ? 1.0 : 0.0);
  }
  for (int64_t r0 = (int64_t)blockIdx.x * 128; r0 < nrows; r0 += (int64_t)gridDim.x * 128) {
    __syncthreads();
    for (int e = tid; e < 128 * NB; e += 128) {
      int r = e / NB, c = e % NB;
      Xs[r][c] = (r0 + r < nrows && c < w) ? X[(r0 + r) * ldx + c] : 0.0;
    }
    __syncthreads();
    double x[NB];
#pragma unroll
    for (int c = 0; c < NB; c++) x[c] = Xs[tid][c];
#pragma unroll
    for (int c = 0; c < NB; c++) {
      double s = x[c];
#pragma unroll
      for (int k = 0; k < c; k++) s = fma(-x[k], Ts[c][k], s);
      x[c] = s / Ts[c][c];
    }
#pragma unroll
    for (int c = 0; c < NB; c++) Xs[tid][c] = x[c];
    __syncthreads();
    for (int e = tid; e < 128 * NB; e += 128) {
      int r = e / NB, c = e % NB;
      if (r0 + r < nrows && c < w) X[(r0 + r) * ldx + c] = Xs[r][c];
    }
  }
}

int trsm_rec(mb_ctx* ctx, const double* Lp, int64_t ldl, int64_t c0, int64_t w, double* X, int64_t ldx,
             int64_t nrows) {
  if (w <= 0) return 0;
  if (w <= NB) {
    int grid = (int)min(ceil_div64(nrows, 128), (int64_t)ctx->n_sm * 8);
    MB_LAUNCH(ctx, trsm_leaf_kernel, grid, 128, 0, Lp + c0 * ldl + c0, ldl, (int)w, X + c0, ldx, nrows);
    return 0;
  }
  int64_t w1 = ((w / 2 + NB - 1) / NB) * NB;
  MB_TRY(trsm_rec(ctx, Lp, ldl, c0, w1, X, ldx, nrows));
  // X[:, c0+w1 : c0+w] -= X[:, c0 : c0+w1] . Lp[c0+w1 : c0+w, c0 : c0+w1]^T
  MB_TRY(mb_gemm_rows_raw(ctx, false, false, nrows, w - w1, w1, -1.0, X + c0, ldx, Lp + (c0 + w1) * ldl + c0, ldl,
                          1.0, X + c0 + w1, ldx));
  return trsm_rec(ctx, Lp, ldl, c0 + w1, w - w1, X, ldx, nrows);
}


// ---- inverse of the 128-wide diagonal blocks of a lower-triangular matrix -----------------------
// One CTA of 32 warps per block, T staged in shared memory ([IB][IB + 1]).  Warp q owns columns
// q, q + 32, q + 64, q + 96 of T^-1: forward substitution down the column, the dot product of each row
// split over the lanes and reduced by shuffles (fixed order), x kept in a per-warp shared-memory strip.
// Output: dense 128 x 128 blocks (zero above the diagonal and beyond w).
constexpr int IB = 128, IBT = 1024, ILD = IB + 1;
constexpr size_t IB_SMEM = ((size_t)IB * ILD + (size_t)(IBT / 32) * IB + IB) * sizeof(double);

// Tsh: factor block (diagonal included), xw: (IBT/32) x IB strip, rdiag: IB reciprocals of the diagonal
__device__ __forceinline__ void tri_inv_block(const double* Tsh, double* xw_all, double* rdiag, int w,
                                              double* __restrict__ out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < IB) rdiag[tid] = (tid < w) ? 1.0 / Tsh[tid * ILD + tid] : 0.0;
  __syncthreads();
  double* xw = xw_all + warp * IB;
  for (int c = warp; c < IB; c += IBT / 32) {
    if (c < w) {
      for (int r = c; r < w; r++) {
        double s = 0.0;
        for (int k = c + lane; k < r; k += 32) s = fma(Tsh[r * ILD + k], xw[k], s);
        s = warp_sum(s);
        const double v = (((r == c) ? 1.0 : 0.0) - s) * rdiag[r];
        if (lane == 0) xw[r] = v;
        __syncwarp();
      }
    }
    for (int r = lane; r < IB; r += 32) out[(int64_t)r * IB + c] = (c < w && r >= c && r < w) ? xw[r] : 0.0;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(IBT)
tri_inv_blocks_kernel(const double* __restrict__ L, int64_t ldl, int64_t m, double* __restrict__ inv) {
  extern __shared__ double Tsh[];  // IB x ILD, then the per-warp strips, then the reciprocal diagonal
  double* xw = Tsh + IB * ILD;
  double* rdiag = xw + (IBT / 32) * IB;
  const int64_t j0 = (int64_t)blockIdx.x * IB;
  const int w = (int)min((int64_t)IB, m - j0);
  for (int e = threadIdx.x; e < IB * IB; e += IBT) {
    const int r = e / IB, k = e % IB;
    Tsh[r * ILD + k] = (r < w && k < w && k <= r) ? L[(j0 + r) * ldl + j0 + k] : 0.0;
  }
  __syncthreads();
  tri_inv_block(Tsh, xw, rdiag, w, inv + (int64_t)blockIdx.x * IB * IB);
}

// X[:, c0:c0+w] <- X[:, c0:c0+w] Lp[c0:c0+w, c0:c0+w]^-T with GEMM leaves on the inverted diagonal blocks.
// `i8`: the large off-diagonal updates X2 -= X1 L21^T run on the tcgen05 int8 digit slices (mb_i8.cu); the choice
// depends on the GLOBAL row count and the block widths only.
int trsm_inv_rec(mb_ctx* ctx, const double* Lp, int64_t ldl, const double* inv, int64_t c0, int64_t w, double* X,
                 int64_t ldx, int64_t nrows, int64_t rows_total) {
  if (w <= 0) return 0;
  if (w <= IB) {
    // in place: a CTA consumes its whole 128 x w input tile before it stores the same tile
    const double* Tinv = inv + (c0 / IB) * IB * IB;
    return mb_gemm_rows_raw(ctx, false, false, nrows, w, w, 1.0, X + c0, ldx, Tinv, IB, 0.0, X + c0, ldx);
  }
  const int64_t w1 = ((w / 2 + IB - 1) / IB) * IB;
  MB_TRY(trsm_inv_rec(ctx, Lp, ldl, inv, c0, w1, X, ldx, nrows, rows_total));
  if (mb_i8_nt_usable(ctx, rows_total, w - w1, w1)) {
    MB_TRY(mb_i8_gemm_nt(ctx, nrows, w - w1, w1, -1.0, X + c0, ldx, Lp + (c0 + w1) * ldl + c0, ldl, 1, X + c0 + w1, ldx));
  } else {
    MB_TRY(mb_gemm_rows_raw(ctx, false, false, nrows, w - w1, w1, -1.0, X + c0, ldx, Lp + (c0 + w1) * ldl + c0, ldl, 1.0,
                            X + c0 + w1, ldx));
  }
  return trsm_inv_rec(ctx, Lp, ldl, inv, c0 + w1, w - w1, X, ldx, nrows, rows_total);
}

int potrf_rec(mb_ctx* ctx, double* A, int64_t lda, int64_t off, int64_t n, int* info) {
  if (n <= 0) return 0;
  double* D = A + off * lda + off;
  if (n <= NB) {
    MB_LAUNCH(ctx, potrf_leaf_kernel, 1, 32, 0, D, lda, (int)n, off, info);
    return 0;
  }
  int64_t n1 = ((n / 2 + NB - 1) / NB) * NB, n2 = n - n1;
  MB_TRY(potrf_rec(ctx, A, lda, off, n1, info));
  double* A21 = A + (off + n1) * lda + off;
  // A21 <- A21 L11^-T
  MB_TRY(trsm_rec(ctx, D, lda, 0, n1, A21, lda, n2));
  // A22 <- A22 - A21 A21^T   (lower tiles only)
  double* A22 = A + (off + n1) * lda + off + n1;
  MB_TRY(mb_gemm_raw(ctx, false, false, n2, n2, n1, -1.0, A21, lda, A21, lda, 1.0, A22, lda, true));
  return potrf_rec(ctx, A, lda, off + n1, n2, info);
}

// ---- leaf: B[0:w, :] <- T^-1 B (trans = 0) or T^-T B (trans = 1), T (w <= 32) lower ----------
// one thread per right-hand-side column (coalesced across the row), the column in registers.
__global__ void __launch_bounds__(128)
trsm_left_leaf_kernel(const double* __restrict__ T, int64_t ldt, int w, int trans, double* __restrict__ B,
                      int64_t ldb, int64_t nrhs) {
  __shared__ double Ts[NB][NB + 1];
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    int r = e / NB, c = e % NB;
    Ts[r][c] = (r < w && c < w) ? T[r * ldt + c] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (col >= nrhs) return;
  double x[NB];
#pragma unroll
  for (int r = 0; r < NB; r++) x[r] = (r < w) ? B[r * ldb + col] : 0.0;
  if (!trans) {
#pragma unroll
    for (int r = 0; r < NB; r++) {
      double s = x[r];
#pragma unroll
      for (int k = 0; k < r; k++) s = fma(-Ts[r][k], x[k], s);
      x[r] = s / Ts[r][r];
    }
  } else {
#pragma unroll
    for (int r = NB - 1; r >= 0; r--) {
      double s = x[r];
#pragma unroll
      for (int k = r + 1; k < NB; k++) s = fma(-Ts[k][r], x[k], s);
      x[r] = s / Ts[r][r];
    }
  }
#pragma unroll
  for (int r = 0; r < NB; r++)
    if (r < w) B[r * ldb + col] = x[r];
}

// B[r0:r0+w, :] <- L[r0:r0+w, r0:r0+w]^-1 (or ^-T) B[r0:r0+w, :], recursive with GEMM updates
int trsm_left_rec(mb_ctx* ctx, const double* L, int64_t ldl, int64_t r0, int64_t w, int trans, double* B,
                  int64_t ldb, int64_t nrhs) {
  if (w <= 0) return 0;
  if (w <= NB) {
    MB_LAUNCH(ctx, trsm_left_leaf_kernel, (int)ceil_div64(nrhs, 128), 128, 0, L + r0 * ldl + r0, ldl, (int)w,
              trans, B + r0 * ldb, ldb, nrhs);
    return 0;
  }
  const int64_t w1 = ((w / 2 + NB - 1) / NB) * NB, w2 = w - w1;
  const double* L21 = L + (r0 + w1) * ldl + r0;
  if (!trans) {
    MB_TRY(trsm_left_rec(ctx, L, ldl, r0, w1, trans, B, ldb, nrhs));
    // B2 -= L21 B1
    MB_TRY(mb_gemm_raw(ctx, false, true, w2, nrhs, w1, -1.0, L21, ldl, B + r0 * ldb, ldb, 1.0,
                       B + (r0 + w1) * ldb, ldb, false));
    return trsm_left_rec(ctx, L, ldl, r0 + w1, w2, trans, B, ldb, nrhs);
  }
  MB_TRY(trsm_left_rec(ctx, L, ldl, r0 + w1, w2, trans, B, ldb, nrhs));
  // B1 -= L21^T B2
  MB_TRY(mb_gemm_raw(ctx, true, true, w1, nrhs, w2, -1.0, L21, ldl, B + (r0 + w1) * ldb, ldb, 1.0, B + r0 * ldb,
                     ldb, false));
  return trsm_left_rec(ctx, L, ldl, r0, w1, trans, B, ldb, nrhs);
}

__global__ void merge_info_kernel(const int* from, int* to) {
  if (*from != 0) atomicCAS(to, 0, *from);
}

__global__ void zero_upper_kernel(double* a, int64_t n, int64_t lda) {
  int64_t i = blockIdx.y, j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && j < n && j > i) a[i * lda + j] = 0.0;
}

// ---- vector triangular solves: one CTA per right-hand side ---------------------------------
// forward:  L x = b     backward: L^T x = b.   x lives in shared memory (m doubles).
__global__ void __launch_bounds__(1024)
trsv_kernel(const double* __restrict__ L, int64_t ldl, int m, double* __restrict__ B, int nrhs, int trans) {
  extern __shared__ double xs[];
  __shared__ double T[NB][NB + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int col = blockIdx.x;
  for (int i = tid; i < m; i += blockDim.x) xs[i] = B[(int64_t)i * nrhs + col];
  __syncthreads();
  const int nblk = (m + NB - 1) / NB;
  for (int bb = 0; bb < nblk; bb++) {
    const int b = trans ? nblk - 1 - bb : bb;
    const int j0 = b * NB, w = min(NB, m - j0);
    // stage the diagonal block
    for (int e = tid; e < NB * NB; e += blockDim.x) {
      int r = e / NB, c = e % NB;
      T[r][c] = (r < w && c < w) ? L[(int64_t)(j0 + r) * ldl + j0 + c] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    if (warp == 0) {
      double xi = (lane < w) ? xs[j0 + lane] : 0.0;
      if (!trans) {
        for (int j = 0; j < w; j++) {
          double v = __shfl_sync(0xffffffffu, xi, j) / T[j][j];
          if (lane == j) xi = v;
          if (lane > j) xi = fma(-T[lane][j], v, xi);
        }
      } else {
        for (int j = w - 1; j >= 0; j--) {
          double v = __shfl_sync(0xffffffffu, xi, j) / T[j][j];
          if (lane == j) xi = v;
          if (lane < j) xi = fma(-T[j][lane], v, xi);
        }
      }
      if (lane < w) xs[j0 + lane] = xi;
    }
    __syncthreads();
    if (!trans) {
      // rows below the block: xs[i] -= sum_jj L[i][j0+jj] xs[j0+jj]
      const double xj = (lane < w) ? xs[j0 + lane] : 0.0;
      for (int i = j0 + NB + warp; i < m; i += nwarps) {
        double v = (lane < w) ? L[(int64_t)i * ldl + j0 + lane] * xj : 0.0;
        v = warp_sum(v);
        if (lane == 0) xs[i] -= v;
      }
    } else {
      // entries before the block: xs[i] -= sum_jj L[j0+jj][i] xs[j0+jj]
      for (int i = tid; i < j0; i += blockDim.x) {
        double s = 0.0;
        for (int jj = 0; jj < w; jj++) s = fma(L[(int64_t)(j0 + jj) * ldl + i], xs[j0 + jj], s);
        xs[i] -= s;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < m; i += blockDim.x) B[(int64_t)i * nrhs + col] = xs[i];
}


// ---- leaf: Cholesky of one (w <= 128) diagonal block in shared memory + the inverse of its factor -----
// Blocked inside the CTA on 32 x 32 sub-blocks (1024 threads, 12 + 10 barriers instead of 256):
//   factor, per 32-column panel p: (A1) warp 0 factors the diagonal block left-looking, its row in registers and
//   the finished rows broadcast from shared memory; (A2) one thread per row below solves its 32 entries against
//   that block; (A3) all threads apply the rank-32 update to the trailing sub-matrix.
//   inverse: (B1) warps 0..3 invert the four diagonal blocks (lane = column, forward substitution in registers);
//   (B2) block row by block row X_ij = -X_ii (sum_k L_ik X_kj), one thread per entry; finished off-diagonal
//   blocks are parked TRANSPOSED in the (otherwise zero) strict upper part of the shared-memory tile.
// Rows / columns >= w are padded with the identity, so every loop is uniform; only the w x w part is stored.
constexpr int SB = 32, LT = 512;  // 512 threads: 128 registers each (the 32-entry rows live in registers)

__global__ void __launch_bounds__(LT)
potrf_inv_leaf128_kernel(double* __restrict__ A, int64_t lda, int w, int64_t global_off, int* info,
                         double* __restrict__ inv_out) {
  extern __shared__ double Ssh[];
  double* xdiag = Ssh + IB * ILD;            // 4 x (32 x 32): inverses of the diagonal sub-blocks
  double* rdiag = xdiag + (IBT / 32) * IB;   // 128 reciprocals of the factor's diagonal
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < IB * IB; e += LT) {
    const int r = e / IB, k = e % IB;
    double v = 0.0;
    if (r < w && k <= r) v = A[(int64_t)r * lda + k];
    else if (r >= w && k == r) v = 1.0;
    Ssh[r * ILD + k] = v;
  }
  __syncthreads();

  for (int p = 0; p < IB / SB; p++) {
    const int c0 = p * SB;
    // (A1) diagonal block, warp 0, lane = row
    if (warp == 0) {
      double row[SB];
#pragma unroll
      for (int k = 0; k < SB; k++) row[k] = Ssh[(c0 + lane) * ILD + c0 + k];
#pragma unroll
      for (int j = 0; j < SB; j++) {
        double s = row[j];
#pragma unroll
        for (int k = 0; k < j; k++) s = fma(-row[k], Ssh[(c0 + j) * ILD + c0 + k], s);  // finished row j, broadcast
        const double djj = __shfl_sync(0xffffffffu, s, j);
        double dinv, d;
        if (!(djj > 0.0)) {
          if (lane == 0 && c0 + j < w) atomicCAS(info, 0, (int)(global_off + c0 + j + 1));
          dinv = d = nan("");
        } else {
          dinv = rsqrt(djj);
          d = djj * dinv;
        }
        const double v = (lane == j) ? d : s * dinv;
        row[j] = v;
        if (lane >= j) Ssh[(c0 + lane) * ILD + c0 + j] = v;
        if (lane == j) rdiag[c0 + j] = dinv;
        __syncwarp();
      }
    }
    __syncthreads();
    // (A2) rows below the diagonal block: X[r][0..31] = S[r][c0..] D^-T, one thread per row
    if (tid < IB - c0 - SB) {
      const int r = c0 + SB + tid;
      double x[SB];
#pragma unroll
      for (int c = 0; c < SB; c++) x[c] = Ssh[r * ILD + c0 + c];
#pragma unroll
      for (int c = 0; c < SB; c++) {
        double s = x[c];
#pragma unroll
        for (int k = 0; k < c; k++) s = fma(-x[k], Ssh[(c0 + c) * ILD + c0 + k], s);
        x[c] = s * rdiag[c0 + c];
      }
#pragma unroll
      for (int c = 0; c < SB; c++) Ssh[r * ILD + c0 + c] = x[c];
    }
    __syncthreads();
    // (A3) trailing update S[i][k] -= sum_c X[i][c] X[k][c] for c0 + 32 <= k <= i
    const int t0 = c0 + SB, nt = IB - t0;
    for (int e = tid; e < nt * nt; e += LT) {
      const int i = t0 + e / nt, k = t0 + e % nt;
      if (k <= i) {
        double s = Ssh[i * ILD + k];
#pragma unroll 8
        for (int c = 0; c < SB; c++) s = fma(-Ssh[i * ILD + c0 + c], Ssh[k * ILD + c0 + c], s);
        Ssh[i * ILD + k] = s;
      }
    }
    __syncthreads();
  }
  // the factor goes back to global memory (explicit zeros above the diagonal)
  for (int e = tid; e < IB * IB; e += LT) {
    const int r = e / IB, k = e % IB;
    if (r < w && k < w) A[(int64_t)r * lda + k] = (k <= r) ? Ssh[r * ILD + k] : 0.0;
  }
  // (B1) inverses of the four diagonal sub-blocks: warp p, lane = column
  if (warp < IB / SB) {
    const int c0 = warp * SB;
    double x[SB];
#pragma unroll
    for (int r = 0; r < SB; r++) {
      double s = (r == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < r; k++) s = fma(-Ssh[(c0 + r) * ILD + c0 + k], x[k], s);  // x[k] = 0 above the lane's column
      x[r] = (r >= lane) ? s * rdiag[c0 + r] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < SB; r++) xdiag[warp * SB * SB + r * SB + lane] = x[r];
  }
  __syncthreads();
  // (B2) off-diagonal blocks, block row i: X_ij = -X_ii (sum_{k=j}^{i-1} L_ik X_kj); each thread owns the two
  // entries (a, bcol) and (a + 16, bcol) of every 32 x 32 block
  const int bcol = tid & 31;
  for (int i = 1; i < IB / SB; i++) {
    double wv[2][IB / SB - 1];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int a = (tid >> 5) + 16 * h;
#pragma unroll
      for (int j = 0; j < IB / SB - 1; j++) {
        wv[h][j] = 0.0;
        if (j < i) {
          double s = 0.0;
          for (int k = j; k < i; k++) {
#pragma unroll 8
            for (int c = 0; c < SB; c++) {
              const double xkj = (k == j) ? xdiag[j * SB * SB + c * SB + bcol]
                                          : Ssh[(SB * j + bcol) * ILD + SB * k + c];  // X_kj[c][bcol], parked transposed
              s = fma(Ssh[(SB * i + a) * ILD + SB * k + c], xkj, s);
            }
          }
          wv[h][j] = s;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int j = 0; j < IB / SB - 1; j++)
        if (j < i) Ssh[(SB * j + bcol) * ILD + SB * i + (tid >> 5) + 16 * h] = wv[h][j];  // W_ij[a][bcol] parked transposed
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int a = (tid >> 5) + 16 * h;
#pragma unroll
      for (int j = 0; j < IB / SB - 1; j++) {
        if (j < i) {
          double s = 0.0;
          for (int c = 0; c <= a; c++) s = fma(xdiag[i * SB * SB + a * SB + c], Ssh[(SB * j + bcol) * ILD + SB * i + c], s);
          wv[h][j] = -s;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int j = 0; j < IB / SB - 1; j++)
        if (j < i) Ssh[(SB * j + bcol) * ILD + SB * i + (tid >> 5) + 16 * h] = wv[h][j];  // X_ij[a][bcol]
    __syncthreads();
  }
  // (B3) dense 128 x 128 inverse (zero above the diagonal and beyond w)
  for (int e = tid; e < IB * IB; e += LT) {
    const int r = e / IB, c = e % IB;
    double v = 0.0;
    if (r < w && c < w && c <= r) {
      const int bi = r / SB, bj = c / SB;
      v = (bi == bj) ? xdiag[bi * SB * SB + (r % SB) * SB + (c % SB)] : Ssh[c * ILD + r];
    }
    inv_out[(int64_t)r * IB + c] = v;
  }
}

// recursive Cholesky with 128-wide leaves; panel solves run as GEMMs on the inverted leaf factors
int potrf128_rec(mb_ctx* ctx, double* A, int64_t lda, int64_t off, int64_t n, int* info, double* inv) {
  if (n <= 0) return 0;
  double* D = A + off * lda + off;
  if (n <= IB) {
    MB_LAUNCH(ctx, potrf_inv_leaf128_kernel, 1, LT, IB_SMEM, D, lda, (int)n, off, info, inv + (off / IB) * IB * IB);
    return 0;
  }
  const int64_t n1 = ((n / 2 + IB - 1) / IB) * IB, n2 = n - n1;
  MB_TRY(potrf128_rec(ctx, A, lda, off, n1, info, inv));
  double* A21 = A + (off + n1) * lda + off;
  MB_TRY(trsm_inv_rec(ctx, D, lda, inv + (off / IB) * IB * IB, 0, n1, A21, lda, n2, 0));
  double* A22 = A + (off + n1) * lda + off + n1;
  MB_TRY(mb_gemm_raw(ctx, false, false, n2, n2, n1, -1.0, A21, lda, A21, lda, 1.0, A22, lda, true));
  return potrf128_rec(ctx, A, lda, off + n1, n2, info, inv);
}

int mb_trsm_ws(mb_ctx* ctx, int64_t m) {
  const size_t need = (size_t)ceil_div64(m, IB) * IB * IB * sizeof(double);
  if (need > ctx->trsm_ws_bytes) {
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    mb_invalidate_graphs(ctx);
    if (ctx->trsm_ws) MB_CUDA(cudaFree(ctx->trsm_ws));
    ctx->trsm_ws = nullptr;
    ctx->trsm_ws_bytes = 0;
    MB_CUDA(mb_dev_malloc(ctx, (void**)&ctx->trsm_ws, need));
    ctx->trsm_ws_bytes = need;
  }
  static mb_per_device_flag configured;
  if (!configured(ctx)) {
    MB_CUDA(cudaFuncSetAttribute(tri_inv_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IB_SMEM));
    MB_CUDA(cudaFuncSetAttribute(potrf_inv_leaf128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IB_SMEM));
    configured(ctx) = true;
  }
  return 0;
}


// ---- blocked vector triangular solve on the inverted 128-blocks ------------------------------------
// forward (L x = b):   for j = 0 .. nb-1:  x_j = Dinv_j b_j ;  b_(>j) -= L_(>j, j) x_j
// backward (L^T x = b): for j = nb-1 .. 0:  x_j = Dinv_j^T b_j ; b_(<j) -= L_(j, <j)^T x_j
// Two small launches per block instead of one CTA walking the whole 200 MB factor.
__global__ void __launch_bounds__(IBT)
trsv_diag_kernel(const double* __restrict__ inv, int w, int trans, double* __restrict__ b) {
  __shared__ double bs[IB], xs[IB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < IB) bs[tid] = (tid < w) ? b[tid] : 0.0;
  __syncthreads();
  if (!trans) {
    // x[r] = sum_c inv[r][c] b[c]: warp per 4 rows, lanes across the row
    for (int r = warp; r < w; r += IBT / 32) {
      double s = 0.0;
      for (int c = lane; c <= r; c += 32) s = fma(inv[(int64_t)r * IB + c], bs[c], s);
      s = warp_sum(s);
      if (lane == 0) xs[r] = s;
    }
  } else if (tid < IB) {
    // x[c] = sum_r inv[r][c] b[r]: thread per column, coalesced across the row
    double s = 0.0;
    for (int r = tid; r < w; r++) s = fma(inv[(int64_t)r * IB + tid], bs[r], s);
    xs[tid] = s;
  }
  __syncthreads();
  if (tid < w) b[tid] = xs[tid];
}
// forward update: rows below block j; one warp per row
__global__ void __launch_bounds__(256)
trsv_update_fwd_kernel(const double* __restrict__ L, int64_t ldl, int64_t row0, int64_t m, int64_t col0, int w,
                       double* __restrict__ b) {
  __shared__ double xs[IB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < IB) xs[tid] = (tid < w) ? b[col0 + tid] : 0.0;
  __syncthreads();
  const int64_t r = row0 + (int64_t)blockIdx.x * 8 + warp;
  if (r >= m) return;
  double s = 0.0;
  for (int c = lane; c < w; c += 32) s = fma(L[r * ldl + col0 + c], xs[c], s);
  s = warp_sum(s);
  if (lane == 0) b[r] -= s;
}
// backward update: columns left of block j; one thread per column
__global__ void __launch_bounds__(256)
trsv_update_bwd_kernel(const double* __restrict__ L, int64_t ldl, int64_t row0, int w, int64_t ncols,
                       double* __restrict__ b) {
  __shared__ double xs[IB];
  const int tid = threadIdx.x;
  if (tid < IB) xs[tid] = (tid < w) ? b[row0 + tid] : 0.0;
  __syncthreads();
  const int64_t c = (int64_t)blockIdx.x * 256 + tid;
  if (c >= ncols) return;
  double s = 0.0;
  for (int r = 0; r < w; r++) s = fma(L[(row0 + r) * ldl + c], xs[r], s);
  b[c] -= s;
}

}  // namespace

int mb_trsm_right_lt_raw(mb_ctx* ctx, const double* Lp, int64_t ldl, int64_t m, double* X, int64_t ldx,
                         int64_t nrows, int64_t rows_total) {
  if (nrows <= 0 || m <= 0) return 0;
  // the algorithm is chosen from the GLOBAL row count, so a row's result does not depend on how the cells are sharded
  if (ctx->opt_trsm == 1 || std::max(nrows, rows_total) < 4 * IB) return trsm_rec(ctx, Lp, ldl, 0, m, X, ldx, nrows);
  // tall right-hand sides: every flop in the DMMA GEMM (diagonal blocks applied as explicit 128 x 128 inverses)
  const int64_t nb = ceil_div64(m, IB);
  MB_TRY(mb_trsm_ws(ctx, m));
  MB_LAUNCH(ctx, tri_inv_blocks_kernel, (int)nb, IBT, IB_SMEM, Lp, ldl, m, ctx->trsm_ws);
  MB_TRY(trsm_inv_rec(ctx, Lp, ldl, ctx->trsm_ws, 0, m, X, ldx, nrows, std::max(nrows, rows_total)));
  return (ctx->opt_i8 && ctx->i8_status) ? mb_i8_check(ctx) : 0;
}

// direct (stream) Cholesky of the n x n matrix at A; the strict upper triangle is zeroed
static int potrf_direct(mb_ctx* ctx, double* A, int64_t n, int64_t lda, int* info_dev) {
  if (ctx->opt_trsm == 1 || n <= 32) {
    MB_TRY(potrf_rec(ctx, A, lda, 0, n, info_dev));
  } else {
    // NB: the TRSM workspace is overwritten leaf by leaf; a later mb_trsm_right_lt recomputes it
    MB_TRY(mb_trsm_ws(ctx, n));
    MB_TRY(potrf128_rec(ctx, A, lda, 0, n, info_dev, ctx->trsm_ws));
  }
  if (n > 0) {
    dim3 grid((unsigned)ceil_div64(n, 256), (unsigned)n);
    MB_CHECK(n < 65536 * 1, "mb_potrf: n=%lld too large for the zero-upper grid", (long long)n);
    MB_LAUNCH(ctx, zero_upper_kernel, grid, 256, 0, A, n, lda);
  }
  return 0;
}

// The recursive factorisation is ~300 small dependent launches (22 ms for n = 5000, about half of it launch
// gaps) and it is replicated on every rank, so it is the Amdahl term of the multi-GPU fit.  For n >= 1024 the
// whole launch sequence is captured ONCE per size into a CUDA graph that works on context-owned buffers
// (matrix copied in and out: 2 x 200 MB, 0.1 ms) and replayed on every later call.
static int potrf_graph(mb_ctx* ctx, double* A, int64_t n, int64_t lda, int* info_dev, bool* done) {
  *done = false;
  const size_t bytes = (size_t)n * n * sizeof(double);
  if (bytes > ctx->potrf_buf_bytes) {
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    mb_invalidate_graphs(ctx);  // they point into the old buffer
    if (ctx->potrf_buf) MB_CUDA(cudaFree(ctx->potrf_buf));
    ctx->potrf_buf = nullptr;
    ctx->potrf_buf_bytes = 0;
    MB_CUDA(mb_dev_malloc(ctx, (void**)&ctx->potrf_buf, bytes));
    ctx->potrf_buf_bytes = bytes;
  }
  if (!ctx->potrf_info) MB_CUDA(mb_dev_malloc(ctx, (void**)&ctx->potrf_info, 256));
  auto it = ctx->potrf_graphs.find(n);
  if (it == ctx->potrf_graphs.end()) {
    // everything that allocates or synchronises happens before the capture starts
    MB_TRY(mb_trsm_ws(ctx, n));
    const int64_t half = ((n / 2 + IB - 1) / IB) * IB + IB;
    MB_TRY(mb_gemm_reserve_ws(ctx, (size_t)16 * half * (half + 2) * sizeof(double)));
    const bool prof = ctx->prof_on;
    ctx->prof_on = false;  // event pairs around captured launches would not be timeable
    const int64_t launches0 = ctx->launches;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
    int rc = (e == cudaSuccess) ? potrf_direct(ctx, ctx->potrf_buf, n, n, ctx->potrf_info) : -1;
    cudaError_t e2 = (e == cudaSuccess) ? cudaStreamEndCapture(ctx->stream, &graph) : e;
    ctx->prof_on = prof;
    const int64_t nodes = ctx->launches - launches0;
    ctx->launches = launches0;
    cudaGraphExec_t exec = nullptr;
    if (rc != 0 || e2 != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      ctx->opt_graph = 0;  // fall back to stream launches for the rest of this context's life
      return 0;
    }
    cudaGraphDestroy(graph);
    it = ctx->potrf_graphs.emplace(n, std::make_pair(exec, nodes)).first;
  }
  MB_CUDA(cudaMemcpy2DAsync(ctx->potrf_buf, (size_t)n * sizeof(double), A, (size_t)lda * sizeof(double),
                            (size_t)n * sizeof(double), (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
  MB_CUDA(cudaMemsetAsync(ctx->potrf_info, 0, sizeof(int), ctx->stream));
  MB_CUDA(cudaGraphLaunch(it->second.first, ctx->stream));
  ctx->launches += it->second.second;
  MB_CUDA(cudaMemcpy2DAsync(A, (size_t)lda * sizeof(double), ctx->potrf_buf, (size_t)n * sizeof(double),
                            (size_t)n * sizeof(double), (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
  // a non-positive pivot found by the replay is OR-ed into the caller's flag (first failing pivot wins there)
  MB_LAUNCH(ctx, merge_info_kernel, 1, 1, 0, ctx->potrf_info, info_dev);
  *done = true;
  return 0;
}

int mb_potrf_raw(mb_ctx* ctx, double* A, int64_t n, int64_t lda, int* info_dev) {
  if (ctx->opt_graph && ctx->opt_trsm != 1 && n >= 1024) {
    bool done = false;
    MB_TRY(potrf_graph(ctx, A, n, lda, info_dev, &done));
    if (done) return 0;
  }
  return potrf_direct(ctx, A, n, lda, info_dev);
}

extern "C" int mb_potrf(mb_ctx* ctx, mb_mat* a) {
  MB_RANGE("mellon_b200: K2 potrf");
  MB_CHECK(ctx && a, "mb_potrf: null argument");
  MB_CHECK(a->rows == a->cols, "mb_potrf: matrix is %lld x %lld, not square", (long long)a->rows,
           (long long)a->cols);
  MB_CUDA(cudaSetDevice(ctx->device));
  if (a->rows == 0) return 0;
  double* scratch;
  MB_TRY(mb_scratch(ctx, 256, &scratch));
  int* info_dev = reinterpret_cast<int*>(scratch);
  MB_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), ctx->stream));
  MB_TRY(mb_potrf_raw(ctx, a->p, a->rows, a->cols, info_dev));
  int info = 0;
  MB_CUDA(cudaMemcpyAsync(&info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  return info;  // > 0: first non-positive pivot (1-based)
}

extern "C" int mb_trsm_right_lt(mb_ctx* ctx, const mb_mat* Lp, mb_mat* X) {
  MB_RANGE("mellon_b200: K3 trsm_right_lt");
  MB_CHECK(ctx && Lp && X, "mb_trsm_right_lt: null argument");
  MB_CHECK(Lp->rows == Lp->cols && X->cols == Lp->rows,
           "mb_trsm_right_lt: Lp is %lld x %lld, X is %lld x %lld", (long long)Lp->rows,
           (long long)Lp->cols, (long long)X->rows, (long long)X->cols);
  MB_CUDA(cudaSetDevice(ctx->device));
  return mb_trsm_right_lt_raw(ctx, Lp->p, Lp->cols, Lp->rows, X->p, X->cols, X->rows,
                              X->global_rows >= 0 ? X->global_rows : X->rows);
}

extern "C" int mb_tri_solve(mb_ctx* ctx, const mb_mat* Lp, int trans, mb_mat* B) {
  MB_RANGE("mellon_b200: tri_solve");
  MB_CHECK(ctx && Lp && B, "mb_tri_solve: null argument");
  MB_CHECK(Lp->rows == Lp->cols && B->rows == Lp->rows, "mb_tri_solve: Lp is %lld x %lld, B has %lld rows",
           (long long)Lp->rows, (long long)Lp->cols, (long long)B->rows);
  MB_CUDA(cudaSetDevice(ctx->device));
  const int64_t m = Lp->rows, nrhs = B->cols;
  if (m == 0 || nrhs == 0) return 0;
  if (nrhs == 1 && m > 2 * IB && ctx->opt_trsm != 1) {
    const int64_t nb = ceil_div64(m, IB);
    MB_TRY(mb_trsm_ws(ctx, m));
    MB_LAUNCH(ctx, tri_inv_blocks_kernel, (int)nb, IBT, IB_SMEM, Lp->p, Lp->cols, m, ctx->trsm_ws);
    const int64_t ldl = Lp->cols;
    for (int64_t jj = 0; jj < nb; jj++) {
      const int64_t j = trans ? nb - 1 - jj : jj;
      const int64_t j0 = j * IB;
      const int w = (int)min((int64_t)IB, m - j0);
      MB_LAUNCH(ctx, trsv_diag_kernel, 1, IBT, 0, ctx->trsm_ws + j * IB * IB, w, trans ? 1 : 0, B->p + j0);
      if (!trans) {
        const int64_t row0 = j0 + w;
        if (row0 < m)
          MB_LAUNCH(ctx, trsv_update_fwd_kernel, (int)ceil_div64(m - row0, 8), 256, 0, Lp->p, ldl, row0, m, j0, w, B->p);
      } else if (j0 > 0) {
        MB_LAUNCH(ctx, trsv_update_bwd_kernel, (int)ceil_div64(j0, 256), 256, 0, Lp->p, ldl, j0, w, j0, B->p);
      }
    }
    return 0;
  }
  const size_t smem = (size_t)m * sizeof(double);
  if (nrhs <= 64 && smem <= 200 * 1024) {
    static mb_per_device_flag configured;
    if (!configured(ctx)) {
      MB_CUDA(cudaFuncSetAttribute(trsv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured(ctx) = true;
    }
    MB_LAUNCH(ctx, trsv_kernel, (int)nrhs, 1024, smem, Lp->p, Lp->cols, (int)m, B->p, (int)nrhs, trans);
    return 0;
  }
  // wide right-hand sides: blocked left solve, the O(m^2 nrhs) work in the DMMA GEMM
  return trsm_left_rec(ctx, Lp->p, Lp->cols, 0, m, trans ? 1 : 0, B->p, nrhs, nrhs);
}
