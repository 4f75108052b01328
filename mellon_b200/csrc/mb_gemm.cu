// FP64 GEMM on the tensor pipe (DMMA m8n8k4, the native f64 MMA shape on sm_100a) with a
// 3-stage cp.async pipeline.  Used by K3 (TRSM updates), K4 (Gram / SYRK), K2 (Cholesky
// trailing updates) and the Nystroem products.
//
//   C[i*ldc + j] = alpha * sum_k A(i,k) * B(j,k) + beta * C[i*ldc + j]
//   A(i,k) = A[i*lda + k] (AK=false)  or  A[k*lda + i] (AK=true,  "k-major")
//   B(j,k) = B[j*ldb + k] (BK=false)  or  B[k*ldb + j] (BK=true)
//
// CTA tile 128x128x32 (one barrier per 32 k: tools/microbench_fp64 shows 34.5 TF with a barrier per 16 k,
// 35.6 TF per 32 k, 37.0 TF without); 16 warps (4 x 4), warp tile 32x32 = 4x4 DMMA tiles (default) or 8 warps (2 x 4),
// warp tile 64x32 (opt_gemm = 2).  Symmetric (lower-only) products split the long k axis so that
// tiles x splits fills whole waves of 148 CTAs; the partial tiles are summed in a fixed order.
#include "mb_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BKT = 32, STAGES = 3;
constexpr int LD_ROWMAJ = BKT + 4;   // [tile_rows][BKT+4]   (stride 36 = 4 mod 16: 4*m + k distinct mod 16)
constexpr int LD_KMAJ = BM + 4;      // [BKT][tile_rows+4]   (stride 132: 4*k + m distinct mod 16)
constexpr int TILE_DOUBLES = BM * LD_ROWMAJ;  // 2560 >= BKT * LD_KMAJ (2112)
constexpr size_t SMEM_BYTES = (size_t)STAGES * 2 * TILE_DOUBLES * sizeof(double);

__device__ __forceinline__ void cp_async16(double* smem, const double* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(double* smem, const double* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Load one operand tile (tile_rows = 128 "row" indices x BKT k indices) into shared memory.
// KMAJ=false: source element (r, k) at src[(r0+r)*ld + k0+k]  -> dst[r*LD_ROWMAJ + k]
// KMAJ=true : source element (r, k) at src[(k0+k)*ld + r0+r]  -> dst[k*LD_KMAJ + r]
template <bool KMAJ, int NTHREADS>
__device__ __forceinline__ void load_tile(double* dst, const double* __restrict__ src, int64_t ld,
                                          int64_t r0, int64_t nr, int64_t k0, int64_t nk, bool vec_ok,
                                          int tid) {
  if (!KMAJ) {
    if (vec_ok) {
#pragma unroll
      for (int it = 0; it < (BM * BKT / 2) / NTHREADS; it++) {
        int e = tid + it * NTHREADS;
        int r = e / (BKT / 2), c = (e % (BKT / 2)) * 2;
        int64_t gr = r0 + r, gk = k0 + c;
        int bytes = 0;
        if (gr < nr) {
          int64_t rem = nk - gk;
          bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
        }
        const double* g = bytes ? src + gr * ld + gk : src;
        cp_async16(dst + r * LD_ROWMAJ + c, g, bytes);
      }
    } else {
#pragma unroll
      for (int it = 0; it < (BM * BKT) / NTHREADS; it++) {
        int e = tid + it * NTHREADS;
        int r = e / BKT, c = e % BKT;
        int64_t gr = r0 + r, gk = k0 + c;
        int bytes = (gr < nr && gk < nk) ? 8 : 0;
        const double* g = bytes ? src + gr * ld + gk : src;
        cp_async8(dst + r * LD_ROWMAJ + c, g, bytes);
      }
    }
  } else {
    if (vec_ok) {
#pragma unroll
      for (int it = 0; it < (BM * BKT / 2) / NTHREADS; it++) {
        int e = tid + it * NTHREADS;
        int k = e / (BM / 2), c = (e % (BM / 2)) * 2;
        int64_t gk = k0 + k, gr = r0 + c;
        int bytes = 0;
        if (gk < nk) {
          int64_t rem = nr - gr;
          bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
        }
        const double* g = bytes ? src + gk * ld + gr : src;
        cp_async16(dst + k * LD_KMAJ + c, g, bytes);
      }
    } else {
#pragma unroll
      for (int it = 0; it < (BM * BKT) / NTHREADS; it++) {
        int e = tid + it * NTHREADS;
        int k = e / BM, c = e % BM;
        int64_t gk = k0 + k, gr = r0 + c;
        int bytes = (gk < nk && gr < nr) ? 8 : 0;
        const double* g = bytes ? src + gk * ld + gr : src;
        cp_async8(dst + k * LD_KMAJ + c, g, bytes);
      }
    }
  }
}


// Per-thread operand loader with everything but the k advance hoisted out of the k loop: each of the 512
// threads owns 4 16-byte chunks of a 128 x 32 operand tile; pointers, shared-memory offsets and edge
// predicates are computed once per output tile, a k-tile costs 4 cp.async + one pointer bump.  (The generic
// `load_tile` recomputed ~300 integer instructions per operand per k-tile between the barrier and the first
// DMMA — the main reason the tensor pipe sat at 81 % busy.)  Needs 16-byte-aligned operands and a full k-tile.
template <bool KMAJ>
struct FastLoader {
  const double* g0;      // chunk 0 of the next k-tile
  int64_t chunk_stride;  // doubles between this thread's consecutive chunks
  int64_t ktile_stride;  // doubles per k-tile advance
  uint32_t s0;           // shared-memory offset (doubles) of chunk 0 inside the operand tile
  uint32_t nb;           // bytes to copy per chunk (4 x 8 bit); 0 => zero fill (edge rows / columns)
  static constexpr uint32_t SCS = KMAJ ? 8 * LD_KMAJ : 32 * LD_ROWMAJ;

  __device__ __forceinline__ void init(const double* src, int64_t ld, int64_t r0, int64_t nr, int64_t kbeg, int tid) {
    nb = 0;
    if (!KMAJ) {
      const int r = tid >> 4, c = (tid & 15) * 2;
      g0 = src + (r0 + r) * ld + kbeg + c;
      chunk_stride = 32 * ld;
      ktile_stride = BKT;
      s0 = r * LD_ROWMAJ + c;
#pragma unroll
      for (int it = 0; it < 4; it++) nb |= ((r0 + r + 32 * it < nr) ? 16u : 0u) << (8 * it);
    } else {
      const int kk = tid >> 6, c = (tid & 63) * 2;
      g0 = src + (kbeg + kk) * ld + r0 + c;
      chunk_stride = 8 * ld;
      ktile_stride = (int64_t)BKT * ld;
      s0 = kk * LD_KMAJ + c;
      const int64_t rem = nr - (r0 + c);
      const uint32_t b = rem >= 2 ? 16u : (rem == 1 ? 8u : 0u);
      nb = b | (b << 8) | (b << 16) | (b << 24);
    }
  }
  __device__ __forceinline__ void load(double* dst) {
#pragma unroll
    for (int it = 0; it < 4; it++) cp_async16(dst + s0 + it * SCS, g0 + it * chunk_stride, (int)((nb >> (8 * it)) & 255u));
    g0 += ktile_stride;
  }
  __device__ __forceinline__ void skip() { g0 += ktile_stride; }
};

template <bool AK, bool BK, int WARPS_M, bool FAST>
__global__ void __launch_bounds__(WARPS_M * 128, 1)
gemm_dmma_kernel(int64_t m, int64_t n, int64_t k_total, double alpha, const double* __restrict__ A, int64_t lda,
                 const double* __restrict__ B, int64_t ldb, double beta, double* __restrict__ C,
                 int64_t ldc, int lower_only, int64_t tiles_n, int64_t n_tiles, int a_vec, int b_vec,
                 int ksplit, int64_t kchunk, double* __restrict__ ws, int64_t ws_stride, int64_t ldw) {
  constexpr int NTHREADS = WARPS_M * 128;
  constexpr int WTM = BM / WARPS_M;  // warp tile rows: 64 or 32
  constexpr int MI = WTM / 8;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp >> 2) * WTM, wn0 = (warp & 3) * 32;
  const int lr = lane >> 2, lk = lane & 3;

  for (int64_t item = blockIdx.x; item < n_tiles * ksplit; item += gridDim.x) {
    const int64_t t = item / ksplit;
    const int ss = (int)(item % ksplit);
    const int64_t kbeg = ss * kchunk;
    const int64_t k = min(k_total, kbeg + kchunk);  // this item contracts [kbeg, k)
    const int64_t nkt = (k - kbeg + BKT - 1) / BKT;
    int64_t bi, bj;
    if (lower_only) {
      // t enumerates (bi, bj) with bj <= bi, row by row
      bi = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
      while ((bi + 1) * (bi + 2) / 2 <= t) bi++;
      while (bi * (bi + 1) / 2 > t) bi--;
      bj = t - bi * (bi + 1) / 2;
    } else {
      bi = t / tiles_n;
      bj = t % tiles_n;
    }
    const int64_t m0 = bi * BM, n0 = bj * BN;

    double acc[MI][4][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // operand loaders: FAST (template) = hoisted loaders for full k-tiles of 16-byte-aligned operands
    FastLoader<AK> la;
    FastLoader<BK> lb;
    if (FAST) {
      la.init(A, lda, m0, m, kbeg, tid);
      lb.init(B, ldb, n0, n, kbeg, tid);
    }

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
      if (s < nkt) {
        double* sa = smem + (size_t)s * 2 * TILE_DOUBLES;
        double* sb = sa + TILE_DOUBLES;
        const int64_t k0 = kbeg + (int64_t)s * BKT;
        if (FAST && k0 + BKT <= k) {
          la.load(sa);
          lb.load(sb);
        } else {
          load_tile<AK, NTHREADS>(sa, A, lda, m0, m, k0, k, a_vec, tid);
          load_tile<BK, NTHREADS>(sb, B, ldb, n0, n, k0, k, b_vec, tid);
        }
      }
      cp_async_commit();
    }

    for (int64_t kt = 0; kt < nkt; kt++) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      {
        const int64_t nt = kt + STAGES - 1;
        if (nt < nkt) {
          double* sa = smem + (size_t)(nt % STAGES) * 2 * TILE_DOUBLES;
          double* sb = sa + TILE_DOUBLES;
          const int64_t k0 = kbeg + nt * BKT;
          if (FAST && k0 + BKT <= k) {
            la.load(sa);
            lb.load(sb);
          } else {
            load_tile<AK, NTHREADS>(sa, A, lda, m0, m, k0, k, a_vec, tid);
            load_tile<BK, NTHREADS>(sb, B, ldb, n0, n, k0, k, b_vec, tid);
          }
        }
        cp_async_commit();
      }
      const double* sa = smem + (size_t)(kt % STAGES) * 2 * TILE_DOUBLES;
      const double* sb = sa + TILE_DOUBLES;
#pragma unroll
      for (int ks = 0; ks < BKT / 4; ks++) {
        double af[MI], bf[4];
        const int kk = ks * 4 + lk;
#pragma unroll
        for (int i = 0; i < MI; i++) {
          int r = wm0 + i * 8 + lr;
          af[i] = AK ? sa[kk * LD_KMAJ + r] : sa[r * LD_ROWMAJ + kk];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          int c = wn0 + j * 8 + lr;
          bf[j] = BK ? sb[kk * LD_KMAJ + c] : sb[c * LD_ROWMAJ + kk];
        }
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    }
    cp_async_wait<0>();
    __syncthreads();

    // epilogue (split-k items store their raw partial tile into the workspace slice of their split)
    double* Cout = (ksplit > 1) ? ws + (int64_t)ss * ws_stride : C;
    const int64_t ldo = (ksplit > 1) ? ldw : ldc;
    const double alpha_e = (ksplit > 1) ? 1.0 : alpha, beta_e = (ksplit > 1) ? 0.0 : beta;
    const bool c_vec = ((ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(Cout) & 15) == 0);
#pragma unroll
    for (int i = 0; i < MI; i++) {
      int64_t row = m0 + wm0 + i * 8 + lr;
      if (row >= m) continue;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int64_t col = n0 + wn0 + j * 8 + lk * 2;
        if (col >= n) continue;
        double* cp = Cout + row * ldo + col;
        double v0 = alpha_e * acc[i][j][0], v1 = alpha_e * acc[i][j][1];
        if (col + 1 < n) {
          if (c_vec) {
            if (beta_e != 0.0) {
              double2 old = *reinterpret_cast<double2*>(cp);
              v0 += beta_e * old.x;
              v1 += beta_e * old.y;
            }
            *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
          } else {
            if (beta_e != 0.0) {
              v0 += beta_e * cp[0];
              v1 += beta_e * cp[1];
            }
            cp[0] = v0;
            cp[1] = v1;
          }
        } else {
          if (beta_e != 0.0) v0 += beta_e * cp[0];
          cp[0] = v0;
        }
      }
    }
  }
}

// ---- DFMA register-tiled reference variant (opt_gemm = 1): A/B measurement only -----------
// 64x64 tile, 256 threads, 4x4 micro-tile, no pipelining: simple and obviously correct.
template <bool AK, bool BK>
__global__ void __launch_bounds__(256)
gemm_dfma_kernel(int64_t m, int64_t n, int64_t k, double alpha, const double* __restrict__ A, int64_t lda,
                 const double* __restrict__ B, int64_t ldb, double beta, double* __restrict__ C,
                 int64_t ldc, int lower_only, int64_t tiles_n, int64_t n_tiles) {
  __shared__ double sa[16][64 + 1], sb[16][64 + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    int64_t bi = t / tiles_n, bj = t % tiles_n;
    if (lower_only && bj > bi) continue;
    const int64_t m0 = bi * 64, n0 = bj * 64;
    double acc[4][4] = {};
    for (int64_t k0 = 0; k0 < k; k0 += 16) {
      for (int e = tid; e < 64 * 16; e += 256) {
        int r, kk;
        if (AK) { kk = e / 64; r = e % 64; } else { r = e / 16; kk = e % 16; }
        int64_t gr = m0 + r, gk = k0 + kk;
        sa[kk][r] = (gr < m && gk < k) ? (AK ? A[gk * lda + gr] : A[gr * lda + gk]) : 0.0;
        if (BK) { kk = e / 64; r = e % 64; } else { r = e / 16; kk = e % 16; }
        gr = n0 + r; gk = k0 + kk;
        sb[kk][r] = (gr < n && gk < k) ? (BK ? B[gk * ldb + gr] : B[gr * ldb + gk]) : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; kk++) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = sa[kk][ty + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = sb[kk][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int64_t row = m0 + ty + 16 * i;
      if (row >= m) continue;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int64_t col = n0 + tx + 16 * j;
        if (col >= n) continue;
        double v = alpha * acc[i][j];
        if (beta != 0.0) v += beta * C[row * ldc + col];
        C[row * ldc + col] = v;
      }
    }
  }
}

// C = alpha * sum_s ws[s] + beta * C over the (lower-triangular tiles of the) output, splits in a fixed order
__global__ void splitk_reduce_kernel(const double* __restrict__ ws, int ksplit, int64_t ws_stride, int64_t ldw, int64_t m,
                                     int64_t n, int64_t ldc, double alpha, double beta, double* __restrict__ C,
                                     int lower_only) {
  const int64_t row = blockIdx.y;
  const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= m || col >= n) return;
  if (lower_only && col / BN > row / BM) return;  // tile never computed
  double s = 0.0;
  for (int q = 0; q < ksplit; q++) s += ws[(int64_t)q * ws_stride + row * ldw + col];
  double v = alpha * s;
  if (beta != 0.0) v += beta * C[row * ldc + col];
  C[row * ldc + col] = v;
}

template <bool AK, bool BK>
int launch_gemm(mb_ctx* ctx, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                const double* B, int64_t ldb, double beta, double* C, int64_t ldc, bool lower_only,
                bool allow_split = true) {
  if (ctx->opt_gemm == 1) {
    int64_t tm = ceil_div64(m, 64), tn = ceil_div64(n, 64), nt = tm * tn;
    int grid = (int)min(nt, (int64_t)ctx->n_sm * 8);
    MB_LAUNCH(ctx, (gemm_dfma_kernel<AK, BK>), grid, 256, 0, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc,
              lower_only ? 1 : 0, tn, nt);
    return 0;
  }
  static mb_per_device_flag configured;
  if (!configured(ctx)) {
    MB_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel<AK, BK, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    MB_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel<AK, BK, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    MB_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel<AK, BK, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    configured(ctx) = true;
  }
  int64_t tm = ceil_div64(m, BM), tn = ceil_div64(n, BN);
  int64_t nt = lower_only ? tm * (tm + 1) / 2 : tm * tn;
  if (lower_only) MB_CHECK(m == n, "gemm lower_only needs a square output");
  // split the contraction axis when the tile count leaves the last wave of CTAs mostly idle: few tiles (the
  // mid-size products inside the Cholesky) split down to 512-deep chunks, many tiles only to 4096-deep ones
  int ksplit = 1;
  if (ctx->opt_gemm != 3 && allow_split) {
    const int64_t sm = ctx->n_sm;
    const int64_t min_chunk = (nt < 4 * sm) ? 512 : 4096;
    double best = (double)nt / (double)(ceil_div64(nt, sm) * sm);
    for (int sp = 2; sp <= 16 && best < 0.97; sp++) {
      if (k / sp < min_chunk) break;
      const int64_t items = nt * sp;
      const double eff = (double)items / (double)(ceil_div64(items, sm) * sm);
      if (eff > best + 0.01) { best = eff; ksplit = sp; }
    }
  }
  int64_t kchunk = k, ws_stride = 0, ldw = 0;
  double* ws = nullptr;
  if (ksplit > 1) {
    kchunk = ceil_div64(ceil_div64(k, ksplit), BKT) * BKT;
    ksplit = (int)ceil_div64(k, kchunk);
    ldw = (n + 1) & ~(int64_t)1;  // dense, even (16-byte rows)
    ws_stride = m * ldw;
    MB_TRY(mb_gemm_reserve_ws(ctx, (size_t)ksplit * ws_stride * sizeof(double)));
    ws = ctx->gemm_ws;
  }
  int grid = (int)min(nt * ksplit, (int64_t)ctx->n_sm);
  int a_vec = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  int b_vec = ((ldb & 1) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  if (ctx->prof_on) ctx->prof_work[MB_PROF_GEMM] += (lower_only ? 1.0 : 2.0) * (double)m * (double)n * (double)k;
  if (ctx->opt_gemm == 2) {
    MB_LAUNCH_P(ctx, MB_PROF_GEMM, (gemm_dmma_kernel<AK, BK, 2, false>), grid, 256, SMEM_BYTES, m, n, k, alpha, A, lda, B, ldb,
                beta, C, ldc, lower_only ? 1 : 0, tn, nt, a_vec, b_vec, ksplit, kchunk, ws, ws_stride, ldw);
  } else if (a_vec && b_vec && ctx->opt_gemm != 4) {
    MB_LAUNCH_P(ctx, MB_PROF_GEMM, (gemm_dmma_kernel<AK, BK, 4, true>), grid, 512, SMEM_BYTES, m, n, k, alpha, A, lda, B, ldb,
                beta, C, ldc, lower_only ? 1 : 0, tn, nt, a_vec, b_vec, ksplit, kchunk, ws, ws_stride, ldw);
  } else {
    MB_LAUNCH_P(ctx, MB_PROF_GEMM, (gemm_dmma_kernel<AK, BK, 4, false>), grid, 512, SMEM_BYTES, m, n, k, alpha, A, lda, B, ldb,
                beta, C, ldc, lower_only ? 1 : 0, tn, nt, a_vec, b_vec, ksplit, kchunk, ws, ws_stride, ldw);
  }
  if (ksplit > 1) {
    dim3 rgrid((unsigned)ceil_div64(n, 256), (unsigned)m);
    MB_LAUNCH(ctx, splitk_reduce_kernel, rgrid, 256, 0, ws, ksplit, ws_stride, ldw, m, n, ldc, alpha, beta, C, lower_only ? 1 : 0);
  }
  return 0;
}

}  // namespace

int mb_gemm_reserve_ws(mb_ctx* ctx, size_t need) {
  if (need > ctx->gemm_ws_bytes) {
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    mb_invalidate_graphs(ctx);
    if (ctx->gemm_ws) MB_CUDA(cudaFree(ctx->gemm_ws));
    ctx->gemm_ws = nullptr;
    ctx->gemm_ws_bytes = 0;
    MB_CUDA(mb_dev_malloc(ctx, (void**)&ctx->gemm_ws, need));
    ctx->gemm_ws_bytes = need;
  }
  return 0;
}

int mb_gemm_raw(mb_ctx* ctx, bool a_kmajor, bool b_kmajor, int64_t m, int64_t n, int64_t k, double alpha,
                const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
                int64_t ldc, bool lower_only) {
  if (m <= 0 || n <= 0) return 0;
  MB_CUDA(cudaSetDevice(ctx->device));
  if (a_kmajor) {
    if (b_kmajor) return launch_gemm<true, true>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
    return launch_gemm<true, false>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
  }
  if (b_kmajor) return launch_gemm<false, true>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
  return launch_gemm<false, false>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
}

int mb_gemm_rows_raw(mb_ctx* ctx, bool a_kmajor, bool b_kmajor, int64_t m, int64_t n, int64_t k, double alpha,
                     const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
                     int64_t ldc) {
  if (m <= 0 || n <= 0) return 0;
  MB_CUDA(cudaSetDevice(ctx->device));
  if (a_kmajor) {
    if (b_kmajor) return launch_gemm<true, true>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, false, false);
    return launch_gemm<true, false>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, false, false);
  }
  if (b_kmajor) return launch_gemm<false, true>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, false, false);
  return launch_gemm<false, false>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, false, false);
}

// C = alpha op(A) op(B) + beta C with row-major matrices.
//   op(A) (m x k): trans_a=0 -> A is m x k (A(i,k)=A[i*lda+k], not k-major); trans_a=1 -> A is k x m (k-major)
//   op(B) (k x n): trans_b=0 -> B is k x n (B(j,k)=B[k*ldb+j], k-major);     trans_b=1 -> B is n x k (not k-major)
extern "C" int mb_gemm(mb_ctx* ctx, int trans_a, int trans_b, double alpha, const mb_mat* A, const mb_mat* B,
                       double beta, mb_mat* C) {
  MB_CHECK(ctx && A && B && C, "mb_gemm: null argument");
  int64_t m = trans_a ? A->cols : A->rows, ka = trans_a ? A->rows : A->cols;
  int64_t n = trans_b ? B->rows : B->cols, kb = trans_b ? B->cols : B->rows;
  MB_CHECK(ka == kb, "mb_gemm: inner dimensions differ (%lld vs %lld)", (long long)ka, (long long)kb);
  MB_CHECK(C->rows == m && C->cols == n, "mb_gemm: output is %lld x %lld, expected %lld x %lld",
           (long long)C->rows, (long long)C->cols, (long long)m, (long long)n);
  if (trans_a && !trans_b && A->global_rows >= 0) {
    // contraction over the cell axis of a sharded operand: fixed chunk tree, summed over all ranks
    MB_CHECK(alpha == 1.0 && beta == 0.0, "mb_gemm: a product contracted over sharded cells takes alpha = 1, beta = 0");
    MB_CHECK(B->global_rows == A->global_rows && B->row_lo == A->row_lo, "mb_gemm: operands are sharded differently");
    mb_chunks g;
    MB_TRY(mb_chunk_grid(ctx, A, &g));
    return mb_gemm_tn_cells(ctx, g, m, n, A->p, A->cols, B->p, B->cols, C->p, C->cols, false);
  }
  if (ka == 0) {
    if (beta == 0.0) return mb_mat_fill(ctx, C, 0.0);
  }
  if (!trans_a && A->global_rows >= 0 && (beta == 0.0 || beta == 1.0) && mb_i8_nt_usable(ctx, A->global_rows, n, ka)) {
    // tall (cells x k) times (k x n): tcgen05 int8 digit slices (decomposition.py:123 / :265, the Nystroem factor Q V)
    const double* Bt = B->p;
    int64_t ldbt = B->cols;
    if (!trans_b) {  // the int8 kernel takes B as n x k (k contiguous): transpose the small operand
      double* tmp;
      MB_TRY(mb_scratch(ctx, (size_t)n * ka * sizeof(double), &tmp));
      mb_mat tv = {tmp, n, ka, ctx, false};
      MB_TRY(mb_mat_transpose(ctx, B, &tv));
      Bt = tmp;
      ldbt = ka;
    }
    MB_TRY(mb_i8_gemm_nt(ctx, m, n, ka, alpha, A->p, A->cols, Bt, ldbt, beta == 1.0 ? 1 : 0, C->p, C->cols));
    return mb_i8_check(ctx);
  }
  if (!trans_a && A->global_rows >= 0)  // output rows are cells: an output row must not depend on the local row count
    return mb_gemm_rows_raw(ctx, false, trans_b == 0, m, n, ka, alpha, A->p, A->cols, B->p, B->cols, beta, C->p, C->cols);
  return mb_gemm_raw(ctx, trans_a != 0, trans_b == 0, m, n, ka, alpha, A->p, A->cols, B->p, B->cols, beta, C->p,
                     C->cols, false);
}
