// Landmark selection, the step before the path: the device side of k-means++ seeding and Lloyd
// iterations as scikit-learn's `k_means(x, k, n_init=1, random_state=seed)` performs them (the reference's
// parameters.py:243-291 calls exactly that).  The random draws, the cumulative sums that turn them into candidate rows
// and the centroid averages stay on the host (mellon_b200/kmeans.py), so the selected rows are scikit-learn's own; the
// O(N T D) / O(N M D) distance work runs here.  The Lloyd assignment step is mb_nn_distances (K1's distance tile with a
// running-minimum epilogue) against the centres.
#include "mb_common.cuh"

namespace {

constexpr int KPP_MAX_T = 16;   // candidates per seeding step: 2 + log(k) <= 16 up to k = 1.2e6

// out[t][i] = min(closest[i], max(xx_i - 2 x_i . c_t + cc_t, 0)); partial[block][t] = sum over the block's rows
__global__ void __launch_bounds__(256)
sqdist_min_kernel(const double* __restrict__ x, const double* __restrict__ xn, int64_t n, int d,
                  const double* __restrict__ c, int T, const double* __restrict__ closest,
                  double* __restrict__ out, double* __restrict__ partial) {
  extern __shared__ double cs[];                 // T x d candidate rows, then T squared norms
  double* cc = cs + (size_t)T * d;
  __shared__ double red[8][KPP_MAX_T];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < T * d; e += 256) cs[e] = c[e];
  __syncthreads();
  if (tid < T) {
    double s = 0.0;
    for (int k = 0; k < d; k++) s = fma(cs[tid * d + k], cs[tid * d + k], s);
    cc[tid] = s;
  }
  __syncthreads();
  const int64_t i = blockIdx.x * 256LL + tid;
  double v[KPP_MAX_T];
#pragma unroll
  for (int t = 0; t < KPP_MAX_T; t++) v[t] = 0.0;
  if (i < n) {
    double dot[KPP_MAX_T];
#pragma unroll
    for (int t = 0; t < KPP_MAX_T; t++) dot[t] = 0.0;
    const double* row = x + i * d;
    for (int k = 0; k < d; k++) {
      const double xv = row[k];
#pragma unroll
      for (int t = 0; t < KPP_MAX_T; t++)
        if (t < T) dot[t] = fma(xv, cs[t * d + k], dot[t]);
    }
    const double xx = xn[i], cl = closest ? closest[i] : 1.7976931348623157e308;
#pragma unroll
    for (int t = 0; t < KPP_MAX_T; t++) {
      if (t < T) {
        const double dist = fmax(xx - 2.0 * dot[t] + cc[t], 0.0);
        v[t] = fmin(cl, dist);
        out[(int64_t)t * n + i] = v[t];
      }
    }
  }
  // block sums in a fixed order: shuffle tree within a warp, then the 8 warps in order
#pragma unroll
  for (int t = 0; t < KPP_MAX_T; t++) {
    if (t < T) {
      const double s = warp_sum(v[t]);
      if (lane == 0) red[warp][t] = s;
    }
  }
  __syncthreads();
  if (tid < T) {
    double s = 0.0;
    for (int w = 0; w < 8; w++) s += red[w][tid];
    partial[(int64_t)blockIdx.x * T + tid] = s;
  }
}

// pot[t] = sum over the blocks, in block order (one thread per candidate: nblocks is a few thousand)
__global__ void pot_sum_kernel(const double* __restrict__ partial, int64_t nblocks, int T, double* __restrict__ pot) {
  const int t = threadIdx.x;
  if (t >= T) return;
  double s = 0.0;
  for (int64_t b = 0; b < nblocks; b++) s += partial[b * T + t];
  pot[t] = s;
}

}  // namespace

extern "C" int mb_sqdist_min(mb_ctx* ctx, const mb_mat* x, const mb_mat* xnorm, const mb_mat* cand, const mb_mat* closest,
                             mb_mat* out, double* pot_host) {
  MB_RANGE("mellon_b200: k-means++ step");
  MB_CHECK(ctx && x && xnorm && cand && out && pot_host, "mb_sqdist_min: null argument");
  const int64_t n = x->rows;
  const int d = (int)x->cols, T = (int)cand->rows;
  MB_CHECK(cand->cols == x->cols && T >= 1 && T <= KPP_MAX_T, "mb_sqdist_min: %d candidates of %lld features (max %d candidates)",
           T, (long long)cand->cols, KPP_MAX_T);
  MB_CHECK(xnorm->rows * xnorm->cols == n && (!closest || closest->rows * closest->cols == n) && out->rows == T &&
               out->cols == n, "mb_sqdist_min: shape mismatch");
  MB_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) {
    for (int t = 0; t < T; t++) pot_host[t] = 0.0;
    return 0;
  }
  const int64_t nblocks = ceil_div64(n, 256);
  double* ws;
  MB_TRY(mb_scratch(ctx, ((size_t)nblocks * T + KPP_MAX_T) * sizeof(double), &ws));
  const size_t smem = ((size_t)T * d + T) * sizeof(double);
  MB_CHECK(smem <= 48 * 1024, "mb_sqdist_min: %d features are too many", d);
  MB_LAUNCH(ctx, sqdist_min_kernel, (unsigned)nblocks, 256, smem, x->p, xnorm->p, n, d, cand->p, T,
            closest ? closest->p : nullptr, out->p, ws);
  MB_LAUNCH(ctx, pot_sum_kernel, 1, 32, 0, ws, nblocks, T, ws + nblocks * T);
  MB_CUDA(cudaMemcpyAsync(pot_host, ws + nblocks * T, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
