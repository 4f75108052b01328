// K4 (Gram over the cell axis + all-reduce), the Ridge initial value, and the symmetric
// eigen-decomposition used by the Nystroem paths.
#include <dlfcn.h>

#include "mb_common.cuh"

extern "C" int mb_gram(mb_ctx* ctx, const mb_mat* L, mb_mat* G) {
  MB_RANGE("mellon_b200: K4 gram");
  MB_CHECK(ctx && L && G, "mb_gram: null argument");
  MB_CHECK(G->rows == L->cols && G->cols == L->cols, "mb_gram: G must be %lld x %lld", (long long)L->cols,
           (long long)L->cols);
  MB_CUDA(cudaSetDevice(ctx->device));
  const int64_t r = L->cols;
  if (r == 0) return 0;
  // G = L^T L: one k-major lower-tile product per chunk of the fixed reduction tree, chunk sums combined by the tree
  // (across ranks when L is a row block of a sharded matrix), then mirrored
  mb_chunks g;
  MB_TRY(mb_chunk_grid(ctx, L, &g));
  MB_TRY(mb_gemm_tn_cells(ctx, g, r, r, L->p, r, L->p, r, G->p, r, true));
  return mb_mat_symmetrize(ctx, G);
}

extern "C" int mb_ridge_init(mb_ctx* ctx, const mb_mat* L, const mb_mat* t, double* z0_host) {
  MB_RANGE("mellon_b200: ridge_init (K4 + K2 + TRSV)");
  MB_CHECK(ctx && L && t && z0_host, "mb_ridge_init: null argument");
  MB_CHECK(t->rows * t->cols == L->rows, "mb_ridge_init: target has %lld entries for %lld cells",
           (long long)(t->rows * t->cols), (long long)L->rows);
  MB_CUDA(cudaSetDevice(ctx->device));
  const int64_t r = L->cols;
  if (r == 0) return 0;
  mb_mat *G = nullptr, *b = nullptr;
  int rc = mb_mat_alloc(ctx, r, r, &G);
  if (rc == 0) rc = mb_mat_alloc(ctx, r, 1, &b);
  if (rc == 0) rc = mb_gram(ctx, L, G);
  if (rc == 0) rc = mb_mat_add_diag(ctx, G, 1.0);  // Ridge(alpha=1)
  if (rc == 0) rc = mb_gemv_t(ctx, L, t, b);
  if (rc == 0) {
    rc = mb_potrf(ctx, G);
    if (rc > 0) {
      mb_set_error("mb_ridge_init: L^T L + I is not positive definite (pivot %d)", rc);
      rc = -4;
    }
  }
  if (rc == 0) rc = mb_tri_solve(ctx, G, 0, b);
  if (rc == 0) rc = mb_tri_solve(ctx, G, 1, b);
  if (rc == 0) rc = mb_mat_download(ctx, b, z0_host, 0, r);
  mb_mat_free(ctx, G);
  mb_mat_free(ctx, b);
  return rc;
}

// ---- symmetric eigen-decomposition ------------------------------------------------------------
// LIBRARY CALL (cuSOLVER Dsyevd, opened with dlopen): the reference's `eigh`
// (decomposition.py:50) is LAPACK syevd; the O(N M r) products around it are this
// library's own kernels.  DESIGN.md lists this as the one non-hand-written device routine.
namespace {
typedef void* cusolverDnHandle_t;
struct CusolverApi {
  void* handle;
  int (*Create)(cusolverDnHandle_t*);
  int (*Destroy)(cusolverDnHandle_t);
  int (*SetStream)(cusolverDnHandle_t, cudaStream_t);
  int (*Dsyevd_bufferSize)(cusolverDnHandle_t, int, int, int, const double*, int, const double*, int*);
  int (*Dsyevd)(cusolverDnHandle_t, int, int, int, double*, int, double*, double*, int, int*);
};
CusolverApi g_cs = {};
cusolverDnHandle_t g_cs_handle = nullptr;

int load_cusolver() {
  if (g_cs.handle) return 0;
  const char* names[] = {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11"};
  void* h = nullptr;
  for (const char* nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
    if (h) break;
  }
  MB_CHECK(h, "could not dlopen libcusolver.so.11: %s", dlerror());
#define SYM(field, name)                                               \
  g_cs.field = reinterpret_cast<decltype(g_cs.field)>(dlsym(h, name)); \
  MB_CHECK(g_cs.field, "libcusolver is missing symbol %s", name);
  SYM(Create, "cusolverDnCreate")
  SYM(Destroy, "cusolverDnDestroy")
  SYM(SetStream, "cusolverDnSetStream")
  SYM(Dsyevd_bufferSize, "cusolverDnDsyevd_bufferSize")
  SYM(Dsyevd, "cusolverDnDsyevd")
#undef SYM
  g_cs.handle = h;
  return 0;
}
}  // namespace

extern "C" int mb_syevd(mb_ctx* ctx, mb_mat* a, mb_mat* w) {
  MB_RANGE("mellon_b200: syevd (cuSOLVER)");
  MB_CHECK(ctx && a && w, "mb_syevd: null argument");
  MB_CHECK(a->rows == a->cols && w->rows * w->cols == a->rows, "mb_syevd: shape mismatch");
  MB_CUDA(cudaSetDevice(ctx->device));
  const int n = (int)a->rows;
  if (n == 0) return 0;
  MB_TRY(load_cusolver());
  if (!g_cs_handle) MB_CHECK(g_cs.Create(&g_cs_handle) == 0, "cusolverDnCreate failed");
  MB_CHECK(g_cs.SetStream(g_cs_handle, ctx->stream) == 0, "cusolverDnSetStream failed");
  // row-major symmetric == column-major symmetric; eigenvectors come back as COLUMNS of the
  // column-major matrix, i.e. ROWS of our row-major buffer -> transpose afterwards.
  int lwork = 0;
  const int JOBZ_VECTOR = 1, UPLO_LOWER = 0;
  MB_CHECK(g_cs.Dsyevd_bufferSize(g_cs_handle, JOBZ_VECTOR, UPLO_LOWER, n, a->p, n, w->p, &lwork) == 0,
           "cusolverDnDsyevd_bufferSize failed");
  double* work;
  MB_TRY(mb_scratch(ctx, ((size_t)lwork + (size_t)n * n + 8) * sizeof(double), &work));
  int* info_dev = reinterpret_cast<int*>(work + lwork);
  double* tmp = work + lwork + 4;
  mb_prof_span sp = {MB_PROF_EIGH, nullptr, nullptr};   // the library call is timed like a kernel class of its own
  if (ctx->prof_on) {
    sp.a = mb_prof_event(ctx);
    sp.b = mb_prof_event(ctx);
    cudaEventRecord(sp.a, ctx->stream);
    ctx->prof_work[MB_PROF_EIGH] += 9.0 * (double)n * (double)n * (double)n;   // ~9 n^3 flops (tridiagonalisation + D&C + back-transform)
  }
  int st = g_cs.Dsyevd(g_cs_handle, JOBZ_VECTOR, UPLO_LOWER, n, a->p, n, w->p, work, lwork, info_dev);
  if (ctx->prof_on) {
    cudaEventRecord(sp.b, ctx->stream);
    ctx->prof_spans.push_back(sp);
  }
  MB_CHECK(st == 0, "cusolverDnDsyevd failed with status %d", st);
  int info = 0;
  MB_CUDA(cudaMemcpyAsync(&info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  MB_CHECK(info == 0, "cusolverDnDsyevd did not converge (info=%d)", info);
  mb_mat tv = {tmp, a->rows, a->cols, ctx, false};
  MB_TRY(mb_mat_transpose(ctx, a, &tv));
  return mb_mat_copy(ctx, &tv, a);
}
