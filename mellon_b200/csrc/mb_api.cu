// Context, device-matrix management and small utility kernels.
#include <stdarg.h>

#include "mb_common.cuh"

static thread_local char g_err[1024] = "";

void mb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* mb_last_error(void) { return g_err; }
static int prof_resolve(mb_ctx* c);
extern "C" int mb_version(void) { return 200; }

#ifndef MB_SOURCE_HASH
#define MB_SOURCE_HASH "unknown"
#endif
// sha256 (first 16 hex digits) of the sources this library was compiled from: __graft_entry__.build() and
// tests/test_abi.py compare it with the tree, so a stale prebuilt library cannot pass for the current code
extern "C" const char* mb_source_hash(void) { return MB_SOURCE_HASH; }

extern "C" int mb_device_count(int* n) {
  MB_CHECK(n != nullptr, "mb_device_count: null output");
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) {
    *n = 0;
    mb_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return -1;
  }
  return 0;
}

extern "C" int mb_ctx_create(int device, mb_ctx** out) {
  MB_CHECK(out != nullptr, "mb_ctx_create: null output");
  int n = 0;
  MB_CUDA(cudaGetDeviceCount(&n));
  MB_CHECK(device >= 0 && device < n, "mb_ctx_create: device %d out of range (have %d)", device, n);
  MB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MB_CUDA(cudaGetDeviceProperties(&prop, device));
  MB_CHECK(prop.major >= 10, "mb_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
           device, prop.major, prop.minor);
  mb_ctx* c = new mb_ctx();
  c->device = device;
  c->n_sm = prop.multiProcessorCount;
  c->cache_cap = (size_t)(0.45 * (double)prop.totalGlobalMem);  // at most 45 % of HBM parked in the block cache
  MB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  MB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 16; i++)
    for (int j = 0; j < 2; j++) MB_CUDA(cudaEventCreate(&c->timer_ev[i][j]));
  MB_CUDA(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
  MB_CUDA(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
  c->rank = 0;
  c->world = 1;
  *out = c;
  return 0;
}

void mb_cache_flush(mb_ctx* c) {
  for (auto& kv : c->block_cache) cudaFree(kv.second);
  c->block_cache.clear();
  c->cached_bytes = 0;
}

void mb_invalidate_graphs(mb_ctx* c) {
  for (auto& kv : c->potrf_graphs) cudaGraphExecDestroy(kv.second.first);
  c->potrf_graphs.clear();
}

cudaError_t mb_dev_malloc(mb_ctx* c, void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess && !c->block_cache.empty()) {
    cudaGetLastError();
    mb_cache_flush(c);
    e = cudaMalloc(p, bytes);
  }
  return e;
}

extern "C" int mb_ctx_destroy(mb_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  mb_comm_destroy(c);
  mb_cache_flush(c);
  if (c->scratch) cudaFree(c->scratch);
  if (c->flush_buf) cudaFree(c->flush_buf);
  if (c->gemm_ws) cudaFree(c->gemm_ws);
  if (c->i8_ws) cudaFree(c->i8_ws);
  if (c->i8_tiles) cudaFree(c->i8_tiles);
  if (c->i8_status) cudaFree(c->i8_status);
  if (c->i8_side) {
    cudaStreamSynchronize(c->i8_side);
    cudaStreamDestroy(c->i8_side);
    cudaEventDestroy(c->i8_ev_start);
    for (int b = 0; b < 2; b++) {
      cudaEventDestroy(c->i8_ev_packed[b]);
      cudaEventDestroy(c->i8_ev_free[b]);
    }
  }
  if (c->trsm_ws) cudaFree(c->trsm_ws);
  mb_invalidate_graphs(c);
  if (c->potrf_buf) cudaFree(c->potrf_buf);
  if (c->potrf_info) cudaFree(c->potrf_info);
  if (c->pinned) cudaFreeHost(c->pinned);
  prof_resolve(c);
  for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
  for (int i = 0; i < 16; i++)
    for (int j = 0; j < 2; j++) cudaEventDestroy(c->timer_ev[i][j]);
  cudaEventDestroy(c->ev_a);
  cudaEventDestroy(c->ev_b);
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->copy_stream);
  delete c;
  return 0;
}

extern "C" int mb_ctx_sync(mb_ctx* c) {
  MB_CHECK(c, "null ctx");
  MB_CUDA(cudaSetDevice(c->device));
  MB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int mb_ctx_info(mb_ctx* c, int* device, int* n_sm, int64_t* free_b, int64_t* total_b) {
  MB_CHECK(c, "null ctx");
  MB_CUDA(cudaSetDevice(c->device));
  size_t f = 0, t = 0;
  MB_CUDA(cudaMemGetInfo(&f, &t));
  if (device) *device = c->device;
  if (n_sm) *n_sm = c->n_sm;
  if (free_b) *free_b = (int64_t)f;
  if (total_b) *total_b = (int64_t)t;
  return 0;
}

extern "C" int64_t mb_ctx_launch_count(mb_ctx* c) { return c ? c->launches : -1; }

extern "C" int mb_timer_start(mb_ctx* c, int slot) {
  MB_CHECK(c && slot >= 0 && slot < 16, "mb_timer_start: bad slot");
  MB_CUDA(cudaEventRecord(c->timer_ev[slot][0], c->stream));
  return 0;
}

extern "C" int mb_timer_stop(mb_ctx* c, int slot, double* ms) {
  MB_CHECK(c && slot >= 0 && slot < 16 && ms, "mb_timer_stop: bad arguments");
  MB_CUDA(cudaEventRecord(c->timer_ev[slot][1], c->stream));
  MB_CUDA(cudaEventSynchronize(c->timer_ev[slot][1]));
  float f = 0;
  MB_CUDA(cudaEventElapsedTime(&f, c->timer_ev[slot][0], c->timer_ev[slot][1]));
  *ms = f;
  return 0;
}

cudaEvent_t mb_prof_event(mb_ctx* c) {
  if (!c->prof_pool.empty()) {
    cudaEvent_t e = c->prof_pool.back();
    c->prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

static int prof_resolve(mb_ctx* c) {
  if (c->prof_spans.empty()) return 0;
  MB_CUDA(cudaStreamSynchronize(c->stream));
  for (const mb_prof_span& sp : c->prof_spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
      c->prof_count[sp.cls]++;
      c->prof_ms[sp.cls] += ms;
    }
    c->prof_pool.push_back(sp.a);
    c->prof_pool.push_back(sp.b);
  }
  c->prof_spans.clear();
  return 0;
}

extern "C" int mb_prof_enable(mb_ctx* c, int on) {
  MB_CHECK(c, "mb_prof_enable: null ctx");
  MB_CUDA(cudaSetDevice(c->device));
  MB_TRY(prof_resolve(c));
  c->prof_on = on != 0;
  return 0;
}

extern "C" int mb_prof_reset(mb_ctx* c) {
  MB_CHECK(c, "mb_prof_reset: null ctx");
  MB_CUDA(cudaSetDevice(c->device));
  MB_TRY(prof_resolve(c));
  for (int i = 0; i < MB_PROF_NCLS; i++) {
    c->prof_count[i] = 0;
    c->prof_ms[i] = 0.0;
    c->prof_work[i] = 0.0;
  }
  return 0;
}

extern "C" int mb_prof_read(mb_ctx* c, int cls, int64_t* count, double* ms, double* work) {
  MB_CHECK(c && cls >= 0 && cls < MB_PROF_NCLS, "mb_prof_read: bad class %d", cls);
  MB_CUDA(cudaSetDevice(c->device));
  MB_TRY(prof_resolve(c));
  if (count) *count = c->prof_count[cls];
  if (ms) *ms = c->prof_ms[cls];
  if (work) *work = c->prof_work[cls];
  return 0;
}

// host-side NVTX ranges for the callers above the ABI (the L-BFGS-B loop, the estimator stages)
extern "C" int mb_range_push(const char* name) {
  nvtxRangePushA(name ? name : "mellon_b200");
  return 0;
}
extern "C" int mb_range_pop(void) {
  nvtxRangePop();
  return 0;
}

extern "C" int mb_set_option(mb_ctx* c, const char* key, int value) {
  MB_CHECK(c && key, "mb_set_option: null");
  if (!strcmp(key, "gemm")) c->opt_gemm = value;
  else if (!strcmp(key, "cov")) c->opt_cov = value;
  else if (!strcmp(key, "trsm")) c->opt_trsm = value;
  else if (!strcmp(key, "graph")) c->opt_graph = value;
  else if (!strcmp(key, "lossgrad")) c->opt_lossgrad = value;
  else if (!strcmp(key, "i8")) c->opt_i8 = value;
  else if (!strcmp(key, "cov_i8")) c->opt_cov_i8 = value;
  else if (!strcmp(key, "i8_issuers")) c->opt_i8_issuers = value;
  else if (!strcmp(key, "i8_overlap")) c->opt_i8_overlap = value;
  else MB_CHECK(false, "mb_set_option: unknown key %s", key);
  return 0;
}

int mb_scratch(mb_ctx* c, size_t bytes, double** out) {
  if (bytes > c->scratch_bytes) {
    MB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->scratch) MB_CUDA(cudaFree(c->scratch));
    c->scratch = nullptr;
    c->scratch_bytes = 0;
    size_t want = bytes + bytes / 4;
    MB_CUDA(mb_dev_malloc(c, (void**)&c->scratch, want));
    c->scratch_bytes = want;
  }
  *out = c->scratch;
  return 0;
}

int mb_pinned(mb_ctx* c, size_t bytes, double** out) {
  if (bytes > c->pinned_bytes) {
    MB_CUDA(cudaStreamSynchronize(c->stream));
    MB_CUDA(cudaStreamSynchronize(c->copy_stream));
    if (c->pinned) MB_CUDA(cudaFreeHost(c->pinned));
    c->pinned = nullptr;
    c->pinned_bytes = 0;
    MB_CUDA(cudaMallocHost(&c->pinned, bytes));
    c->pinned_bytes = bytes;
  }
  *out = c->pinned;
  return 0;
}

// ---- utility kernels --------------------------------------------------------------------
__global__ void k_fill(double* p, int64_t n, double v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) p[i] = v;
}

__global__ void k_add_diag(double* a, int64_t n, int64_t ld, double v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i * ld + i] += v;
}

__global__ void k_add_diag_vec(double* a, int64_t n, int64_t ld, const double* __restrict__ v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i * ld + i] += v[i];
}

__global__ void k_scale_rows(double* a, int64_t rows, int64_t cols, const double* __restrict__ s) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
  int64_t n = rows * cols;
  for (; i < n; i += st) a[i] *= s[i / cols];
}

__global__ void k_transpose(const double* __restrict__ src, double* __restrict__ dst, int64_t rows,
                            int64_t cols) {
  __shared__ double tile[32][33];
  int64_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[c * rows + r] = tile[threadIdx.x][j];
  }
}

__global__ void k_scale_cols(double* a, int64_t rows, int64_t cols, const double* __restrict__ s) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
  int64_t n = rows * cols;
  for (; i < n; i += st) a[i] *= s[i % cols];
}

__global__ void k_copy_cols(const double* __restrict__ src, int64_t rows, int64_t scols, int64_t c0,
                            int64_t ncols, double* __restrict__ dst) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
  int64_t n = rows * ncols;
  for (; i < n; i += st) {
    int64_t r = i / ncols, c = i % ncols;
    dst[i] = src[r * scols + c0 + c];
  }
}

__global__ void k_symmetrize(double* a, int64_t n) {
  // upper(i, j) <- lower(j, i), tiled through shared memory so both sides coalesce
  __shared__ double tile[32][33];
  int64_t bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;  // read lower tiles only
  int64_t r0 = bi * 32, c0 = bj * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t r = r0 + j, c = c0 + threadIdx.x;
    if (r < n && c < n) tile[j][threadIdx.x] = a[r * n + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t r = c0 + j, c = r0 + threadIdx.x;  // transposed position
    if (r < n && c < n && c > r) a[r * n + c] = tile[threadIdx.x][j];
  }
}

__global__ void k_scale(double* a, int64_t n, double s) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) a[i] *= s;
}

// one warp per row, fixed-order lane sums + shuffle tree
__global__ void k_row_sumsq(const double* __restrict__ a, int64_t rows, int64_t cols, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < rows; i += nwarps) {
    const double* row = a + i * cols;
    double s = 0.0;
    for (int64_t c = lane; c < cols; c += 32) s = fma(row[c], row[c], s);
    s = warp_sum(s);
    if (lane == 0) out[i] = s;
  }
}

__global__ void k_flush(double* p, int64_t n, double v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) p[i] = v + (double)i;
}

extern "C" int mb_flush_l2(mb_ctx* c) {
  MB_CHECK(c, "null ctx");
  MB_CUDA(cudaSetDevice(c->device));
  if (!c->flush_buf) {
    c->flush_bytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
    MB_CUDA(cudaMalloc(&c->flush_buf, c->flush_bytes));
  }
  MB_LAUNCH(c, k_flush, c->n_sm * 8, 512, 0, c->flush_buf, (int64_t)(c->flush_bytes / 8), 1.0);
  return 0;
}

extern "C" int mb_host_alloc(int64_t bytes, void** out) {
  MB_CHECK(out && bytes >= 0, "mb_host_alloc: bad argument");
  *out = nullptr;
  if (bytes == 0) return 0;
  MB_CUDA(cudaMallocHost(out, (size_t)bytes));
  return 0;
}

extern "C" int mb_host_free(void* p) {
  if (p) MB_CUDA(cudaFreeHost(p));
  return 0;
}

// ---- matrices ------------------------------------------------------------------------------
extern "C" int mb_mat_alloc(mb_ctx* c, int64_t rows, int64_t cols, mb_mat** out) {
  MB_CHECK(c && out, "mb_mat_alloc: null argument");
  MB_CHECK(rows >= 0 && cols >= 0, "mb_mat_alloc: negative shape (%lld, %lld)", (long long)rows,
           (long long)cols);
  MB_CUDA(cudaSetDevice(c->device));
  mb_mat* m = new mb_mat();
  m->rows = rows;
  m->cols = cols;
  m->ctx = c;
  m->owns = true;
  m->p = nullptr;
  size_t bytes = (size_t)rows * (size_t)cols * sizeof(double);
  if (bytes > 0) {
    // a cached block of the same size class (at most 1/8 larger)?
    auto it = c->block_cache.lower_bound(bytes);
    if (it != c->block_cache.end() && it->first <= bytes + bytes / 8) {
      m->p = it->second;
      m->alloc_bytes = it->first;
      c->cached_bytes -= it->first;
      c->block_cache.erase(it);
      *out = m;
      return 0;
    }
    cudaError_t e = cudaMalloc(&m->p, bytes);
    if (e != cudaSuccess && !c->block_cache.empty()) {
      cudaGetLastError();
      mb_cache_flush(c);
      e = cudaMalloc(&m->p, bytes);
    }
    m->alloc_bytes = bytes;
    if (e != cudaSuccess) {
      delete m;
      mb_set_error("mb_mat_alloc: cudaMalloc of %zu bytes (%lld x %lld f64) failed: %s", bytes,
                   (long long)rows, (long long)cols, cudaGetErrorString(e));
      cudaGetLastError();
      return -1;
    }
  }
  *out = m;
  return 0;
}

extern "C" int mb_mat_free(mb_ctx* c, mb_mat* m) {
  if (!m) return 0;
  if (c) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
  }
  if (m->owns && m->p) {
    if (c && m->alloc_bytes >= ((size_t)1 << 20) && c->cached_bytes + m->alloc_bytes <= c->cache_cap) {
      c->block_cache.emplace(m->alloc_bytes, m->p);
      c->cached_bytes += m->alloc_bytes;
    } else {
      cudaFree(m->p);
    }
  }
  delete m;
  return 0;
}

extern "C" int mb_mat_shape(const mb_mat* m, int64_t* rows, int64_t* cols) {
  MB_CHECK(m, "mb_mat_shape: null matrix");
  if (rows) *rows = m->rows;
  if (cols) *cols = m->cols;
  return 0;
}

extern "C" int mb_mat_upload(mb_ctx* c, mb_mat* m, const double* host, int64_t row0, int64_t nrows) {
  MB_CHECK(c && m && (host || nrows == 0), "mb_mat_upload: null argument");
  MB_CHECK(row0 >= 0 && nrows >= 0 && row0 + nrows <= m->rows,
           "mb_mat_upload: rows [%lld, %lld) outside matrix with %lld rows", (long long)row0,
           (long long)(row0 + nrows), (long long)m->rows);
  MB_CUDA(cudaSetDevice(c->device));
  size_t bytes = (size_t)nrows * m->cols * sizeof(double);
  if (bytes == 0) return 0;
  MB_CUDA(cudaMemcpyAsync(m->p + row0 * m->cols, host, bytes, cudaMemcpyHostToDevice, c->stream));
  MB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int mb_mat_download(mb_ctx* c, const mb_mat* m, double* host, int64_t row0, int64_t nrows) {
  MB_CHECK(c && m && (host || nrows == 0), "mb_mat_download: null argument");
  MB_CHECK(row0 >= 0 && nrows >= 0 && row0 + nrows <= m->rows,
           "mb_mat_download: rows [%lld, %lld) outside matrix with %lld rows", (long long)row0,
           (long long)(row0 + nrows), (long long)m->rows);
  MB_CUDA(cudaSetDevice(c->device));
  size_t bytes = (size_t)nrows * m->cols * sizeof(double);
  if (bytes == 0) return 0;
  MB_CUDA(cudaMemcpyAsync(host, m->p + row0 * m->cols, bytes, cudaMemcpyDeviceToHost, c->stream));
  MB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int mb_mat_copy(mb_ctx* c, const mb_mat* src, mb_mat* dst) {
  MB_CHECK(c && src && dst, "mb_mat_copy: null argument");
  MB_CHECK(src->rows == dst->rows && src->cols == dst->cols, "mb_mat_copy: shape mismatch");
  MB_CUDA(cudaSetDevice(c->device));
  size_t bytes = (size_t)src->rows * src->cols * sizeof(double);
  if (bytes) MB_CUDA(cudaMemcpyAsync(dst->p, src->p, bytes, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

extern "C" int mb_mat_fill(mb_ctx* c, mb_mat* m, double v) {
  MB_CHECK(c && m, "mb_mat_fill: null argument");
  MB_CUDA(cudaSetDevice(c->device));
  int64_t n = m->rows * m->cols;
  if (n == 0) return 0;
  int grid = (int)min((int64_t)c->n_sm * 8, ceil_div64(n, 256));
  MB_LAUNCH(c, k_fill, grid, 256, 0, m->p, n, v);
  return 0;
}

extern "C" int mb_mat_transpose(mb_ctx* c, const mb_mat* src, mb_mat* dst) {
  MB_CHECK(c && src && dst, "mb_mat_transpose: null argument");
  MB_CHECK(src->rows == dst->cols && src->cols == dst->rows, "mb_mat_transpose: shape mismatch");
  MB_CUDA(cudaSetDevice(c->device));
  if (src->rows == 0 || src->cols == 0) return 0;
  dim3 grid((unsigned)ceil_div64(src->cols, 32), (unsigned)ceil_div64(src->rows, 32));
  MB_CHECK(grid.y < 65536, "mb_mat_transpose: too many rows (%lld)", (long long)src->rows);
  MB_LAUNCH(c, k_transpose, grid, dim3(32, 8), 0, src->p, dst->p, src->rows, src->cols);
  return 0;
}

extern "C" int mb_mat_add_diag(mb_ctx* c, mb_mat* a, double v) {
  MB_CHECK(c && a, "mb_mat_add_diag: null argument");
  MB_CHECK(a->rows == a->cols, "mb_mat_add_diag: matrix is %lld x %lld, not square",
           (long long)a->rows, (long long)a->cols);
  MB_CUDA(cudaSetDevice(c->device));
  if (a->rows == 0) return 0;
  MB_LAUNCH(c, k_add_diag, (int)ceil_div64(a->rows, 256), 256, 0, a->p, a->rows, a->cols, v);
  return 0;
}

extern "C" int mb_mat_add_diag_vec(mb_ctx* c, mb_mat* a, const mb_mat* v) {
  MB_CHECK(c && a && v, "mb_mat_add_diag_vec: null argument");
  MB_CHECK(a->rows == a->cols, "mb_mat_add_diag_vec: matrix is %lld x %lld, not square",
           (long long)a->rows, (long long)a->cols);
  MB_CHECK(v->rows * v->cols == a->rows, "mb_mat_add_diag_vec: %lld diagonal entries for %lld rows",
           (long long)(v->rows * v->cols), (long long)a->rows);
  MB_CUDA(cudaSetDevice(c->device));
  if (a->rows == 0) return 0;
  MB_LAUNCH(c, k_add_diag_vec, (int)ceil_div64(a->rows, 256), 256, 0, a->p, a->rows, a->cols, v->p);
  return 0;
}

extern "C" int mb_mat_scale_rows(mb_ctx* c, mb_mat* a, const mb_mat* s) {
  MB_CHECK(c && a && s, "mb_mat_scale_rows: null argument");
  MB_CHECK(s->rows * s->cols == a->rows, "mb_mat_scale_rows: scale has %lld entries for %lld rows",
           (long long)(s->rows * s->cols), (long long)a->rows);
  MB_CUDA(cudaSetDevice(c->device));
  int64_t n = a->rows * a->cols;
  if (n == 0) return 0;
  int grid = (int)min((int64_t)c->n_sm * 8, ceil_div64(n, 256));
  MB_LAUNCH(c, k_scale_rows, grid, 256, 0, a->p, a->rows, a->cols, s->p);
  return 0;
}

extern "C" int mb_mat_scale_cols(mb_ctx* c, mb_mat* a, const mb_mat* s) {
  MB_CHECK(c && a && s, "mb_mat_scale_cols: null argument");
  MB_CHECK(s->rows * s->cols == a->cols, "mb_mat_scale_cols: scale has %lld entries for %lld columns",
           (long long)(s->rows * s->cols), (long long)a->cols);
  MB_CUDA(cudaSetDevice(c->device));
  int64_t n = a->rows * a->cols;
  if (n == 0) return 0;
  int grid = (int)min((int64_t)c->n_sm * 8, ceil_div64(n, 256));
  MB_LAUNCH(c, k_scale_cols, grid, 256, 0, a->p, a->rows, a->cols, s->p);
  return 0;
}

extern "C" int mb_mat_copy_cols(mb_ctx* c, const mb_mat* src, int64_t c0, int64_t ncols, mb_mat* dst) {
  MB_CHECK(c && src && dst, "mb_mat_copy_cols: null argument");
  MB_CHECK(c0 >= 0 && ncols >= 0 && c0 + ncols <= src->cols && dst->rows == src->rows &&
               dst->cols == ncols,
           "mb_mat_copy_cols: bad column range / destination shape");
  MB_CUDA(cudaSetDevice(c->device));
  int64_t n = src->rows * ncols;
  if (n == 0) return 0;
  int grid = (int)min((int64_t)c->n_sm * 8, ceil_div64(n, 256));
  MB_LAUNCH(c, k_copy_cols, grid, 256, 0, src->p, src->rows, src->cols, c0, ncols, dst->p);
  return 0;
}

extern "C" int mb_mat_copy_rows(mb_ctx* c, const mb_mat* src, int64_t r0, int64_t nrows, mb_mat* dst) {
  MB_CHECK(c && src && dst, "mb_mat_copy_rows: null argument");
  MB_CHECK(r0 >= 0 && nrows >= 0 && r0 + nrows <= src->rows && dst->rows == nrows && dst->cols == src->cols,
           "mb_mat_copy_rows: rows [%lld, %lld) of a %lld x %lld matrix into %lld x %lld", (long long)r0,
           (long long)(r0 + nrows), (long long)src->rows, (long long)src->cols, (long long)dst->rows, (long long)dst->cols);
  MB_CUDA(cudaSetDevice(c->device));
  if (nrows * src->cols == 0) return 0;
  MB_CUDA(cudaMemcpyAsync(dst->p, src->p + r0 * src->cols, (size_t)(nrows * src->cols) * sizeof(double),
                          cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

extern "C" int mb_mat_symmetrize(mb_ctx* c, mb_mat* a) {
  MB_CHECK(c && a, "mb_mat_symmetrize: null argument");
  MB_CHECK(a->rows == a->cols, "mb_mat_symmetrize: not square");
  MB_CUDA(cudaSetDevice(c->device));
  if (a->rows == 0) return 0;
  unsigned t = (unsigned)ceil_div64(a->rows, 32);
  MB_LAUNCH(c, k_symmetrize, dim3(t, t), dim3(32, 8), 0, a->p, a->rows);
  return 0;
}

extern "C" int mb_mat_scale(mb_ctx* c, mb_mat* a, double s) {
  MB_CHECK(c && a, "mb_mat_scale: null argument");
  MB_CUDA(cudaSetDevice(c->device));
  int64_t n = a->rows * a->cols;
  if (n == 0) return 0;
  int grid = (int)min((int64_t)c->n_sm * 8, ceil_div64(n, 256));
  MB_LAUNCH(c, k_scale, grid, 256, 0, a->p, n, s);
  return 0;
}

// a <- a (+ | *) b, a (+ | *) value, or a ** value, elementwise
__global__ void k_combine(double* __restrict__ a, const double* __restrict__ b, int64_t n, int op, double value) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) {
    const double x = a[i], y = b ? b[i] : value;
    a[i] = (op == MB_OP_ADD) ? x + y : (op == MB_OP_MUL) ? x * y : pow(x, value);
  }
}

extern "C" int mb_mat_combine(mb_ctx* c, int op, mb_mat* a, const mb_mat* b, double value) {
  MB_CHECK(c && a, "mb_mat_combine: null argument");
  MB_CHECK(op == MB_OP_ADD || op == MB_OP_MUL || op == MB_OP_POW, "mb_mat_combine: op %d is not ADD / MUL / POW", op);
  MB_CHECK(!b || (op != MB_OP_POW && b->rows == a->rows && b->cols == a->cols), "mb_mat_combine: operand mismatch");
  MB_CUDA(cudaSetDevice(c->device));
  int64_t n = a->rows * a->cols;
  if (n == 0) return 0;
  int grid = (int)min((int64_t)c->n_sm * 8, ceil_div64(n, 256));
  MB_LAUNCH(c, k_combine, grid, 256, 0, a->p, b ? b->p : nullptr, n, op, value);
  return 0;
}

extern "C" int mb_mat_row_sumsq(mb_ctx* c, const mb_mat* a, mb_mat* out) {
  MB_CHECK(c && a && out, "mb_mat_row_sumsq: null argument");
  MB_CHECK(out->rows * out->cols == a->rows, "mb_mat_row_sumsq: output has %lld entries for %lld rows",
           (long long)(out->rows * out->cols), (long long)a->rows);
  MB_CUDA(cudaSetDevice(c->device));
  if (a->rows == 0) return 0;
  if (a->cols == 0) return mb_mat_fill(c, out, 0.0);
  int grid = (int)min((int64_t)c->n_sm * 8, ceil_div64(a->rows, 8));
  MB_LAUNCH(c, k_row_sumsq, grid, 256, 0, a->p, a->rows, a->cols, out->p);
  return 0;
}
