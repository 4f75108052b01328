// Shared internals of libmellon_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/mellon_b200.h"

struct mb_mat {
  double* p;
  int64_t rows, cols;
  mb_ctx* ctx;
  bool owns;
  size_t alloc_bytes = 0;  // size of the device block behind p (>= rows * cols * 8 when it came from the cache)
  // row block [row_lo, row_lo + rows) of a matrix of `global_rows` rows whose cell axis is sharded over the ranks
  // (mb_mat_set_shard); -1 = not sharded: reductions over its rows stay on this rank
  int64_t global_rows = -1;
  int64_t row_lo = 0;
};

// ---- fixed reduction tree over the cell axis ------------------------------------------------------------------
// Every sum over cells (Gram, L^T t, loss + gradient, Hessian diagonal) is formed the same way whatever the number
// of ranks: the GLOBAL cell axis is cut into MB_NCHUNK chunks of ceil(G / MB_NCHUNK) rows, each chunk is summed in
// an order that depends on the chunk alone, and the chunk sums are combined by one pairwise tree over the chunk
// index.  Rank boundaries fall on chunk boundaries (backend.row_block), so a rank owns a contiguous range of leaves.
constexpr int MB_NCHUNK = 32;
struct mb_chunks {
  int64_t G;        // global rows
  int64_t cr;       // rows per chunk
  int64_t row_lo;   // first global row held locally
  int64_t rows;     // local rows
  int c_lo, c_hi;   // local chunk range [c_lo, c_hi) in global chunk ids
  bool sharded;     // combine across ranks
  // local row range of global chunk c (empty when c lies beyond G)
  inline void range(int c, int64_t* i0, int64_t* i1) const {
    int64_t a = std::min<int64_t>(G, (int64_t)c * cr), b = std::min<int64_t>(G, (int64_t)(c + 1) * cr);
    *i0 = a - row_lo;
    *i1 = b - row_lo;
  }
};
int mb_chunk_grid(mb_ctx* ctx, const mb_mat* m, mb_chunks* out);
// leaves: MB_NCHUNK x count doubles (global chunk slots; the producer writes exact zeros into the slots this rank
// does not own); out <- tree sum over all chunks of all ranks (all-reduce of the zero-padded leaves when sharded)
int mb_tree_reduce_small(mb_ctx* ctx, const mb_chunks& g, double* leaves, int64_t count, double* out);
// tree over matrices too large to gather: `acc` holds this rank's tree sum of its own leaves on entry and the global
// tree sum on return (butterfly exchange for power-of-two worlds whose ranks own aligned subtrees)
int mb_tree_combine_ranks(mb_ctx* ctx, const mb_chunks& g, double* acc, int64_t count);
int mb_axpy_raw(mb_ctx* ctx, double* dst, const double* src, int64_t count);
// int8 digit-slice Gram of one block of cells (mb_i8.cu): out (r x r dense, tiles on / below the diagonal) = A^T A
bool mb_i8_gram_usable(mb_ctx* ctx, int64_t rows, int64_t r);
int mb_i8_gram_leaf(mb_ctx* ctx, const double* A, int64_t lda, int64_t rows, int64_t r, double* out);
// a sequence of Gram leaves over one factor that is complete on the main stream at `begin`: their packs may run ahead
int mb_i8_gram_begin(mb_ctx* ctx);
void mb_i8_gram_end(mb_ctx* ctx);
int mb_i8_check(mb_ctx* ctx);
// K1 on int8 digit slices (mb_cov_i8.cu); *done = false: shape / kind outside that kernel
int mb_cov_i8_build(mb_ctx* ctx, int kind, double c, const mb_mat* x, const mb_mat* y, const int* dims_host, int d,
                    double* out, int64_t ldo, bool* done, int64_t self_offset = 0, int64_t* nn_idx = nullptr);
// C = beta C + alpha A B^T on int8 digit slices (A: n x k, B: p x k row-major; beta 0 or 1); rows_total: the GLOBAL
// number of rows of the row-sharded operand (the path is chosen from it, not from the local row count)
bool mb_i8_nt_usable(mb_ctx* ctx, int64_t rows_total, int64_t p, int64_t k);
int mb_i8_gemm_nt(mb_ctx* ctx, int64_t n, int64_t p, int64_t k, double alpha, const double* A, int64_t lda,
                  const double* B, int64_t ldb, int beta, double* C, int64_t ldc);  // dst = dst + src (elementwise)

// kernel classes for the in-library stopwatch (mb_prof_*): CUDA events around each launch
enum { MB_PROF_COV = 0, MB_PROF_MATVEC = 1, MB_PROF_GEMM = 2, MB_PROF_LOSSGRAD = 3, MB_PROF_OTHER = 4,
       MB_PROF_I8 = 5, MB_PROF_EIGH = 6, MB_PROF_NCLS = 7 };

struct mb_prof_span {
  int cls;
  cudaEvent_t a, b;
};

struct mb_ctx {
  int device = 0;
  int n_sm = 0;
  cudaStream_t stream = nullptr;       // compute stream (everything is ordered on it)
  cudaStream_t copy_stream = nullptr;  // H2D/D2H overlap for the streaming predictor
  cudaEvent_t timer_ev[16][2] = {};
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  int64_t launches = 0;
  // scratch
  double* scratch = nullptr;           // reusable device scratch
  size_t scratch_bytes = 0;
  double* pinned = nullptr;            // reusable pinned host staging
  size_t pinned_bytes = 0;
  double* flush_buf = nullptr;
  size_t flush_bytes = 0;
  // device block cache: freed matrices are kept (after a stream sync) and handed back to allocations of the
  // same size class, so a fit does not pay cudaMalloc / cudaFree of its 40 GB factor every time
  std::multimap<size_t, double*> block_cache;
  size_t cached_bytes = 0, cache_cap = 0;
  // CUDA-graph replay of the launch-bound Cholesky (fixed internal buffers so one graph per size serves every call)
  double* potrf_buf = nullptr;
  size_t potrf_buf_bytes = 0;
  int* potrf_info = nullptr;
  std::map<int64_t, std::pair<cudaGraphExec_t, int64_t>> potrf_graphs;  // n -> (exec, kernel nodes)
  int opt_graph = 1;     // 1 = replay the Cholesky (n >= 1024) from a captured graph
  double* trsm_ws = nullptr;           // inverted 128 x 128 diagonal blocks of the TRSM
  size_t trsm_ws_bytes = 0;
  double* gemm_ws = nullptr;           // split-k partial tiles (own buffer: GEMMs run inside scratch users)
  size_t gemm_ws_bytes = 0;
  // int8 digit-slice GEMMs (mb_i8.cu): digit operands / scales / partials, tile list, error flag
  void* i8_ws = nullptr;
  size_t i8_ws_bytes = 0;
  int2* i8_tiles = nullptr;
  int64_t i8_tiles_r = -1;
  int i8_ntiles = 0;
  int* i8_status = nullptr;
  // option "i8_overlap": the pack of the NEXT slab / chunk runs on a side stream under the GEMM of the current one (two
  // operand buffers, events for the hand-offs)
  cudaStream_t i8_side = nullptr;
  cudaEvent_t i8_ev_start = nullptr, i8_ev_packed[2] = {nullptr, nullptr}, i8_ev_free[2] = {nullptr, nullptr};
  int i8_seq = -1;       // Gram leaves since mb_i8_gram_begin (-1: no sequence open, leaves run on the main stream)
  size_t i8_seq_one = 0; // size of one operand buffer of the open sequence
  int opt_i8_overlap = 0;  // 1 = packs of the next slab / chunk on the side stream (measured: no gain, see mb_i8.cu)
  int opt_cov_i8 = 1;    // 1 = K1 of one exponential-family leaf on tcgen05 int8 digit slices (large shapes), 2 = always, 0 = DMMA kernel
  int opt_i8_issuers = 4;  // MMA-issuing warps of the int8 GEMM kernels (1, 2 or 4)
  int opt_i8 = 1;        // 1 = Gram products of large factors on tcgen05 kind::i8 digit slices, 0 = FP64 DMMA tiles
  // NCCL
  void* comm = nullptr;
  int rank = 0, world = 1;
  bool solo = false;     // mb_comm_solo: behave as a single rank although a communicator is attached
  // options
  int opt_gemm = 0;      // 0 = DMMA tiles (default), 1 = DFMA register tiles
  int opt_cov = 0;       // 0 = DMMA tile kernel (exp-family leaf) else 1; 1 = DFMA register-tile kernel; 2 = general kernel
  int opt_trsm = 0;      // 0 = GEMM leaves on inverted 128-blocks (tall X), 1 = 32-wide substitution leaves
  int opt_lossgrad = 0;  // 0 = TMA-ring streaming single pass, 1 = two-pass, 2 = register-fused single pass
  // per-kernel-class stopwatch
  bool prof_on = false;
  std::vector<mb_prof_span> prof_spans;      // recorded, not yet resolved
  std::vector<cudaEvent_t> prof_pool;        // recycled events
  int64_t prof_count[MB_PROF_NCLS] = {};
  double prof_ms[MB_PROF_NCLS] = {};
  double prof_work[MB_PROF_NCLS] = {};  // algorithmic bytes (HBM-bound classes) or flops (GEMM)
};

cudaEvent_t mb_prof_event(mb_ctx* ctx);

void mb_set_error(const char* fmt, ...);

#define MB_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) {                                                       \
      mb_set_error("%s:%d CUDA error: %s (%s)", __FILE__, __LINE__,                \
                   cudaGetErrorString(_e), #call);                                 \
      return -1;                                                                   \
    }                                                                              \
  } while (0)

#define MB_CHECK(cond, ...)                                                        \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      mb_set_error(__VA_ARGS__);                                                   \
      return -2;                                                                   \
    }                                                                              \
  } while (0)

#define MB_TRY(call)                                                               \
  do {                                                                             \
    int _r = (call);                                                               \
    if (_r != 0) return _r;                                                        \
  } while (0)

// Every kernel launch goes through this so the context can count them and catch
// launch-configuration errors at the call site.
#define MB_LAUNCH(ctx, kernel, grid, block, smem, ...)                             \
  do {                                                                             \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);               \
    (ctx)->launches++;                                                             \
    MB_CUDA(cudaGetLastError());                                                   \
  } while (0)

// cudaFuncSetAttribute applies to the CURRENT device: one flag per (call site, device), so that contexts on several
// devices of one process each configure their kernels
struct mb_per_device_flag {
  bool done[64] = {};
  bool& operator()(const mb_ctx* c) { return done[c->device & 63]; }
};

// launch on another stream of the context (the int8 side stream)
#define MB_LAUNCH_ON(ctx, strm, kernel, grid, block, smem, ...)                   \
  do {                                                                             \
    kernel<<<(grid), (block), (smem), (strm)>>>(__VA_ARGS__);                      \
    (ctx)->launches++;                                                             \
    MB_CUDA(cudaGetLastError());                                                   \
  } while (0)

// Same, with the launch bracketed by CUDA events on the stream when profiling is on.
#define MB_LAUNCH_P(ctx, cls, kernel, grid, block, smem, ...)                      \
  do {                                                                             \
    mb_prof_span _sp = {(cls), nullptr, nullptr};                                  \
    if ((ctx)->prof_on) {                                                          \
      _sp.a = mb_prof_event(ctx);                                                  \
      _sp.b = mb_prof_event(ctx);                                                  \
      cudaEventRecord(_sp.a, (ctx)->stream);                                       \
    }                                                                              \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);               \
    (ctx)->launches++;                                                             \
    if ((ctx)->prof_on) {                                                          \
      cudaEventRecord(_sp.b, (ctx)->stream);                                       \
      (ctx)->prof_spans.push_back(_sp);                                            \
    }                                                                              \
    MB_CUDA(cudaGetLastError());                                                   \
  } while (0)

// NVTX range over an entry point (free unless a profiler is attached): one timeline row per stage of a fit
struct mb_nvtx_range {
  explicit mb_nvtx_range(const char* name) { nvtxRangePushA(name); }
  ~mb_nvtx_range() { nvtxRangePop(); }
};
#define MB_RANGE(name) mb_nvtx_range _mb_nvtx_range_(name)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// scratch management (grows monotonically; contents undefined)
int mb_scratch(mb_ctx* ctx, size_t bytes, double** out);
void mb_cache_flush(mb_ctx* ctx);  // release every cached device block
// cudaMalloc that empties the block cache and retries once before giving up
cudaError_t mb_dev_malloc(mb_ctx* ctx, void** p, size_t bytes);
int mb_pinned(mb_ctx* ctx, size_t bytes, double** out);

// internal helpers implemented across translation units
int mb_mat_view(mb_ctx* ctx, double* p, int64_t rows, int64_t cols, mb_mat* out);

// GEMM on raw pointers with leading dimensions (row-major).
//   a_kmajor=0: A element (i,k) at A[i*lda + k];  a_kmajor=1: at A[k*lda + i]
//   b_kmajor=0: B element (j,k) at B[j*ldb + k];  b_kmajor=1: at B[k*ldb + j]
//   C[i*ldc + j] = alpha * sum_k A(i,k) B(j,k) + beta * C[i*ldc + j]
//   lower_only: skip output tiles strictly above the diagonal (SYRK-style).
int mb_gemm_raw(mb_ctx* ctx, bool a_kmajor, bool b_kmajor, int64_t m, int64_t n, int64_t k,
                double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                double beta, double* C, int64_t ldc, bool lower_only);
// same for products whose output rows are cells (m = local cells): never split along k, so that an output row does
// not depend on how many rows this rank happens to hold
int mb_gemm_rows_raw(mb_ctx* ctx, bool a_kmajor, bool b_kmajor, int64_t m, int64_t n, int64_t k,
                     double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                     double beta, double* C, int64_t ldc);
// C (m x n) = A^T B summed over the rows (cells) of A (k x m) and B (k x n) through the fixed chunk tree of `g`;
// lower_only: A == B, lower tiles + mirror.  The result is the GLOBAL sum on every rank when g.sharded.
int mb_gemm_tn_cells(mb_ctx* ctx, const mb_chunks& g, int64_t m, int64_t n, const double* A, int64_t lda,
                     const double* B, int64_t ldb, double* C, int64_t ldc, bool lower_only);

int mb_potrf_raw(mb_ctx* ctx, double* A, int64_t n, int64_t lda, int* info_dev_accum);
int mb_trsm_right_lt_raw(mb_ctx* ctx, const double* Lp, int64_t ldl, int64_t m, double* X,
                         int64_t ldx, int64_t nrows, int64_t rows_total);
int mb_allreduce_raw(mb_ctx* ctx, double* p, int64_t count);
// exchange `count` doubles with rank `peer` (ncclSend + ncclRecv in one group) on the context's stream
int mb_sendrecv_raw(mb_ctx* ctx, const double* send, double* recv, int64_t count, int peer);
// broadcast `count` doubles from rank `root`
int mb_bcast_raw(mb_ctx* ctx, double* p, int64_t count, int root);
void mb_invalidate_graphs(mb_ctx* ctx);  // captured graphs hold raw workspace pointers: drop them when one moves
int mb_gemm_reserve_ws(mb_ctx* ctx, size_t bytes);  // grow the split-k workspace up front (no cudaMalloc under capture)

// warp / block reductions -------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
