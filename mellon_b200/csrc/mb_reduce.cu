// The fixed reduction tree over the cell axis (see mb_common.cuh): chunk grid of a (possibly sharded) matrix, the
// pairwise tree over MB_NCHUNK chunk sums, and its combination across ranks.  Because the leaves are GLOBAL chunks
// and rank boundaries fall on chunk boundaries, the bits of every reduced quantity are the same for 1, 2, 3, ... ranks.
#include "mb_common.cuh"

extern "C" int mb_mat_set_shard(mb_mat* m, int64_t global_rows, int64_t row_lo) {
  MB_CHECK(m, "mb_mat_set_shard: null matrix");
  if (global_rows < 0) {
    m->global_rows = -1;
    m->row_lo = 0;
    return 0;
  }
  MB_CHECK(row_lo >= 0 && row_lo + m->rows <= global_rows, "mb_mat_set_shard: rows [%lld, %lld) outside 0..%lld",
           (long long)row_lo, (long long)(row_lo + m->rows), (long long)global_rows);
  m->global_rows = global_rows;
  m->row_lo = row_lo;
  return 0;
}

extern "C" int mb_row_block(int64_t global_rows, int rank, int world, int64_t* row_lo, int64_t* row_hi,
                            int64_t* chunk_rows) {
  MB_CHECK(global_rows >= 0 && world >= 1 && rank >= 0 && rank < world, "mb_row_block: bad argument");
  const int64_t cr = std::max<int64_t>(1, ceil_div64(global_rows, MB_NCHUNK));
  const int64_t c_lo = (int64_t)rank * MB_NCHUNK / world, c_hi = (int64_t)(rank + 1) * MB_NCHUNK / world;
  if (row_lo) *row_lo = std::min(global_rows, c_lo * cr);
  if (row_hi) *row_hi = std::min(global_rows, c_hi * cr);
  if (chunk_rows) *chunk_rows = cr;
  return 0;
}

int mb_chunk_grid(mb_ctx* ctx, const mb_mat* m, mb_chunks* out) {
  const bool marked = m->global_rows >= 0;
  out->G = marked ? m->global_rows : m->rows;
  out->cr = std::max<int64_t>(1, ceil_div64(out->G, MB_NCHUNK));
  out->row_lo = marked ? m->row_lo : 0;
  out->rows = m->rows;
  out->sharded = marked && ctx->comm != nullptr && ctx->world > 1 && !ctx->solo;
  const int64_t end = out->row_lo + m->rows;
  if (out->sharded) {
    // the canonical layout: every rank can derive every other rank's leaves from (G, world)
    int64_t lo, hi;
    MB_TRY(mb_row_block(out->G, ctx->rank, ctx->world, &lo, &hi, nullptr));
    MB_CHECK(lo == out->row_lo && hi == end,
             "sharded matrix holds rows [%lld, %lld) but rank %d of %d owns [%lld, %lld) of %lld rows (mb_row_block)",
             (long long)out->row_lo, (long long)end, ctx->rank, ctx->world, (long long)lo, (long long)hi,
             (long long)out->G);
    out->c_lo = (int)((int64_t)ctx->rank * MB_NCHUNK / ctx->world);
    out->c_hi = (int)((int64_t)(ctx->rank + 1) * MB_NCHUNK / ctx->world);
  } else {
    // one rank sums everything it holds: all MB_NCHUNK leaves are local (those beyond the last row are empty)
    MB_CHECK(out->row_lo == 0 && end == out->G,
             "matrix marked as rows [%lld, %lld) of %lld but no communicator is attached", (long long)out->row_lo,
             (long long)end, (long long)out->G);
    out->c_lo = 0;
    out->c_hi = MB_NCHUNK;
  }
  return 0;
}

namespace {

// out[j] = pairwise tree over the MB_NCHUNK leaves of column j:  ((l0 + l1) + (l2 + l3)) + ...
__global__ void tree_leaves_kernel(const double* __restrict__ leaves, int64_t count, double* __restrict__ out) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= count) return;
  double v[MB_NCHUNK];
#pragma unroll
  for (int c = 0; c < MB_NCHUNK; c++) v[c] = leaves[(int64_t)c * count + j];
#pragma unroll
  for (int w = MB_NCHUNK / 2; w >= 1; w >>= 1) {
#pragma unroll
    for (int i = 0; i < w; i++) v[i] = v[2 * i] + v[2 * i + 1];
  }
  out[j] = v[0];
}

__global__ void add_kernel(double* __restrict__ dst, const double* __restrict__ a, const double* __restrict__ b,
                           int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t st = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) dst[i] = a[i] + b[i];
}

struct Node {
  int start, size;   // leaves [start, start + size), size a power of two, start % size == 0
  double* buf;       // nullptr: every leaf lies beyond the last row (an exact zero)
};

// maximal aligned subtrees covering [lo, hi), left to right
void canonical_nodes(int lo, int hi, std::vector<Node>* out) {
  while (lo < hi) {
    int size = 1;
    while (lo % (2 * size) == 0 && lo + 2 * size <= hi) size *= 2;
    out->push_back({lo, size, nullptr});
    lo += size;
  }
}

struct Pool {
  double* base;
  int64_t stride;
  std::vector<int> free_slots;
  double* get() {
    if (free_slots.empty()) return nullptr;
    int s = free_slots.back();
    free_slots.pop_back();
    return base + (int64_t)s * stride;
  }
  void put(double* p) {
    if (p) free_slots.push_back((int)((p - base) / stride));
  }
};

int launch_add(mb_ctx* ctx, double* dst, const double* a, const double* b, int64_t n) {
  int grid = (int)std::min<int64_t>(ceil_div64(n, 256), (int64_t)ctx->n_sm * 16);
  MB_LAUNCH(ctx, add_kernel, std::max(grid, 1), 256, 0, dst, a, b, n);
  return 0;
}

// push a node; while the two topmost nodes are siblings, replace them by their parent (left + right)
int push_merge(mb_ctx* ctx, std::vector<Node>* st, Node nd, int64_t count, Pool* pool) {
  st->push_back(nd);
  while (st->size() >= 2) {
    Node& a = (*st)[st->size() - 2];
    Node& b = (*st)[st->size() - 1];
    if (!(a.size == b.size && a.start % (2 * a.size) == 0 && b.start == a.start + a.size)) break;
    Node p = {a.start, 2 * a.size, nullptr};
    if (a.buf && b.buf) {
      MB_TRY(launch_add(ctx, a.buf, a.buf, b.buf, count));
      p.buf = a.buf;
      pool->put(b.buf);
    } else {
      p.buf = a.buf ? a.buf : b.buf;
    }
    st->pop_back();
    st->pop_back();
    st->push_back(p);
  }
  return 0;
}

}  // namespace

int mb_axpy_raw(mb_ctx* ctx, double* dst, const double* src, int64_t count) {
  return launch_add(ctx, dst, dst, src, count);
}

int mb_tree_reduce_small(mb_ctx* ctx, const mb_chunks& g, double* leaves, int64_t count, double* out) {
  if (count == 0) return 0;
  // slots this rank does not own were written as exact zeros by the producer, so the all-reduce below only moves
  // data: a sum with zeros is exact
  if (g.sharded) MB_TRY(mb_allreduce_raw(ctx, leaves, (int64_t)MB_NCHUNK * count));
  MB_LAUNCH(ctx, tree_leaves_kernel, (int)ceil_div64(count, 128), 128, 0, leaves, count, out);
  return 0;
}

int mb_gemm_tn_cells(mb_ctx* ctx, const mb_chunks& g, int64_t m, int64_t n, const double* A, int64_t lda,
                     const double* B, int64_t ldb, double* C, int64_t ldc, bool lower_only) {
  if (m <= 0 || n <= 0) return 0;
  MB_CHECK(ldc == n, "mb_gemm_tn_cells: dense output expected");
  const int64_t count = m * n;
  const int world = g.sharded ? ctx->world : 1, me = g.sharded ? ctx->rank : 0;
  auto leaf_empty = [&](int c) { return (int64_t)c * g.cr >= g.G; };
  // symmetric products of large factors run on the tcgen05 int8 digit slices; the choice depends on the GLOBAL chunk
  // size and the width only, so it is the same for every number of ranks
  const bool i8 = lower_only && A == B && lda == ldb && m == n && mb_i8_gram_usable(ctx, std::min(g.cr, g.G), m);

  std::vector<Node> mine;
  canonical_nodes(g.c_lo, g.c_hi, &mine);
  // scratch: stack of the local phase (<= 6) + finished own nodes + stack of the global phase + one incoming buffer
  const int nbuf = (int)mine.size() + 8;
  double* base;
  MB_TRY(mb_scratch(ctx, (size_t)nbuf * count * sizeof(double), &base));
  Pool pool = {base, count, {}};
  for (int s = nbuf - 1; s >= 0; s--) pool.free_slots.push_back(s);

  // local phase: every own node is the tree over its leaves
  struct SeqGuard {                          // closes the leaf sequence on every way out
    mb_ctx* c;
    ~SeqGuard() { mb_i8_gram_end(c); }
  } seq_guard = {ctx};
  if (i8) MB_TRY(mb_i8_gram_begin(ctx));     // A is complete on the stream: the leaves may pack ahead of their MMA kernels
  for (Node& nd : mine) {
    std::vector<Node> st;
    for (int c = nd.start; c < nd.start + nd.size; c++) {
      Node leaf = {c, 1, nullptr};
      int64_t i0, i1;
      g.range(c, &i0, &i1);
      if (!leaf_empty(c) && i1 > i0) {
        leaf.buf = pool.get();
        MB_CHECK(leaf.buf, "mb_gemm_tn_cells: buffer pool exhausted");
        if (i8) {
          MB_TRY(mb_i8_gram_leaf(ctx, A + i0 * lda, lda, i1 - i0, m, leaf.buf));
        } else {
          MB_TRY(mb_gemm_raw(ctx, true, true, m, n, i1 - i0, 1.0, A + i0 * lda, lda, B + i0 * ldb, ldb, 0.0, leaf.buf, n,
                             lower_only));
        }
      }
      MB_TRY(push_merge(ctx, &st, leaf, count, &pool));
    }
    MB_CHECK(st.size() == 1, "mb_gemm_tn_cells: local tree did not close");
    nd.buf = st[0].buf;
  }
  mb_i8_gram_end(ctx);

  // global phase: all ranks walk the canonical nodes of every rank in order; the owner broadcasts its node
  std::vector<Node> st;
  for (int rho = 0; rho < world; rho++) {
    std::vector<Node> theirs;
    if (rho == me) {
      theirs = mine;
    } else {
      canonical_nodes((int)((int64_t)rho * MB_NCHUNK / world), (int)((int64_t)(rho + 1) * MB_NCHUNK / world), &theirs);
    }
    for (Node nd : theirs) {
      bool empty = true;
      for (int c = nd.start; c < nd.start + nd.size; c++) empty = empty && leaf_empty(c);
      if (empty) {
        if (rho == me) pool.put(nd.buf);
        nd.buf = nullptr;
      } else if (world > 1) {
        if (rho != me) {
          nd.buf = pool.get();
          MB_CHECK(nd.buf, "mb_gemm_tn_cells: buffer pool exhausted");
        }
        MB_CHECK(nd.buf, "mb_gemm_tn_cells: rank %d holds no data for leaves [%d, %d)", rho, nd.start, nd.start + nd.size);
        MB_TRY(mb_bcast_raw(ctx, nd.buf, count, rho));
      }
      MB_TRY(push_merge(ctx, &st, nd, count, &pool));
    }
  }
  MB_CHECK(st.size() == 1 && st[0].start == 0 && st[0].size == MB_NCHUNK, "mb_gemm_tn_cells: tree did not close");
  if (i8) MB_TRY(mb_i8_check(ctx));
  if (st[0].buf) {
    MB_CUDA(cudaMemcpyAsync(C, st[0].buf, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    MB_CUDA(cudaMemsetAsync(C, 0, (size_t)count * sizeof(double), ctx->stream));
  }
  return 0;
}
