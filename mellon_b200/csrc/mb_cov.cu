// K1 / K7: fused pairwise-distance + covariance-kernel evaluation.
//
//   STORE : K(i, j) = prog(x_i, y_j) written once to HBM (never the distance matrix).
//   MATVEC: out(i) = mu + sum_j prog(x_i, y_j) w_j   (K never materialised).
//
// Reference arithmetic being restated (paths relative to /root/reference/mellon):
//   util.py:351-366   sq = xx - 2 xy + yy + 1e-12 ; dist = sqrt(max(sq, 0))
//   cov.py            the six `k` methods ; base_cov.py Add/Mul/Pow.k
//
// Fast path (one leaf): persistent CTA per 128-row panel; the x panel and the 64-row landmark
// tiles arrive in shared memory by 1-D bulk TMA (cp.async.bulk + mbarrier, landmark tiles
// double-buffered); 8x4 register micro-tiles accumulate x.y on the FP64 pipe; the epilogue
// evaluates the kernel in registers.
#include "mb_common.cuh"
#include "mb_math.cuh"

namespace {

constexpr int BM = 128, BN = 64, NT = 256;
enum { MODE_STORE = 0, MODE_MATVEC = 1, MODE_NNMIN = 2 };

struct LeafParams {
  int kind;
  double c1;     // Matern: sqrt(nu2)/ls ; ExpQuad: 0.5/ls^2 ; Exponential: 0.5/ls ; RatQuad: 1/(2 a ls^2) ; Linear: 1/ls
  double alpha;  // RatQuad
};

struct Leaf {
  int kind;
  double ls, alpha;
  std::vector<int> dims;  // absolute columns; empty => all
  bool all_dims;
};

struct Program {
  std::vector<Leaf> leaves;
  int n_ops;
  int op[MB_MAX_OPS];        // mb_kop_code
  int leaf_idx[MB_MAX_OPS];  // for LEAF ops
  double value[MB_MAX_OPS];  // CONST / POW
};

struct DevProgram {
  int n_ops, n_leaves;
  int op[MB_MAX_OPS];
  int leaf_idx[MB_MAX_OPS];
  double value[MB_MAX_OPS];
  LeafParams leaf[MB_MAX_LEAVES];
};

LeafParams leaf_params(const Leaf& l) {
  LeafParams p;
  p.kind = l.kind;
  p.alpha = l.alpha;
  switch (l.kind) {
    case MB_K_MATERN32: p.c1 = sqrt(3.0) / l.ls; break;
    case MB_K_MATERN52: p.c1 = sqrt(5.0) / l.ls; break;
    case MB_K_EXPQUAD: p.c1 = 0.5 / (l.ls * l.ls); break;
    case MB_K_EXPONENTIAL: p.c1 = 0.5 / l.ls; break;
    case MB_K_RATQUAD: p.c1 = 1.0 / (2.0 * l.alpha * l.ls * l.ls); break;
    case MB_K_DISTANCE: p.c1 = 1.0; break;
    default: p.c1 = 1.0 / l.ls; break;
  }
  return p;
}

int parse_program(const mb_kprog* prog, int64_t n_cols, Program* out) {
  MB_CHECK(prog && prog->ops && prog->n_ops > 0 && prog->n_ops <= MB_MAX_OPS,
           "covariance program must have 1..%d ops", MB_MAX_OPS);
  out->leaves.clear();
  out->n_ops = prog->n_ops;
  int depth = 0;
  for (int i = 0; i < prog->n_ops; i++) {
    const mb_kop& o = prog->ops[i];
    out->op[i] = o.op;
    out->value[i] = o.value;
    out->leaf_idx[i] = -1;
    switch (o.op) {
      case MB_OP_LEAF: {
        MB_CHECK((int)out->leaves.size() < MB_MAX_LEAVES, "covariance program has more than %d leaves",
                 MB_MAX_LEAVES);
        MB_CHECK(o.kind >= 0 && o.kind <= MB_K_DISTANCE, "unknown kernel kind %d", o.kind);
        MB_CHECK(o.ls > 0.0 || o.kind == MB_K_LINEAR || o.kind == MB_K_DISTANCE,
                 "length scale must be positive, got %g", o.ls);
        Leaf l;
        l.kind = o.kind;
        l.ls = o.ls;
        l.alpha = o.alpha;
        l.all_dims = o.dim_cnt < 0;
        if (!l.all_dims) {
          MB_CHECK(prog->dims && o.dim_off >= 0 && o.dim_off + o.dim_cnt <= prog->n_dims,
                   "leaf dims outside the program's dims table");
          for (int d = 0; d < o.dim_cnt; d++) {
            int col = prog->dims[o.dim_off + d];
            MB_CHECK(col >= 0 && col < n_cols, "active dim %d outside input with %lld columns", col,
                     (long long)n_cols);
            l.dims.push_back(col);
          }
          if ((int64_t)l.dims.size() == n_cols) {
            bool ident = true;
            for (int d = 0; d < (int)l.dims.size(); d++) ident &= (l.dims[d] == d);
            l.all_dims = ident;
          }
        }
        out->leaf_idx[i] = (int)out->leaves.size();
        out->leaves.push_back(l);
        depth++;
        break;
      }
      case MB_OP_CONST: depth++; break;
      case MB_OP_ADD:
      case MB_OP_MUL:
        MB_CHECK(depth >= 2, "malformed covariance program (binary op on stack depth %d)", depth);
        depth--;
        break;
      case MB_OP_POW: MB_CHECK(depth >= 1, "malformed covariance program (pow on empty stack)"); break;
      default: MB_CHECK(false, "unknown covariance op %d", o.op);
    }
    MB_CHECK(depth <= MB_STACK_DEPTH, "covariance program needs stack depth > %d", MB_STACK_DEPTH);
  }
  MB_CHECK(depth == 1, "malformed covariance program (final stack depth %d)", depth);
  return 0;
}

void to_dev_program(const Program& p, DevProgram* d) {
  d->n_ops = p.n_ops;
  d->n_leaves = (int)p.leaves.size();
  for (int i = 0; i < p.n_ops; i++) {
    d->op[i] = p.op[i];
    d->leaf_idx[i] = p.leaf_idx[i];
    d->value[i] = p.value[i];
  }
  for (int l = 0; l < d->n_leaves; l++) d->leaf[l] = leaf_params(p.leaves[l]);
}

// ---- per-element kernel evaluation ---------------------------------------------------------
// dot = x.y over the leaf's active dims, xx / yy the squared norms over the same dims.
template <int KIND>
__device__ __forceinline__ double eval_leaf(double dot, double xx, double yy, double c1, double alpha) {
  if (KIND == MB_K_LINEAR) return dot * c1;
  double sq = fma(-2.0, dot, xx);  // xx - 2 xy (2 xy is exact, so this rounds like the reference)
  sq = sq + yy;
  sq = sq + 1e-12;
  sq = fmax(sq, 0.0);
  if (KIND == MB_K_EXPQUAD) return exp(-sq * c1);
  if (KIND == MB_K_RATQUAD) return pow(fma(sq, c1, 1.0), -alpha);
  double dist = sqrt(sq);
  if (KIND == MB_K_DISTANCE) return dist;
  if (KIND == MB_K_EXPONENTIAL) return exp(-dist * c1);
  double r = dist * c1;
  double e = exp(-r);
  if (KIND == MB_K_MATERN32) return (r + 1.0) * e;
  return (r + r * r * (1.0 / 3.0) + 1.0) * e;  // MATERN52
}

__device__ __forceinline__ double eval_leaf_dyn(int kind, double dot, double xx, double yy, double c1,
                                                double alpha) {
  switch (kind) {
    case MB_K_MATERN32: return eval_leaf<MB_K_MATERN32>(dot, xx, yy, c1, alpha);
    case MB_K_MATERN52: return eval_leaf<MB_K_MATERN52>(dot, xx, yy, c1, alpha);
    case MB_K_EXPQUAD: return eval_leaf<MB_K_EXPQUAD>(dot, xx, yy, c1, alpha);
    case MB_K_EXPONENTIAL: return eval_leaf<MB_K_EXPONENTIAL>(dot, xx, yy, c1, alpha);
    case MB_K_RATQUAD: return eval_leaf<MB_K_RATQUAD>(dot, xx, yy, c1, alpha);
    case MB_K_DISTANCE: return eval_leaf<MB_K_DISTANCE>(dot, xx, yy, c1, alpha);
    default: return eval_leaf<MB_K_LINEAR>(dot, xx, yy, c1, alpha);
  }
}

// postfix evaluation on a 4-deep register stack (ops are warp-uniform)
__device__ __forceinline__ double eval_program(const DevProgram& P, const double* leafval) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int i = 0; i < P.n_ops; i++) {
    const int op = P.op[i];
    if (op == MB_OP_LEAF || op == MB_OP_CONST) {
      double v = P.value[i];
      if (op == MB_OP_LEAF) {
        const int li = P.leaf_idx[i];
        v = leafval[0];
        if (li == 1) v = leafval[1];
        if (li == 2) v = leafval[2];
        if (li == 3) v = leafval[3];
      }
      s3 = s2; s2 = s1; s1 = s0; s0 = v;
    } else if (op == MB_OP_POW) {
      s0 = pow(s0, P.value[i]);
    } else {
      s0 = (op == MB_OP_ADD) ? (s1 + s0) : (s1 * s0);
      s1 = s2; s2 = s3;
    }
  }
  return s0;
}

// ---- packing: gather a leaf's active columns, zero-pad to ldp, squared row norms ------------
__global__ void pack_kernel(const double* __restrict__ a, int64_t n, int64_t cols, const int* __restrict__ dims,
                            int d, int ldp, double scale, double* __restrict__ packed, double* __restrict__ norm) {
  // one warp per row
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    double s = 0.0;
    for (int c = lane; c < ldp; c += 32) {
      double v = 0.0;
      if (c < d) v = a[i * cols + (dims ? dims[c] : c)] * scale;
      if (packed) packed[i * ldp + c] = v;
      s = fma(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) norm[i] = s;
  }
}

// ---- mbarrier / bulk-copy (TMA) helpers ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MB_DONE;\n"
      "bra MB_WAIT;\n"
      "MB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- fast path: one leaf ---------------------------------------------------------------------
template <int KIND, int MODE>
__global__ void __launch_bounds__(NT, 1)
cov_tile_kernel(const double* __restrict__ xp, const double* __restrict__ xnorm, int64_t n,
                const double* __restrict__ yp, const double* __restrict__ ynorm, int64_t m, int ldp,
                LeafParams par, double* __restrict__ out, int64_t ldo, const double* __restrict__ w, double mu,
                int64_t self_offset = 0, int64_t* __restrict__ nn_idx = nullptr) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* xs = reinterpret_cast<double*>(smem_raw);
  double* ys0 = xs + (size_t)BM * ldp;
  double* ys1 = ys0 + (size_t)BN * ldp;
  __shared__ __align__(8) uint64_t bar_x, bar_y[2];
  __shared__ double mv_red[2][BM];
  __shared__ int64_t nn_red[2][BM];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = lane & 7, ty = lane >> 3;
  const int wr0 = (warp >> 1) * 32, wc0 = (warp & 1) * 32;

  // zero the tiles once so partially filled tiles never hold NaN garbage
  for (int e = tid; e < (BM + 2 * BN) * ldp; e += NT) xs[e] = 0.0;
  if (tid == 0) {
    mbar_init(&bar_x, 1);
    mbar_init(&bar_y[0], 1);
    mbar_init(&bar_y[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // order the generic-proxy zero fill before async-proxy (TMA) writes to the same bytes
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();

  uint32_t px = 0, py0 = 0, py1 = 0;
  const int64_t n_panels = (n + BM - 1) / BM;
  const int64_t n_ytiles = (m + BN - 1) / BN;
  const uint32_t row_bytes = (uint32_t)ldp * 8u;

  for (int64_t panel = blockIdx.x; panel < n_panels; panel += gridDim.x) {
    const int64_t row0 = panel * BM;
    const int rows_valid = (int)min((int64_t)BM, n - row0);
    if (tid == 0) {
      mbar_expect_tx(&bar_x, rows_valid * row_bytes);
      bulk_g2s(xs, xp + row0 * ldp, rows_valid * row_bytes, &bar_x);
      const int cv = (int)min((int64_t)BN, m);
      mbar_expect_tx(&bar_y[0], cv * row_bytes);
      bulk_g2s(ys0, yp, cv * row_bytes, &bar_y[0]);
    }
    double xn[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int64_t r = row0 + wr0 + ty + 4 * i;
      xn[i] = (r < n) ? xnorm[r] : 0.0;
    }
    double rowacc[8];
    int64_t rowidx[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      rowacc[i] = (MODE == MODE_NNMIN) ? __longlong_as_double(0x7ff0000000000000LL) : 0.0;
      rowidx[i] = -1;
    }
    mbar_wait(&bar_x, px);
    px ^= 1;

    for (int64_t jt = 0; jt < n_ytiles; jt++) {
      const int buf = (int)(jt & 1);
      if (tid == 0 && jt + 1 < n_ytiles) {
        const int64_t c0n = (jt + 1) * BN;
        const int cv = (int)min((int64_t)BN, m - c0n);
        uint64_t* b = buf ? &bar_y[0] : &bar_y[1];
        mbar_expect_tx(b, cv * row_bytes);
        bulk_g2s(buf ? ys0 : ys1, yp + c0n * ldp, cv * row_bytes, b);
      }
      if (buf == 0) { mbar_wait(&bar_y[0], py0); py0 ^= 1; }
      else          { mbar_wait(&bar_y[1], py1); py1 ^= 1; }
      const double* ys = buf ? ys1 : ys0;
      const int64_t col0 = jt * BN;

      double acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0;

      const double* xb = xs + (size_t)(wr0 + ty) * ldp;
      const double* yb = ys + (size_t)(wc0 + tx) * ldp;
      const int xstep = 4 * ldp, ystep = 8 * ldp;
#pragma unroll 5
      for (int k = 0; k < ldp; k += 2) {
        double2 xv[8], yv[4];
#pragma unroll
        for (int i = 0; i < 8; i++) xv[i] = *reinterpret_cast<const double2*>(xb + i * xstep + k);
#pragma unroll
        for (int j = 0; j < 4; j++) yv[j] = *reinterpret_cast<const double2*>(yb + j * ystep + k);
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = fma(xv[i].y, yv[j].y, fma(xv[i].x, yv[j].x, acc[i][j]));
      }

      double yn[4], wv[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int64_t c = col0 + wc0 + tx + 8 * j;
        yn[j] = (c < m) ? ynorm[c] : 0.0;
        if (MODE == MODE_MATVEC) wv[j] = (c < m) ? w[c] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int64_t r = row0 + wr0 + ty + 4 * i;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int64_t c = col0 + wc0 + tx + 8 * j;
          if (MODE == MODE_NNMIN) {
            // running minimum of the squared distance over every point but the cell itself
            const double sq = (xn[i] - 2.0 * acc[i][j]) + yn[j];
            if (c < m && c != r + self_offset && sq < rowacc[i]) { rowacc[i] = sq; rowidx[i] = c; }
            continue;
          }
          double v = eval_leaf<KIND>(acc[i][j], xn[i], yn[j], par.c1, par.alpha);
          if (MODE == MODE_STORE) {
            if (r < n && c < m) out[r * ldo + c] = v;
          } else {
            if (c < m) rowacc[i] = fma(v, wv[j], rowacc[i]);
          }
        }
      }
      __syncthreads();  // everyone is done with ys[buf] (and xs on the last tile)
    }

    if (MODE == MODE_NNMIN) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        double s = rowacc[i];
        int64_t id = rowidx[i];
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          double s2 = __shfl_xor_sync(0xffffffffu, s, o);
          int64_t id2 = __shfl_xor_sync(0xffffffffu, id, o);
          if (s2 < s || (s2 == s && id2 >= 0 && (id < 0 || id2 < id))) { s = s2; id = id2; }
        }
        if (tx == 0) { mv_red[warp & 1][wr0 + ty + 4 * i] = s; nn_red[warp & 1][wr0 + ty + 4 * i] = id; }
      }
      __syncthreads();
      if (tid < BM) {
        int64_t r = row0 + tid;
        double s = mv_red[0][tid], s2 = mv_red[1][tid];
        int64_t id = nn_red[0][tid], id2 = nn_red[1][tid];
        if (s2 < s || (s2 == s && id2 >= 0 && (id < 0 || id2 < id))) { s = s2; id = id2; }
        if (r < n) { out[r * ldo] = s; nn_idx[r] = id; }
      }
      __syncthreads();
    }
    if (MODE == MODE_MATVEC) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        double s = rowacc[i];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (tx == 0) mv_red[warp & 1][wr0 + ty + 4 * i] = s;
      }
      __syncthreads();
      if (tid < BM) {
        int64_t r = row0 + tid;
        if (r < n) out[r * ldo] = mu + (mv_red[0][tid] + mv_red[1][tid]);
      }
      __syncthreads();
    }
  }
}


// ---- tensor-pipe path: one leaf of the exponential family --------------------------------------
// The x.y contraction runs as FP64 MMA (mma.sync m8n8k4, DMMA) on operands that the pack kernel
// has pre-scaled by the leaf's distance scale (sqrt(5)/ls for Matern52, ...), so the accumulator is
// the scaled dot product and sq' = xn + yn - 2 acc is already (r_scaled)^2.  DMMA shares the FP64
// pipe with DFMA (tools/microbench_fp64: 37.1 TF either way, no overlap) but needs 12 shared-memory
// loads per 16 MMAs instead of 12 per 64 FMAs, which is what bounded the register-tile kernel.
// The epilogue is the lean sqrt / exp of mb_math.cuh (23 FP64 instructions per Matern52 element).
//
// CTA = 128 cells x 64 landmarks, 8 warps as 4 x 2, warp tile 32 x 32 = 4 x 4 MMA tiles; the cell
// panel and the double-buffered landmark tiles arrive by 1-D bulk TMA; two CTAs share an SM.
constexpr int MBM = 128, MBN = 64, MNT = 256;
__device__ const double g_exp2_tab[64] = MB_EXP2_TABLE_INIT;

template <int KIND>
__device__ __forceinline__ double eval_scaled(double sq, const double* tab, double alpha) {
  sq = mbmath::clamp_tiny(sq);
  if (KIND == MB_K_EXPQUAD) return mbmath::exp_neg(sq, tab);
  const double r = mbmath::sqrt_pos(sq);
  const double e = mbmath::exp_neg(r, tab);
  if (KIND == MB_K_EXPONENTIAL) return e;
  if (KIND == MB_K_MATERN32) return fma(r, e, e);
  return fma(fma(r, 1.0 / 3.0, 1.0), r, 1.0) * e;  // MATERN52
}

template <int KIND, int MODE>
__global__ void __launch_bounds__(MNT, 2)
cov_mma_kernel(const double* __restrict__ xp, const double* __restrict__ xnorm, int64_t n,
               const double* __restrict__ yp, const double* __restrict__ ynorm, int64_t m, int ldp, double eps_scaled,
               double* __restrict__ out, int64_t ldo, const double* __restrict__ w, double mu) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* xs = reinterpret_cast<double*>(smem_raw);
  double* ys0 = xs + (size_t)MBM * ldp;
  double* ys1 = ys0 + (size_t)MBN * ldp;
  __shared__ __align__(8) uint64_t bar_x, bar_y[2];
  __shared__ double tab[64];
  __shared__ double mv_red[2][MBM];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lk = lane & 3;
  const int wr0 = (warp >> 1) * 32, wc0 = (warp & 1) * 32;

  if (tid < 64) tab[tid] = g_exp2_tab[tid];
  for (int e = tid; e < (MBM + 2 * MBN) * ldp; e += MNT) xs[e] = 0.0;
  if (tid == 0) {
    mbar_init(&bar_x, 1);
    mbar_init(&bar_y[0], 1);
    mbar_init(&bar_y[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();

  uint32_t px = 0, py0 = 0, py1 = 0;
  const int64_t n_panels = (n + MBM - 1) / MBM;
  const int64_t n_ytiles = (m + MBN - 1) / MBN;
  const uint32_t row_bytes = (uint32_t)ldp * 8u;
  const bool vec_store = (MODE == MODE_STORE) && ((ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);

  for (int64_t panel = blockIdx.x; panel < n_panels; panel += gridDim.x) {
    const int64_t row0 = panel * MBM;
    const int rows_valid = (int)min((int64_t)MBM, n - row0);
    if (tid == 0) {
      mbar_expect_tx(&bar_x, rows_valid * row_bytes);
      bulk_g2s(xs, xp + row0 * ldp, rows_valid * row_bytes, &bar_x);
      const int cv = (int)min((int64_t)MBN, m);
      mbar_expect_tx(&bar_y[0], cv * row_bytes);
      bulk_g2s(ys0, yp, cv * row_bytes, &bar_y[0]);
    }
    double xn[4], rowacc[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int64_t r = row0 + wr0 + i * 8 + lr;
      xn[i] = (r < n) ? xnorm[r] + eps_scaled : 0.0;
      rowacc[i] = 0.0;
    }
    mbar_wait(&bar_x, px);
    px ^= 1;

    for (int64_t jt = 0; jt < n_ytiles; jt++) {
      const int buf = (int)(jt & 1);
      if (tid == 0 && jt + 1 < n_ytiles) {
        const int64_t c0n = (jt + 1) * MBN;
        const int cv = (int)min((int64_t)MBN, m - c0n);
        uint64_t* b = buf ? &bar_y[0] : &bar_y[1];
        mbar_expect_tx(b, cv * row_bytes);
        bulk_g2s(buf ? ys0 : ys1, yp + c0n * ldp, cv * row_bytes, b);
      }
      const int64_t col0 = jt * MBN;
      // norms / weights of this thread's 8 columns (in flight while the tile is awaited)
      double yn[4][2], wv[4][2];
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int64_t c = col0 + wc0 + j * 8 + 2 * lk + h;
          yn[j][h] = (c < m) ? ynorm[c] : 0.0;
          if (MODE == MODE_MATVEC) wv[j][h] = (c < m) ? w[c] : 0.0;
        }
      if (buf == 0) { mbar_wait(&bar_y[0], py0); py0 ^= 1; }
      else          { mbar_wait(&bar_y[1], py1); py1 ^= 1; }
      const double* ys = buf ? ys1 : ys0;

      double acc[4][4][2];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

      const double* xb = xs + (size_t)(wr0 + lr) * ldp + lk;
      const double* yb = ys + (size_t)(wc0 + lr) * ldp + lk;
      const int step8 = 8 * ldp;
#pragma unroll 2
      for (int k = 0; k < ldp; k += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; i++) af[i] = xb[i * step8 + k];
#pragma unroll
        for (int j = 0; j < 4; j++) bf[j] = yb[j * step8 + k];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                         : "d"(af[i]), "d"(bf[j]));
      }

#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int64_t r = row0 + wr0 + i * 8 + lr;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int64_t c = col0 + wc0 + j * 8 + 2 * lk;
          const double v0 = eval_scaled<KIND>(fma(-2.0, acc[i][j][0], xn[i] + yn[j][0]), tab, 0.0);
          const double v1 = eval_scaled<KIND>(fma(-2.0, acc[i][j][1], xn[i] + yn[j][1]), tab, 0.0);
          if (MODE == MODE_STORE) {
            if (r < n) {
              double* o = out + r * ldo + c;
              if (vec_store && c + 1 < m) {
                *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
              } else {
                if (c < m) o[0] = v0;
                if (c + 1 < m) o[1] = v1;
              }
            }
          } else {
            rowacc[i] = fma(v1, wv[j][1], fma(v0, wv[j][0], rowacc[i]));  // wv = 0 beyond m
          }
        }
      }
      __syncthreads();  // everyone is done with ys[buf] (and xs on the last tile)
    }

    if (MODE == MODE_MATVEC) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        double s = rowacc[i];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (lk == 0) mv_red[warp & 1][wr0 + i * 8 + lr] = s;
      }
      __syncthreads();
      if (tid < MBM) {
        const int64_t r = row0 + tid;
        if (r < n) out[r * ldo] = mu + (mv_red[0][tid] + mv_red[1][tid]);
      }
      __syncthreads();
    }
  }
}

// exact Euclidean distance of each row to its selected neighbour: sqrt(sum (x_i - y_j)^2), one warp per row
__global__ void nn_exact_kernel(const double* __restrict__ x, int64_t n, const double* __restrict__ y, int d,
                                const int64_t* __restrict__ idx, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const int64_t j = idx[i];
    double s = 0.0;
    if (j >= 0)
      for (int c = lane; c < d; c += 32) {
        double t = x[i * d + c] - y[j * d + c];
        s = fma(t, t, s);
      }
    s = warp_sum(s);
    if (lane == 0) out[i] = (j >= 0) ? sqrt(s) : __longlong_as_double(0x7ff8000000000000LL);
  }
}

// ---- general path: up to MB_MAX_LEAVES leaves, arbitrary widths ---------------------------------
struct GenLeafPtrs {
  const double* xp[MB_MAX_LEAVES];
  const double* xn[MB_MAX_LEAVES];
  const double* yp[MB_MAX_LEAVES];
  const double* yn[MB_MAX_LEAVES];
  int ldp[MB_MAX_LEAVES];
};

template <int MODE>
__global__ void __launch_bounds__(256)
cov_general_kernel(DevProgram P, GenLeafPtrs G, int64_t n, int64_t m, double* __restrict__ out, int64_t ldo,
                   const double* __restrict__ w, double mu) {
  constexpr int TB = 64, KC = 32;
  __shared__ double xs[TB][KC + 1], ys[TB][KC + 1];
  __shared__ double mv_red[16][TB];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t n_rt = (n + TB - 1) / TB, n_ct = (m + TB - 1) / TB;
  for (int64_t rt = blockIdx.x; rt < n_rt; rt += gridDim.x) {
    const int64_t row0 = rt * TB;
    double rowacc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t ct = 0; ct < n_ct; ct++) {
      const int64_t col0 = ct * TB;
      double leafval[MB_MAX_LEAVES][4][4];
#pragma unroll
      for (int l = 0; l < MB_MAX_LEAVES; l++) {
        if (l < P.n_leaves) {
          double acc[4][4] = {};
          const int ldp = G.ldp[l];
          for (int k0 = 0; k0 < ldp; k0 += KC) {
            __syncthreads();
            for (int e = tid; e < TB * KC; e += 256) {
              int r = e / KC, k = e % KC;
              int64_t gr = row0 + r, gc = col0 + r;
              xs[r][k] = (gr < n && k0 + k < ldp) ? G.xp[l][gr * ldp + k0 + k] : 0.0;
              ys[r][k] = (gc < m && k0 + k < ldp) ? G.yp[l][gc * ldp + k0 + k] : 0.0;
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < KC; k++) {
              double a[4], b[4];
#pragma unroll
              for (int i = 0; i < 4; i++) a[i] = xs[ty + 16 * i][k];
#pragma unroll
              for (int j = 0; j < 4; j++) b[j] = ys[tx + 16 * j][k];
#pragma unroll
              for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; i++) {
            int64_t r = row0 + ty + 16 * i;
            double xx = (r < n) ? G.xn[l][r] : 0.0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
              int64_t c = col0 + tx + 16 * j;
              double yy = (c < m) ? G.yn[l][c] : 0.0;
              leafval[l][i][j] = eval_leaf_dyn(P.leaf[l].kind, acc[i][j], xx, yy, P.leaf[l].c1, P.leaf[l].alpha);
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int64_t r = row0 + ty + 16 * i;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          int64_t c = col0 + tx + 16 * j;
          double lv[MB_MAX_LEAVES];
#pragma unroll
          for (int l = 0; l < MB_MAX_LEAVES; l++) lv[l] = (l < P.n_leaves) ? leafval[l][i][j] : 0.0;
          double v = eval_program(P, lv);
          if (MODE == MODE_STORE) {
            if (r < n && c < m) out[r * ldo + c] = v;
          } else {
            if (c < m) rowacc[i] = fma(v, w[c], rowacc[i]);
          }
        }
      }
    }
    if (MODE == MODE_MATVEC) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; i++) mv_red[tx][ty + 16 * i] = rowacc[i];
      __syncthreads();
      if (tid < TB) {
        double s = 0.0;
        for (int t = 0; t < 16; t++) s += mv_red[t][tid];
        int64_t r = row0 + tid;
        if (r < n) out[r * ldo] = mu + s;
      }
    }
  }
}

// diag(i) = prog(x_i, x_i): xy == xx so sq = (xx - 2xx) + xx + 1e-12 = 1e-12 exactly.
__global__ void cov_diag_kernel(DevProgram P, GenLeafPtrs G, int64_t n, double* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double lv[MB_MAX_LEAVES];
#pragma unroll
  for (int l = 0; l < MB_MAX_LEAVES; l++) {
    lv[l] = 0.0;
    if (l < P.n_leaves) {
      double xx = G.xn[l][i];
      lv[l] = eval_leaf_dyn(P.leaf[l].kind, xx, xx, xx, P.leaf[l].c1, P.leaf[l].alpha);
    }
  }
  out[i] = eval_program(P, lv);
}

// ---- host orchestration ---------------------------------------------------------------------------
struct Operand {
  const double* p;
  const double* norm;
  int ldp;
};

int ldp_for(int d) {
  int l = d + (d & 1);       // even
  if ((l & 3) == 0) l += 2;  // ldp/2 odd: conflict-free 16-byte shared loads across consecutive rows
  return l;
}

// leading dimension of the DMMA operand tiles: a multiple of 4 (k-step) that is 4 mod 8, so the 8 rows x 4 k
// fragment loads of a half-warp fall into 16 distinct 8-byte banks
int ldp_mma(int d) {
  int l = (d + 3) & ~3;
  if ((l & 7) == 0) l += 4;
  return l;
}

// distance scale folded into the packed operands of the tensor-pipe path (0 => kind not handled there)
double mma_scale(const Leaf& l) {
  switch (l.kind) {
    case MB_K_MATERN32: return sqrt(3.0) / l.ls;
    case MB_K_MATERN52: return sqrt(5.0) / l.ls;
    case MB_K_EXPQUAD: return sqrt(0.5) / l.ls;
    case MB_K_EXPONENTIAL: return 0.5 / l.ls;
    default: return 0.0;
  }
}

struct Plan {
  bool mma = false;      // single exponential-family leaf: operands packed pre-scaled for cov_mma_kernel
  double eps_scaled = 0.0;  // 1e-12 * scale^2
  bool same = false;  // y is x (symmetric landmark covariance)
  double in_cols = 0.0;  // columns of the input matrices
  Program prog;
  DevProgram dprog;
  std::vector<Operand> xo, yo;
  int* dims_dev[MB_MAX_LEAVES];
};

// Lay out scratch and pack every leaf's operands.  `same` => y is x (pack once).
int prepare(mb_ctx* ctx, const mb_kprog* kp, const mb_mat* x, const mb_mat* y, Plan* plan, bool allow_mma = false) {
  MB_CHECK(x->cols == y->cols, "covariance inputs have %lld and %lld columns", (long long)x->cols,
           (long long)y->cols);
  MB_TRY(parse_program(kp, x->cols, &plan->prog));
  to_dev_program(plan->prog, &plan->dprog);
  const bool same = (x->p == y->p && x->rows == y->rows);
  plan->same = same;
  plan->in_cols = (double)x->cols;
  const int nl = (int)plan->prog.leaves.size();
  double scale = 1.0;
  if (allow_mma && ctx->opt_cov == 0 && nl == 1 && plan->prog.n_ops == 1) {
    const Leaf& lf = plan->prog.leaves[0];
    const int d0 = lf.all_dims ? (int)x->cols : (int)lf.dims.size();
    const double sc = mma_scale(lf);
    if (sc > 0.0 && (size_t)(MBM + 2 * MBN) * ldp_mma(d0) * sizeof(double) <= 200 * 1024) {
      plan->mma = true;
      plan->eps_scaled = 1e-12 * sc * sc;
      scale = sc;
    }
  }
  size_t total = 0;
  std::vector<size_t> off_xp(nl), off_xn(nl), off_yp(nl), off_yn(nl), off_dims(nl);
  std::vector<bool> direct(nl);
  for (int l = 0; l < nl; l++) {
    const Leaf& lf = plan->prog.leaves[l];
    int d = lf.all_dims ? (int)x->cols : (int)lf.dims.size();
    MB_CHECK(d > 0, "covariance leaf with zero active dims");
    int ldp = plan->mma ? ldp_mma(d) : ldp_for(d);
    direct[l] = !plan->mma && lf.all_dims && ldp == d;
    auto take = [&](size_t doubles) {
      size_t o = total;
      total += (doubles + 15) & ~(size_t)15;
      return o;
    };
    off_dims[l] = take(((size_t)d + 1) / 2 + 2);
    off_xn[l] = take((size_t)x->rows);
    off_xp[l] = direct[l] ? 0 : take((size_t)x->rows * ldp);
    if (!same) {
      off_yn[l] = take((size_t)y->rows);
      off_yp[l] = direct[l] ? 0 : take((size_t)y->rows * ldp);
    }
  }
  double* s;
  MB_TRY(mb_scratch(ctx, (total + 16) * sizeof(double), &s));
  plan->xo.resize(nl);
  plan->yo.resize(nl);
  for (int l = 0; l < nl; l++) {
    const Leaf& lf = plan->prog.leaves[l];
    int d = lf.all_dims ? (int)x->cols : (int)lf.dims.size();
    int ldp = plan->mma ? ldp_mma(d) : ldp_for(d);
    int* dims_dev = nullptr;
    if (!lf.all_dims) {
      dims_dev = reinterpret_cast<int*>(s + off_dims[l]);
      MB_CUDA(cudaMemcpyAsync(dims_dev, lf.dims.data(), sizeof(int) * d, cudaMemcpyHostToDevice, ctx->stream));
    }
    auto pack = [&](const mb_mat* a, size_t offp, size_t offn, Operand* o) -> int {
      double* packed = direct[l] ? nullptr : s + offp;
      double* norm = s + offn;
      if (a->rows > 0) {
        int grid = (int)min((int64_t)ctx->n_sm * 8, ceil_div64(a->rows, 8));
        MB_LAUNCH(ctx, pack_kernel, grid, 256, 0, a->p, a->rows, a->cols, dims_dev, d, ldp, scale, packed, norm);
      }
      o->p = direct[l] ? a->p : packed;
      o->norm = norm;
      o->ldp = ldp;
      return 0;
    };
    MB_TRY(pack(x, off_xp[l], off_xn[l], &plan->xo[l]));
    if (same) plan->yo[l] = plan->xo[l];
    else MB_TRY(pack(y, off_yp[l], off_yn[l], &plan->yo[l]));
  }
  // the std::vector<int> dims were copied with an async memcpy from pageable memory, which the
  // runtime stages before returning, so `plan->prog` may be destroyed after this call.
  return 0;
}

// stopwatch class of a launch + its algorithmic HBM bytes: inputs once, output once
template <int MODE>
int prof_class(mb_ctx* ctx, const Plan& plan, int64_t n, int64_t m) {
  const int cls = (MODE == MODE_MATVEC) ? MB_PROF_MATVEC : (plan.same ? MB_PROF_OTHER : MB_PROF_COV);
  if (ctx->prof_on) {
    double d = 0.0;
    for (const Leaf& l : plan.prog.leaves) d = std::max<double>(d, l.all_dims ? 0.0 : (double)l.dims.size());
    double bytes = (MODE == MODE_MATVEC) ? 8.0 * ((double)n + (double)m) : 8.0 * (double)n * (double)m;
    ctx->prof_work[cls] += bytes + 8.0 * plan.in_cols * ((double)n + (plan.same ? 0.0 : (double)m));
    (void)d;
  }
  return cls;
}

template <int MODE>
int launch_fast(mb_ctx* ctx, const Plan& plan, int64_t n, int64_t m, double* out, int64_t ldo, const double* w,
                double mu, bool* done) {
  *done = false;
  if (plan.prog.n_ops != 1 || plan.prog.leaves.size() != 1) return 0;
  const Operand& xo = plan.xo[0];
  const Operand& yo = plan.yo[0];
  const int ldp = xo.ldp;
  size_t smem = (size_t)(BM + 2 * BN) * ldp * sizeof(double);
  if (smem > 200 * 1024) return 0;  // very wide inputs take the general path
  if ((reinterpret_cast<uintptr_t>(xo.p) & 15) || (reinterpret_cast<uintptr_t>(yo.p) & 15)) return 0;
  LeafParams par = plan.dprog.leaf[0];
  int64_t n_panels = ceil_div64(n, BM);
  int grid = (int)min(n_panels, (int64_t)ctx->n_sm);
  const int prof_cls = prof_class<MODE>(ctx, plan, n, m);
#define MB_COV_CASE(K)                                                                                        \
  case K: {                                                                                                   \
    static mb_per_device_flag cfg;                                                                                     \
    if (!cfg(ctx)) {                                                                                               \
      MB_CUDA(cudaFuncSetAttribute(cov_tile_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                   200 * 1024));                                                              \
      cfg(ctx) = true;                                                                                             \
    }                                                                                                         \
    MB_LAUNCH_P(ctx, prof_cls, (cov_tile_kernel<K, MODE>), grid, NT, smem, xo.p, xo.norm, n, yo.p, yo.norm, m, ldp, par,  \
              out, ldo, w, mu);                                                                               \
    break;                                                                                                    \
  }
  switch (par.kind) {
    MB_COV_CASE(MB_K_MATERN32)
    MB_COV_CASE(MB_K_MATERN52)
    MB_COV_CASE(MB_K_EXPQUAD)
    MB_COV_CASE(MB_K_EXPONENTIAL)
    MB_COV_CASE(MB_K_RATQUAD)
    MB_COV_CASE(MB_K_LINEAR)
    MB_COV_CASE(MB_K_DISTANCE)
    default: MB_CHECK(false, "unknown kernel kind %d", par.kind);
  }
#undef MB_COV_CASE
  *done = true;
  return 0;
}


template <int MODE>
int launch_mma(mb_ctx* ctx, const Plan& plan, int64_t n, int64_t m, double* out, int64_t ldo, const double* w,
               double mu, bool* done) {
  *done = false;
  if (!plan.mma) return 0;
  const Operand& xo = plan.xo[0];
  const Operand& yo = plan.yo[0];
  const int ldp = xo.ldp;
  const size_t smem = (size_t)(MBM + 2 * MBN) * ldp * sizeof(double);
  const int kind = plan.dprog.leaf[0].kind;
  const int64_t n_panels = ceil_div64(n, MBM);
  const int grid = (int)min(n_panels, (int64_t)ctx->n_sm * 2);
  const int prof_cls = prof_class<MODE>(ctx, plan, n, m);
#define MB_MMA_CASE(K)                                                                                       \
  case K: {                                                                                                  \
    static mb_per_device_flag cfg;                                                                                    \
    if (!cfg(ctx)) {                                                                                              \
      MB_CUDA(cudaFuncSetAttribute(cov_mma_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                   200 * 1024));                                                             \
      cfg(ctx) = true;                                                                                            \
    }                                                                                                        \
    MB_LAUNCH_P(ctx, prof_cls, (cov_mma_kernel<K, MODE>), grid, MNT, smem, xo.p, xo.norm, n, yo.p, yo.norm, m, ldp, \
                plan.eps_scaled, out, ldo, w, mu);                                                           \
    break;                                                                                                   \
  }
  switch (kind) {
    MB_MMA_CASE(MB_K_MATERN32)
    MB_MMA_CASE(MB_K_MATERN52)
    MB_MMA_CASE(MB_K_EXPQUAD)
    MB_MMA_CASE(MB_K_EXPONENTIAL)
    default: return 0;
  }
#undef MB_MMA_CASE
  *done = true;
  return 0;
}

template <int MODE>
int launch_general(mb_ctx* ctx, const Plan& plan, int64_t n, int64_t m, double* out, int64_t ldo, const double* w,
                   double mu) {
  GenLeafPtrs G;
  memset(&G, 0, sizeof(G));
  for (size_t l = 0; l < plan.xo.size(); l++) {
    G.xp[l] = plan.xo[l].p;
    G.xn[l] = plan.xo[l].norm;
    G.yp[l] = plan.yo[l].p;
    G.yn[l] = plan.yo[l].norm;
    G.ldp[l] = plan.xo[l].ldp;
  }
  int grid = (int)min(ceil_div64(n, 64), (int64_t)ctx->n_sm * 4);
  const int prof_cls = prof_class<MODE>(ctx, plan, n, m);
  MB_LAUNCH_P(ctx, prof_cls, cov_general_kernel<MODE>, grid, 256, 0, plan.dprog, G, n, m, out, ldo, w, mu);
  return 0;
}

int cov_build_impl(mb_ctx* ctx, const mb_kprog* prog, const mb_mat* x, const mb_mat* y, double* out, int64_t ldo) {
  if (x->rows == 0 || y->rows == 0) return 0;
  // tensor-core route (mb_cov_i8.cu): ONE exponential-family leaf, cells x landmarks (x != y), large shapes, D <= 64
  if (ctx->opt_cov == 0 && ctx->opt_cov_i8 != 0 && x->p != y->p) {
    Program pr;
    MB_TRY(parse_program(prog, x->cols, &pr));
    if (pr.leaves.size() == 1 && pr.n_ops == 1) {
      const Leaf& lf = pr.leaves[0];
      const double sc = mma_scale(lf);
      if (sc > 0.0) {
        bool done = false;
        const int d0 = lf.all_dims ? (int)x->cols : (int)lf.dims.size();
        MB_TRY(mb_cov_i8_build(ctx, lf.kind, sc, x, y, lf.all_dims ? nullptr : lf.dims.data(), d0, out, ldo, &done));
        if (done) {
          if (ctx->prof_on)   // algorithmic bytes: K written once, both inputs read once
            ctx->prof_work[MB_PROF_COV] += 8.0 * ((double)x->rows * y->rows + (double)x->cols * ((double)x->rows + y->rows));
          return mb_i8_check(ctx);   // protocol time-out / non-finite operand flag (one stream sync per K1 call)
        }
      }
    }
  }
  Plan plan;
  MB_TRY(prepare(ctx, prog, x, y, &plan, true));
  bool done = false;
  MB_TRY(launch_mma<MODE_STORE>(ctx, plan, x->rows, y->rows, out, ldo, nullptr, 0.0, &done));
  if (!done && ctx->opt_cov != 2) MB_TRY(launch_fast<MODE_STORE>(ctx, plan, x->rows, y->rows, out, ldo, nullptr, 0.0, &done));
  if (!done) MB_TRY(launch_general<MODE_STORE>(ctx, plan, x->rows, y->rows, out, ldo, nullptr, 0.0));
  return 0;
}

}  // namespace

extern "C" int mb_cov_build(mb_ctx* ctx, const mb_kprog* prog, const mb_mat* x, const mb_mat* y, mb_mat* K) {
  MB_RANGE("mellon_b200: K1 cov_build");
  MB_CHECK(ctx && prog && x && y && K, "mb_cov_build: null argument");
  MB_CHECK(K->rows == x->rows && K->cols == y->rows, "mb_cov_build: K is %lld x %lld, expected %lld x %lld",
           (long long)K->rows, (long long)K->cols, (long long)x->rows, (long long)y->rows);
  MB_CUDA(cudaSetDevice(ctx->device));
  return cov_build_impl(ctx, prog, x, y, K->p, K->cols);
}

extern "C" int mb_cov_diag(mb_ctx* ctx, const mb_kprog* prog, const mb_mat* x, mb_mat* out) {
  MB_CHECK(ctx && prog && x && out, "mb_cov_diag: null argument");
  MB_CHECK(out->rows * out->cols == x->rows, "mb_cov_diag: output has %lld entries for %lld rows",
           (long long)(out->rows * out->cols), (long long)x->rows);
  MB_CUDA(cudaSetDevice(ctx->device));
  if (x->rows == 0) return 0;
  Plan plan;
  MB_TRY(prepare(ctx, prog, x, x, &plan));
  GenLeafPtrs G;
  memset(&G, 0, sizeof(G));
  for (size_t l = 0; l < plan.xo.size(); l++) G.xn[l] = plan.xo[l].norm;
  MB_LAUNCH(ctx, cov_diag_kernel, (int)ceil_div64(x->rows, 256), 256, 0, plan.dprog, G, x->rows, out->p);
  return 0;
}

extern "C" int mb_cov_matvec(mb_ctx* ctx, const mb_kprog* prog, const mb_mat* xq, const mb_mat* base,
                             const mb_mat* w, double mu, mb_mat* out) {
  MB_RANGE("mellon_b200: K7 cov_matvec");
  MB_CHECK(ctx && prog && xq && base && w && out, "mb_cov_matvec: null argument");
  MB_CHECK(w->rows == base->rows, "mb_cov_matvec: weights have %lld rows for %lld base points",
           (long long)w->rows, (long long)base->rows);
  MB_CHECK(out->rows == xq->rows && out->cols == w->cols, "mb_cov_matvec: output must be (%lld, %lld)",
           (long long)xq->rows, (long long)w->cols);
  MB_CUDA(cudaSetDevice(ctx->device));
  if (xq->rows == 0 || w->cols == 0) return 0;
  if (base->rows == 0) return mb_mat_fill(ctx, out, mu);
  const int64_t p = w->cols;
  if (p == 1) {
    Plan plan;
    MB_TRY(prepare(ctx, prog, xq, base, &plan, true));
    bool done = false;
    MB_TRY(launch_mma<MODE_MATVEC>(ctx, plan, xq->rows, base->rows, out->p, 1, w->p, mu, &done));
    if (!done && ctx->opt_cov != 2)
      MB_TRY(launch_fast<MODE_MATVEC>(ctx, plan, xq->rows, base->rows, out->p, 1, w->p, mu, &done));
    if (!done) MB_TRY(launch_general<MODE_MATVEC>(ctx, plan, xq->rows, base->rows, out->p, 1, w->p, mu));
    return 0;
  }
  // several output columns: build K for row chunks and contract on the tensor pipe
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(xq->rows, (int64_t)(256 << 20) / (8 * base->rows)));
  mb_mat* Kc = nullptr;
  MB_TRY(mb_mat_alloc(ctx, chunk, base->rows, &Kc));
  int rc = 0;
  for (int64_t r0 = 0; r0 < xq->rows && rc == 0; r0 += chunk) {
    int64_t nr = min(chunk, xq->rows - r0);
    mb_mat xv = {xq->p + r0 * xq->cols, nr, xq->cols, ctx, false};
    rc = cov_build_impl(ctx, prog, &xv, base, Kc->p, base->rows);
    if (rc == 0) {
      mb_mat ov = {out->p + r0 * p, nr, p, ctx, false};
      rc = mb_mat_fill(ctx, &ov, mu);
      // out = K w + mu : A = Kc (nr x m, not k-major), B(j,k) = w[k*p + j] (k-major)
      if (rc == 0) rc = mb_gemm_raw(ctx, false, true, nr, p, base->rows, 1.0, Kc->p, base->rows, w->p, p, 1.0, ov.p, p, false);
    }
  }
  mb_mat_free(ctx, Kc);
  return rc;
}

extern "C" int mb_predict_mean(mb_ctx* ctx, const mb_kprog* prog, const double* xq_host, int64_t nq, int64_t d,
                               const mb_mat* base, const mb_mat* w, double mu, double* out_host) {
  MB_RANGE("mellon_b200: K7 predict_mean (host queries)");
  MB_CHECK(ctx && prog && base && w && (nq == 0 || (xq_host && out_host)), "mb_predict_mean: null argument");
  MB_CHECK(d == base->cols, "mb_predict_mean: queries have %lld features, base points %lld", (long long)d,
           (long long)base->cols);
  MB_CUDA(cudaSetDevice(ctx->device));
  if (nq == 0) return 0;
  const int64_t p = w->cols;
  const int64_t chunk = std::min<int64_t>(nq, 1 << 18);
  mb_mat *xb[2] = {nullptr, nullptr}, *ob[2] = {nullptr, nullptr};
  cudaEvent_t up[2], done[2];
  int rc = 0;
  for (int b = 0; b < 2 && rc == 0; b++) {
    rc = mb_mat_alloc(ctx, chunk, d, &xb[b]);
    if (rc == 0) rc = mb_mat_alloc(ctx, chunk, p, &ob[b]);
    cudaEventCreateWithFlags(&up[b], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming);
  }
  // copy stream: H2D of chunk c+1 overlaps the kernel of chunk c; D2H results ride the same stream
  const int64_t nchunks = ceil_div64(nq, chunk);
  auto upload = [&](int64_t c) -> int {
    int b = (int)(c & 1);
    int64_t r0 = c * chunk, nr = min(chunk, nq - r0);
    MB_CUDA(cudaMemcpyAsync(xb[b]->p, xq_host + r0 * d, (size_t)nr * d * sizeof(double), cudaMemcpyHostToDevice,
                            ctx->copy_stream));
    MB_CUDA(cudaEventRecord(up[b], ctx->copy_stream));
    return 0;
  };
  if (rc == 0) rc = upload(0);
  for (int64_t c = 0; c < nchunks && rc == 0; c++) {
    int b = (int)(c & 1);
    int64_t r0 = c * chunk, nr = min(chunk, nq - r0);
    if (c + 1 < nchunks) {
      // buffer b^1 was last read by the kernel of chunk c-1, which the compute stream has
      // already been told to finish before its D2H; make the copy stream wait for it.
      if (c >= 1) cudaStreamWaitEvent(ctx->copy_stream, done[b ^ 1], 0);
      rc = upload(c + 1);
      if (rc) break;
    }
    cudaStreamWaitEvent(ctx->stream, up[b], 0);
    mb_mat xv = {xb[b]->p, nr, d, ctx, false};
    mb_mat ov = {ob[b]->p, nr, p, ctx, false};
    rc = mb_cov_matvec(ctx, prog, &xv, base, w, mu, &ov);
    if (rc) break;
    if (cudaMemcpyAsync(out_host + r0 * p, ob[b]->p, (size_t)nr * p * sizeof(double), cudaMemcpyDeviceToHost,
                        ctx->stream) != cudaSuccess) {
      mb_set_error("mb_predict_mean: D2H copy failed");
      rc = -1;
      break;
    }
    cudaEventRecord(done[b], ctx->stream);
  }
  cudaStreamSynchronize(ctx->copy_stream);
  cudaStreamSynchronize(ctx->stream);
  for (int b = 0; b < 2; b++) {
    mb_mat_free(ctx, xb[b]);
    mb_mat_free(ctx, ob[b]);
    cudaEventDestroy(up[b]);
    cudaEventDestroy(done[b]);
  }
  return rc;
}

extern "C" int mb_nn_distances(mb_ctx* ctx, const mb_mat* x, const mb_mat* all, int64_t self_offset, mb_mat* dist,
                               int64_t* idx_host) {
  MB_RANGE("mellon_b200: nn_distances");
  MB_CHECK(ctx && x && all && dist, "mb_nn_distances: null argument");
  MB_CHECK(x->cols == all->cols, "mb_nn_distances: feature counts differ (%lld vs %lld)", (long long)x->cols,
           (long long)all->cols);
  MB_CHECK(dist->rows * dist->cols == x->rows, "mb_nn_distances: output has %lld entries for %lld rows",
           (long long)(dist->rows * dist->cols), (long long)x->rows);
  MB_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = x->rows, m = all->rows;
  if (n == 0) return 0;
  mb_kop op = {MB_OP_LEAF, MB_K_DISTANCE, 1.0, 1.0, 0.0, 0, -1};
  mb_kprog prog = {1, 0, &op, nullptr};
  int64_t* idx_dev = nullptr;
  MB_CUDA(mb_dev_malloc(ctx, (void**)&idx_dev, sizeof(int64_t) * n));
  int rc = 0;
  // tensor-core route (mb_cov_i8.cu, exact int8 digit-slice contraction + running-minimum epilogue): D <= 64, large shapes
  bool done_i8 = false;
  rc = mb_cov_i8_build(ctx, MB_K_DISTANCE, 1.0, x, all, nullptr, (int)x->cols, dist->p, 1, &done_i8, self_offset, idx_dev);
  if (rc == 0 && !done_i8) {
    Plan plan;
    rc = prepare(ctx, &prog, x, all, &plan);
    bool done = false;
    if (rc == 0) {
      const Operand& xo = plan.xo[0];
      const Operand& yo = plan.yo[0];
      size_t smem = (size_t)(BM + 2 * BN) * xo.ldp * sizeof(double);
      if (smem > 200 * 1024) {
        mb_set_error("mb_nn_distances: %lld features are too many for the tile kernel", (long long)x->cols);
        rc = -2;
      } else {
        static mb_per_device_flag cfg;   
        if (!cfg(ctx)) {
          cudaFuncSetAttribute(cov_tile_kernel<MB_K_DISTANCE, MODE_NNMIN>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          cfg(ctx) = true;
        }
        int grid = (int)min(ceil_div64(n, BM), (int64_t)ctx->n_sm);
        cov_tile_kernel<MB_K_DISTANCE, MODE_NNMIN><<<grid, NT, smem, ctx->stream>>>(
            xo.p, xo.norm, n, yo.p, yo.norm, m, xo.ldp, plan.dprog.leaf[0], dist->p, 1, nullptr, 0.0, self_offset,
            idx_dev);
        ctx->launches++;
        done = cudaGetLastError() == cudaSuccess;
        if (!done) { mb_set_error("mb_nn_distances: launch failed"); rc = -1; }
      }
    }
  }
  if (rc == 0) {
    int grid = (int)min((int64_t)ctx->n_sm * 8, ceil_div64(n, 8));
    nn_exact_kernel<<<grid, 256, 0, ctx->stream>>>(x->p, n, all->p, (int)x->cols, idx_dev, dist->p);
    ctx->launches++;
    if (idx_host) cudaMemcpyAsync(idx_host, idx_dev, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, ctx->stream);
  }
  cudaStreamSynchronize(ctx->stream);
  cudaFree(idx_dev);
  if (rc == 0 && cudaGetLastError() != cudaSuccess) { mb_set_error("mb_nn_distances: kernel failed"); rc = -1; }
  if (rc == 0 && done_i8) rc = mb_i8_check(ctx);
  return rc;
}

extern "C" int mb_cov_chol(mb_ctx* ctx, const mb_kprog* prog, const mb_mat* xu, double diag_add, mb_mat* Lp) {
  MB_CHECK(ctx && prog && xu && Lp, "mb_cov_chol: null argument");
  MB_CHECK(Lp->rows == xu->rows && Lp->cols == xu->rows, "mb_cov_chol: Lp must be %lld x %lld",
           (long long)xu->rows, (long long)xu->rows);
  MB_TRY(mb_cov_build(ctx, prog, xu, xu, Lp));
  MB_TRY(mb_mat_add_diag(ctx, Lp, diag_add));
  return mb_potrf(ctx, Lp);
}

extern "C" int mb_lowrank_standard(mb_ctx* ctx, const mb_kprog* prog, const mb_mat* x, const mb_mat* xu,
                                   const mb_mat* Lp, mb_mat* L) {
  MB_CHECK(ctx && prog && x && xu && Lp && L, "mb_lowrank_standard: null argument");
  MB_TRY(mb_cov_build(ctx, prog, x, xu, L));
  return mb_trsm_right_lt(ctx, Lp, L);
}
