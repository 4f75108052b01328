// K5 / K6 and the O(N r) passes over L: MAP objective value + gradient, transform,
// Hessian diagonal, L^T t.  HBM-bound streaming kernels; every reduction is a fixed-order
// tree so results are bit-reproducible run to run and identical on every rank after the
// all-reduce.
#include "mb_common.cuh"

namespace {

constexpr int ROW_THREADS = 256;   // 8 warps per CTA, one row per warp at a time
enum { MODE_TRANSFORM = 0, MODE_GRAD = 1, MODE_HESS = 2 };

// Segments: the unit of work of every pass that sums over cells.  A segment is a run of `sr` consecutive GLOBAL rows
// inside one chunk of the fixed reduction tree (mb_common.cuh); its rows are summed in row order by one CTA, the
// `spc` segment sums of a chunk are added in segment order, and the chunk sums go through the pairwise tree.  None of
// this depends on which rank holds the rows, so the bits of the result do not depend on the number of ranks.
constexpr int SEG_PER_CHUNK = 37;   // 32 chunks x 37 = 1184 = 8 x 148 SMs: whole waves on 1, 2, 4 and 8 ranks
constexpr int SEG_MIN_ROWS = 16;
struct SegGrid {
  int64_t cr, sr, G, row_lo;
  int spc, c_lo, c_hi, nseg;
};
SegGrid make_seg_grid(const mb_chunks& g) {
  SegGrid s;
  s.cr = g.cr;
  s.G = g.G;
  s.row_lo = g.row_lo;
  s.sr = std::max<int64_t>(SEG_MIN_ROWS, ceil_div64(g.cr, SEG_PER_CHUNK));
  s.spc = (int)ceil_div64(g.cr, s.sr);
  s.c_lo = g.c_lo;
  s.c_hi = g.c_hi;
  s.nseg = (g.c_hi - g.c_lo) * s.spc;
  return s;
}
// local row range [i0, i1) of local segment `seg` (possibly empty)
__device__ __forceinline__ void seg_rows(const SegGrid& s, int seg, int64_t* i0, int64_t* i1) {
  const int c = s.c_lo + seg / s.spc, j = seg % s.spc;
  const int64_t cend = min(s.G, (int64_t)(c + 1) * s.cr);
  const int64_t a = min(cend, (int64_t)c * s.cr + (int64_t)j * s.sr), b = min(cend, a + s.sr);
  *i0 = a - s.row_lo;
  *i1 = b - s.row_lo;
}

// f_i = L[i,:] . z + mu ; per mode:
//   TRANSFORM: out_f[i] = f_i
//   GRAD:      wv[i] = exp(f_i + V_i) - 1 ; lv[i] = f_i - A_i
//   HESS:      wv[i] = exp(f_i + V_i)
template <int MODE>
__global__ void __launch_bounds__(ROW_THREADS)
rowdot_kernel(const double* __restrict__ L, int64_t n, int r, const double* __restrict__ z, double mu,
              const double* __restrict__ V, double* __restrict__ out, double* __restrict__ lv) {
  extern __shared__ double zs[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = tid; c < r; c += ROW_THREADS) zs[c] = z[c];
  __syncthreads();
  const int64_t warps_total = (int64_t)gridDim.x * (ROW_THREADS / 32);
  const bool vec = ((r & 1) == 0) && ((reinterpret_cast<uintptr_t>(L) & 15) == 0);
  for (int64_t i = (int64_t)blockIdx.x * (ROW_THREADS / 32) + warp; i < n; i += warps_total) {
    const double* row = L + i * r;
    double s0 = 0.0, s1 = 0.0;
    if (vec) {
      const double2* row2 = reinterpret_cast<const double2*>(row);
      const int r2 = r >> 1;
      int c = lane;
      for (; c + 96 < r2; c += 128) {
        double2 a = row2[c], b = row2[c + 32], d = row2[c + 64], e = row2[c + 96];
        s0 = fma(a.x, zs[2 * c], s0);            s1 = fma(a.y, zs[2 * c + 1], s1);
        s0 = fma(b.x, zs[2 * (c + 32)], s0);     s1 = fma(b.y, zs[2 * (c + 32) + 1], s1);
        s0 = fma(d.x, zs[2 * (c + 64)], s0);     s1 = fma(d.y, zs[2 * (c + 64) + 1], s1);
        s0 = fma(e.x, zs[2 * (c + 96)], s0);     s1 = fma(e.y, zs[2 * (c + 96) + 1], s1);
      }
      for (; c < r2; c += 32) {
        double2 a = row2[c];
        s0 = fma(a.x, zs[2 * c], s0);
        s1 = fma(a.y, zs[2 * c + 1], s1);
      }
    } else {
      for (int c = lane; c < r; c += 32) s0 = fma(row[c], zs[c], s0);
    }
    double f = warp_sum(s0 + s1) + mu;
    if (lane == 0) {
      if (MODE == MODE_TRANSFORM) {
        out[i] = f;
      } else {
        double A = exp(f + V[i]);
        out[i] = (MODE == MODE_GRAD) ? (A - 1.0) : A;
        if (MODE == MODE_GRAD) lv[i] = f - A;
      }
    }
  }
}

// lpartial[seg] = sum of lv over the rows of the segment: one warp per segment, lane-strided then a shuffle tree
__global__ void __launch_bounds__(256)
seg_sum_kernel(const double* __restrict__ lv, const SegGrid sg, double* __restrict__ lpartial) {
  const int seg = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (seg >= sg.nseg) return;
  int64_t i0, i1;
  seg_rows(sg, seg, &i0, &i1);
  double s = 0.0;
  for (int64_t i = i0 + lane; i < i1; i += 32) s += lv[i];
  s = warp_sum(s);
  if (lane == 0) lpartial[seg] = s;
}

// partial[seg][c] = sum_{i in segment} wv[i] * (SQ ? L[i,c]^2 : L[i,c])
template <bool SQ>
__global__ void __launch_bounds__(256)
colsum_kernel(const double* __restrict__ L, int64_t n, int r, const double* __restrict__ wv,
              const SegGrid sg, double* __restrict__ partial) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 2;
  int64_t i0, i1;
  seg_rows(sg, blockIdx.y, &i0, &i1);
  if (c >= r) return;
  const bool pair = (c + 1 < r);
  const bool vec = pair && ((r & 1) == 0) && ((reinterpret_cast<uintptr_t>(L) & 15) == 0);
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  int64_t i = i0;
  if (vec) {
    for (; i + 3 < i1; i += 4) {
      double2 v0 = *reinterpret_cast<const double2*>(L + i * r + c);
      double2 v1 = *reinterpret_cast<const double2*>(L + (i + 1) * r + c);
      double2 v2 = *reinterpret_cast<const double2*>(L + (i + 2) * r + c);
      double2 v3 = *reinterpret_cast<const double2*>(L + (i + 3) * r + c);
      double w0 = wv[i], w1 = wv[i + 1], w2 = wv[i + 2], w3 = wv[i + 3];
      if (SQ) {
        v0.x *= v0.x; v0.y *= v0.y; v1.x *= v1.x; v1.y *= v1.y;
        v2.x *= v2.x; v2.y *= v2.y; v3.x *= v3.x; v3.y *= v3.y;
      }
      a0 = fma(w0, v0.x, a0); a1 = fma(w0, v0.y, a1);
      b0 = fma(w1, v1.x, b0); b1 = fma(w1, v1.y, b1);
      a0 = fma(w2, v2.x, a0); a1 = fma(w2, v2.y, a1);
      b0 = fma(w3, v3.x, b0); b1 = fma(w3, v3.y, b1);
    }
  }
  for (; i < i1; i++) {
    double w = wv[i];
    double x = L[i * r + c], y = pair ? L[i * r + c + 1] : 0.0;
    if (SQ) { x *= x; y *= y; }
    a0 = fma(w, x, a0);
    a1 = fma(w, y, a1);
  }
  double* p = partial + (int64_t)blockIdx.y * r;
  p[c] = a0 + b0;
  if (pair) p[c + 1] = a1 + b1;
}

// leaves[c][j] = sum over the segments of chunk c, in segment order, of partial[seg][j]  (j < r), and of
// lpartial[seg] for j == r when has_l; chunks this rank does not hold get an exact zero.  FUSE (one rank): the
// pairwise tree over the chunk sums is taken right here and out[j] written instead.
template <bool FUSE>
__global__ void seg_reduce_kernel(const double* __restrict__ partial, const double* __restrict__ lpartial, int r,
                                  int has_l, const SegGrid sg, double* __restrict__ dst) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int count = r + (has_l ? 1 : 0);
  if (j >= count) return;
  double v[MB_NCHUNK];
#pragma unroll
  for (int c = 0; c < MB_NCHUNK; c++) {
    double s = 0.0;
    if (c >= sg.c_lo && c < sg.c_hi) {
      const int s0 = (c - sg.c_lo) * sg.spc;
      if (j < r) {
        for (int q = 0; q < sg.spc; q++) s += partial[(int64_t)(s0 + q) * r + j];
      } else {
        for (int q = 0; q < sg.spc; q++) s += lpartial[s0 + q];
      }
    }
    v[c] = s;
  }
  if (FUSE) {
#pragma unroll
    for (int w = MB_NCHUNK / 2; w >= 1; w >>= 1) {
#pragma unroll
      for (int i = 0; i < w; i++) v[i] = v[2 * i] + v[2 * i + 1];
    }
    dst[j] = v[0];
  } else {
#pragma unroll
    for (int c = 0; c < MB_NCHUNK; c++) dst[(int64_t)c * count + j] = v[c];
  }
}

// ---- fused single pass (default): L is read from HBM exactly once per evaluation -----------
// CTA = 512 threads; thread t owns columns {2t, 2t+1} + 1024 q (q < CP column pairs) and keeps
// their gradient accumulators in registers.  Rows are processed RG at a time: the CTA loads
// the RG x r slab straight into registers (coalesced 16-byte loads), block-reduces the RG row
// dots, evaluates A = exp(f + V) and accumulates (A-1) * L from the same registers.
template <int CP, int RG, bool SQ>
__global__ void __launch_bounds__(512, 1)
fused_rows_kernel(const double* __restrict__ L, int64_t n, int r, const double* __restrict__ z, double mu,
                  const double* __restrict__ V, const SegGrid sg, double* __restrict__ partial,
                  double* __restrict__ lpartial) {
  constexpr int T = 512, NW = T / 32;
  __shared__ double red[2][RG][NW];
  __shared__ double fsh[2][RG];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t i0, i1;
  seg_rows(sg, blockIdx.x, &i0, &i1);

  double2 zr[CP], g[CP];
#pragma unroll
  for (int q = 0; q < CP; q++) {
    int c = 2 * tid + 2 * T * q;
    zr[q].x = (c < r) ? z[c] : 0.0;
    zr[q].y = (c + 1 < r) ? z[c + 1] : 0.0;
    g[q] = make_double2(0.0, 0.0);
  }
  double lsum = 0.0;
  int buf = 0;
  for (int64_t ib = i0; ib < i1; ib += RG, buf ^= 1) {
    double2 v[RG][CP];
#pragma unroll
    for (int rr = 0; rr < RG; rr++) {
      const int64_t i = ib + rr;
#pragma unroll
      for (int q = 0; q < CP; q++) {
        int c = 2 * tid + 2 * T * q;
        v[rr][q] = (i < i1 && c < r) ? *reinterpret_cast<const double2*>(L + i * r + c)
                                     : make_double2(0.0, 0.0);
      }
    }
#pragma unroll
    for (int rr = 0; rr < RG; rr++) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < CP; q++) s = fma(v[rr][q].x, zr[q].x, fma(v[rr][q].y, zr[q].y, s));
      s = warp_sum(s);
      if (lane == 0) red[buf][rr][warp] = s;
    }
    __syncthreads();
    if (tid < RG) {
      double f = mu;
#pragma unroll
      for (int w = 0; w < NW; w++) f += red[buf][tid][w];
      const int64_t i = ib + tid;
      double wgt = 0.0;
      if (i < i1) {
        double A = exp(f + V[i]);
        wgt = SQ ? A : (A - 1.0);
        if (!SQ) lsum += f - A;
      }
      fsh[buf][tid] = wgt;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < RG; rr++) {
      const double wgt = fsh[buf][rr];
#pragma unroll
      for (int q = 0; q < CP; q++) {
        double x = v[rr][q].x, y = v[rr][q].y;
        if (SQ) { x *= x; y *= y; }
        g[q].x = fma(wgt, x, g[q].x);
        g[q].y = fma(wgt, y, g[q].y);
      }
    }
  }
  double* p = partial + (int64_t)blockIdx.x * r;
#pragma unroll
  for (int q = 0; q < CP; q++) {
    int c = 2 * tid + 2 * T * q;
    if (c < r) p[c] = g[q].x;
    if (c + 1 < r) p[c + 1] = g[q].y;
  }
  if (!SQ) {
    // lsum lives in threads 0..RG-1 of warp 0
    if (warp == 0) {
      double s = warp_sum(tid < RG ? lsum : 0.0);
      if (lane == 0) lpartial[blockIdx.x] = s;
    }
  }
}


// ---- streaming single pass (default): bulk-TMA ring, L read from HBM exactly once ------------
// One persistent CTA per SM owns a contiguous block of rows.  Thread 0 keeps `ns - 2` slabs of
// `rb` rows (rb * r contiguous doubles = ONE cp.async.bulk each) in flight into a shared-memory
// ring; the 16 warps take the row dots of slab k from shared memory, meet at ONE __syncthreads,
// then warp 0 turns the dots into weights (A - 1 or A) while everyone accumulates the gradient
// of slab k-1 from the copy still sitting in the ring.  Column c = 2 t + 1024 q belongs to
// thread t, so the gradient lives in registers and every shared-memory access is a conflict-free
// 16-byte load.  Summation order is fixed => bit-reproducible.
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void s_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SW_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SW_DONE;\n"
      "bra SW_WAIT;\n"
      "SW_DONE:\n"
      "}\n" ::"r"(s_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void s_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}

constexpr int ST = 448, SNW = ST / 32, SMAX_RB = 8, SMAX_NS = 8;  // 14 compute warps
constexpr int ST_ALL = ST + 64;  // + 1 scalar warp + 1 producer warp = 512 threads (128 registers each)
constexpr int SRED = 16;          // partial-sum slots per row (slots >= SNW stay zero)

__device__ __forceinline__ void s_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s_u32(bar)) : "memory");
}

// Warp roles (no __syncthreads in the row loop; everything is mbarrier hand-offs):
//   producer warp : waits for a ring stage to be released by the 14 compute warps, re-arms its
//                   `full` barrier and issues the bulk copy of the next slab into it;
//   compute warps : slab k   -> partial row dots (columns c = 2t + 896q of thread t), warp-reduced,
//                               posted to red[k&1] and signalled on dot_ready[k&1];
//                   slab k-1 -> gradient / Hessian accumulation with the weights the scalar warp
//                               published on w_ready[(k-1)&1] one iteration earlier, then the stage is
//                               released on empty[];
//   scalar warp   : sums the 14 warp partials of each row in a fixed order, f = mu + dot,
//                   A = exp(f + V), weight = A - 1 (or A), loss partial sum; publishes wsh[k&1].
// The serial exp chain of the scalar warp therefore overlaps the compute warps' next slab instead
// of stalling them (the single-barrier version sat at 2700 clk per 40 KB row, 4.4 TB/s).
template <int CP, bool SQ, bool HOLD>
__global__ void __launch_bounds__(ST_ALL, 1)
stream_rows_kernel(const double* __restrict__ L, int64_t n, int r, const double* __restrict__ z, double mu,
                   const double* __restrict__ V, const SegGrid sg, int rb, int ns,
                   double* __restrict__ partial, double* __restrict__ lpartial) {
  extern __shared__ __align__(128) unsigned char stream_smem[];
  double* ring = reinterpret_cast<double*>(stream_smem);
  __shared__ __align__(8) uint64_t full[SMAX_NS], empty[SMAX_NS], dot_ready[2], w_ready[2];
  __shared__ double red[2][SMAX_RB][SRED];
  __shared__ double wsh[2][SMAX_RB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t i0, i1;
  seg_rows(sg, blockIdx.x, &i0, &i1);
  const int64_t nslab = (i1 > i0) ? (i1 - i0 + rb - 1) / rb : 0;
  const size_t stage_doubles = (size_t)rb * r;

  for (int e = tid; e < 2 * SMAX_RB * SRED; e += ST_ALL) (&red[0][0][0])[e] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < ns; s++) {
      s_mbar_init(&full[s], 1);
      s_mbar_init(&empty[s], SNW);
    }
    for (int b = 0; b < 2; b++) {
      s_mbar_init(&dot_ready[b], SNW);
      s_mbar_init(&w_ready[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == SNW + 1) {
    // ---- producer ----
    if (lane == 0) {
      for (int64_t k = 0; k < nslab; k++) {
        const int s = (int)(k % ns);
        if (k >= ns) s_mbar_wait(&empty[s], (uint32_t)(((k / ns) - 1) & 1));
        const int64_t row = i0 + k * rb;
        const uint32_t bytes = (uint32_t)(min((int64_t)rb, i1 - row) * r * sizeof(double));
        s_mbar_expect_tx(&full[s], bytes);
        s_bulk_g2s(ring + s * stage_doubles, L + row * r, bytes, &full[s]);
      }
    }
    return;
  }

  if (warp == SNW) {
    // ---- scalar warp: lane l serves row l >> 2, summing the partials of warps 4 (l & 3) .. +3 ----
    const int rr = lane >> 2, part = lane & 3;
    double lsum = 0.0;
    for (int64_t k = 0; k < nslab; k++) {
      const int pb = (int)(k & 1);
      const int64_t row0 = i0 + k * rb;
      const int rows_k = (int)min((int64_t)rb, i1 - row0);
      const double vk = (rr < rows_k) ? V[row0 + rr] : 0.0;  // in flight while the dots are awaited
      s_mbar_wait(&dot_ready[pb], (uint32_t)((k >> 1) & 1));
      double s = 0.0;
      if (rr < rows_k) s = ((red[pb][rr][4 * part] + red[pb][rr][4 * part + 1]) + red[pb][rr][4 * part + 2]) +
                           red[pb][rr][4 * part + 3];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (part == 0 && rr < rows_k) {
        const double f = mu + s;
        const double A = exp(f + vk);
        wsh[pb][rr] = SQ ? A : (A - 1.0);
        if (!SQ) lsum += f - A;
      }
      __syncwarp();
      if (lane == 0) s_mbar_arrive(&w_ready[pb]);
    }
    if (!SQ) {
      // lanes 0, 4, ..., 28 hold the row-wise partial sums: fixed-order tree
      double t = (part == 0) ? lsum : 0.0;
      t = warp_sum(t);
      if (lane == 0) lpartial[blockIdx.x] = t;
    }
    return;
  }

  // ---- compute warps ----
  double2 zr[CP], g[CP];
#pragma unroll
  for (int q = 0; q < CP; q++) {
    const int c = 2 * tid + 2 * ST * q;
    zr[q].x = (c < r) ? z[c] : 0.0;
    zr[q].y = (c + 1 < r) ? z[c + 1] : 0.0;
    g[q] = make_double2(0.0, 0.0);
  }
  if (HOLD) {
    // one row per slab (rb == 1): the row is read from shared memory ONCE, its stage is released at
    // once, and the copy held in registers serves the gradient one iteration later
    double2 vcur[CP], vprev[CP];
#pragma unroll
    for (int q = 0; q < CP; q++) vprev[q] = make_double2(0.0, 0.0);
    for (int64_t k = 0; k <= nslab; k++) {
      const int pb = (int)(k & 1);
      if (k < nslab) {
        s_mbar_wait(&full[k % ns], (uint32_t)((k / ns) & 1));
        const double* row = ring + (size_t)(k % ns) * stage_doubles;
#pragma unroll
        for (int q = 0; q < CP; q++) {
          const int c = 2 * tid + 2 * ST * q;
          vcur[q] = (c < r) ? *reinterpret_cast<const double2*>(row + c) : make_double2(0.0, 0.0);
        }
        __syncwarp();
        if (lane == 0) s_mbar_arrive(&empty[k % ns]);
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int q = 0; q < CP; q++) {
          s0 = fma(vcur[q].x, zr[q].x, s0);
          s1 = fma(vcur[q].y, zr[q].y, s1);
        }
        const double s = warp_sum(s0 + s1);
        if (lane == 0) {
          red[pb][0][warp] = s;
          s_mbar_arrive(&dot_ready[pb]);
        }
      }
      if (k >= 1) {
        s_mbar_wait(&w_ready[pb ^ 1], (uint32_t)(((k - 1) >> 1) & 1));
        const double wgt = wsh[pb ^ 1][0];
#pragma unroll
        for (int q = 0; q < CP; q++) {
          double2 v = vprev[q];
          if (SQ) { v.x *= v.x; v.y *= v.y; }
          g[q].x = fma(wgt, v.x, g[q].x);
          g[q].y = fma(wgt, v.y, g[q].y);
        }
      }
#pragma unroll
      for (int q = 0; q < CP; q++) vprev[q] = vcur[q];
    }
  } else {
    for (int64_t k = 0; k <= nslab; k++) {
      const int pb = (int)(k & 1);
      if (k < nslab) {
        const int rows_k = (int)min((int64_t)rb, i1 - (i0 + k * rb));
        s_mbar_wait(&full[k % ns], (uint32_t)((k / ns) & 1));
        const double* slab = ring + (size_t)(k % ns) * stage_doubles;
        for (int rr = 0; rr < rows_k; rr++) {
          const double* row = slab + (size_t)rr * r;
          double s0 = 0.0, s1 = 0.0;
  #pragma unroll
          for (int q = 0; q < CP; q++) {
            const int c = 2 * tid + 2 * ST * q;
            if (c < r) {
              const double2 v = *reinterpret_cast<const double2*>(row + c);
              s0 = fma(v.x, zr[q].x, s0);
              s1 = fma(v.y, zr[q].y, s1);
            }
          }
          const double s = warp_sum(s0 + s1);
          if (lane == 0) red[pb][rr][warp] = s;
        }
        __syncwarp();
        if (lane == 0) s_mbar_arrive(&dot_ready[pb]);
      }
      if (k >= 1) {
        const int64_t kp = k - 1;
        const int rows_p = (int)min((int64_t)rb, i1 - (i0 + kp * rb));
        s_mbar_wait(&w_ready[pb ^ 1], (uint32_t)((kp >> 1) & 1));
        const double* slab = ring + (size_t)(kp % ns) * stage_doubles;
        for (int rr = 0; rr < rows_p; rr++) {
          const double wgt = wsh[pb ^ 1][rr];
          const double* row = slab + (size_t)rr * r;
  #pragma unroll
          for (int q = 0; q < CP; q++) {
            const int c = 2 * tid + 2 * ST * q;
            if (c < r) {
              double2 v = *reinterpret_cast<const double2*>(row + c);
              if (SQ) { v.x *= v.x; v.y *= v.y; }
              g[q].x = fma(wgt, v.x, g[q].x);
              g[q].y = fma(wgt, v.y, g[q].y);
            }
          }
        }
        __syncwarp();
        if (lane == 0) s_mbar_arrive(&empty[kp % ns]);
      }
    }
  }
  double* p = partial + (int64_t)blockIdx.x * r;
#pragma unroll
  for (int q = 0; q < CP; q++) {
    const int c = 2 * tid + 2 * ST * q;
    if (c < r) *reinterpret_cast<double2*>(p + c) = g[q];
  }
}

struct PassBuffers {
  double* zdev;      // r
  double* wv;        // n   (two-pass route: per-row weights)
  double* lv;        // n   (two-pass route: per-row loss terms)
  double* partial;   // nseg * r
  double* lpartial;  // nseg
  double* leaves;    // MB_NCHUNK * (r + 1)
  double* out;       // r + 1
};

int carve(mb_ctx* ctx, int64_t n, int r, int64_t nseg, PassBuffers* pb) {
  const size_t rp = (size_t)r + 2, np = (size_t)n + 2, sp = (size_t)nseg + 2;
  size_t need = (rp + 2 * np + (size_t)nseg * r + sp + (size_t)MB_NCHUNK * rp + rp + 8) * sizeof(double);
  double* s;
  MB_TRY(mb_scratch(ctx, need, &s));
  pb->zdev = s;
  pb->wv = pb->zdev + (rp & ~(size_t)1);
  pb->lv = pb->wv + (np & ~(size_t)1);
  pb->partial = pb->lv + (np & ~(size_t)1);
  pb->lpartial = pb->partial + (size_t)nseg * r + ((size_t)nseg * r & 1);
  pb->leaves = pb->lpartial + (sp & ~(size_t)1);
  pb->out = pb->leaves + (size_t)MB_NCHUNK * rp;
  return 0;
}

// segment partials -> chunk leaves -> tree -> pb->out (count = r, + 1 when the loss slot is carried)
int finish_reduce(mb_ctx* ctx, const mb_chunks& g, const SegGrid& sg, int r, bool has_l, PassBuffers* pb) {
  const int count = r + (has_l ? 1 : 0);
  const int grid = (int)ceil_div64(count, 128);
  if (!g.sharded) {
    MB_LAUNCH(ctx, seg_reduce_kernel<true>, grid, 128, 0, pb->partial, pb->lpartial, r, has_l ? 1 : 0, sg, pb->out);
    return 0;
  }
  MB_LAUNCH(ctx, seg_reduce_kernel<false>, grid, 128, 0, pb->partial, pb->lpartial, r, has_l ? 1 : 0, sg, pb->leaves);
  return mb_tree_reduce_small(ctx, g, pb->leaves, count, pb->out);
}

template <int MODE>
int launch_rowdot(mb_ctx* ctx, const mb_mat* L, const double* zdev, double mu, const double* V, double* out,
                  double* lv) {
  int grid = (int)min((int64_t)ctx->n_sm * 4, ceil_div64(L->rows, ROW_THREADS / 32));
  grid = max(grid, 1);
  size_t smem = (size_t)L->cols * sizeof(double);
  MB_CHECK(smem <= 160 * 1024, "rank %lld too large for the row-dot kernel (max 20480)", (long long)L->cols);
  static mb_per_device_flag configured[3];
  if (!configured[MODE](ctx)) {
    MB_CUDA(cudaFuncSetAttribute(rowdot_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured[MODE](ctx) = true;
  }
  MB_LAUNCH(ctx, rowdot_kernel<MODE>, grid, ROW_THREADS, smem, L->p, L->rows, (int)L->cols, zdev, mu, V, out, lv);
  return 0;
}

// two-pass route: (A-1 | A | t) weights in wv, then column sums per segment (+ the per-row loss terms in lv)
int colsum(mb_ctx* ctx, const mb_mat* L, const double* wv, bool sq, PassBuffers* pb, const SegGrid& sg, const double* lv) {
  const int r = (int)L->cols;
  if (sg.nseg == 0) return 0;
  dim3 grid((unsigned)ceil_div64(r, 512), (unsigned)sg.nseg);
  if (sq) MB_LAUNCH(ctx, colsum_kernel<true>, grid, 256, 0, L->p, L->rows, r, wv, sg, pb->partial);
  else MB_LAUNCH(ctx, colsum_kernel<false>, grid, 256, 0, L->p, L->rows, r, wv, sg, pb->partial);
  if (lv) MB_LAUNCH(ctx, seg_sum_kernel, (int)ceil_div64(sg.nseg, 8), 256, 0, lv, sg, pb->lpartial);
  return 0;
}

template <bool SQ>
int launch_fused(mb_ctx* ctx, const mb_mat* L, const double* zdev, double mu, const double* V, PassBuffers* pb,
                 const SegGrid& sg, bool* done) {
  const int r = (int)L->cols;
  const int64_t n = L->rows;
  *done = false;
  if ((r & 1) || (reinterpret_cast<uintptr_t>(L->p) & 15)) return 0;
  const int cp = (int)ceil_div64(r, 1024);
#define MB_FUSED(CPV, RGV)                                                                                \
  {                                                                                                       \
    if (ctx->prof_on) ctx->prof_work[MB_PROF_LOSSGRAD] += 8.0 * ((double)n * r + (double)n);                \
    MB_LAUNCH_P(ctx, MB_PROF_LOSSGRAD, (fused_rows_kernel<CPV, RGV, SQ>), sg.nseg, 512, 0, L->p, n, r, zdev, mu, V, sg, \
              pb->partial, pb->lpartial);                                                                 \
    *done = true;                                                                                         \
  }
  if (cp == 1) MB_FUSED(1, 8)
  else if (cp == 2) MB_FUSED(2, 8)
  else if (cp <= 3) MB_FUSED(3, 4)
  else if (cp <= 4) MB_FUSED(4, 4)
  else if (cp <= 5) MB_FUSED(5, 4)
  else if (cp <= 6) MB_FUSED(6, 2)
  else if (cp <= 8) MB_FUSED(8, 2)
#undef MB_FUSED
  return 0;
}


template <bool SQ>
int launch_stream(mb_ctx* ctx, const mb_mat* L, const double* zdev, double mu, const double* V, PassBuffers* pb,
                  const SegGrid& sg, bool* done) {
  const int r = (int)L->cols;
  const int64_t n = L->rows;
  *done = false;
  if ((r & 1) || (reinterpret_cast<uintptr_t>(L->p) & 15) || r > 8 * 1024) return 0;
  const size_t row_bytes = (size_t)r * sizeof(double);
  int rb = 1;
  while (rb < SMAX_RB && (size_t)(2 * rb) * row_bytes <= 40 * 1024) rb *= 2;
  const size_t stage = (size_t)rb * row_bytes;
  const size_t budget = 200 * 1024;
  int ns = (int)std::min<size_t>(SMAX_NS, budget / stage);
  if (ns < 3) return 0;
  const int grid = sg.nseg;
  const size_t smem = (size_t)ns * stage;
  const int cp = (int)ceil_div64(r, 2 * ST);
  const bool hold = (rb == 1 && cp <= 6);  // row copy in registers: 4 * cp more registers
#define MB_STREAM(CPV)                                                                                      \
  case CPV: {                                                                                               \
    static mb_per_device_flag cfg;                                                                                   \
    if (!cfg(ctx)) {                                                                                             \
      MB_CUDA(cudaFuncSetAttribute(stream_rows_kernel<CPV, SQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   (int)budget));                                                           \
      MB_CUDA(cudaFuncSetAttribute(stream_rows_kernel<CPV, SQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   (int)budget));                                                           \
      cfg(ctx) = true;                                                                                           \
    }                                                                                                       \
    if (ctx->prof_on) ctx->prof_work[MB_PROF_LOSSGRAD] += 8.0 * ((double)n * r + (double)n);                \
    if (hold) {                                                                                             \
      MB_LAUNCH_P(ctx, MB_PROF_LOSSGRAD, (stream_rows_kernel<CPV, SQ, true>), grid, ST_ALL, smem, L->p, n, r, zdev, mu, \
                  V, sg, rb, ns, pb->partial, pb->lpartial);                                                \
    } else {                                                                                                \
      MB_LAUNCH_P(ctx, MB_PROF_LOSSGRAD, (stream_rows_kernel<CPV, SQ, false>), grid, ST_ALL, smem, L->p, n, r, zdev, mu, \
                  V, sg, rb, ns, pb->partial, pb->lpartial);                                                \
    }                                                                                                       \
    break;                                                                                                  \
  }
  switch (cp) {
    MB_STREAM(1) MB_STREAM(2) MB_STREAM(3) MB_STREAM(4) MB_STREAM(5) MB_STREAM(6) MB_STREAM(7) MB_STREAM(8)
    MB_STREAM(9) MB_STREAM(10)
    default: return 0;
  }
#undef MB_STREAM
  *done = true;
  return 0;
}

// one pass over the local rows of L with weights exp(f + V) [- 1]: gradient / Hessian-diagonal sums (and the loss
// terms) per segment, then the fixed tree; the result (r [+ 1] doubles) is in pb->out on every rank
template <bool SQ>
int objective_pass(mb_ctx* ctx, const mb_mat* L, const mb_mat* V, double mu, const double* z_host, PassBuffers* pb) {
  const int64_t n = L->rows;
  const int r = (int)L->cols;
  mb_chunks g;
  MB_TRY(mb_chunk_grid(ctx, L, &g));
  const SegGrid sg = make_seg_grid(g);
  MB_TRY(carve(ctx, n, r, sg.nseg, pb));
  MB_CUDA(cudaMemcpyAsync(pb->zdev, z_host, (size_t)r * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (sg.nseg > 0) {
    bool done = false;
    if (ctx->opt_lossgrad == 0) MB_TRY(launch_stream<SQ>(ctx, L, pb->zdev, mu, V->p, pb, sg, &done));
    if (!done && ctx->opt_lossgrad != 1) MB_TRY(launch_fused<SQ>(ctx, L, pb->zdev, mu, V->p, pb, sg, &done));
    if (!done) {
      if (n > 0) MB_TRY((launch_rowdot<SQ ? MODE_HESS : MODE_GRAD>(ctx, L, pb->zdev, mu, V->p, pb->wv, pb->lv)));
      MB_TRY(colsum(ctx, L, pb->wv, SQ, pb, sg, SQ ? nullptr : pb->lv));
    }
  }
  return finish_reduce(ctx, g, sg, r, !SQ, pb);
}

}  // namespace

extern "C" int mb_transform(mb_ctx* ctx, const mb_mat* L, const double* z_host, double mu, double* f_host) {
  MB_RANGE("mellon_b200: transform");
  MB_CHECK(ctx && L && z_host && (f_host || L->rows == 0), "mb_transform: null argument");
  MB_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = L->rows;
  const int r = (int)L->cols;
  if (n == 0) return 0;
  PassBuffers pb;
  MB_TRY(carve(ctx, n, r, 1, &pb));
  MB_CUDA(cudaMemcpyAsync(pb.zdev, z_host, (size_t)r * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  MB_TRY(launch_rowdot<MODE_TRANSFORM>(ctx, L, pb.zdev, mu, nullptr, pb.wv, nullptr));
  MB_CUDA(cudaMemcpyAsync(f_host, pb.wv, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int mb_loss_grad(mb_ctx* ctx, const mb_mat* L, const mb_mat* V, double sum_vdr, double mu, double k,
                            const double* z_host, double* loss, double* grad_host) {
  MB_RANGE("mellon_b200: K5 loss_grad");
  MB_CHECK(ctx && L && V && z_host && loss && grad_host, "mb_loss_grad: null argument");
  MB_CHECK(V->rows * V->cols == L->rows, "mb_loss_grad: V has %lld entries for %lld cells",
           (long long)(V->rows * V->cols), (long long)L->rows);
  MB_CUDA(cudaSetDevice(ctx->device));
  const int r = (int)L->cols;
  PassBuffers pb;
  MB_TRY(objective_pass<false>(ctx, L, V, mu, z_host, &pb));
  double* host;
  MB_TRY(mb_pinned(ctx, (size_t)(r + 1) * sizeof(double), &host));
  MB_CUDA(cudaMemcpyAsync(host, pb.out, (size_t)(r + 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  double zz = 0.0;
  for (int c = 0; c < r; c++) {
    zz += z_host[c] * z_host[c];
    grad_host[c] = z_host[c] + host[c];
  }
  // loss = -(prior + likelihood) = 1/2|z|^2 + k/2 log(2 pi) - (sum(f - A) + sum Vdr)
  *loss = 0.5 * zz + 0.5 * k * 1.8378770664093453 - (host[r] + sum_vdr);
  return 0;
}

extern "C" int mb_hess_diag(mb_ctx* ctx, const mb_mat* L, const mb_mat* V, double mu, const double* z_host,
                            double* diag_host) {
  MB_RANGE("mellon_b200: K6 hess_diag");
  MB_CHECK(ctx && L && V && z_host && diag_host, "mb_hess_diag: null argument");
  MB_CHECK(V->rows * V->cols == L->rows, "mb_hess_diag: V has %lld entries for %lld cells",
           (long long)(V->rows * V->cols), (long long)L->rows);
  MB_CUDA(cudaSetDevice(ctx->device));
  const int r = (int)L->cols;
  PassBuffers pb;
  MB_TRY(objective_pass<true>(ctx, L, V, mu, z_host, &pb));
  double* host;
  MB_TRY(mb_pinned(ctx, (size_t)r * sizeof(double), &host));
  MB_CUDA(cudaMemcpyAsync(host, pb.out, (size_t)r * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int c = 0; c < r; c++) diag_host[c] = 1.0 + host[c];
  return 0;
}

// b = L^T t  [tree reduce over the cells of all ranks]
extern "C" int mb_gemv_t(mb_ctx* ctx, const mb_mat* L, const mb_mat* t, mb_mat* b) {
  MB_RANGE("mellon_b200: gemv_t");
  MB_CHECK(ctx && L && t && b, "mb_gemv_t: null argument");
  MB_CHECK(t->rows * t->cols == L->rows && b->rows * b->cols == L->cols, "mb_gemv_t: shape mismatch");
  MB_CUDA(cudaSetDevice(ctx->device));
  const int r = (int)L->cols;
  if (r == 0) return 0;
  mb_chunks g;
  MB_TRY(mb_chunk_grid(ctx, L, &g));
  const SegGrid sg = make_seg_grid(g);
  PassBuffers pb;
  MB_TRY(carve(ctx, 0, r, sg.nseg, &pb));
  MB_TRY(colsum(ctx, L, t->p, false, &pb, sg, nullptr));
  MB_TRY(finish_reduce(ctx, g, sg, r, false, &pb));
  MB_CUDA(cudaMemcpyAsync(b->p, pb.out, (size_t)r * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}
