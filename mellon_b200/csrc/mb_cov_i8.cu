// K1 on the 5th-generation tensor cores: the x.y contraction of the fused distance + covariance build as tcgen05.mma
// kind::i8 over 8-bit digit slices (exact to float64 rounding, see mb_i8.cu for the arithmetic), accumulators in TMEM,
// so that the FP64 pipe — which bounds the DMMA kernel of mb_cov.cu at ~25 ms for 1e6 x 5000 — only runs the epilogue
// (recombination, sqrt, exp, polynomial).  util.py:351-366 + cov.py `k` of one exponential-family leaf.
//
// Operands: every row of x (cells) and y (landmarks) is pre-scaled by the leaf's distance scale c (sqrt(5)/ls, ...),
// gets one power-of-two scale 2^E from its own largest component and becomes 54-bit fixed point, 7 balanced 8-bit
// digits per component; the squared norm is taken of the ROUNDED values, so sq = xn + yn - 2 x.y stays consistent.
// Digits are K-major, K padded to a multiple of 32 (D <= 64), per panel of 128 cells / 64 landmarks
//   [k-step][slice 7][k16 chunk 2][rows][16 B]      (+ the 64 norms and 64 scales behind every landmark panel)
// so ONE bulk copy fills the cell panel (56 KB) and one a landmark stage (29 KB).
//
// CTA (persistent, one per SM, 20 warps):
//   warp 0    producer: cell panel once per 128 cells, landmark stages through a 3-deep ring (cp.async.bulk + mbarrier)
//   warp 1    MMA issuer (one elected lane): per 128 x 64 tile the 28 digit pairs t + u <= 6 x K / 32 MMAs, grouped by
//             g = t + u into 7 TMEM accumulators of 64 columns, issued so that consecutive MMAs hit different
//             accumulators; the next tile is issued as soon as the epilogue has DRAINED this one (t_empty), so the
//             tensor pipe runs under the epilogue's sqrt / exp
//   warps 4-19 epilogue: warp w owns TMEM lanes 32 (w % 4) .. + 31 (the rows) and 16 columns; drains the 7 groups with
//             tcgen05.ld, folds them pairwise in int32 (|G_g| < 2^23), Horner in float64, then the lean sqrt / exp /
//             polynomial of mb_math.cuh and 128-byte row segments straight to global memory.
#include "mb_common.cuh"
#include "mb_math.cuh"

namespace {

constexpr int DB = 8, NS = 7;
constexpr long long DHALF = 1LL << (DB - 1), DMASK = (1LL << DB) - 1;
constexpr int KS = 32;
constexpr int TM = 128, TN = 64;
constexpr int XSLICE = 2 * TM * 16, XKSTEP = NS * XSLICE;   // 4 KB per slice, 28 KB per k-step of a cell panel
constexpr int YSLICE = 2 * TN * 16, YKSTEP = NS * YSLICE;   // 2 KB per slice, 14 KB per k-step of a landmark panel
constexpr int YCONST = 2 * TN * 8;                          // 64 norms + 64 scales behind a landmark panel
constexpr int NYB = 3;                                      // landmark stages in flight
constexpr int NEPI = 16;                                    // epilogue warps
constexpr int NTHREADS = (4 + NEPI) * 32;                   // control warpgroup + epilogue warps
constexpr int EC = TN / (NEPI / 4);                         // 16 columns per epilogue thread
constexpr int NISS = 2;                                     // MMA-issuing warps (one thread issues an MMA per ~80 clk)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ const double g_tab_i8[64] = MB_EXP2_TABLE_INIT;

// ---- pack: one block per panel of P rows; thread = (row, k-step) -------------------------------------------------
// out: digits of the panel, and (norm, scale) per row: scale = 2^(E - 54) (x) or the same (y; the kernel multiplies)
template <int P>
__global__ void __launch_bounds__(2 * P)
pack_kernel(const double* __restrict__ x, int64_t n, int64_t ldx, const int* __restrict__ dims, int d, int ksteps,
            double c, int8_t* __restrict__ digits, int64_t panel_bytes, double* __restrict__ norm,
            double* __restrict__ scale, int consts_in_panel, int* __restrict__ status) {
  __shared__ double red[2][P];
  const int row = threadIdx.x % P, ks = threadIdx.x / P;      // ks in {0, 1}
  const int64_t i = blockIdx.x * (int64_t)P + row;
  double v[KS];
  double m = 0.0;
#pragma unroll
  for (int k = 0; k < KS; k++) {
    const int kk = ks * KS + k;
    double t = 0.0;
    if (i < n && ks < ksteps && kk < d) t = c * x[i * ldx + (dims ? dims[kk] : kk)];
    v[k] = t;
    m = fmax(m, fabs(t));
  }
  red[ks][row] = m;
  __syncthreads();
  m = fmax(red[0][row], red[1][row]);
  __syncthreads();
  int E = 0;
  if (m > 0.0) frexp(m, &E);                                  // |v| < 2^E
  if (!(m < 1.7e308)) atomicExch(status, 2);
  int8_t* blk = digits + blockIdx.x * panel_bytes + (int64_t)ks * (NS * 2 * P * 16);
  double nrm = 0.0;
  uint32_t dig[NS][8];
#pragma unroll
  for (int t = 0; t < NS; t++)
#pragma unroll
    for (int q = 0; q < 8; q++) dig[t][q] = 0u;
#pragma unroll
  for (int k = 0; k < KS; k++) {
    long long q = llrint(ldexp(v[k], 54 - E));
    const double vq = ldexp((double)q, E - 54);
    nrm = fma(vq, vq, nrm);
#pragma unroll
    for (int t = NS - 1; t >= 0; t--) {
      const long long dd = ((q + DHALF) & DMASK) - DHALF;
      q = (q - dd) >> DB;
      dig[t][k >> 2] |= (uint32_t)(uint8_t)(int8_t)dd << (8 * (k & 3));
    }
  }
  if (ks < ksteps) {
#pragma unroll
    for (int t = 0; t < NS; t++) {
      *reinterpret_cast<uint4*>(blk + t * (2 * P * 16) + 0 * (P * 16) + row * 16) = make_uint4(dig[t][0], dig[t][1], dig[t][2], dig[t][3]);
      *reinterpret_cast<uint4*>(blk + t * (2 * P * 16) + 1 * (P * 16) + row * 16) = make_uint4(dig[t][4], dig[t][5], dig[t][6], dig[t][7]);
    }
  }
  red[ks][row] = nrm;
  __syncthreads();
  if (ks == 0) {
    const double tot = red[0][row] + red[1][row];
    const double sc = (i < n) ? ldexp(1.0, E - 54) : 0.0;
    if (consts_in_panel) {
      double* cst = reinterpret_cast<double*>(digits + blockIdx.x * panel_bytes + (int64_t)ksteps * (NS * 2 * P * 16));
      cst[row] = (i < n) ? tot : 0.0;
      cst[P + row] = sc;
    } else if (i < n) {
      norm[i] = tot;
      scale[i] = sc;
    }
  }
}

// ---- tcgen05 / mbarrier helpers --------------------------------------------------------------------------------
__device__ __forceinline__ void umma_i8_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi),
      "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
constexpr uint32_t DESC_HI = ((128u >> 4) & 0x3FFFu) | (1u << 14);   // SBO = 128 B, descriptor version 1
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* status) {
  uint32_t ok = 0;
  long long spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    if (!ok) {
      if (++spins > 20000000LL) { atomicExch(status, 1); return false; }
      if ((spins & 1023) == 0 && *(volatile int*)status == 1) return false;
    }
  }
  return true;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}

// (double)v for an int32 v with ONE FP64-pipe instruction: the bits 0x43300000'(v ^ 0x80000000) are the double
// 2^52 + 2^31 + v, so a subtraction finishes it (I2F.F64 runs at a fraction of the DADD rate)
__device__ __forceinline__ double i2d(int32_t v) {
  return __hiloint2double(0x43300000, v ^ (int32_t)0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
}

template <int KIND>
__device__ __forceinline__ double eval_scaled(double sq, const double* tab) {
  sq = mbmath::clamp_tiny(sq);
  if (KIND == MB_K_EXPQUAD) return mbmath::exp_neg(sq, tab);
  const double r = mbmath::sqrt_pos(sq);
  const double e = mbmath::exp_neg(r, tab);
  if (KIND == MB_K_EXPONENTIAL) return e;
  if (KIND == MB_K_MATERN32) return fma(r, e, e);
  return fma(fma(r, 1.0 / 3.0, 1.0), r, 1.0) * e;  // MATERN52
}

struct CovI8Args {
  const int8_t* xd;       // cell panels: ksteps * XKSTEP bytes each
  const int8_t* yd;       // landmark panels: ksteps * YKSTEP + YCONST bytes each
  const double* xnorm;
  const double* xscale;
  int64_t n, m;
  int ksteps;
  double eps_scaled;      // 1e-12 c^2 (util.py:365: + 1e-12 inside the square root)
  double* out;
  int64_t ldo;
  int vec;                // out is 16-byte aligned and ldo even: 16-byte stores
  int64_t self_offset;    // NN mode: column r + self_offset is the cell itself and is skipped
  int64_t* nn_idx;        // NN mode: index of the nearest column per row (out receives the squared distance)
  int* status;
};
enum { I8_STORE = 0, I8_NNMIN = 1 };

template <int KIND, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
cov_i8_kernel(const CovI8Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t x_full, x_empty, y_full[NYB], y_empty[NYB], t_full, t_empty;
  __shared__ uint32_t tmem_base_sh;
  __shared__ double tab[64];
  __shared__ double nn_val[MODE == I8_NNMIN ? NEPI / 4 : 1][MODE == I8_NNMIN ? TM : 1];
  __shared__ int64_t nn_col[MODE == I8_NNMIN ? NEPI / 4 : 1][MODE == I8_NNMIN ? TM : 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int xbytes = a.ksteps * XKSTEP, ybytes = a.ksteps * YKSTEP + YCONST;
  unsigned char* xs = smem;
  unsigned char* ys = smem + xbytes;
  const int64_t n_panels = (a.n + TM - 1) / TM, n_tiles = (a.m + TN - 1) / TN;

  if (tid < 64) tab[tid] = g_tab_i8[tid];
  if (tid == 0) {
    mbar_init(&x_full, 1);
    mbar_init(&x_empty, NISS);
    for (int s = 0; s < NYB; s++) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], NISS + NEPI); }
    mbar_init(&t_full, NISS);
    mbar_init(&t_empty, NEPI);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s_u32(&tmem_base_sh)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_sh;

  if (warp == 0) {
    // ---- producer ----
    if (lane == 0) {
      int64_t it = 0;        // landmark stages issued so far (across panels)
      int64_t pcount = 0;    // panels issued so far
      for (int64_t panel = blockIdx.x; panel < n_panels; panel += gridDim.x, pcount++) {
        if (pcount > 0 && !mbar_wait(&x_empty, (uint32_t)((pcount - 1) & 1), a.status)) break;
        mbar_expect_tx(&x_full, (uint32_t)xbytes);
        bulk_g2s(xs, a.xd + panel * (int64_t)xbytes, (uint32_t)xbytes, &x_full);
        bool ok = true;
        for (int64_t jt = 0; jt < n_tiles && ok; jt++, it++) {
          const int s = (int)(it % NYB);
          if (it >= NYB) ok = mbar_wait(&y_empty[s], (uint32_t)(((it / NYB) - 1) & 1), a.status);
          if (!ok) break;
          mbar_expect_tx(&y_full[s], (uint32_t)ybytes);
          bulk_g2s(ys + (size_t)s * ybytes, a.yd + jt * (int64_t)ybytes, (uint32_t)ybytes, &y_full[s]);
        }
        if (!ok) break;
      }
    }
  } else if (warp <= NISS) {
    // ---- MMA issuers (warps 1 .. NISS): whole warp runs the loops (uniform operands), one elected lane issues its
    // share of the tile's MMAs (every NISS-th one) ----
    const int iw = warp - 1;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    const uint32_t xs0 = s_u32(xs), ys0 = s_u32(ys);
    int64_t it = 0, pcount = 0;
    bool ok = true;
    for (int64_t panel = blockIdx.x; panel < n_panels && ok; panel += gridDim.x, pcount++) {
      if (lane == 0) ok = mbar_wait(&x_full, (uint32_t)(pcount & 1), a.status);
      ok = __shfl_sync(0xffffffffu, ok, 0);
      for (int64_t jt = 0; jt < n_tiles && ok; jt++, it++) {
        const int s = (int)(it % NYB);
        if (lane == 0) ok = mbar_wait(&y_full[s], (uint32_t)((it / NYB) & 1), a.status);
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t ysb = ys0 + (uint32_t)s * (uint32_t)ybytes;
        // the epilogue has drained the previous tile's accumulators
        if (it > 0) {
          if (lane == 0) ok = mbar_wait(&t_empty, (uint32_t)((it - 1) & 1), a.status);
          ok = __shfl_sync(0xffffffffu, ok, 0);
        }
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        if (ok && elect_one()) {
          // groups {6,3,2,0} belong to issuer 0, {5,4,1} to issuer 1: an accumulator is written by ONE thread (MMAs of
          // different threads are not ordered, and a group's first MMA overwrites)
          for (int ks = 0; ks < a.ksteps; ks++) {
#pragma unroll
            for (int t = 0; t < NS; t++) {
              const uint32_t alo = desc_lo(xs0 + (uint32_t)((ks * NS + t) * XSLICE), TM * 16);
#pragma unroll
              for (int g = NS - 1; g >= t; g--) {
                const int owner = (g == 6 || g == 3 || g == 2 || g == 0) ? 0 : 1;
                if (owner == iw) {
                  const uint32_t blo = desc_lo(ysb + (uint32_t)((ks * NS + (g - t)) * YSLICE), TN * 16);
                  umma_i8_lh(tmem_base + (uint32_t)(g * TN), alo, DESC_HI, blo, DESC_HI, idesc, (ks == 0 && t == 0) ? 0u : 1u);
                }
              }
            }
          }
          umma_commit(&t_full);
        }
        __syncwarp();
        if (ok && elect_one()) {
          umma_commit(&y_empty[s]);                       // the tile's MMAs have read the stage
          if (jt + 1 == n_tiles) umma_commit(&x_empty);   // ... and the cell panel
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue ----
    const int ew = warp - 4, quad = warp & 3, cb = ew >> 2;        // TMEM lane quadrant = warp % 4; 16 columns
    const int row = quad * 32 + lane, c0 = cb * EC;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
    int64_t it = 0;
    bool ok = true;
    for (int64_t panel = blockIdx.x; panel < n_panels && ok; panel += gridDim.x) {
      const int64_t gi = panel * TM + row;
      const double xn = (gi < a.n) ? a.xnorm[gi] + a.eps_scaled : 0.0;
      // 2 x.y = H 256^6 sx sy 2: fold the constant factors into the row scale
      const double sx = (gi < a.n) ? -2.0 * a.xscale[gi] * (double)(1LL << (DB * (NS - 1))) : 0.0;
      double best = __longlong_as_double(0x7ff0000000000000LL);   // NN mode: running minimum over this thread's columns
      int64_t best_col = -1;
      for (int64_t jt = 0; jt < n_tiles && ok; jt++, it++) {
        const int s = (int)(it % NYB);
        const uint32_t par = (uint32_t)(it & 1);
        double H[EC];
        // drain the 7 group accumulators of this thread's 16 columns: fold (0,1), (2,3), (4,5) in int32, then 6
        ok = ok && mbar_wait(&t_full, par, a.status);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int p = 0; p < 4; p++) {
          int32_t v0[EC], v1[EC];
          tmem_ld16(lane_addr + (uint32_t)(2 * p * TN), v0);
          if (p < 3) tmem_ld16(lane_addr + (uint32_t)((2 * p + 1) * TN), v1);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
          for (int e = 0; e < EC; e++) {
            if (p == 0) H[e] = i2d(v0[e] * 256 + v1[e]);
            else if (p < 3) H[e] = fma(H[e], 65536.0, i2d(v0[e] * 256 + v1[e]));
            else H[e] = fma(H[e], 256.0, i2d(v0[e]));
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty);               // the issuer may start the next tile
        // column constants of this tile (shared memory, broadcast reads), then the covariance
        const double* cst = reinterpret_cast<const double*>(ys + (size_t)s * ybytes + (size_t)a.ksteps * YKSTEP);
        ok = ok && mbar_wait(&y_full[s], (uint32_t)((it / NYB) & 1), a.status);   // visibility of the constants
        if (MODE == I8_NNMIN) {
          // exact nearest neighbour: running minimum of xn + yn - 2 x.y over every column but the cell itself; columns
          // come in increasing order, so `<` keeps the lowest index among equal distances
#pragma unroll
          for (int e = 0; e < EC; e++) {
            const int64_t gj = jt * TN + c0 + e;
            const double sq = fma(H[e] * sx, cst[TN + c0 + e], xn + cst[c0 + e]);
            if (gj < a.m && gj != gi + a.self_offset && sq < best) { best = sq; best_col = gj; }
          }
        } else {
        // 8 columns at a time: evaluate and store (a thread owns 16 consecutive columns of ONE row: 128 bytes)
          double* orow = a.out + gi * a.ldo + jt * TN + c0;
#pragma unroll
          for (int h = 0; h < EC; h += 8) {
            double val[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
              const double yn = cst[c0 + h + e], sy = cst[TN + c0 + h + e];
              const double sq = fma(H[h + e] * sx, sy, xn + yn);            // xn + yn - 2 x.y (+ eps)
              val[e] = eval_scaled<KIND>(sq, tab);
            }
            if (gi < a.n) {
#pragma unroll
              for (int e = 0; e < 8; e += 2) {
                const int64_t gj = jt * TN + c0 + h + e;
                if (gj + 1 < a.m && a.vec) {
                  *reinterpret_cast<double2*>(orow + h + e) = make_double2(val[e], val[e + 1]);
                } else {
                  if (gj < a.m) orow[h + e] = val[e];
                  if (gj + 1 < a.m) orow[h + e + 1] = val[e + 1];
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&y_empty[s]);          // done with the stage's constants
      }
      if (MODE == I8_NNMIN) {
        // the four warps of a lane quadrant hold the minima over their own 16 columns of every tile: combine them in
        // column-block order, lowest index first among equal distances (named barrier over the 16 epilogue warps)
        nn_val[cb][row] = best;
        nn_col[cb][row] = best_col;
        asm volatile("bar.sync 1, %0;\n" ::"n"(NEPI * 32) : "memory");
        if (cb == 0 && gi < a.n) {
          double v = nn_val[0][row];
          int64_t c = nn_col[0][row];
#pragma unroll
          for (int q = 1; q < NEPI / 4; q++) {
            const double v2 = nn_val[q][row];
            const int64_t c2 = nn_col[q][row];
            if (v2 < v || (v2 == v && c2 >= 0 && (c < 0 || c2 < c))) { v = v2; c = c2; }
          }
          a.out[gi * a.ldo] = v;
          a.nn_idx[gi] = c;
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(NEPI * 32) : "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}

template <int KIND, int MODE>
int launch_cov_i8(mb_ctx* ctx, const CovI8Args& a, size_t smem) {
  static mb_per_device_flag configured;
  if (!configured(ctx)) {
    MB_CUDA((cudaFuncSetAttribute(cov_i8_kernel<KIND, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)));
    configured(ctx) = true;
  }
  const int64_t n_panels = ceil_div64(a.n, TM);
  const int grid = (int)std::min<int64_t>(n_panels, ctx->n_sm);
  if (MODE == I8_NNMIN) {
    MB_LAUNCH(ctx, (cov_i8_kernel<KIND, MODE>), grid, NTHREADS, smem, a);
  } else {
    MB_LAUNCH_P(ctx, MB_PROF_COV, (cov_i8_kernel<KIND, MODE>), grid, NTHREADS, smem, a);
  }
  return 0;
}

}  // namespace

// K(i, j) = k_kind(c |x_i - y_j|) for one exponential-family leaf over the columns `dims` (NULL: the first d columns).
// *done = false when the shape is outside this kernel (d > 64, tiny problems): the caller falls back to the DMMA kernel.
// kind == MB_K_DISTANCE: exact nearest-neighbour search instead (c = 1): out(i) = min_j |x_i - y_j|^2 over j != i +
// self_offset, nn_idx(i) its index (device pointer).
int mb_cov_i8_build(mb_ctx* ctx, int kind, double c, const mb_mat* x, const mb_mat* y, const int* dims_host, int d,
                    double* out, int64_t ldo, bool* done, int64_t self_offset, int64_t* nn_idx) {
  *done = false;
  if (ctx->opt_cov_i8 == 0 || d < 1 || d > 2 * KS) return 0;
  if (!(kind == MB_K_MATERN32 || kind == MB_K_MATERN52 || kind == MB_K_EXPQUAD || kind == MB_K_EXPONENTIAL ||
        (kind == MB_K_DISTANCE && nn_idx)))
    return 0;
  const int64_t n = x->rows, m = y->rows;
  if (ctx->opt_cov_i8 != 2 && (n < 4096 || m < 256)) return 0;
  if (n == 0 || m == 0) return 0;
  MB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->i8_status) {
    MB_CUDA(cudaMalloc(&ctx->i8_status, sizeof(int)));
    MB_CUDA(cudaMemset(ctx->i8_status, 0, sizeof(int)));
  }
  const int ksteps = (d + KS - 1) / KS;
  const int64_t xp = ceil_div64(n, TM), yp = ceil_div64(m, TN);
  const int64_t xbytes = (int64_t)ksteps * XKSTEP, ybytes = (int64_t)ksteps * YKSTEP + YCONST;
  auto al = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  const size_t off_y = al((size_t)xp * xbytes), off_xn = off_y + al((size_t)yp * ybytes),
               off_xs = off_xn + al((size_t)n * 8), off_dm = off_xs + al((size_t)n * 8), total = off_dm + al((size_t)d * 4);
  double* wsd;
  MB_TRY(mb_scratch(ctx, total, &wsd));
  unsigned char* ws = reinterpret_cast<unsigned char*>(wsd);
  int8_t* xd = reinterpret_cast<int8_t*>(ws);
  int8_t* yd = reinterpret_cast<int8_t*>(ws + off_y);
  double* xnorm = reinterpret_cast<double*>(ws + off_xn);
  double* xscale = reinterpret_cast<double*>(ws + off_xs);
  int* dims_dev = nullptr;
  if (dims_host) {
    dims_dev = reinterpret_cast<int*>(ws + off_dm);
    MB_CUDA(cudaMemcpyAsync(dims_dev, dims_host, sizeof(int) * d, cudaMemcpyHostToDevice, ctx->stream));
  }
  MB_LAUNCH(ctx, (pack_kernel<TM>), (unsigned)xp, 2 * TM, 0, x->p, n, x->cols, dims_dev, d, ksteps, c, xd, xbytes, xnorm, xscale, 0,
            ctx->i8_status);
  MB_LAUNCH(ctx, (pack_kernel<TN>), (unsigned)yp, 2 * TN, 0, y->p, m, y->cols, dims_dev, d, ksteps, c, yd, ybytes, nullptr, nullptr, 1,
            ctx->i8_status);
  CovI8Args a;
  a.xd = xd;
  a.yd = yd;
  a.xnorm = xnorm;
  a.xscale = xscale;
  a.n = n;
  a.m = m;
  a.ksteps = ksteps;
  a.eps_scaled = (kind == MB_K_DISTANCE) ? 0.0 : 1e-12 * c * c;
  a.self_offset = self_offset;
  a.nn_idx = nn_idx;
  a.out = out;
  a.ldo = ldo;
  a.vec = ((ldo & 1) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
  a.status = ctx->i8_status;
  const size_t smem = (size_t)xbytes + (size_t)NYB * ybytes;
  switch (kind) {
    case MB_K_MATERN32: MB_TRY((launch_cov_i8<MB_K_MATERN32, I8_STORE>(ctx, a, smem))); break;
    case MB_K_MATERN52: MB_TRY((launch_cov_i8<MB_K_MATERN52, I8_STORE>(ctx, a, smem))); break;
    case MB_K_EXPQUAD: MB_TRY((launch_cov_i8<MB_K_EXPQUAD, I8_STORE>(ctx, a, smem))); break;
    case MB_K_DISTANCE: MB_TRY((launch_cov_i8<MB_K_DISTANCE, I8_NNMIN>(ctx, a, smem))); break;
    default: MB_TRY((launch_cov_i8<MB_K_EXPONENTIAL, I8_STORE>(ctx, a, smem))); break;
  }
  *done = true;
  return 0;
}
