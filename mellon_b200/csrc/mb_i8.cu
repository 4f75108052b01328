// Float64 products on the 5th-generation tensor cores: tcgen05.mma kind::i8 over digit slices (Ozaki scheme),
// accumulators in TMEM, operands staged by bulk TMA copies.  FP64 has no tcgen05 kind, and the legacy DMMA path
// (mb_gemm.cu) tops out at 37 TF; the int8 pipe delivers the same products, to the same accuracy, 2.4x faster.
//
// Gram matrix G = A^T A over a block of cells (K4: parameters.py:895-896 via sklearn Ridge; conditional.py:61).
// Contraction over the cells, so every COLUMN j of A gets one power-of-two scale 2^E_j (|A_ij| < 2^E_j over the
// group of cells) and each entry becomes a 54-bit fixed-point integer q = rint(A_ij 2^(54 - E_j)), written as
// NS = 7 balanced 8-bit digits d_0 (most significant) .. d_6 in [-128, 127].  A float64 product is the sum of the int8
// products d_t d_u with t + u <= 6 (28 pairs; the dropped ones are below 2^-56 of the scale product, i.e. below
// float64 rounding of the result).  Products with equal g = t + u share one int32 accumulator, so a 128 x 64 output
// tile keeps 7 accumulators of 64 TMEM columns (448 of the 512), resident over a whole group of <= 16384 cells
// (|G_g| <= 7 * 2^14 * 16384 < 2^31: the integer sums are EXACT, whatever their order).  The flush folds
// H = sum_g G_g 256^(6-g) in float64 (Horner) and scales by 256^6 2^(E_i + E_j - 108).
//
// Data flow per group of cells:
//   colmax_kernel     column maxima -> exponents
//   pack_cols_kernel  float64 row-major -> digits, TRANSPOSED to K-major, in two tile-contiguous layouts
//                     A: [128-col panel][k-step of 32 cells][slice 7][k16 chunk 2][128 cols][16 B]   28 KB per (panel, k-step)
//                     B: [ 64-col panel][k-step           ][slice 7][k16 chunk 2][ 64 cols][16 B]   14 KB per (panel, k-step)
//                     so one 1-D bulk copy (cp.async.bulk) fills an operand stage
//   gram_i8_kernel    one CTA per (group, lower 128 x 64 tile): warp 0 = producer (4-stage ring of 42 KB), warps 1..NI
//                     = MMA issuers: the whole warp runs the loop so that the MMA operands are warp-uniform, one
//                     elected lane issues; every accumulator group belongs to ONE issuer; then warps 1-4 flush
//                     (tcgen05.ld lane quadrant = warp % 4).
// Measured (profiles/bench_kernels_r02.txt, profiles/ncu_*_i8_r02.csv): the N = 64 MMA is bound by the tensor core's
// shared-memory operand reads (48 clk per 128 x 64 x 32 MMA = 2/3 of the int8 peak), one thread issues an MMA per
// ~80 clk, so the issuers share the 28 MMAs of a k-step.
#include "mb_common.cuh"

namespace {

constexpr int DB = 8;                          // digit width
constexpr int NS = 7;                          // digit slices per value: 54-bit fixed point
constexpr int KS = 32;                         // cells per k-step (one kind::i8 MMA: K = 32)
constexpr int TA = 128, TB = 64;               // output tile: 128 rows (A panel) x 64 columns (B panel)
constexpr int ASLICE = 2 * TA * 16, ABLOCK = NS * ASLICE;   // 4 KB per slice, 28 KB per (A panel, k-step)
constexpr int BSLICE = 2 * TB * 16, BBLOCK = NS * BSLICE;   // 2 KB per slice, 14 KB per (B panel, k-step)
constexpr int NST = 4;                         // operand stages in flight (a fifth one changes nothing: profiles/bench_kernels_r02.txt)
constexpr int NT = 5 * 32;                     // producer warp + issuer warp (also flushes) + 3 more flush warps
constexpr int SMEM_TOTAL = NST * (ABLOCK + BBLOCK);         // 168 KB
constexpr int KC = 16384;                      // cells per group: |G_g| <= 7 * 2^28 < 2^31

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- float64 -> digit slices ------------------------------------------------------------------------------------
// q = rint(v 2^(54 - E)), |q| <= 2^54.  Adding 128 to every balanced digit turns them into the plain base-256 digits of
// the non-negative number q + sum_t 128 256^t, so the digits are the BYTES of (q + DBIAS) ^ DBIAS (the xor takes the
// 128 off again, in two's complement): byte b is slice NS - 1 - b.  One add and one xor instead of a 7-step carry loop.
constexpr unsigned long long DBIAS = 0x0080808080808080ull;

__device__ __noinline__ long long fixed54_tiny(double v, int sh) { return llrint(ldexp(v, sh)); }

// q[c] = rint(x[c] 2^(54 - E)) for 16 values with one exponent.  The scale 2^(54 - E) is a double (built from its
// exponent bits; the product is exact) unless the whole row / column is below 2^-960: ldexp then, same integers.
__device__ __forceinline__ void fixed54x16(const double (&x)[16], int E, long long (&q)[16]) {
  if (E >= -960) {
    const double s = __longlong_as_double((long long)(1023 + 54 - E) << 52);
#pragma unroll
    for (int c = 0; c < 16; c++) q[c] = __double2ll_rn(x[c] * s);
  } else {
#pragma unroll
    for (int c = 0; c < 16; c++) q[c] = fixed54_tiny(x[c], 54 - E);
  }
}

// digits of four values -> one 32-bit word per slice, value c in byte c (the operand layout's 4 consecutive k)
__device__ __forceinline__ void digits_of4(const long long* q, uint32_t (&w)[NS]) {
  uint32_t lo[4], hi[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const unsigned long long v = ((unsigned long long)q[c] + DBIAS) ^ DBIAS;
    lo[c] = (uint32_t)v;
    hi[c] = (uint32_t)(v >> 32);
  }
  uint32_t a = __byte_perm(lo[0], lo[1], 0x5140), b = __byte_perm(lo[2], lo[3], 0x5140);   // bytes 0, 1 of the pair
  w[6] = __byte_perm(a, b, 0x5410);
  w[5] = __byte_perm(a, b, 0x7632);
  a = __byte_perm(lo[0], lo[1], 0x7362), b = __byte_perm(lo[2], lo[3], 0x7362);            // bytes 2, 3
  w[4] = __byte_perm(a, b, 0x5410);
  w[3] = __byte_perm(a, b, 0x7632);
  a = __byte_perm(hi[0], hi[1], 0x5140), b = __byte_perm(hi[2], hi[3], 0x5140);            // bytes 4, 5
  w[2] = __byte_perm(a, b, 0x5410);
  w[1] = __byte_perm(a, b, 0x7632);
  a = __byte_perm(hi[0], hi[1], 0x7362), b = __byte_perm(hi[2], hi[3], 0x7362);            // byte 6 (byte 7 is zero)
  w[0] = __byte_perm(a, b, 0x5410);
}

// ---- column scales: max |A_ij| over the group, as the bit pattern of a non-negative double (ordered like uint64) ----
__global__ void colmax_kernel(const double* __restrict__ A, int64_t rows, int64_t r, int64_t ld,
                              unsigned long long* __restrict__ cmax) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= r) return;
  const int64_t per = (rows + gridDim.y - 1) / gridDim.y, i0 = blockIdx.y * per, i1 = min(rows, i0 + per);
  double m = 0.0;
  for (int64_t i = i0; i < i1; i++) m = fmax(m, fabs(A[i * ld + j]));
  atomicMax(cmax + j, (unsigned long long)__double_as_longlong(m));
}

// ---- pack: one block per (128-column panel, k-step); thread = (column, 16-cell chunk) -----------------------------
// reads 32 rows x 1 KB (coalesced), writes 16-byte pieces that are consecutive across the threads of a warp
__global__ void __launch_bounds__(256)
pack_cols_kernel(const double* __restrict__ A, int64_t rows, int64_t r, int64_t ld,
                 const unsigned long long* __restrict__ cmax, int64_t nks, int8_t* __restrict__ Ad,
                 int8_t* __restrict__ Bd, double* __restrict__ scale, int* __restrict__ status) {
  const int64_t pa = blockIdx.x, ks = blockIdx.y;
  const int col = threadIdx.x & 127, chunk = threadIdx.x >> 7;
  const int64_t j = pa * TA + col;
  int E = 0;
  if (j < r) {
    const double m = __longlong_as_double((long long)cmax[j]);
    if (m > 0.0) frexp(m, &E);                               // m = f 2^E, 0.5 <= f < 1  =>  |v| < 2^E
    if (!(m < 1.7e308)) atomicExch(status, 2);               // inf / nan in the operand: no fixed-point image
    if (ks == 0 && chunk == 0) scale[j] = ldexp(1.0, E - 54);
  } else if (ks == 0 && chunk == 0) {
    scale[j] = 0.0;
  }
  double x[16];
#pragma unroll
  for (int c = 0; c < 16; c++) {
    const int64_t i = ks * KS + chunk * 16 + c;
    x[c] = (j < r && i < rows) ? A[i * ld + j] : 0.0;
  }
  long long q[16];
  fixed54x16(x, E, q);
  uint32_t dig[NS][4];
#pragma unroll
  for (int g = 0; g < 4; g++) {
    uint32_t w[NS];
    digits_of4(q + 4 * g, w);
#pragma unroll
    for (int t = 0; t < NS; t++) dig[t][g] = w[t];
  }
  int8_t* ab = Ad + (pa * nks + ks) * (int64_t)ABLOCK;
  int8_t* bb = Bd + ((2 * pa + (col >> 6)) * nks + ks) * (int64_t)BBLOCK;
#pragma unroll
  for (int t = 0; t < NS; t++) {
    const uint4 v = make_uint4(dig[t][0], dig[t][1], dig[t][2], dig[t][3]);
    *reinterpret_cast<uint4*>(ab + t * ASLICE + chunk * (TA * 16) + col * 16) = v;
    *reinterpret_cast<uint4*>(bb + t * BSLICE + chunk * (TB * 16) + (col & 63) * 16) = v;
  }
}

// ---- tcgen05 helpers -------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle, K-major: 8-row groups 128 B apart (SBO), the two 16-byte K chunks of
// a K = 32 int8 operand LBO apart
// tcgen05.mma kind::i8 with the shared-memory matrix descriptors given as (low, high) 32-bit halves: the issue loop only
// adds the slice offset to the low word (start address >> 4) instead of rebuilding 64-bit descriptors.  The scalar code around an MMA is what limits
// one issuing thread (~144 clk per MMA with make_desc in the loop, an order of magnitude less this way).
__device__ __forceinline__ void umma_i8_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi),
      "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TENSOR MEMORY (128 lanes x 8 columns of 32 bit = one 128 x 32 int8 slice), B from shared memory: the
// MMA then reads 2 KB of shared memory instead of 6 KB and becomes MAC-bound (32 clk) instead of operand-read-bound (48)
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
// shared memory (matrix descriptor, same canonical layout as an MMA operand) -> tensor memory, 128 lanes x 256 bit;
// executes in issue order with the MMAs of the same thread
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t tmem_dst, uint32_t s_lo, uint32_t s_hi) {
  asm volatile("{\n\t.reg .b64 ds;\n\tmov.b64 ds, {%1, %2};\n\ttcgen05.cp.cta_group::1.128x256b [%0], ds;\n\t}\n" ::"r"(tmem_dst),
               "r"(s_lo), "r"(s_hi)
               : "memory");
}

__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
constexpr uint32_t DESC_HI = ((128u >> 4) & 0x3FFFu) | (1u << 14);   // SBO = 128 B, descriptor version 1 (bit 46)

// true on exactly one lane of a converged warp; unlike `lane == 0` the compiler KNOWS a single thread follows, so the
// MMA operands go to uniform registers directly instead of through a per-instruction waterfall loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void umma_commit(uint64_t* bar);
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* status, unsigned backoff_ns = 0);

// The 28 MMAs of one k-step, acc_g += A_t B_(g-t)^T for t <= g <= 6, shared among NI issuing warps BY GROUP: every
// accumulator g belongs to one issuer (MMAs of different threads are not ordered, and the first MMA of a group must
// come before the others: it overwrites).  One thread issues an MMA every ~80 clk (uniform-datapath set-up of the five
// operands), the tensor pipe takes 48 clk per 128 x 64 x 32 MMA, so several issuers are needed to keep it busy.
//   NI = 4: {6} {5,0} {4,1} {3,2} (7 MMAs each);  NI = 2: {6,3,2,0} (15) {5,4,1} (13);  NI = 1: everything
__host__ __device__ constexpr int group_owner(int ni, int g) {
  return ni == 1 ? 0 : ni == 2 ? ((g == 6 || g == 3 || g == 2 || g == 0) ? 0 : 1)
                               : (g == 6 ? 0 : (g == 5 || g == 0) ? 1 : (g == 4 || g == 1) ? 2 : 3);
}
template <int NI, int W>
__device__ __forceinline__ void issue_kstep(uint32_t a_lo, uint32_t b_lo, uint32_t tmem_base, uint32_t idesc, uint32_t fresh) {
#pragma unroll
  for (int t = 0; t < NS; t++) {
#pragma unroll
    for (int g = NS - 1; g >= t; g--)
      if (group_owner(NI, g) == W)
        umma_i8_lh(tmem_base + (uint32_t)(g * TB), a_lo + (uint32_t)(t * (ASLICE >> 4)), DESC_HI,
                   b_lo + (uint32_t)((g - t) * (BSLICE >> 4)), DESC_HI, idesc, (t == 0) ? fresh : 1u);
  }
}
// NI = 0: ONE issuer with the A slices staged in tensor memory (columns 448 .. 503, behind the 7 accumulators): 7 copies
// shared -> tensor memory, then the 28 MMAs with A from there.  Copies and MMAs of one thread execute in issue order, so
// the next k-step's copies cannot overtake the MMAs that still read the slices.
constexpr uint32_t A_TMEM_COL = NS * TB;       // 448
__device__ __forceinline__ void issue_kstep_atmem(uint32_t a_lo, uint32_t b_lo, uint32_t tmem_base, uint32_t idesc, uint32_t fresh) {
#pragma unroll
  for (int t = 0; t < NS; t++) tmem_cp_128x256b(tmem_base + A_TMEM_COL + (uint32_t)(8 * t), a_lo + (uint32_t)(t * (ASLICE >> 4)), DESC_HI);
#pragma unroll
  for (int t = 0; t < NS; t++) {
#pragma unroll
    for (int g = NS - 1; g >= t; g--)
      umma_i8_ts(tmem_base + (uint32_t)(g * TB), tmem_base + A_TMEM_COL + (uint32_t)(8 * t),
                 b_lo + (uint32_t)((g - t) * (BSLICE >> 4)), DESC_HI, idesc, (t == 0) ? fresh : 1u);
  }
}

template <int NI>
__device__ __forceinline__ void issue_kstep_w(int w, uint32_t a_lo, uint32_t b_lo, uint32_t tmem_base, uint32_t idesc,
                                              uint32_t fresh) {
  if (NI == 0) { issue_kstep_atmem(a_lo, b_lo, tmem_base, idesc, fresh); return; }
  if (w == 0) issue_kstep<NI == 0 ? 1 : NI, 0>(a_lo, b_lo, tmem_base, idesc, fresh);
  else if (NI > 1 && w == 1) issue_kstep<NI == 0 ? 1 : NI, 1 % (NI == 0 ? 1 : NI)>(a_lo, b_lo, tmem_base, idesc, fresh);
  else if (NI > 2 && w == 2) issue_kstep<NI == 0 ? 1 : NI, 2 % (NI == 0 ? 1 : NI)>(a_lo, b_lo, tmem_base, idesc, fresh);
  else if (NI > 3) issue_kstep<NI == 0 ? 1 : NI, 3 % (NI == 0 ? 1 : NI)>(a_lo, b_lo, tmem_base, idesc, fresh);
}

// the issuing loop of warp `w` (0-based among the NI issuers): the WHOLE warp runs it so that the MMA operands are
// warp-uniform (uniform registers, no per-instruction waterfall); one lane polls the barrier, one elected lane issues
template <int NI>
__device__ __forceinline__ void issuer_loop(int w, int lane, int64_t nks, unsigned char* smem, uint64_t* full, uint64_t* empty,
                                            uint64_t* done, uint32_t tmem_base, int* status) {
  bool ok = true;
  // instruction descriptor: D = s32, A = B = signed 8 bit, K-major both, N = 64, M = 128
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TB >> 3) << 17) | ((uint32_t)(TA >> 4) << 24);
  const uint32_t smem0 = s_u32(smem);
  for (int64_t ks = 0; ks < nks && ok; ks++) {
    const int s = (int)(ks % NST);
    if (lane == 0) ok = mbar_wait(&full[s], (uint32_t)((ks / NST) & 1), status);
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) break;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t sa = smem0 + (uint32_t)s * (ABLOCK + BBLOCK), sb = sa + ABLOCK;
    if (elect_one()) {
      issue_kstep_w<NI>(w, desc_lo(sa, TA * 16), desc_lo(sb, TB * 16), tmem_base, idesc, (ks == 0) ? 0u : 1u);
      umma_commit(&empty[s]);     // arrives once every MMA of this issuer that reads stage s has completed
    }
    __syncwarp();
  }
  if (elect_one()) umma_commit(done);
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
// bounded wait: a protocol bug shows up as an error code instead of a hung GPU
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* status, unsigned backoff_ns) {
  uint32_t ok = 0;
  long long spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    if (!ok) {
      if (backoff_ns) __nanosleep(backoff_ns);
      if (++spins > 20000000LL) { atomicExch(status, 1); return false; }
      if ((spins & 1023) == 0 && *(volatile int*)status == 1) return false;   // some wait timed out already: drain quickly
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

struct GramArgs {
  const int8_t* Ad;       // [group][A panel][k-step] blocks
  const int8_t* Bd;       // [group][B panel][k-step] blocks
  const double* scale;    // [group][rpad]
  const int2* tiles;      // (pa, pb) of the lower 128 x 64 tiles, pa-major
  int ntiles;
  int64_t rows, grows;    // cells in the leaf, cells per group (a multiple of KS)
  int64_t nks_max;        // k-steps of a full group (stride of the digit arrays)
  int64_t a_gstride, b_gstride, rpad;
  int64_t r;
  double* out;            // [group] r x r dense (ld = r); group g stores its own partial (no accumulation)
  int64_t out_gstride;
  int* status;
};

// out_g[pa*128 .. , pb*64 ..] = (A_g^T A_g) tile for the K-major digit operands of group g
template <int NI>
__global__ void __launch_bounds__(NT, 1)
gram_i8_kernel(const GramArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[NST], empty[NST], done;
  __shared__ uint32_t tmem_base_sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = blockIdx.x / a.ntiles, tile = blockIdx.x % a.ntiles;
  const int pa = a.tiles[tile].x, pb = a.tiles[tile].y;
  const int64_t g_rows = min(a.grows, a.rows - (int64_t)grp * a.grows);
  const int64_t nks = (g_rows + KS - 1) / KS;

  if (tid == 0) {
    for (int s = 0; s < NST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NI == 0 ? 1 : NI); }
    mbar_init(&done, NI == 0 ? 1 : NI);
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
    // ---- producer: two bulk copies per k-step into a 4-deep ring ----
    if (lane == 0) {
      const int8_t* asrc = a.Ad + grp * a.a_gstride + (int64_t)pa * a.nks_max * ABLOCK;
      const int8_t* bsrc = a.Bd + grp * a.b_gstride + (int64_t)pb * a.nks_max * BBLOCK;
      for (int64_t ks = 0; ks < nks; ks++) {
        const int s = (int)(ks % NST);
        if (ks >= NST && !mbar_wait(&empty[s], (uint32_t)(((ks / NST) - 1) & 1), a.status)) break;
        unsigned char* stage = smem + s * (ABLOCK + BBLOCK);
        mbar_expect_tx(&full[s], ABLOCK + BBLOCK);
        bulk_g2s(stage, asrc + ks * ABLOCK, ABLOCK, &full[s]);
        bulk_g2s(stage + ABLOCK, bsrc + ks * BBLOCK, BBLOCK, &full[s]);
      }
    }
  } else {
    // ---- issuers: warps 1 .. NI (see issuer_loop); warps 1-4 flush ----
    if (warp <= (NI == 0 ? 1 : NI)) issuer_loop<NI>(warp - 1, lane, nks, smem, full, empty, &done, tmem_base, a.status);
    __syncwarp();
    // ---- flush: warps 1-4, TMEM lane quadrant = warp % 4 (row of the tile), 16 columns at a time ----
    if (mbar_wait(&done, 0, a.status, 256)) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const int quad = warp & 3, row = quad * 32 + lane;
      const int64_t gi = (int64_t)pa * TA + row;
      const double* sc = a.scale + grp * a.rpad;
      double* out = a.out + grp * a.out_gstride;
      // Horner gives H = sum_g G_g 256^(6-g); the product is H 256^6 2^(E_i + E_j - 108): 256^6 goes into the row scale
      const double si = (gi < a.r) ? sc[gi] * (double)(1LL << (DB * (NS - 1))) : 0.0;
      for (int c0 = 0; c0 < TB; c0 += 16) {
        double h[16];
#pragma unroll
        for (int g = 0; g < NS; g++) {
          int32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * TB + c0), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
          for (int e = 0; e < 16; e++) h[e] = (g == 0) ? (double)v[e] : fma(h[e], (double)(1 << DB), (double)v[e]);
        }
        if (gi < a.r) {
          if (nks == 0) {
#pragma unroll
            for (int e = 0; e < 16; e++) h[e] = 0.0;
          }
#pragma unroll
          for (int e = 0; e < 16; e++) {
            const int64_t gj = (int64_t)pb * TB + c0 + e;
            if (gj < a.r) out[gi * a.r + gj] = h[e] * si * sc[gj];
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}

// out = p_0 + p_1 + ... (fixed order) on the lower 128 x 64 tiles' rows: rows >= cols / tile granularity is not needed,
// the strict upper triangle is mirrored afterwards, so the whole matrix is summed
__global__ void sum_groups_kernel(const double* __restrict__ parts, int ng, int64_t stride, int64_t n, double* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t st = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) {
    double s = parts[i];
    for (int g = 1; g < ng; g++) s += parts[(int64_t)g * stride + i];
    out[i] = s;
  }
}

// ==== general product C = beta C + alpha A B^T (A: n x k, B: p x k, both row-major, k contiguous) =====================
// TRSM updates X2 -= X1 L21^T (decomposition.py:209) and the Nystroem factor Q V (decomposition.py:265).  The
// contraction index is the column, so every ROW of A and of B gets its own power-of-two scale (from the row's own
// maximum: an output row depends on nothing but its own row of A and on B), digits stay row-major (no transpose):
//   A: [128-row panel][k-step of 32 columns][slice 7][k16 chunk 2][128 rows][16 B]
//   B: [ 64-row panel][k-step              ][slice 7][k16 chunk 2][ 64 rows][16 B]

// one warp per row: exponent E with |a_ij| < 2^E over the k columns, scale[i] = 2^(E - 54)
__global__ void __launch_bounds__(256)
rowmax_kernel(const double* __restrict__ A, int64_t rows, int64_t k, int64_t ld, int* __restrict__ expo,
              double* __restrict__ scale, int64_t rows_pad, int* __restrict__ status) {
  const int64_t i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= rows_pad) return;
  double m = 0.0;
  if (i < rows)
    for (int64_t c = lane; c < k; c += 32) m = fmax(m, fabs(A[i * ld + c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) {
    int E = 0;
    if (m > 0.0) frexp(m, &E);
    if (!(m < 1.7e308)) atomicExch(status, 2);
    expo[i] = E;
    scale[i] = (i < rows) ? ldexp(1.0, E - 54) : 0.0;
  }
}

// block = (panel of P rows, k-step of 32 columns); thread = (row, 16-column chunk)
template <int P>
__global__ void __launch_bounds__(2 * P)
pack_rows_kernel(const double* __restrict__ X, int64_t rows, int64_t k, int64_t ld, int64_t nks,
                 const int* __restrict__ expo, int8_t* __restrict__ Xd) {
  const int64_t panel = blockIdx.x, ks = blockIdx.y;
  const int row = threadIdx.x % P, chunk = threadIdx.x / P;
  const int64_t i = panel * P + row;
  const int E = (i < rows) ? expo[i] : 0;
  const int64_t k0 = ks * KS + chunk * 16;
  const double* src = X + i * ld + k0;
  double x[16];
  if (i < rows && k0 + 16 <= k && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // 16-byte loads: the 128 B of a thread are one cache line, fetched with 8 requests instead of 16
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(src) + c);
      x[2 * c] = v.x;
      x[2 * c + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < 16; c++) x[c] = (i < rows && k0 + c < k) ? src[c] : 0.0;
  }
  long long q[16];
  fixed54x16(x, E, q);
  uint32_t dig[NS][4];
#pragma unroll
  for (int g = 0; g < 4; g++) {
    uint32_t w[NS];
    digits_of4(q + 4 * g, w);
#pragma unroll
    for (int t = 0; t < NS; t++) dig[t][g] = w[t];
  }
  int8_t* blk = Xd + (panel * nks + ks) * (int64_t)(NS * 2 * P * 16);
#pragma unroll
  for (int t = 0; t < NS; t++)
    *reinterpret_cast<uint4*>(blk + t * (2 * P * 16) + chunk * (P * 16) + row * 16) = make_uint4(dig[t][0], dig[t][1], dig[t][2], dig[t][3]);
}

struct NtArgs {
  const int8_t* Ad;     // [A panel][k-step] blocks of ABLOCK
  const int8_t* Bd;     // [B panel][k-step] blocks of BBLOCK
  const double* sa;     // row scales of A (n padded to 128)
  const double* sb;     // row scales of B (p padded to 64)
  int64_t nks;          // k-steps of the contraction
  int npb;              // B panels (output column tiles)
  int64_t n, p;
  double alpha;
  int beta;             // 0: C = alpha A B^T ; 1: C += alpha A B^T
  double* C;
  int64_t ldc;
  int* status;
};

// one CTA per 128 x 64 output tile; B panel index fastest, so the CTAs in flight share an A panel through L2
template <int NI>
__global__ void __launch_bounds__(NT, 1)
gemm_nt_i8_kernel(const NtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[NST], empty[NST], done;
  __shared__ uint32_t tmem_base_sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t pa = blockIdx.x / a.npb;
  const int pb = (int)(blockIdx.x % a.npb);
  const int64_t nks = a.nks;

  if (tid == 0) {
    for (int s = 0; s < NST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NI == 0 ? 1 : NI); }
    mbar_init(&done, NI == 0 ? 1 : NI);
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
    if (lane == 0) {
      const int8_t* asrc = a.Ad + pa * nks * ABLOCK;
      const int8_t* bsrc = a.Bd + (int64_t)pb * nks * BBLOCK;
      for (int64_t ks = 0; ks < nks; ks++) {
        const int s = (int)(ks % NST);
        if (ks >= NST && !mbar_wait(&empty[s], (uint32_t)(((ks / NST) - 1) & 1), a.status)) break;
        unsigned char* stage = smem + s * (ABLOCK + BBLOCK);
        mbar_expect_tx(&full[s], ABLOCK + BBLOCK);
        bulk_g2s(stage, asrc + ks * ABLOCK, ABLOCK, &full[s]);
        bulk_g2s(stage + ABLOCK, bsrc + ks * BBLOCK, BBLOCK, &full[s]);
      }
    }
  } else {
    if (warp <= (NI == 0 ? 1 : NI)) issuer_loop<NI>(warp - 1, lane, nks, smem, full, empty, &done, tmem_base, a.status);
    __syncwarp();
    if (mbar_wait(&done, 0, a.status, 256)) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const int quad = warp & 3, row = quad * 32 + lane;
      const int64_t gi = pa * TA + row;
      const double si = (gi < a.n) ? a.alpha * a.sa[gi] * (double)(1LL << (DB * (NS - 1))) : 0.0;
      double* crow = a.C + gi * a.ldc + (int64_t)pb * TB;
      for (int c0 = 0; c0 < TB; c0 += 16) {
        double h[16];
#pragma unroll
        for (int g = 0; g < NS; g++) {
          int32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * TB + c0), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
          for (int e = 0; e < 16; e++) h[e] = (g == 0) ? (double)v[e] : fma(h[e], (double)(1 << DB), (double)v[e]);
        }
        if (gi < a.n) {
#pragma unroll
          for (int e = 0; e < 16; e++) {
            const int64_t gj = (int64_t)pb * TB + c0 + e;
            if (gj < a.p) {
              const double v = (nks > 0) ? h[e] * si * a.sb[gj] : 0.0;
              crow[c0 + e] = a.beta ? crow[c0 + e] + v : v;
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}

int ensure_ws(mb_ctx* ctx, size_t bytes) {
  if (bytes > ctx->i8_ws_bytes) {
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->i8_ws) MB_CUDA(cudaFree(ctx->i8_ws));
    ctx->i8_ws = nullptr;
    ctx->i8_ws_bytes = 0;
    MB_CUDA(mb_dev_malloc(ctx, (void**)&ctx->i8_ws, bytes));
    ctx->i8_ws_bytes = bytes;
  }
  return 0;
}


// side stream + events, created on first use; false while the main stream is being captured into a graph
int side_setup(mb_ctx* ctx, bool* usable) {
  *usable = false;
  if (!ctx->opt_i8_overlap) return 0;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(ctx->stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return 0;
  }
  if (!ctx->i8_side) {
    int lo = 0, hi = 0;
    MB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t sd = nullptr;
    MB_CUDA(cudaStreamCreateWithPriority(&sd, cudaStreamNonBlocking, hi));   // the packs feed the next GEMM: run them first
    MB_CUDA(cudaEventCreateWithFlags(&ctx->i8_ev_start, cudaEventDisableTiming));
    for (int b = 0; b < 2; b++) {
      MB_CUDA(cudaEventCreateWithFlags(&ctx->i8_ev_packed[b], cudaEventDisableTiming));
      MB_CUDA(cudaEventCreateWithFlags(&ctx->i8_ev_free[b], cudaEventDisableTiming));
    }
    ctx->i8_side = sd;
  }
  *usable = true;
  return 0;
}

// everything enqueued on the main stream so far (the operands) is visible to the side stream
int side_fork(mb_ctx* ctx) {
  MB_CUDA(cudaEventRecord(ctx->i8_ev_start, ctx->stream));
  MB_CUDA(cudaStreamWaitEvent(ctx->i8_side, ctx->i8_ev_start, 0));
  return 0;
}

}  // namespace

bool mb_i8_gram_usable(mb_ctx* ctx, int64_t rows, int64_t r) {
  if (ctx->opt_i8 == 2) return r >= 1 && rows >= 1;   // forced (tests, sanitizer runs on small shapes)
  return ctx->opt_i8 != 0 && r >= 512 && rows >= 2048;
}

// Open a sequence of Gram leaves over a factor that is complete on the main stream NOW: with "i8_overlap" set, the
// leaves pack their digits on the side stream into alternating buffers until mb_i8_gram_end, so the (HBM-bound) pack of
// leaf c + 1 runs under the MMA kernel of leaf c.  Same kernels on the same data: the results do not change.
// Measured (profiles/bench_kernels_r02.txt (e)): 1-2 % on the Gram, nothing on the TRSM updates - the pack's loads and
// the MMA operands go through the same L1 / shared-memory port, which is what bounds the MMA kernels - so it is OFF by
// default and kept as a measured variant.
int mb_i8_gram_begin(mb_ctx* ctx) {
  bool usable = false;
  MB_TRY(side_setup(ctx, &usable));
  if (!usable) return 0;
  MB_TRY(side_fork(ctx));
  ctx->i8_seq = 0;
  ctx->i8_seq_one = 0;
  return 0;
}
void mb_i8_gram_end(mb_ctx* ctx) { ctx->i8_seq = -1; }

// out (r x r dense; the tiles on and below the diagonal are written) = A^T A for the `rows` x r block at A
int mb_i8_gram_leaf(mb_ctx* ctx, const double* A, int64_t lda, int64_t rows, int64_t r, double* out) {
  MB_CUDA(cudaSetDevice(ctx->device));
  const int64_t npa = ceil_div64(r, TA), npb = 2 * npa, rpad = npa * TA;
  const int ng = (int)ceil_div64(rows, KC);
  const int64_t grows = ceil_div64(ceil_div64(rows, ng), KS) * KS, nks_max = grows / KS;
  // tile list (pa-major: CTAs that run together share an A panel through L2), cached per r
  if (ctx->i8_tiles_r != r) {
    std::vector<int2> ht;
    for (int pa = 0; pa < npa; pa++)
      for (int pb = 0; pb < npb; pb++)
        if (64 * pb < 128 * (pa + 1) && (int64_t)pb * TB < r) ht.push_back(make_int2(pa, pb));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->i8_tiles) MB_CUDA(cudaFree(ctx->i8_tiles));
    MB_CUDA(cudaMalloc(&ctx->i8_tiles, ht.size() * sizeof(int2)));
    MB_CUDA(cudaMemcpy(ctx->i8_tiles, ht.data(), ht.size() * sizeof(int2), cudaMemcpyHostToDevice));
    ctx->i8_tiles_r = r;
    ctx->i8_ntiles = (int)ht.size();
  }
  if (!ctx->i8_status) {
    MB_CUDA(cudaMalloc(&ctx->i8_status, sizeof(int)));
    MB_CUDA(cudaMemset(ctx->i8_status, 0, sizeof(int)));
  }
  static mb_per_device_flag configured;
  if (!configured(ctx)) {
    MB_CUDA(cudaFuncSetAttribute(gram_i8_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    MB_CUDA(cudaFuncSetAttribute(gram_i8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    MB_CUDA(cudaFuncSetAttribute(gram_i8_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    MB_CUDA(cudaFuncSetAttribute(gram_i8_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured(ctx) = true;
  }
  // workspace: one operand buffer {digits (A and B layouts), scales, column maxima} (two inside a sequence with the
  // packs on the side stream) + per-group partial products
  const bool ahead = ctx->i8_seq >= 0;                      // inside mb_i8_gram_begin .. end: packs on the side stream
  const int buf = ahead ? (ctx->i8_seq & 1) : 0;
  cudaStream_t ps = ahead ? ctx->i8_side : ctx->stream;
  const size_t a_g = (size_t)npa * nks_max * ABLOCK, b_g = (size_t)npb * nks_max * BBLOCK;
  const size_t part = (ng > 1) ? (size_t)r * r * sizeof(double) : 0;
  auto al = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  const size_t off_b = al((size_t)ng * a_g), off_s = off_b + al((size_t)ng * b_g),
               off_m = off_s + al((size_t)ng * rpad * 8), need = off_m + al((size_t)ng * rpad * 8);
  // the two buffers keep ONE size over a sequence (the leaves of a tree differ in rows: the last chunk is shorter), or
  // the second buffer of a small leaf would start inside the first buffer of the large one before it
  size_t one = need;
  if (ahead) {
    if (need > ctx->i8_seq_one) {
      if (ctx->i8_seq > 0) MB_CUDA(cudaStreamSynchronize(ctx->stream));   // a larger leaf late in a sequence: drain first
      ctx->i8_seq_one = need;
    }
    one = ctx->i8_seq_one;
  }
  const size_t nbuf = ahead ? 2 : 1, total = nbuf * one + (size_t)ng * part;
  MB_TRY(ensure_ws(ctx, total));
  unsigned char* ws = reinterpret_cast<unsigned char*>(ctx->i8_ws) + (size_t)buf * one;
  int8_t* Ad = reinterpret_cast<int8_t*>(ws);
  int8_t* Bd = reinterpret_cast<int8_t*>(ws + off_b);
  double* scale = reinterpret_cast<double*>(ws + off_s);
  unsigned long long* cmax = reinterpret_cast<unsigned long long*>(ws + off_m);
  double* parts = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ctx->i8_ws) + nbuf * one);

  if (ahead && ctx->i8_seq >= 2) MB_CUDA(cudaStreamWaitEvent(ps, ctx->i8_ev_free[buf], 0));   // its last reader is done
  MB_CUDA(cudaMemsetAsync(cmax, 0, (size_t)ng * rpad * 8, ps));
  for (int g = 0; g < ng; g++) {
    const int64_t i0 = (int64_t)g * grows, grw = std::min(grows, rows - i0), nks = ceil_div64(grw, KS);
    MB_LAUNCH_ON(ctx, ps, colmax_kernel, dim3((unsigned)ceil_div64(r, 128), 64), 128, 0, A + i0 * lda, grw, r, lda,
                 cmax + (int64_t)g * rpad);
    MB_LAUNCH_ON(ctx, ps, pack_cols_kernel, dim3((unsigned)npa, (unsigned)nks), 256, 0, A + i0 * lda, grw, r, lda,
                 cmax + (int64_t)g * rpad, nks_max, Ad + (size_t)g * a_g, Bd + (size_t)g * b_g, scale + (int64_t)g * rpad,
                 ctx->i8_status);
  }
  if (ahead) {
    MB_CUDA(cudaEventRecord(ctx->i8_ev_packed[buf], ps));
    MB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->i8_ev_packed[buf], 0));
  }
  GramArgs ga;
  ga.Ad = Ad;
  ga.Bd = Bd;
  ga.scale = scale;
  ga.tiles = ctx->i8_tiles;
  ga.ntiles = ctx->i8_ntiles;
  ga.rows = rows;
  ga.grows = grows;
  ga.nks_max = nks_max;
  ga.a_gstride = (int64_t)a_g;
  ga.b_gstride = (int64_t)b_g;
  ga.rpad = rpad;
  ga.r = r;
  ga.out = (ng > 1) ? parts : out;
  ga.out_gstride = (int64_t)r * r;
  ga.status = ctx->i8_status;
  if (ctx->prof_on) ctx->prof_work[MB_PROF_I8] += (double)r * (double)r * (double)rows;  // SYRK: n r^2 flops
  const unsigned ggrid = (unsigned)(ng * ctx->i8_ntiles);
  if (ctx->opt_i8_issuers == 0) MB_LAUNCH_P(ctx, MB_PROF_I8, gram_i8_kernel<0>, ggrid, NT, SMEM_TOTAL, ga);
  else if (ctx->opt_i8_issuers == 1) MB_LAUNCH_P(ctx, MB_PROF_I8, gram_i8_kernel<1>, ggrid, NT, SMEM_TOTAL, ga);
  else if (ctx->opt_i8_issuers == 2) MB_LAUNCH_P(ctx, MB_PROF_I8, gram_i8_kernel<2>, ggrid, NT, SMEM_TOTAL, ga);
  else MB_LAUNCH_P(ctx, MB_PROF_I8, gram_i8_kernel<4>, ggrid, NT, SMEM_TOTAL, ga);
  if (ahead) {
    MB_CUDA(cudaEventRecord(ctx->i8_ev_free[buf], ctx->stream));
    ctx->i8_seq++;
  }
  if (ng > 1) {
    const int64_t n = (int64_t)r * r;
    MB_LAUNCH(ctx, sum_groups_kernel, (int)std::min<int64_t>(ceil_div64(n, 256), (int64_t)ctx->n_sm * 16), 256, 0, parts, ng,
              (int64_t)r * r, n, out);
  }
  return 0;
}


bool mb_i8_nt_usable(mb_ctx* ctx, int64_t rows_total, int64_t p, int64_t k) {
  if (ctx->opt_i8 == 2) return k >= 1 && k <= KC;
  return ctx->opt_i8 != 0 && rows_total >= 8192 && p >= 128 && k >= 256 && k <= KC;
}

// C (n x p, ldc) = beta C + alpha A B^T ; A: n x k (lda), B: p x k (ldb), row-major.  beta is 0 or 1.
// Rows of A are processed in slabs of 65536 so the digit workspace stays ~1 GB; a row of C depends on its own row of A
// and on B alone (per-row scales from the row's own maximum), never on the slab or on the number of rows.
int mb_i8_gemm_nt(mb_ctx* ctx, int64_t n, int64_t p, int64_t k, double alpha, const double* A, int64_t lda,
                  const double* B, int64_t ldb, int beta, double* C, int64_t ldc) {
  if (n <= 0 || p <= 0) return 0;
  MB_CHECK(k > 0 && k <= KC, "mb_i8_gemm_nt: contraction length %lld outside 1..%d", (long long)k, KC);
  MB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->i8_status) {
    MB_CUDA(cudaMalloc(&ctx->i8_status, sizeof(int)));
    MB_CUDA(cudaMemset(ctx->i8_status, 0, sizeof(int)));
  }
  static mb_per_device_flag configured;
  if (!configured(ctx)) {
    MB_CUDA(cudaFuncSetAttribute(gemm_nt_i8_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    MB_CUDA(cudaFuncSetAttribute(gemm_nt_i8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    MB_CUDA(cudaFuncSetAttribute(gemm_nt_i8_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    MB_CUDA(cudaFuncSetAttribute(gemm_nt_i8_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured(ctx) = true;
  }
  const int64_t SLAB = 65536;
  const int64_t nks = ceil_div64(k, KS), npb = ceil_div64(p, TB), ppad = npb * TB;
  const int64_t slab_rows = std::min(n, SLAB), slab_pad = ceil_div64(slab_rows, TA) * TA;
  // more than one slab: the pack of slab s + 1 runs on the side stream under the MMA kernel of slab s (two A buffers)
  bool ahead = false;
  if (n > SLAB) MB_TRY(side_setup(ctx, &ahead));
  const int nbuf = ahead ? 2 : 1;
  auto al = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  const size_t b_bytes = (size_t)npb * nks * BBLOCK, a_bytes = (size_t)(slab_pad / TA) * nks * ABLOCK;
  const size_t a_one = al(a_bytes) + al((size_t)slab_pad * 8) + al((size_t)slab_pad * 4);   // digits, scales, exponents
  const size_t off_a = al(b_bytes), off_sb = off_a + nbuf * a_one, off_eb = off_sb + al((size_t)ppad * 8),
               total = off_eb + al((size_t)ppad * 4);
  MB_TRY(ensure_ws(ctx, total));
  unsigned char* ws = reinterpret_cast<unsigned char*>(ctx->i8_ws);
  int8_t* Bd = reinterpret_cast<int8_t*>(ws);
  double* sb = reinterpret_cast<double*>(ws + off_sb);
  int* eb = reinterpret_cast<int*>(ws + off_eb);

  if (ahead) MB_TRY(side_fork(ctx));                        // A is complete on the main stream here
  cudaStream_t ps = ahead ? ctx->i8_side : ctx->stream;
  MB_LAUNCH(ctx, rowmax_kernel, (unsigned)ceil_div64(ppad, 8), 256, 0, B, p, k, ldb, eb, sb, ppad, ctx->i8_status);
  MB_LAUNCH(ctx, (pack_rows_kernel<TB>), dim3((unsigned)npb, (unsigned)nks), 2 * TB, 0, B, p, k, ldb, nks, eb, Bd);
  int slab = 0;
  for (int64_t r0 = 0; r0 < n; r0 += SLAB, slab++) {
    const int64_t rows = std::min(SLAB, n - r0), npa = ceil_div64(rows, TA), rpad = npa * TA;
    const int buf = ahead ? (slab & 1) : 0;
    unsigned char* wa = ws + off_a + (size_t)buf * a_one;
    int8_t* Ad = reinterpret_cast<int8_t*>(wa);
    double* sa = reinterpret_cast<double*>(wa + al(a_bytes));
    int* ea = reinterpret_cast<int*>(wa + al(a_bytes) + al((size_t)slab_pad * 8));
    if (ahead && slab >= 2) MB_CUDA(cudaStreamWaitEvent(ps, ctx->i8_ev_free[buf], 0));      // its last reader is done
    MB_LAUNCH_ON(ctx, ps, rowmax_kernel, (unsigned)ceil_div64(rpad, 8), 256, 0, A + r0 * lda, rows, k, lda, ea, sa, rpad,
                 ctx->i8_status);
    MB_LAUNCH_ON(ctx, ps, (pack_rows_kernel<TA>), dim3((unsigned)npa, (unsigned)nks), 2 * TA, 0, A + r0 * lda, rows, k, lda, nks,
                 ea, Ad);
    if (ahead) {
      MB_CUDA(cudaEventRecord(ctx->i8_ev_packed[buf], ps));
      MB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->i8_ev_packed[buf], 0));
    }
    NtArgs na;
    na.Ad = Ad;
    na.Bd = Bd;
    na.sa = sa;
    na.sb = sb;
    na.nks = nks;
    na.npb = (int)npb;
    na.n = rows;
    na.p = p;
    na.alpha = alpha;
    na.beta = beta;
    na.C = C + r0 * ldc;
    na.ldc = ldc;
    na.status = ctx->i8_status;
    if (ctx->prof_on) ctx->prof_work[MB_PROF_I8] += 2.0 * (double)rows * (double)p * (double)k;
    const unsigned ngrid = (unsigned)(npa * npb);
    if (ctx->opt_i8_issuers == 0) MB_LAUNCH_P(ctx, MB_PROF_I8, gemm_nt_i8_kernel<0>, ngrid, NT, SMEM_TOTAL, na);
    else if (ctx->opt_i8_issuers == 1) MB_LAUNCH_P(ctx, MB_PROF_I8, gemm_nt_i8_kernel<1>, ngrid, NT, SMEM_TOTAL, na);
    else if (ctx->opt_i8_issuers == 2) MB_LAUNCH_P(ctx, MB_PROF_I8, gemm_nt_i8_kernel<2>, ngrid, NT, SMEM_TOTAL, na);
    else MB_LAUNCH_P(ctx, MB_PROF_I8, gemm_nt_i8_kernel<4>, ngrid, NT, SMEM_TOTAL, na);
    if (ahead) MB_CUDA(cudaEventRecord(ctx->i8_ev_free[buf], ctx->stream));
  }
  return 0;
}

// the kernels report protocol time-outs / non-finite operands through a device flag: read it (one sync)
int mb_i8_check(mb_ctx* ctx) {
  if (!ctx->i8_status) return 0;
  int st = 0;
  MB_CUDA(cudaMemcpyAsync(&st, ctx->i8_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (st != 0) {
    cudaMemsetAsync(ctx->i8_status, 0, sizeof(int), ctx->stream);
    MB_CHECK(st != 1, "int8 digit-slice GEMM: an mbarrier wait timed out (protocol error)");
    MB_CHECK(false, "int8 digit-slice GEMM: non-finite value in the operand");
  }
  return 0;
}
