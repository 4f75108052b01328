// tcgen05 `kind::i8` (int8 x int8 -> int32 in TMEM) on B200: does it run, is the operand/descriptor layout what we
// think it is, and how fast is it at the narrow N the Ozaki-slice covariance kernel would use?  
//
//   1. correctness: one 128 x N x 32 MMA on random int8 operands in the no-swizzle K-major canonical layout
//      ((8 rows x 16 B) core matrices, SBO between 8-row groups, LBO between the two 16-byte K chunks), read back
//      with tcgen05.ld and compared with the host;
//   2. throughput: a single thread issues back-to-back accumulating MMAs on every SM, tcgen05.commit -> mbarrier.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_umma_i8 microbench_umma_i8.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address, 16-byte units
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // leading-dimension byte offset
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride-dimension byte offset
  d |= (uint64_t)1 << 46;                            // descriptor version 1 (Blackwell)
  return d;                                          // layout_type 0 = SWIZZLE_NONE, base_offset 0
}
__host__ __device__ inline uint32_t make_idesc_i8(int M, int N) {
  uint32_t d = 0;
  d |= 2u << 4;                  // c_format = S32
  d |= 1u << 7;                  // a_format = signed 8 bit
  d |= 1u << 10;                 // b_format = signed 8 bit
  d |= (uint32_t)(N >> 3) << 17; // n_dim
  d |= (uint32_t)(M >> 4) << 24; // m_dim
  return d;                      // K-major A and B, dense, no saturate
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, long long max_cycles) {
  const long long t0 = clock64();
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    if (!ok && clock64() - t0 > max_cycles) return false;
  }
  return true;
}

// mode 0: correctness (out = D, 128 x N int32); mode 1: throughput (cycles[blockIdx.x] = elapsed clocks for `iters` MMAs)
__global__ void __launch_bounds__(128, 1)
umma_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int N, int mode, int iters, int32_t* __restrict__ out,
            long long* __restrict__ cycles, int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                 // [2 chunks][128 rows][16 B]
  unsigned char* sB = smem + 4096;          // [2 chunks][N rows][16 B]
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bars4[4];
  __shared__ uint32_t tmem_base_sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // operands: global row-major [rows][32] int8 -> canonical K-major no-swizzle layout
  for (int e = tid; e < 128 * 32; e += 128) { const int r = e >> 5, k = e & 31; sA[(k >> 4) * (128 * 16) + r * 16 + (k & 15)] = A[e]; }
  for (int e = tid; e < N * 32; e += 128)   { const int r = e >> 5, k = e & 31; sB[(k >> 4) * (N * 16) + r * 16 + (k & 15)] = B[e]; }
  if (tid == 0) {
    for (int q = 0; q < 4; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s_u32(&bars4[q])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s_u32(&tmem_base_sh)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  // generic-proxy writes of the operands must be visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_sh;

  const uint64_t da = make_desc(s_u32(sA), 128 * 16, 128);
  const uint64_t db = make_desc(s_u32(sB), (uint32_t)N * 16, 128);
  const uint32_t idesc = make_idesc_i8(128, N);

  long long t0 = 0, t1 = 0;
  bool ok = true;
  if (tid == 0 && mode < 2) {
    t0 = clock64();
    if (mode == 0) {
      umma_i8(tmem_base, da, db, idesc, 0);
    } else {
      // rotate over 512 / N accumulator column groups like the real kernel would
      const int groups = 512 / N;
      for (int it = 0; it < iters; it++) umma_i8(tmem_base + (uint32_t)((it % groups) * N), da, db, idesc, it >= groups);
    }
    umma_commit(&bar);
    ok = mbar_wait_bounded(&bar, 0, 4000000000LL);
    t1 = clock64();
    if (!ok) atomicExch(status, 1);
    if (mode == 1) cycles[blockIdx.x] = t1 - t0;
  }
  if (mode >= 2) {
    // mode = 2 + log2(issuers): `issuers` warps each issue iters / issuers MMAs into their own TMEM column range
    const int issuers = 1 << (mode - 2);
    __syncthreads();
    if (lane == 0 && warp < issuers) {
      const int per = 512 / issuers, groups = per / N > 0 ? per / N : 1;
      const long long s0 = clock64();
      for (int it = 0; it < iters / issuers; it++)
        umma_i8(tmem_base + (uint32_t)(warp * per + (it % groups) * N), da, db, idesc, it >= groups);
      umma_commit(&bars4[warp]);
      const bool done = mbar_wait_bounded(&bars4[warp], 0, 4000000000LL);
      if (!done) atomicExch(status, 1);
      if (warp == 0) cycles[blockIdx.x] = clock64() - s0;
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (mode == 0 && blockIdx.x == 0) {
    // warp w reads TMEM lanes 32 w .. 32 w + 31 (= rows of D), 8 columns at a time
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      const int row = warp * 32 + lane;
      for (int j = 0; j < 8; j++) out[row * N + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sm = p.multiProcessorCount;
  printf("device %s, %d SMs, cc %d.%d, clock %.0f MHz\n", p.name, sm, p.major, p.minor, p.clockRate * 1e-3);
  int8_t *A, *B; int32_t* out; long long* cyc; int* status;
  CK(cudaMalloc(&A, 128 * 32)); CK(cudaMalloc(&B, 256 * 32)); CK(cudaMalloc(&out, 128 * 256 * 4));
  CK(cudaMalloc(&cyc, sizeof(long long) * sm)); CK(cudaMalloc(&status, 4));
  int8_t hA[128 * 32], hB[256 * 32];
  srand(1);
  for (int i = 0; i < 128 * 32; i++) hA[i] = (int8_t)(rand() % 128 - 64);
  for (int i = 0; i < 256 * 32; i++) hB[i] = (int8_t)(rand() % 128 - 64);
  CK(cudaMemcpy(A, hA, sizeof(hA), cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, hB, sizeof(hB), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  static int32_t hout[128 * 256];
  for (int N : {32, 64, 128, 256}) {
    CK(cudaMemset(status, 0, 4)); CK(cudaMemset(out, 0xff, 128 * 256 * 4));
    umma_kernel<<<1, 128, 16384>>>(A, B, N, 0, 1, out, cyc, status);
    CK(cudaDeviceSynchronize());
    int st; CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hout, out, 128 * N * 4, cudaMemcpyDeviceToHost));
    long bad = 0; int first = -1;
    for (int i = 0; i < 128; i++) for (int j = 0; j < N; j++) {
      int ref = 0; for (int k = 0; k < 32; k++) ref += (int)hA[i * 32 + k] * (int)hB[j * 32 + k];
      if (hout[i * N + j] != ref) { if (first < 0) first = i * N + j; bad++; }
    }
    printf("correctness M=128 N=%3d K=32 int8: %s (%ld of %d wrong%s)", N, st ? "TIMEOUT" : (bad ? "MISMATCH" : "exact"), bad, 128 * N,
           st ? ", mbarrier never completed" : "");
    if (bad && first >= 0) {
      int i = first / N, j = first % N, ref = 0; for (int k = 0; k < 32; k++) ref += (int)hA[i * 32 + k] * (int)hB[j * 32 + k];
      printf("  first at (%d,%d): got %d want %d", i, j, hout[first], ref);
    }
    printf("\n");
  }
  const int iters = 4096;
  static long long hc[1024];
  for (int N : {32, 64, 128, 256}) {
    for (int grid : {1, sm}) {
      CK(cudaMemset(status, 0, 4));
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      umma_kernel<<<grid, 128, 16384>>>(A, B, N, 1, iters, out, cyc, status);  // warm-up
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      umma_kernel<<<grid, 128, 16384>>>(A, B, N, 1, iters, out, cyc, status);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      int st; CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hc, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
      double avg = 0; for (int i = 0; i < grid; i++) avg += (double)hc[i]; avg /= grid;
      const double macs_per_mma = 128.0 * N * 32.0;
      printf("throughput N=%3d on %3d SM(s): %s %.1f clk per MMA, %.0f int8 MAC/clk/SM, kernel %.3f ms => %.1f TOPS (2 ops per MAC) chip-wide\n",
             N, grid, st ? "TIMEOUT" : "", avg / iters, macs_per_mma * iters / avg, ms, 2.0 * macs_per_mma * iters * grid / (ms * 1e-3) * 1e-12);
    }
  }
  for (int N : {32, 64, 128}) {
    for (int lg : {0, 1, 2}) {
      CK(cudaMemset(status, 0, 4));
      umma_kernel<<<sm, 128, 16384>>>(A, B, N, 2 + lg, iters, out, cyc, status);
      CK(cudaDeviceSynchronize());
      int st; CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hc, cyc, sizeof(long long) * sm, cudaMemcpyDeviceToHost));
      double avg = 0; for (int i = 0; i < sm; i++) avg += (double)hc[i]; avg /= sm;
      printf("multi-issuer N=%3d, %d issuing warps: %s %.1f clk per MMA (SM aggregate), %.0f int8 MAC/clk/SM\n", N, 1 << lg,
             st ? "TIMEOUT" : "", avg / iters, 128.0 * N * 32.0 * iters / avg);
    }
  }
  return 0;
}
