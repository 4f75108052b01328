// How fast can one pass over a 40 GB row-major matrix be read on B200, and with which access shape?
// Decides the layout of the K5 streaming kernel (mb_infer.cu).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_stream microbench_stream.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nW_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra W_DONE;\nbra W_WAIT;\nW_DONE:\n}\n" ::"r"(s_u32(bar)),
      "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}

// (1) grid-stride LDG.128
__global__ void read_gridstride(const double2* __restrict__ in, double* out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i < n; i += st) { double2 v = in[i]; s += v.x + v.y; }
  if (s == 123.456) out[0] = s;
}
// (2) per-CTA contiguous chunk, LDG.128, UNROLL loads in flight per thread
template <int UNROLL>
__global__ void read_chunked(const double2* __restrict__ in, double* out, size_t n) {
  size_t per = (n + gridDim.x - 1) / gridDim.x;
  size_t b = blockIdx.x * per, e = min(n, b + per);
  double s = 0;
  size_t i = b + threadIdx.x;
  for (; i + (size_t)(UNROLL - 1) * blockDim.x < e; i += (size_t)UNROLL * blockDim.x) {
    double2 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) v[u] = in[i + (size_t)u * blockDim.x];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) s += v[u].x + v[u].y;
  }
  for (; i < e; i += blockDim.x) { double2 v = in[i]; s += v.x + v.y; }
  if (s == 123.456) out[0] = s;
}
// (3) bulk-copy ring; `interleave`: slab k of CTA b is global slab (k * gridDim + b) instead of a contiguous chunk.
// consume: 0 = nobody touches the data, 1 = all threads read the slab once from shared memory (LDS.128)
__global__ void read_bulk_ring(const double* __restrict__ in, double* out, size_t n_slabs, int slab_bytes, int ns,
                               int inflight, int interleave, int consume) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[32];
  const int tid = threadIdx.x;
  size_t per = (n_slabs + gridDim.x - 1) / gridDim.x;
  size_t my = 0;
  if (interleave) my = (n_slabs > blockIdx.x) ? (n_slabs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  else { size_t b = blockIdx.x * per; my = b < n_slabs ? min(per, n_slabs - b) : 0; }
  auto slab_addr = [&](size_t k) -> const double* {
    size_t g = interleave ? (k * gridDim.x + blockIdx.x) : (blockIdx.x * per + k);
    return in + g * (size_t)(slab_bytes / 8);
  };
  if (tid == 0) {
    for (int s = 0; s < ns; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (tid == 0)
    for (size_t k = 0; k < my && k < (size_t)inflight; k++) {
      mbar_expect_tx(&full[k % ns], slab_bytes);
      bulk_g2s(smem + (k % ns) * (size_t)slab_bytes, slab_addr(k), slab_bytes, &full[k % ns]);
    }
  double s = 0;
  for (size_t k = 0; k < my; k++) {
    mbar_wait(&full[k % ns], (uint32_t)((k / ns) & 1));
    if (consume) {
      const double2* p = reinterpret_cast<const double2*>(smem + (k % ns) * (size_t)slab_bytes);
      for (int c = tid; c < slab_bytes / 16; c += blockDim.x) { double2 v = p[c]; s += v.x + v.y; }
    }
    __syncthreads();
    if (tid == 0 && k + inflight < my) {
      size_t kk = k + inflight;
      mbar_expect_tx(&full[kk % ns], slab_bytes);
      bulk_g2s(smem + (kk % ns) * (size_t)slab_bytes, slab_addr(kk), slab_bytes, &full[kk % ns]);
    }
  }
  if (s == 123.456) out[0] = s;
}

template <class F>
float timeit(F f, int reps = 3) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sm = p.multiProcessorCount;
  const size_t rows = 1000000, r = 5000;
  const size_t bytes = rows * r * 8;  // 40 GB
  double* a; CK(cudaMalloc(&a, bytes)); CK(cudaMemset(a, 0, bytes));
  double* out; CK(cudaMalloc(&out, 4096));
  const size_t n2 = bytes / 16;
  printf("device %s, %d SMs; matrix %zu x %zu f64 = %.1f GB\n", p.name, sm, rows, r, bytes * 1e-9);
  float ms;
  for (int mult : {4, 8, 16}) {
    ms = timeit([&] { read_gridstride<<<sm * mult, 512>>>((const double2*)a, out, n2); });
    printf("grid-stride LDG.128  %2d CTA/SM x512      : %7.3f ms  %7.1f GB/s\n", mult, ms, bytes / ms * 1e-6);
  }
  ms = timeit([&] { read_chunked<4><<<sm, 512>>>((const double2*)a, out, n2); });
  printf("chunked LDG.128 1 CTA/SM x512 unroll 4   : %7.3f ms  %7.1f GB/s\n", ms, bytes / ms * 1e-6);
  ms = timeit([&] { read_chunked<8><<<sm, 512>>>((const double2*)a, out, n2); });
  printf("chunked LDG.128 1 CTA/SM x512 unroll 8   : %7.3f ms  %7.1f GB/s\n", ms, bytes / ms * 1e-6);
  ms = timeit([&] { read_chunked<8><<<sm, 1024>>>((const double2*)a, out, n2); });
  printf("chunked LDG.128 1 CTA/SM x1024 unroll 8  : %7.3f ms  %7.1f GB/s\n", ms, bytes / ms * 1e-6);
  ms = timeit([&] { read_chunked<8><<<sm * 2, 512>>>((const double2*)a, out, n2); });
  printf("chunked LDG.128 2 CTA/SM x512 unroll 8   : %7.3f ms  %7.1f GB/s\n", ms, bytes / ms * 1e-6);
  ms = timeit([&] { read_chunked<8><<<sm * 4, 512>>>((const double2*)a, out, n2); });
  printf("chunked LDG.128 4 CTA/SM x512 unroll 8   : %7.3f ms  %7.1f GB/s\n", ms, bytes / ms * 1e-6);
  CK(cudaFuncSetAttribute(read_bulk_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  struct Cfg { int slab, ns, inflight, cta_per_sm, interleave, consume; };
  const Cfg cfgs[] = {
      {40000, 5, 3, 1, 0, 0}, {40000, 5, 3, 1, 0, 1}, {40000, 5, 5, 1, 0, 0}, {40000, 5, 3, 1, 1, 0},
      {40000, 5, 3, 1, 1, 1}, {20000, 10, 8, 1, 0, 0}, {20000, 10, 8, 1, 0, 1}, {8000, 25, 20, 1, 0, 0},
      {8000, 25, 20, 1, 0, 1}, {8000, 25, 20, 1, 1, 1}, {40000, 2, 2, 2, 0, 0}, {40000, 2, 2, 2, 0, 1},
      {20000, 5, 4, 2, 0, 1}, {20000, 5, 4, 2, 1, 1}, {8000, 12, 10, 2, 0, 1}, {4000, 25, 20, 2, 0, 1},
      {16000, 12, 10, 1, 0, 1}, {16000, 12, 10, 1, 1, 1},
  };
  for (const Cfg& c : cfgs) {
    const size_t n_slabs = bytes / c.slab;
    const size_t smem = (size_t)c.slab * c.ns;
    ms = timeit([&] {
      read_bulk_ring<<<sm * c.cta_per_sm, 512, smem>>>(a, out, n_slabs, c.slab, c.ns, c.inflight, c.interleave, c.consume);
    });
    printf("bulk ring slab %5d B x%2d stages, %2d in flight, %d CTA/SM, %s, %s: %7.3f ms  %7.1f GB/s\n", c.slab, c.ns,
           c.inflight, c.cta_per_sm, c.interleave ? "interleaved" : "contiguous ", c.consume ? "LDS-consumed" : "untouched   ",
           ms, bytes / ms * 1e-6);
  }
  return 0;
}
