// Microbenchmarks that decide the FP64 kernel design on B200 (sm_100a):
//   DFMA peak, DMMA (mma.sync f64) peak for m8n8k4 / m16n8k8 / m16n8k16, DFMA+DMMA mixed, HBM copy/write.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp64 microbench_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void dmma1688_kernel(double* out, int iters, double a, double b) {
  double c[ILP][4], av[4] = {a, a, b, b}, bv[2] = {b, a};
#pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma1688(c[i], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void dmma16816_kernel(double* out, int iters, double a, double b) {
  double c[ILP][4], av[8] = {a, a, b, b, a, b, a, b}, bv[4] = {b, a, a, b};
#pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma16816(c[i], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: NM dmma884 + NF dfma per iteration, independent chains
template <int NM, int NF>
__global__ void mixed_kernel(double* out, int iters, double a, double b) {
  double c[NM + 1][2], f[NF + 1];
#pragma unroll
  for (int i = 0; i < NM; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
  for (int i = 0; i < NF; i++) f[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NM; i++) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = fma(f[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NM; i++) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NF; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// GEMM-shaped inner loops: fragments come from shared memory with rotating registers, 32x32 warp tile.
//   SHAPE 0: m8n8k4   16 MMAs + 8 LDS.64 per k4 step
//   SHAPE 1: m16n8k8   8 MMAs + 16 LDS.64 per k8 step
//   SHAPE 2: m16n8k16  8 MMAs + 32 LDS.64 per k16 step
// SYNC_EVERY > 0 adds a __syncthreads every SYNC_EVERY k16-equivalents (the k-tile barrier of the GEMM).
template <int SHAPE, int SYNC_EVERY>
__global__ void __launch_bounds__(512, 1) gemm_loop_kernel(double* out, int iters) {
  __shared__ double sa[128 * 20], sb[128 * 20];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < 128 * 20; e += blockDim.x) { sa[e] = 1e-3 * e; sb[e] = 2e-3 * e; }
  __syncthreads();
  const int lr = lane >> 2, lk = lane & 3;
  const double* pa = sa + ((warp >> 2) * 32 + lr) * 20 + lk;
  const double* pb = sb + ((warp & 3) * 32 + lr) * 20 + lk;
  double acc[4][4][2];   // m8n8k4 view
  double acc4[2][4][4];  // m16n8kX view
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc4[i][j][0] = acc4[i][j][1] = acc4[i][j][2] = acc4[i][j][3] = 0.0;
  for (int it = 0; it < iters; it++) {
    if (SHAPE == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; i++) af[i] = pa[i * 160 + ks * 4];
#pragma unroll
        for (int j = 0; j < 4; j++) bf[j] = pb[j * 160 + ks * 4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    } else if (SHAPE == 1) {
#pragma unroll
      for (int ks = 0; ks < 2; ks++) {
        double af[2][4], bf[4][2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
          af[i][0] = pa[(i * 16) * 20 + ks * 8];      af[i][1] = pa[(i * 16 + 8) * 20 + ks * 8];
          af[i][2] = pa[(i * 16) * 20 + ks * 8 + 4];  af[i][3] = pa[(i * 16 + 8) * 20 + ks * 8 + 4];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) { bf[j][0] = pb[j * 160 + ks * 8]; bf[j][1] = pb[j * 160 + ks * 8 + 4]; }
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma1688(acc4[i][j], af[i], bf[j]);
      }
    } else {
      double af[2][8], bf[4][4];
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int q = 0; q < 4; q++) { af[i][2 * q] = pa[(i * 16) * 20 + q * 4]; af[i][2 * q + 1] = pa[(i * 16 + 8) * 20 + q * 4]; }
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) bf[j][q] = pb[j * 160 + q * 4];
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma16816(acc4[i][j], af[i], bf[j]);
    }
    if (SYNC_EVERY > 0 && (it % SYNC_EVERY) == SYNC_EVERY - 1) __syncthreads();
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) s += acc[i][j][0] + acc[i][j][1];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) s += acc4[i][j][0] + acc4[i][j][1] + acc4[i][j][2] + acc4[i][j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// exp / sqrt throughput
__global__ void exp_kernel(double* out, int iters, double a) {
  double x[4] = {a + threadIdx.x * 1e-3, a * 2, a * 3, a * 4}, s = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) { s += exp(-x[i]); x[i] += 1e-6; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void sqrt_kernel(double* out, int iters, double a) {
  double x[4] = {a + threadIdx.x * 1e-3, a * 2, a * 3, a * 4}, s = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) { s += sqrt(x[i]); x[i] += 1e-6; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#include "../mellon_b200/csrc/mb_math.cuh"
// lean Matern52 epilogue (mb_math.cuh): sq -> (1 + r + r^2/3) exp(-r), r = sqrt(sq)
__global__ void lean_matern_kernel(double* out, int iters, double a) {
  __shared__ double tab[64];
  { const double t[64] = MB_EXP2_TABLE_INIT; if (threadIdx.x < 64) tab[threadIdx.x] = t[threadIdx.x]; }
  __syncthreads();
  double x[4] = {a + threadIdx.x * 1e-3, a * 2, a * 3, a * 4}, s = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      double sq = mbmath::clamp_tiny(x[i]);
      double r = mbmath::sqrt_pos(sq);
      double e = mbmath::exp_neg(r, tab);
      s += fma(fma(r, 1.0 / 3.0, 1.0), r, 1.0) * e;
      x[i] += 1e-6;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void libm_matern_kernel(double* out, int iters, double a) {
  double x[4] = {a + threadIdx.x * 1e-3, a * 2, a * 3, a * 4}, s = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      double sq = fmax(x[i], 0.0);
      double r = sqrt(sq);
      double e = exp(-r);
      s += (r + r * r * (1.0 / 3.0) + 1.0) * e;
      x[i] += 1e-6;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// int32 -> f64: cvt.rn.f64.s32 vs the magic-number trick (integer ops + 1 DADD)
__global__ void i2d_cvt_kernel(double* out, int iters, int a) {
  int v[8]; double s[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { v[i] = a + threadIdx.x + i; s[i] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) { s[i] += (double)v[i]; v[i] += 3; }
  }
  double t = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) t += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
__global__ void i2d_magic_kernel(double* out, int iters, int a) {
  int v[8]; double s[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { v[i] = a + threadIdx.x + i; s[i] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      s[i] += __hiloint2double(0x43300000 ^ 0x00080000, v[i] ^ 0x80000000);  // 2^52+2^51 + (v + 2^31) - ... folded below
      v[i] += 3;
    }
  }
  double t = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) t += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
// accuracy of the hardware rsqrt seed and of sqrt_pos / exp_neg against libdevice
__global__ void accuracy_kernel(double* out) {
  __shared__ double tab[64];
  { const double t[64] = MB_EXP2_TABLE_INIT; if (threadIdx.x < 64) tab[threadIdx.x] = t[threadIdx.x]; }
  __syncthreads();
  double me = 0, ms = 0, mseed = 0;
  for (int i = 0; i < 20000; i++) {
    double u = (threadIdx.x * 20000.0 + i + 0.37) / (blockDim.x * 20000.0);
    double r = u * 60.0;
    double a = mbmath::exp_neg(r, tab), b = exp(-r);
    me = fmax(me, fabs(a - b) / b);
    double s = (1.0 + 3.0 * u) * exp2((double)((i % 200) - 100));
    double q = mbmath::sqrt_pos(s), q0 = sqrt(s);
    ms = fmax(ms, fabs(q - q0) / q0);
    double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    mseed = fmax(mseed, fabs(y * q0 - 1.0));
  }
  out[threadIdx.x * 3 + 0] = me; out[threadIdx.x * 3 + 1] = ms; out[threadIdx.x * 3 + 2] = mseed;
}

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) out[i] = in[i];
}
__global__ void write_kernel(double2* __restrict__ out, size_t n, double v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) out[i] = make_double2(v, v);
}
__global__ void read_kernel(const double2* __restrict__ in, double* out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i < n; i += st) { double2 v = in[i]; s += v.x + v.y; }
  if (s == 123.456) out[0] = s;
}

template <class F>
float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sm = p.multiProcessorCount;
  printf("device %s, %d SMs, cc %d.%d\n", p.name, sm, p.major, p.minor);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sm * 8 * 1024));
  const int iters = 20000;
  for (int warps : {4, 8, 16}) {
    int threads = warps * 32;
    dim3 g(sm), b(threads);
    float ms = timeit([&] { dfma_kernel<16><<<g, b>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 16 * iters * (double)threads * sm;
    printf("DFMA       warps/SM=%2d: %8.2f TFLOP/s\n", warps, fl / ms * 1e-9);
    ms = timeit([&] { dmma884_kernel<16><<<g, b>>>(out, iters, 1.0000001, 1e-9); });
    fl = 2.0 * 256 * 16 * iters * (double)warps * sm;
    printf("DMMA 884   warps/SM=%2d: %8.2f TFLOP/s\n", warps, fl / ms * 1e-9);
    ms = timeit([&] { dmma1688_kernel<8><<<g, b>>>(out, iters, 1.0000001, 1e-9); });
    fl = 2.0 * 1024 * 8 * iters * (double)warps * sm;
    printf("DMMA 1688  warps/SM=%2d: %8.2f TFLOP/s\n", warps, fl / ms * 1e-9);
    ms = timeit([&] { dmma16816_kernel<8><<<g, b>>>(out, iters, 1.0000001, 1e-9); });
    fl = 2.0 * 2048 * 8 * iters * (double)warps * sm;
    printf("DMMA 16816 warps/SM=%2d: %8.2f TFLOP/s\n", warps, fl / ms * 1e-9);
  }
  {
    dim3 g(sm), b(512);
    float ms = timeit([&] { mixed_kernel<8, 8><<<g, b>>>(out, iters, 1.0000001, 1e-9); });
    double fl_m = 2.0 * 256 * 8 * iters * 16.0 * sm, fl_f = 2.0 * 8 * iters * 512.0 * sm;
    printf("MIXED 8 dmma884 + 8 dfma / iter (16 warps): %.3f ms  -> DMMA %.2f TF + DFMA %.2f TF = %.2f TF\n", ms,
           fl_m / ms * 1e-9, fl_f / ms * 1e-9, (fl_m + fl_f) / ms * 1e-9);
    ms = timeit([&] { mixed_kernel<8, 32><<<g, b>>>(out, iters, 1.0000001, 1e-9); });
    fl_f = 2.0 * 32 * iters * 512.0 * sm;
    printf("MIXED 8 dmma884 + 32 dfma / iter (16 warps): %.3f ms -> DMMA %.2f TF + DFMA %.2f TF = %.2f TF\n", ms,
           fl_m / ms * 1e-9, fl_f / ms * 1e-9, (fl_m + fl_f) / ms * 1e-9);
    ms = timeit([&] { mixed_kernel<8, 0><<<g, b>>>(out, iters, 1.0000001, 1e-9); });
    printf("ONLY  8 dmma884 / iter (16 warps): %.3f ms -> %.2f TF\n", ms, fl_m / ms * 1e-9);
    ms = timeit([&] { mixed_kernel<0, 32><<<g, b>>>(out, iters, 1.0000001, 1e-9); });
    printf("ONLY  32 dfma / iter (16 warps): %.3f ms -> %.2f TF\n", ms, fl_f / ms * 1e-9);
  }

  {
    dim3 g(sm), b(512);
    const int it3 = 20000;
    const double fl = 2.0 * 32 * 32 * 16 * (double)it3 * 16 * sm;  // per iteration each warp does a 32x32x16 product
    float ms = timeit([&] { gemm_loop_kernel<0, 0><<<g, b>>>(out, it3); });
    printf("GEMM-loop m8n8k4   smem frags, no sync : %.2f TF\n", fl / ms * 1e-9);
    ms = timeit([&] { gemm_loop_kernel<0, 1><<<g, b>>>(out, it3); });
    printf("GEMM-loop m8n8k4   smem frags, sync/k16: %.2f TF\n", fl / ms * 1e-9);
    ms = timeit([&] { gemm_loop_kernel<0, 2><<<g, b>>>(out, it3); });
    printf("GEMM-loop m8n8k4   smem frags, sync/k32: %.2f TF\n", fl / ms * 1e-9);
    ms = timeit([&] { gemm_loop_kernel<1, 0><<<g, b>>>(out, it3); });
    printf("GEMM-loop m16n8k8  smem frags, no sync : %.2f TF\n", fl / ms * 1e-9);
    ms = timeit([&] { gemm_loop_kernel<1, 1><<<g, b>>>(out, it3); });
    printf("GEMM-loop m16n8k8  smem frags, sync/k16: %.2f TF\n", fl / ms * 1e-9);
    ms = timeit([&] { gemm_loop_kernel<2, 0><<<g, b>>>(out, it3); });
    printf("GEMM-loop m16n8k16 smem frags, no sync : %.2f TF\n", fl / ms * 1e-9);
    ms = timeit([&] { gemm_loop_kernel<2, 1><<<g, b>>>(out, it3); });
    printf("GEMM-loop m16n8k16 smem frags, sync/k16: %.2f TF\n", fl / ms * 1e-9);
  }
  {
    dim3 g(sm * 2), b(512);
    int it2 = 2000;
    float ms = timeit([&] { exp_kernel<<<g, b>>>(out, it2, 0.37); });
    printf("exp():  %.2f G evals/s\n", 4.0 * it2 * 512.0 * sm * 2 / ms * 1e-6);
    ms = timeit([&] { sqrt_kernel<<<g, b>>>(out, it2, 0.37); });
    printf("sqrt(): %.2f G evals/s\n", 4.0 * it2 * 512.0 * sm * 2 / ms * 1e-6);
  }

  {
    dim3 g(sm * 2), b(512);
    int it2 = 2000;
    float ms = timeit([&] { lean_matern_kernel<<<g, b>>>(out, it2, 0.37); });
    printf("lean Matern52 epilogue: %.2f G evals/s\n", 4.0 * it2 * 512.0 * sm * 2 / ms * 1e-6);
    ms = timeit([&] { libm_matern_kernel<<<g, b>>>(out, it2, 0.37); });
    printf("libm Matern52 epilogue: %.2f G evals/s\n", 4.0 * it2 * 512.0 * sm * 2 / ms * 1e-6);
    ms = timeit([&] { i2d_cvt_kernel<<<g, b>>>(out, it2, 5); });
    printf("cvt.f64.s32 + dadd: %.2f G/s\n", 8.0 * it2 * 512.0 * sm * 2 / ms * 1e-6);
    ms = timeit([&] { i2d_magic_kernel<<<g, b>>>(out, it2, 5); });
    printf("magic i2d + dadd:   %.2f G/s\n", 8.0 * it2 * 512.0 * sm * 2 / ms * 1e-6);
    accuracy_kernel<<<1, 256>>>(out); CK(cudaDeviceSynchronize());
    double h[768]; CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
    double me = 0, msq = 0, mseed = 0;
    for (int i = 0; i < 256; i++) { me = fmax(me, h[3*i]); msq = fmax(msq, h[3*i+1]); mseed = fmax(mseed, h[3*i+2]); }
    printf("accuracy on device: exp_neg max rel %.3e, sqrt_pos max rel %.3e, rsqrt seed max rel %.3e\n", me, msq, mseed);
  }
  {
    size_t bytes = (size_t)8 << 30;  // 8 GiB each
    double2 *a, *b2; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b2, bytes));
    CK(cudaMemset(a, 0, bytes)); CK(cudaMemset(b2, 0, bytes));
    size_t n = bytes / sizeof(double2);
    float ms = timeit([&] { copy_kernel<<<sm * 16, 512>>>(a, b2, n); });
    printf("HBM copy  (r+w): %.1f GB/s\n", 2.0 * bytes / ms * 1e-6);
    ms = timeit([&] { write_kernel<<<sm * 16, 512>>>(b2, n, 1.0); });
    printf("HBM write      : %.1f GB/s\n", 1.0 * bytes / ms * 1e-6);
    ms = timeit([&] { read_kernel<<<sm * 16, 512>>>(a, out, n); });
    printf("HBM read       : %.1f GB/s\n", 1.0 * bytes / ms * 1e-6);
    ms = timeit([&] { CK(cudaMemsetAsync(b2, 0, bytes)); });
    printf("cudaMemset     : %.1f GB/s\n", 1.0 * bytes / ms * 1e-6);
  }
  return 0;
}
