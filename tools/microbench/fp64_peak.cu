// FP64 pipe microbenchmark for sm_100a: DFMA vs DMMA (m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peak fp64_peak.cu
// Reports TFLOP/s for each instruction shape with all SMs busy; used to pick the contraction primitive
// and as the FP64 roofline cross-check next to cuBLAS DGEMM (tools/microbench/dgemm_peak.py).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a0) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  double a = a0, b = 1.0 - a0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
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
__global__ void k_dmma884(double* out, int iters, double a0) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = a0, b = 1.0 - a0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int SHAPE>
__global__ void k_dmma16(double* out, int iters, double a0) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = a0 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 1.0 - a0 * i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (SHAPE == 4) dmma1684(c[i], a, b[0]);
      if (SHAPE == 8) dmma1688(c[i], a, b);
      if (SHAPE == 16) dmma16816(c[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double timeit(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best * 1e-3;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz L2 %d B\n", p.name, sms, p.clockRate, p.l2CacheSize);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 4096;
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = warps * 32; int blocks = sms * 2;
    if (warps * 2 > 64) blocks = sms;
    double nthreads = double(blocks) * threads;
    {
      double t = timeit([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 0.5); }, 5);
      printf("DFMA          warps/blk %2d blocks %d : %.2f TFLOP/s\n", warps, blocks, 2.0 * 8 * iters * nthreads / t * 1e-12);
    }
    {
      double t = timeit([&] { k_dmma884<8><<<blocks, threads>>>(out, iters, 0.5); }, 5);
      printf("DMMA m8n8k4   warps/blk %2d blocks %d : %.2f TFLOP/s\n", warps, blocks, 2.0 * 256 * 8 * iters * (nthreads / 32) / t * 1e-12);
    }
    {
      double t = timeit([&] { k_dmma16<4, 4><<<blocks, threads>>>(out, iters, 0.5); }, 5);
      printf("DMMA m16n8k4  warps/blk %2d blocks %d : %.2f TFLOP/s\n", warps, blocks, 2.0 * 512 * 4 * iters * (nthreads / 32) / t * 1e-12);
    }
    {
      double t = timeit([&] { k_dmma16<4, 8><<<blocks, threads>>>(out, iters, 0.5); }, 5);
      printf("DMMA m16n8k8  warps/blk %2d blocks %d : %.2f TFLOP/s\n", warps, blocks, 2.0 * 1024 * 4 * iters * (nthreads / 32) / t * 1e-12);
    }
    {
      double t = timeit([&] { k_dmma16<4, 16><<<blocks, threads>>>(out, iters, 0.5); }, 5);
      printf("DMMA m16n8k16 warps/blk %2d blocks %d : %.2f TFLOP/s\n", warps, blocks, 2.0 * 2048 * 4 * iters * (nthreads / 32) / t * 1e-12);
    }
  }
  return 0;
}
