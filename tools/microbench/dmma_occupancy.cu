// How many warps per SM sub-partition does the FP64 tensor pipe (DMMA.8x8x4) need to stay saturated?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_occupancy dmma_occupancy.cu
// One block per SM, W warps per block (W / 4 warps per sub-partition), ILP independent accumulator tiles per warp.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k(double* out, int iters, double a0) {
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
template <int ILP>
void run(double* out, int sms, int warps) {
  const int iters = 8192 * 8 / ILP;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<ILP><<<sms, warps * 32>>>(out, iters, 0.5); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0)); k<ILP><<<sms, warps * 32>>>(out, iters, 0.5); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  printf("warps/SM %2d (%.2f per sub-partition) ILP %2d : %.2f TFLOP/s\n", warps, warps / 4.0, ILP,
         2.0 * 256 * ILP * iters * double(sms) * warps / (best * 1e-3) * 1e-12);
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 1024));
  for (int warps : {1, 2, 4, 8, 12, 16}) { run<8>(out, sms, warps); run<32>(out, sms, warps); }
  return 0;
}
