// Dependent-issue latency of FP64 instructions on sm_100a (one thread, one chain): nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double* out, long long* cycles, int iters, double x0, double y0) {
  double x = x0, y = y0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) x = fma(x, y, y);                      // DFMA
    if (OP == 1) x = x * y;                             // DMUL
    if (OP == 2) x = x + y;                             // DADD
    if (OP == 3) x = double(__double2float_rn(x)) + y;  // F2F.F32.F64 + F2F.F64.F32 + DADD
    if (OP == 4) x = sqrt(x) + y;                       // library sqrt + DADD
    if (OP == 5) x = 1.0 / x + y;                       // library divide + DADD
    if (OP == 6) x = double(rsqrtf(float(x))) + y;      // cvt + MUFU.RSQ + cvt + DADD
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 8);
  const char* names[] = {"DFMA", "DMUL", "DADD", "F2F round trip + DADD", "sqrt + DADD", "1/x + DADD", "cvt+rsqrtf+cvt + DADD"};
  const int iters = 4096;
  for (int op = 0; op < 7; ++op) {
    for (int threads = 1; threads <= 32; threads *= 32) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: chain<0><<<1, threads>>>(out, cyc, iters, 1.0000001, 0.9999999); break;
          case 1: chain<1><<<1, threads>>>(out, cyc, iters, 1.0000001, 0.9999999); break;
          case 2: chain<2><<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9); break;
          case 3: chain<3><<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9); break;
          case 4: chain<4><<<1, threads>>>(out, cyc, iters, 2.0, 1.0); break;
          case 5: chain<5><<<1, threads>>>(out, cyc, iters, 2.0, 1.0); break;
          case 6: chain<6><<<1, threads>>>(out, cyc, iters, 2.0, 1.0); break;
        }
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("%-28s threads %2d : %.1f cycles per iteration\n", names[op], threads, double(h) / iters);
    }
  }
  return 0;
}
