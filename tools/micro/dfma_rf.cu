// Microbenchmark: does a 3-distinct-operand DFMA stream run below the 2-cycle FP64 pipe rate, and do more
// warps per SM sub-partition hide it?  Also measures the plain FP64 FMA peak on this part (not in
// MEASURED_PEAKS.json).    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_rf dfma_rf.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int NCH, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k(double* out, const double* in, int iters) {
  double x[NCH], a[NCH], b[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) { x[i] = in[i] + threadIdx.x; a[i] = in[NCH + i]; b[i] = in[2 * NCH + i]; }
  const double s = in[100];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 8; rep++) {
#pragma unroll
      for (int i = 0; i < NCH; i++) {
        if (MODE == 0) x[i] = fma(a[i], b[(i + rep) % NCH], x[i]);          // 3 distinct register operands
        else if (MODE == 1) x[i] = fma(s, b[(i + rep) % NCH], x[i]);        // one operand shared by consecutive FMAs
        else if (MODE == 2) x[i] = x[i] * b[(i + rep) % NCH];               // DMUL, 2 operands
        else x[i] = fma(s, x[i], b[(i + rep) % NCH]);
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) r += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE, int NCH, int THREADS>
void run(const char* name, double* out, double* in) {
  const int iters = 2000 * 16 / NCH, blocks = 148;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, NCH, THREADS><<<blocks, THREADS>>>(out, in, 10);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("%-28s threads/SM=%4d  launch failed: %s\n", name, THREADS, cudaGetErrorString(err)); return; }
  cudaEventRecord(e0);
  k<MODE, NCH, THREADS><<<blocks, THREADS>>>(out, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)blocks * THREADS * iters * 8.0 * NCH;
  printf("%-28s threads/SM=%4d chains/thread=%2d  %.3f ms  %.2f Tinstr/s  = %.2f TFLOP/s (FMA=2)\n", name, THREADS, NCH, ms,
         n / ms * 1e-9, 2 * n / ms * 1e-9);
}
int main() {
  double *out, *in; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&in, 1024 * 8);
  double h[1024]; for (int i = 0; i < 1024; i++) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0, 16, 256>("dfma 3 distinct operands", out, in);
  run<1, 16, 256>("dfma shared multiplicand", out, in);
  run<2, 16, 256>("dmul", out, in);
  run<3, 16, 256>("dfma shared, acc in slot B", out, in);
  run<0, 16, 512>("dfma 3 distinct operands", out, in);
  run<1, 16, 512>("dfma shared multiplicand", out, in);
  run<2, 16, 512>("dmul", out, in);
  run<0, 8, 512>("dfma 3 distinct operands", out, in);
  run<0, 8, 1024>("dfma 3 distinct operands", out, in);
  run<1, 8, 1024>("dfma shared multiplicand", out, in);
  run<2, 8, 1024>("dmul", out, in);
  return 0;
}
