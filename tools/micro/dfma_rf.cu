// Microbenchmark: does a 3-distinct-operand DFMA stream run below the 2-cycle FP64 pipe rate?
// Also measures the plain FP64 FMA peak on this part (not in MEASURED_PEAKS.json).
#include <cstdio>
#include <cuda_runtime.h>
#define NCH 16
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double* out, const double* in, int iters) {
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
template <int MODE>
void run(const char* name, int threads, double* out, double* in) {
  const int iters = 2000, blocks = 148;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, in, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)blocks * threads * iters * 8.0 * NCH;
  printf("%-28s threads/SM=%4d  %.3f ms  %.2f Tinstr/s  = %.2f TFLOP/s (FMA=2)\n", name, threads, ms, n / ms * 1e-9, 2 * n / ms * 1e-9);
}
int main() {
  double *out, *in; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&in, 1024 * 8);
  double h[1024]; for (int i = 0; i < 1024; i++) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int t : {256, 512, 1024}) {
    if (t == 256) { run<0>("dfma 3 distinct operands", t, out, in); run<1>("dfma shared multiplicand", t, out, in); run<2>("dmul", t, out, in); run<3>("dfma shared, acc in slot B", t, out, in); }
    if (t == 512) { run<0>("dfma 3 distinct operands", t, out, in); run<1>("dfma shared multiplicand", t, out, in); run<2>("dmul", t, out, in); }
    if (t == 1024) { run<0>("dfma 3 distinct operands", t, out, in); run<1>("dfma shared multiplicand", t, out, in); }
  }
  return 0;
}
