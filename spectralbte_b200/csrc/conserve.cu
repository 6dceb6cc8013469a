// spectralbte_b200/csrc/conserve.cu -- K4 / K5: conservation projection, moments, Maxwellian split
// and the time-integration glue.
//
// References (relative to /root/reference):
//   conserveAllMoments             src/conserve.c:207-264 (reduce b = C Q, solve, Q -= C^T lambda)
//   solveWithCCt                   src/conserve.c:173-202 (applied per cell by one thread)
//   getDensity/BulkVelocity/Temp.  src/momentRoutines.c:58-72,116-142,168-183
//   getEnergy                      src/momentRoutines.c:146-165
//   find_maxwellians               src/collisions.c:91-106
//   Euler / Heun updates           exec/boltz.c:206-241 (0D), :296-343 (1D)
// One CTA per cell; all reductions are fixed-order trees (no atomics) so results are deterministic.
#include "common.cuh"
#include "internal.h"

namespace sbte {

constexpr int RED_THREADS = 512;

__device__ __forceinline__ void decode(int idx, int N, int& i, int& j, int& k) {
  i = idx / (N * N);
  j = (idx / N) % N;
  k = idx % N;
}

// b[0..4] = sum over the cell of Q * w3 dv^3 {1, v_i, v_j, v_k, |v|^2/2}
__device__ __forceinline__ void functionals(const double* __restrict__ Q, const double* __restrict__ v,
                                            const double* __restrict__ wt, int N, double dv3, double (&b)[5],
                                            double* scratch) {
  const int n3 = N * N * N;
#pragma unroll
  for (int a = 0; a < 5; a++) b[a] = 0.0;
  for (int idx = threadIdx.x; idx < n3; idx += blockDim.x) {
    int i, j, k;
    decode(idx, N, i, j, k);
    const double pre = wt[i] * wt[j] * wt[k] * dv3;
    const double q = Q[idx];
    b[0] += q * pre;
    b[1] += q * (pre * v[i]);
    b[2] += q * (pre * v[j]);
    b[3] += q * (pre * v[k]);
    b[4] += q * (pre * 0.5 * (v[i] * v[i] + v[j] * v[j] + v[k] * v[k]));
  }
  block_reduce_sum<5>(b, scratch);
}

__global__ void __launch_bounds__(RED_THREADS)
conserve_kernel(double* __restrict__ Qall, const double* __restrict__ v, const double* __restrict__ wt, int N,
                double dv, ConsLU lu) {
  __shared__ double scratch[5 * 32];
  __shared__ double lam[5];
  const int n3 = N * N * N;
  double* Q = Qall + (long)blockIdx.x * n3;
  const double dv3 = dv * dv * dv;
  double b[5];
  functionals(Q, v, wt, N, dv3, b, scratch);
  if (threadIdx.x == 0) {
    const int n = 5;
    for (int k = 0; k < n - 1; k++) {
      const int p = lu.piv[k];
      if (p != k) { const double t = b[p]; b[p] = b[k]; b[k] = t; }
      for (int i = k + 1; i < n; i++) b[i] -= lu.a[i * n + k] * b[k];
    }
    b[n - 1] = b[n - 1] / lu.a[(n - 1) * n + (n - 1)];
    for (int i = n - 2; i >= 0; i--) {
      double sum = 0.0;
      for (int j = i + 1; j < n; j++) sum += lu.a[i * n + j] * b[j];
      b[i] = 1.0 / lu.a[i * n + i] * (b[i] - sum);
    }
    for (int a = 0; a < n; a++) lam[a] = b[a];
  }
  __syncthreads();
  const double l0 = lam[0], l1 = lam[1], l2 = lam[2], l3 = lam[3], l4 = lam[4];
  for (int idx = threadIdx.x; idx < n3; idx += blockDim.x) {
    int i, j, k;
    decode(idx, N, i, j, k);
    const double pre = wt[i] * wt[j] * wt[k] * dv3;
    Q[idx] -= (pre * l0 + (pre * v[i]) * l1 + (pre * v[j]) * l2 + (pre * v[k]) * l3 +
               (pre * 0.5 * (v[i] * v[i] + v[j] * v[j] + v[k] * v[k])) * l4);
  }
}

void launch_conserve(sbte_ctx* c, double* Q, int batch) {
  conserve_kernel<<<batch, RED_THREADS, 0, c->stream>>>(Q, c->d_v, c->d_wt, c->N, c->dv, c->lu);
  c->launches += 1;
}

__global__ void __launch_bounds__(RED_THREADS)
functionals_kernel(const double* __restrict__ Qall, double* __restrict__ out, const double* __restrict__ v,
                   const double* __restrict__ wt, int N, double dv) {
  __shared__ double scratch[5 * 32];
  const int n3 = N * N * N;
  double b[5];
  functionals(Qall + (long)blockIdx.x * n3, v, wt, N, dv * dv * dv, b, scratch);
  if (threadIdx.x == 0)
    for (int a = 0; a < 5; a++) out[blockIdx.x * 5 + a] = b[a];
}

void launch_moment_functionals(sbte_ctx* c, const double* Q, double* b5, int batch) {
  functionals_kernel<<<batch, RED_THREADS, 0, c->stream>>>(Q, b5, c->d_v, c->d_wt, c->N, c->dv);
  c->launches += 1;
}

// out = a*x + b*y + (s*Q)/Kn, evaluated in the reference's operation order
__global__ void update_kernel(double* __restrict__ out, double a, const double* __restrict__ x, double b,
                              const double* __restrict__ y, double s, double Kn, const double* __restrict__ Q,
                              long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    double base = (a == 1.0) ? x[i] : a * x[i];
    if (y) base = base + b * y[i];
    out[i] = base + s * Q[i] / Kn;
  }
}

void launch_update(sbte_ctx* c, double* out, double a, const double* x, double b, const double* y, double s,
                   double Kn, const double* Q, long n) {
  const int threads = 256;
  long blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  update_kernel<<<(unsigned)blocks, threads, 0, c->stream>>>(out, a, x, b, y, s, Kn, Q, n);
  c->launches += 1;
}

// rho, u, T (three dependent weighted reductions, as the reference computes them) + energy split
__device__ __forceinline__ void cell_moments(const double* __restrict__ f, const double* __restrict__ v,
                                             const double* __restrict__ wt, int N, double dv3, double& rho,
                                             double (&u)[3], double& T, double (&e)[2], double* scratch) {
  const int n3 = N * N * N;
  double r1[3] = {0.0, 0.0, 0.0};  // rho, Epos, Eneg
  for (int idx = threadIdx.x; idx < n3; idx += blockDim.x) {
    int i, j, k;
    decode(idx, N, i, j, k);
    const double w = dv3 * wt[i] * wt[j] * wt[k];
    r1[0] += w * f[idx];
    const double en = w * f[idx] * (v[i] * v[i] + v[j] * v[j] + v[k] * v[k]);
    if (en > 0) r1[1] += en; else r1[2] -= en;
  }
  block_reduce_sum<3>(r1, scratch);
  rho = r1[0]; e[0] = r1[1]; e[1] = r1[2];
  double r2[3] = {0.0, 0.0, 0.0};
  for (int idx = threadIdx.x; idx < n3; idx += blockDim.x) {
    int i, j, k;
    decode(idx, N, i, j, k);
    const double w = dv3 * wt[i] * wt[j] * wt[k] / rho;
    r2[0] += (v[i] * w) * f[idx];
    r2[1] += (v[j] * w) * f[idx];
    r2[2] += (v[k] * w) * f[idx];
  }
  block_reduce_sum<3>(r2, scratch);
  u[0] = r2[0]; u[1] = r2[1]; u[2] = r2[2];
  double r3[1] = {0.0};
  for (int idx = threadIdx.x; idx < n3; idx += blockDim.x) {
    int i, j, k;
    decode(idx, N, i, j, k);
    const double t = (v[i] - u[0]) * (v[i] - u[0]) + (v[j] - u[1]) * (v[j] - u[1]) + (v[k] - u[2]) * (v[k] - u[2]);
    r3[0] += t * dv3 * wt[i] * wt[j] * wt[k] * f[idx] / (3.0 * rho);
  }
  block_reduce_sum<1>(r3, scratch);
  T = r3[0];
}

__global__ void __launch_bounds__(RED_THREADS)
moments_kernel(const double* __restrict__ fall, double* __restrict__ out, const double* __restrict__ v,
               const double* __restrict__ wt, int N, double dv) {
  __shared__ double scratch[5 * 32];
  const int n3 = N * N * N;
  double rho, u[3], T, e[2];
  cell_moments(fall + (long)blockIdx.x * n3, v, wt, N, dv * dv * dv, rho, u, T, e, scratch);
  if (threadIdx.x == 0) {
    double* o = out + (long)blockIdx.x * 8;
    o[0] = rho; o[1] = u[0]; o[2] = u[1]; o[3] = u[2]; o[4] = T; o[5] = e[0]; o[6] = e[1]; o[7] = rho * T;
  }
}

void launch_moments(sbte_ctx* c, const double* f, double* mom8, int batch) {
  moments_kernel<<<batch, RED_THREADS, 0, c->stream>>>(f, mom8, c->d_v, c->d_wt, c->N, c->dv);
  c->launches += 1;
}

// M = Maxwellian with the moments of f;  g = f - Msub  (Msub == nullptr -> M itself).  The reference
// subtracts M_i from BOTH distributions (src/collisions.c:104), hence the explicit Msub.
__global__ void __launch_bounds__(RED_THREADS)
maxwellian_split_kernel(const double* __restrict__ f, const double* __restrict__ Msub, double* __restrict__ M,
                        double* __restrict__ g, const double* __restrict__ v, const double* __restrict__ wt, int N,
                        double dv) {
  __shared__ double scratch[5 * 32];
  const int n3 = N * N * N;
  double rho, u[3], T, e[2];
  cell_moments(f, v, wt, N, dv * dv * dv, rho, u, T, e, scratch);
  const double pre = rho * pow(0.5 / (M_PI * T), 1.5);
  for (int idx = threadIdx.x; idx < n3; idx += blockDim.x) {
    int i, j, k;
    decode(idx, N, i, j, k);
    const double m = pre * exp(-(0.5 / T) * ((v[i] - u[0]) * (v[i] - u[0]) + (v[j] - u[1]) * (v[j] - u[1]) +
                                            (v[k] - u[2]) * (v[k] - u[2])));
    M[idx] = m;
    g[idx] = f[idx] - (Msub ? Msub[idx] : m);
  }
}

void launch_maxwellian_split(sbte_ctx* c, const double* f, const double* Msub, double* M, double* g) {
  maxwellian_split_kernel<<<1, RED_THREADS, 0, c->stream>>>(f, Msub, M, g, c->d_v, c->d_wt, c->N, c->dv);
  c->launches += 1;
}

}  // namespace sbte
