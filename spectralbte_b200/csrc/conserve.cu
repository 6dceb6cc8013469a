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
// One CTA per cell for slabs; a single cell (0D) is spread over a thread-block cluster of RED_CL CTAs whose
// partial sums are combined in rank order through distributed shared memory.  All reductions are fixed-order
// trees (no atomics) so results are deterministic.
#include <cooperative_groups.h>

#include "common.cuh"
#include "internal.h"

namespace sbte {

constexpr int RED_THREADS = 512;
constexpr int RED_CL = 8;   // CTAs per cell in the single-cell (cluster) variants

// this CTA's share [lo, hi) of the n3 nodes of its cell, and the cell index
template <int CL>
__device__ __forceinline__ void cell_slice(int n3, int& cell, int& lo, int& hi) {
  if constexpr (CL == 1) { cell = blockIdx.x; lo = 0; hi = n3; }
  else {
    const int r = blockIdx.x % CL;
    cell = blockIdx.x / CL;
    lo = (int)((long)n3 * r / CL);
    hi = (int)((long)n3 * (r + 1) / CL);
  }
}

// block sum, then (CL > 1) the sum over the cluster's CTAs in rank order; `xch` is a shared NV-double mailbox
// that must not be reused by a later call (no trailing barrier)
template <int NV, int CL>
__device__ __forceinline__ void reduce_all(double (&v)[NV], double* scratch, double* xch) {
  block_reduce_sum<NV>(v, scratch);
  if constexpr (CL > 1) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int a = 0; a < NV; a++) xch[a] = v[a];
    }
    cluster.sync();
#pragma unroll
    for (int a = 0; a < NV; a++) v[a] = 0.0;
    for (int r = 0; r < CL; r++) {
      const double* remote = cluster.map_shared_rank(xch, r);
#pragma unroll
      for (int a = 0; a < NV; a++) v[a] += remote[a];
    }
  }
}
template <int CL>
__device__ __forceinline__ void cluster_exit_barrier() {
  if constexpr (CL > 1) cooperative_groups::this_cluster().sync();   // mailboxes stay valid until everyone has read
}

template <typename K, typename... A>
static void launch_maybe_cluster(K kern, int ctas, int cl, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(RED_THREADS);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cl > 1 ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, args...);
}

__device__ __forceinline__ void decode(int idx, int N, int& i, int& j, int& k) {
  i = idx / (N * N);
  j = (idx / N) % N;
  k = idx % N;
}

// b[0..4] = sum over the cell of Q * w3 dv^3 {1, v_i, v_j, v_k, |v|^2/2}
template <int CL>
__device__ __forceinline__ void functionals(const double* __restrict__ Q, const double* __restrict__ v,
                                            const double* __restrict__ wt, int N, double dv3, double (&b)[5],
                                            double* scratch, double* xch, int lo, int hi) {
#pragma unroll
  for (int a = 0; a < 5; a++) b[a] = 0.0;
#pragma unroll 4
  for (int idx = lo + threadIdx.x; idx < hi; idx += blockDim.x) {
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
  reduce_all<5, CL>(b, scratch, xch);
}

template <int CL>
__global__ void __launch_bounds__(RED_THREADS)
conserve_kernel(double* __restrict__ Qall, const double* __restrict__ v, const double* __restrict__ wt, int N,
                double dv, ConsLU lu) {
  __shared__ double scratch[5 * 32];
  __shared__ double lam[5];
  __shared__ double xch[5];
  const int n3 = N * N * N;
  int cell, lo, hi;
  cell_slice<CL>(n3, cell, lo, hi);
  double* Q = Qall + (long)cell * n3;
  const double dv3 = dv * dv * dv;
  double b[5];
  functionals<CL>(Q, v, wt, N, dv3, b, scratch, xch, lo, hi);
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
#pragma unroll 4
  for (int idx = lo + threadIdx.x; idx < hi; idx += blockDim.x) {
    int i, j, k;
    decode(idx, N, i, j, k);
    const double pre = wt[i] * wt[j] * wt[k] * dv3;
    Q[idx] -= (pre * l0 + (pre * v[i]) * l1 + (pre * v[j]) * l2 + (pre * v[k]) * l3 +
               (pre * 0.5 * (v[i] * v[i] + v[j] * v[j] + v[k] * v[k])) * l4);
  }
  cluster_exit_barrier<CL>();
}

// a few cells (0D): a cluster per cell; slabs: a CTA per cell
static inline bool use_cluster(int batch) { return batch <= 4; }

void launch_conserve(sbte_ctx* c, double* Q, int batch) {
  if (use_cluster(batch))
    launch_maybe_cluster(conserve_kernel<RED_CL>, batch * RED_CL, RED_CL, c->stream, Q, (const double*)c->d_v,
                         (const double*)c->d_wt, c->N, c->dv, c->lu);
  else
    conserve_kernel<1><<<batch, RED_THREADS, 0, c->stream>>>(Q, c->d_v, c->d_wt, c->N, c->dv, c->lu);
  c->launches += 1;
}

__global__ void __launch_bounds__(RED_THREADS)
functionals_kernel(const double* __restrict__ Qall, double* __restrict__ out, const double* __restrict__ v,
                   const double* __restrict__ wt, int N, double dv) {
  __shared__ double scratch[5 * 32];
  const int n3 = N * N * N;
  double b[5];
  functionals<1>(Qall + (long)blockIdx.x * n3, v, wt, N, dv * dv * dv, b, scratch, nullptr, 0, n3);
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
template <int CL>
__device__ __forceinline__ void cell_moments(const double* __restrict__ f, const double* __restrict__ v,
                                             const double* __restrict__ wt, int N, double dv3, double& rho,
                                             double (&u)[3], double& T, double (&e)[2], double* scratch,
                                             double* xch /* 7 doubles */, int lo, int hi) {
  double r1[3] = {0.0, 0.0, 0.0};  // rho, Epos, Eneg
#pragma unroll 4
  for (int idx = lo + threadIdx.x; idx < hi; idx += blockDim.x) {
    int i, j, k;
    decode(idx, N, i, j, k);
    const double w = dv3 * wt[i] * wt[j] * wt[k];
    r1[0] += w * f[idx];
    const double en = w * f[idx] * (v[i] * v[i] + v[j] * v[j] + v[k] * v[k]);
    if (en > 0) r1[1] += en; else r1[2] -= en;
  }
  reduce_all<3, CL>(r1, scratch, xch);
  rho = r1[0]; e[0] = r1[1]; e[1] = r1[2];
  double r2[3] = {0.0, 0.0, 0.0};
#pragma unroll 4
  for (int idx = lo + threadIdx.x; idx < hi; idx += blockDim.x) {
    int i, j, k;
    decode(idx, N, i, j, k);
    const double w = dv3 * wt[i] * wt[j] * wt[k] / rho;
    r2[0] += (v[i] * w) * f[idx];
    r2[1] += (v[j] * w) * f[idx];
    r2[2] += (v[k] * w) * f[idx];
  }
  reduce_all<3, CL>(r2, scratch, xch + 3);
  u[0] = r2[0]; u[1] = r2[1]; u[2] = r2[2];
  double r3[1] = {0.0};
#pragma unroll 4
  for (int idx = lo + threadIdx.x; idx < hi; idx += blockDim.x) {
    int i, j, k;
    decode(idx, N, i, j, k);
    const double t = (v[i] - u[0]) * (v[i] - u[0]) + (v[j] - u[1]) * (v[j] - u[1]) + (v[k] - u[2]) * (v[k] - u[2]);
    r3[0] += t * dv3 * wt[i] * wt[j] * wt[k] * f[idx] / (3.0 * rho);
  }
  reduce_all<1, CL>(r3, scratch, xch + 6);
  T = r3[0];
}

template <int CL>
__global__ void __launch_bounds__(RED_THREADS)
moments_kernel(const double* __restrict__ fall, double* __restrict__ out, const double* __restrict__ v,
               const double* __restrict__ wt, int N, double dv) {
  __shared__ double scratch[5 * 32];
  __shared__ double xch[7];
  const int n3 = N * N * N;
  int cell, lo, hi;
  cell_slice<CL>(n3, cell, lo, hi);
  double rho, u[3], T, e[2];
  cell_moments<CL>(fall + (long)cell * n3, v, wt, N, dv * dv * dv, rho, u, T, e, scratch, xch, lo, hi);
  if (threadIdx.x == 0 && lo == 0) {
    double* o = out + (long)cell * 8;
    o[0] = rho; o[1] = u[0]; o[2] = u[1]; o[3] = u[2]; o[4] = T; o[5] = e[0]; o[6] = e[1]; o[7] = rho * T;
  }
  cluster_exit_barrier<CL>();
}

void launch_moments(sbte_ctx* c, const double* f, double* mom8, int batch) {
  if (use_cluster(batch))
    launch_maybe_cluster(moments_kernel<RED_CL>, batch * RED_CL, RED_CL, c->stream, f, mom8, (const double*)c->d_v,
                         (const double*)c->d_wt, c->N, c->dv);
  else
    moments_kernel<1><<<batch, RED_THREADS, 0, c->stream>>>(f, mom8, c->d_v, c->d_wt, c->N, c->dv);
  c->launches += 1;
}

// M = Maxwellian with the moments of f;  g = f - Msub  (Msub == nullptr -> M itself).  The reference
// subtracts M_i from BOTH distributions (src/collisions.c:104), hence the explicit Msub.
__global__ void __launch_bounds__(RED_THREADS)
maxwellian_split_kernel(const double* __restrict__ f, const double* __restrict__ Msub, double* __restrict__ M,
                        double* __restrict__ g, const double* __restrict__ v, const double* __restrict__ wt, int N,
                        double dv) {
  __shared__ double scratch[5 * 32];
  __shared__ double xch[7];
  const int n3 = N * N * N;
  int cell, lo, hi;
  cell_slice<RED_CL>(n3, cell, lo, hi);   // always one cluster: a single cell
  double rho, u[3], T, e[2];
  cell_moments<RED_CL>(f, v, wt, N, dv * dv * dv, rho, u, T, e, scratch, xch, lo, hi);
  const double pre = rho * pow(0.5 / (M_PI * T), 1.5);
#pragma unroll 4
  for (int idx = lo + threadIdx.x; idx < hi; idx += blockDim.x) {
    int i, j, k;
    decode(idx, N, i, j, k);
    const double m = pre * exp(-(0.5 / T) * ((v[i] - u[0]) * (v[i] - u[0]) + (v[j] - u[1]) * (v[j] - u[1]) +
                                            (v[k] - u[2]) * (v[k] - u[2])));
    M[idx] = m;
    g[idx] = f[idx] - (Msub ? Msub[idx] : m);
  }
  cluster_exit_barrier<RED_CL>();
}

void launch_maxwellian_split(sbte_ctx* c, const double* f, const double* Msub, double* M, double* g) {
  launch_maybe_cluster(maxwellian_split_kernel, RED_CL, RED_CL, c->stream, f, Msub, M, g, (const double*)c->d_v,
                       (const double*)c->d_wt, c->N, c->dv);
  c->launches += 1;
}

}  // namespace sbte
