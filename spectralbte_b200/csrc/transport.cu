// spectralbte_b200/csrc/transport.cu -- K6 / K7: 1D upwind transport and diffuse walls on device slabs.
//
// References (relative to /root/reference):
//   upwindOne   src/transportroutines.c:94-238   (ghost fill :107-172, stencil :203-216)
//   upwindTwo   src/transportroutines.c:241-470  (extrapolated ghosts :271-297, wall faces :351-404,
//                                                 minmod MUSCL stencil :406-468)
//   advectTwo   src/transportroutines.c:477-492  (two upwindTwo passes, then the average)
//   minmod      src/transportroutines.c:80-90
//   setDiffuseReflectionBC  src/boundaryConditions.c:39-84
// Slabs are contiguous [cell][N^3]; threads run along the velocity index (coalesced), the x-stencil
// strides by whole cells.  v_x index i is the slowest velocity index: i < N/2 moves left.
#include "common.cuh"
#include "transport.h"

namespace sbte {

__device__ __forceinline__ double minmod3(double a, double b, double d) {
  if (a > 0 && b > 0 && d > 0) return fmin(fmin(a, b), d);
  if (a < 0 && b < 0 && d < 0) return fmax(fmax(a, b), d);
  return 0.0;
}

// ---------------------------------------------------------------- diffuse wall (K7)
// sigma_W from the outgoing half of `in`, Maxwellian at TW into the incoming half of `out`.
__global__ void __launch_bounds__(512)
diffuse_bc_kernel(const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ v,
                  const double* __restrict__ wt, int N, double hv, double TW, int bdry) {
  __shared__ double scratch[32];
  const int nn = N * N, half = (N / 2) * nn;
  const int obeg = (bdry == 0) ? 0 : half, oend = (bdry == 0) ? half : N * nn;   // outgoing
  const int ibeg = (bdry == 0) ? half : 0, iend = (bdry == 0) ? N * nn : half;   // incoming
  double s[1] = {0.0};
  for (int idx = obeg + threadIdx.x; idx < oend; idx += blockDim.x) {
    const int i = idx / nn, j = (idx / N) % N, k = idx % N;
    s[0] += v[i] * wt[i] * wt[j] * wt[k] * hv * hv * hv * in[idx];
  }
  block_reduce_sum<1>(s, scratch);
  double sig = s[0];
  if (bdry == 0) sig *= -sqrt(2.0 * M_PI * 1.0 / (1.0 * TW));
  else sig *= sqrt(2.0 * M_PI * 1.0 / (1.0 * TW));
  const double pre = sig * pow(0.5 * 1.0 / (M_PI * 1.0 * TW), 1.5);
  __syncthreads();  // `in` may alias `out`: all reads of the outgoing half are done (disjoint halves anyway)
  for (int idx = ibeg + threadIdx.x; idx < iend; idx += blockDim.x) {
    const int i = idx / nn, j = (idx / N) % N, k = idx % N;
    out[idx] = pre * exp(-0.5 * 1.0 / (1.0 * TW) * (v[i] * v[i] + v[j] * v[j] + v[k] * v[k]));
  }
}

void launch_diffuse_bc(cudaStream_t st, const double* in, double* out, const double* v, const double* wt, int N,
                       double hv, double TW, int bdry) {
  diffuse_bc_kernel<<<1, 512, 0, st>>>(in, out, v, wt, N, hv, TW, bdry);
}

// ---------------------------------------------------------------- first order (K6a)
// Ghost cells may live on the neighbouring GPU: peerL / peerR (null = use the local ghost cells) point at the
// neighbour rank's boundary cells mapped over NVLink (CUDA IPC), so the stencil itself performs the halo
// loads; peerL[g*n3 + p] is my ghost cell g, peerR[g*n3 + p] my ghost cell nX + order + g.
template <int ORDER>
__device__ __forceinline__ const double* cell_src(const double* __restrict__ f, const double* __restrict__ peerL,
                                                  const double* __restrict__ peerR, long n3, int nX, int l) {
  if (l < ORDER && peerL) return peerL + (long)l * n3;
  if (l >= nX + ORDER && peerR) return peerR + (long)(l - nX - ORDER) * n3;
  return f + (long)l * n3;
}

__global__ void __launch_bounds__(256)
upwind_one_kernel(const double* __restrict__ f, double* __restrict__ fc, const double* __restrict__ v,
                  const double* __restrict__ dx, int N, int nX, double dt, const double* __restrict__ peerL,
                  const double* __restrict__ peerR) {
  const long n3 = (long)N * N * N;
  const int l = blockIdx.y + 1;
  const double* fl = f + (long)l * n3;
  const double* fm = cell_src<1>(f, peerL, peerR, n3, nX, l - 1);
  const double* fp = cell_src<1>(f, peerL, peerR, n3, nX, l + 1);
  for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n3; p += (long)gridDim.x * blockDim.x) {
    const int i = (int)(p / (N * N));
    const double cfl = dt * v[i] / dx[l];
    double r;
    if (i < N / 2) r = (1.0 + cfl) * fl[p] - cfl * fp[p];
    else r = (1.0 - cfl) * fl[p] + cfl * fm[p];
    fc[(long)l * n3 + p] = r;
  }
}

void launch_upwind_one(cudaStream_t st, const double* f, double* fc, const double* v, const double* dx, int N, int nX,
                       double dt, const double* peerL, const double* peerR) {
  const long n3 = (long)N * N * N;
  dim3 grid((unsigned)((n3 + 255) / 256), nX);
  upwind_one_kernel<<<grid, 256, 0, st>>>(f, fc, v, dx, N, nX, dt, peerL, peerR);
}

// ---------------------------------------------------------------- cross-GPU ordering for peer halos
// Each rank owns two counters in IPC-shared device memory: ready (source array of pass e is complete) and
// done (my reads of the neighbours' arrays in pass e are complete).  Plain stream order on each GPU plus
// these two tiny kernels replaces the host-synchronised message exchange.
// Flags of one rank: {ready, done, epoch, error}.  epoch counts this rank's upwind passes ON THE DEVICE, so the same
// launches can be replayed from a CUDA graph; every rank issues the same sequence of passes.
//   begin: epoch++ ; ready = epoch ("my source array is complete") ; wait until each neighbour is ready for
//          this pass and has finished reading my cells in the previous one (done >= epoch - 1)
//   end:   done = epoch ("I no longer read my neighbours' source array of this pass")
//   quiesce: wait until the neighbours' done reaches my epoch (before overwriting cells they may be reading)
// Waits are bounded by `timeout` clock cycles (0 = wait for ever; the slab passes SBTE_HALO_TIMEOUT_S, default 120 s).
// A rank may legitimately be late (writing files, instantiating a graph, stopped in a debugger), so running out of
// time is NOT a trap: the waiting rank raises the error word flags[3], stops waiting -- every later wait of this slab
// returns at once -- and the host reports the failure at its next synchronising call (sbte_slab_moments / _download /
// _halo_state).  The CUDA context survives and no GPU hangs.
__device__ __forceinline__ void spin_until(volatile int* my, const volatile int* L, const volatile int* R, int vr, int vd,
                                           long long timeout) {
  if (my[3] == 0) {
    const long long t0 = clock64();
    while ((L && (L[0] < vr || L[1] < vd)) || (R && (R[0] < vr || R[1] < vd))) {
      if (timeout > 0 && clock64() - t0 > timeout) { my[3] = 1; break; }
      __nanosleep(64);
    }
  }
  __threadfence_system();
}
__global__ void halo_begin_kernel(int* my, const int* nbL, const int* nbR, long long timeout) {
  volatile int* m = reinterpret_cast<volatile int*>(my);
  const int e = m[2] + 1;
  m[2] = e;
  __threadfence_system();
  m[0] = e;
  __threadfence_system();
  spin_until(m, reinterpret_cast<const volatile int*>(nbL), reinterpret_cast<const volatile int*>(nbR), e, e - 1, timeout);
}
__global__ void halo_end_kernel(int* my) {
  __threadfence_system();
  volatile int* m = reinterpret_cast<volatile int*>(my);
  m[1] = m[2];
  __threadfence_system();
}
__global__ void halo_quiesce_kernel(int* my, const int* nbL, const int* nbR, long long timeout) {
  volatile int* m = reinterpret_cast<volatile int*>(my);
  spin_until(m, reinterpret_cast<const volatile int*>(nbL), reinterpret_cast<const volatile int*>(nbR), 0, m[2], timeout);
}
void launch_halo_begin(cudaStream_t st, int* my, const int* nbL, const int* nbR, long long timeout) {
  halo_begin_kernel<<<1, 1, 0, st>>>(my, nbL, nbR, timeout);
}
void launch_halo_end(cudaStream_t st, int* my) { halo_end_kernel<<<1, 1, 0, st>>>(my); }
void launch_halo_quiesce(cudaStream_t st, int* my, const int* nbL, const int* nbR, long long timeout) {
  halo_quiesce_kernel<<<1, 1, 0, st>>>(my, nbL, nbR, timeout);
}

// ---------------------------------------------------------------- second order (K6b)
// ghost by linear extrapolation: f[dst] = 2 f[a] - f[b]
__global__ void extrapolate_kernel(double* __restrict__ f, long n3, int dst, int a, int b) {
  for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n3; p += (long)gridDim.x * blockDim.x)
    f[dst * n3 + p] = 2 * f[a * n3 + p] - f[b * n3 + p];
}
void launch_extrapolate(cudaStream_t st, double* f, long n3, int dst, int a, int b) {
  extrapolate_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>(f, n3, dst, a, b);
}

// wall face values from the limited slope in the wall cell `l` (2 on the left, nX+1 on the right):
//   left  wall: outgoing half (i <  N/2): face = f_l - dx/2 * s ; no-flux fill (i >= N/2): f_l + dx/2 * s
//   right wall: outgoing half (i >= N/2): face = f_l + dx/2 * s ; no-flux fill (i <  N/2): f_l - dx/2 * s
// `fill_noflux` = 1 writes both halves (no wall model); 0 writes only the outgoing half (the diffuse
// kernel then fills the incoming half).
__global__ void wall_face_kernel(const double* __restrict__ f, double* __restrict__ face, const double* __restrict__ x,
                                 const double* __restrict__ dx, int N, int l, int right, int fill_noflux) {
  const long n3 = (long)N * N * N;
  const long half = (long)(N / 2) * N * N;
  for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n3; p += (long)gridDim.x * blockDim.x) {
    const bool lower = p < half;                 // i < N/2
    const bool outgoing = right ? !lower : lower;
    if (!outgoing && !fill_noflux) continue;
    const double fm = f[(long)(l - 1) * n3 + p], f0 = f[(long)l * n3 + p], fp = f[(long)(l + 1) * n3 + p];
    const double s = minmod3((f0 - fm) / (x[l] - x[l - 1]), (fp - f0) / (x[l + 1] - x[l]),
                             (fp - fm) / (x[l + 1] - x[l - 1]));
    // both walls: lower half (i < N/2) takes f0 - dx/2 s, upper half takes f0 + dx/2 s
    const double val = lower ? f0 - 0.5 * dx[l] * s : f0 + 0.5 * dx[l] * s;
    face[p] = val;
  }
}
void launch_wall_face(cudaStream_t st, const double* f, double* face, const double* x, const double* dx, int N, int l,
                      int right, int fill_noflux) {
  const long n3 = (long)N * N * N;
  wall_face_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>(f, face, x, dx, N, l, right, fill_noflux);
}

// MUSCL / minmod stencil for the owned cells l = 2 .. nX+1. left_wall / right_wall: this rank holds
// the physical boundary and uses the wall faces fl / fr instead of the neighbour reconstruction.
//
// Each thread marches one velocity node through UP_CH consecutive cells: the limited slope of cell l is the
// "own" slope of cell l and the upwind-neighbour slope of cell l+1 (v_x > 0) or l-1 (v_x < 0), so carrying it
// along the march computes every slope once (3 FP64 divisions per node and pass instead of 6) with exactly the
// reference's expressions (src/transportroutines.c:411-416,441-446), i.e. the same bits as the per-cell form.
constexpr int UP_CH = 8;

__device__ __forceinline__ double slope_at(double fm, double f0, double fp, const double* __restrict__ x, int l) {
  return minmod3((f0 - fm) / (x[l] - x[l - 1]), (fp - f0) / (x[l + 1] - x[l]), (fp - fm) / (x[l + 1] - x[l - 1]));
}

__global__ void __launch_bounds__(256)
upwind_two_kernel(const double* f, double* __restrict__ fc, const double* __restrict__ fl,
                  const double* __restrict__ fr, const double* __restrict__ v, const double* __restrict__ x,
                  const double* __restrict__ dx, int N, int nX, double dt, int left_wall, int right_wall,
                  const double* peerL, const double* peerR, double force) {
  const long n3 = (long)N * N * N;
  const int h = N / 2;
  const int l0 = 2 + blockIdx.y * UP_CH;
  const int l1 = min(l0 + UP_CH, nX + 2);
  const long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (p >= n3) return;
  const int i = (int)(p / (N * N));
  const int j = (int)((p / N) % N);
  const double hv = 0.5 * dt * v[i];
  auto F = [&](int l) { return __ldcg(cell_src<2>(f, peerL, peerR, n3, nX, l) + p); };
  auto forced = [&](double r, int l) {   // Poiseuille forcing (:428-436,457-465): d/dv_y of the pass input
    if (force == 0.0) return r;
    const double* c0 = f + (long)l * n3;
    if (j == 0) return r - force * c0[p + N];
    if (j == N - 1) return r - force * c0[p - N];
    return r - force * (c0[p + N] - c0[p - N]);
  };
  if (i >= h) {   // information travels to the right: face value of cell l-1 is the upwind one
    double fm = F(l0 - 1), f0 = F(l0);
    double sprev = 0.0;
    if (!(l0 == 2 && left_wall)) sprev = slope_at(F(l0 - 2), fm, f0, x, l0 - 1);
#pragma unroll
    for (int q = 0; q < UP_CH; q++) {
      const int l = l0 + q;
      if (l >= l1) break;
      const double fp = F(l + 1);
      const double s1 = slope_at(fm, f0, fp, x, l);
      const double cfl = hv / dx[l];
      double r;
      if (l == 2 && left_wall) r = f0 - cfl * (f0 + 0.5 * dx[l] * s1 - fl[p]);
      else r = f0 - cfl * (f0 + 0.5 * dx[l] * s1 - (fm + 0.5 * dx[l - 1] * sprev));
      fc[(long)l * n3 + p] = forced(r, l);
      fm = f0; f0 = fp; sprev = s1;
    }
  } else {        // information travels to the left: face value of cell l+1 is the upwind one
    double fm = F(l0 - 1), f0 = F(l0), fp = F(l0 + 1);
    double s1 = slope_at(fm, f0, fp, x, l0);
#pragma unroll
    for (int q = 0; q < UP_CH; q++) {
      const int l = l0 + q;
      if (l >= l1) break;
      const double cfl = hv / dx[l];
      double r, fpp = 0.0, s2 = 0.0;
      if (l == nX + 1 && right_wall) r = f0 - cfl * (fr[p] - (f0 - 0.5 * dx[l] * s1));
      else {
        fpp = F(l + 2);
        s2 = slope_at(f0, fp, fpp, x, l + 1);
        r = f0 - cfl * (fp - 0.5 * dx[l + 1] * s2 - (f0 - 0.5 * dx[l] * s1));
      }
      fc[(long)l * n3 + p] = forced(r, l);
      fm = f0; f0 = fp; fp = fpp; s1 = s2;
    }
  }
}
void launch_upwind_two(cudaStream_t st, const double* f, double* fc, const double* fl, const double* fr,
                       const double* v, const double* x, const double* dx, int N, int nX, double dt, int left_wall,
                       int right_wall, const double* peerL, const double* peerR, double force) {
  const long n3 = (long)N * N * N;
  dim3 grid((unsigned)((n3 + 255) / 256), (unsigned)((nX + UP_CH - 1) / UP_CH));
  upwind_two_kernel<<<grid, 256, 0, st>>>(f, fc, fl, fr, v, x, dx, N, nX, dt, left_wall, right_wall, peerL, peerR,
                                          force);
}

// fc = 0.5 * (f + fc) on n contiguous doubles (src/transportroutines.c:487-491)
__global__ void average_kernel(const double* __restrict__ f, double* __restrict__ fc, long n) {
  for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n; p += (long)gridDim.x * blockDim.x)
    fc[p] = 0.5 * (f[p] + fc[p]);
}
void launch_average(cudaStream_t st, const double* f, double* fc, long n) {
  long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  average_kernel<<<(unsigned)blocks, 256, 0, st>>>(f, fc, n);
}

// With lazy module loading (the CUDA default) the first launch of a kernel loads it, and the driver cannot do that
// while another kernel is spinning on the device: a pass that waits for a neighbour driven from the same process
// (several slabs on one GPU, or one process driving several GPUs) would then wait for a kernel that cannot be
// loaded.  Loading every transport / halo kernel up front removes that window.
int preload_transport_kernels() {
  cudaFuncAttributes a;
  const void* fns[] = {(const void*)diffuse_bc_kernel, (const void*)upwind_one_kernel, (const void*)upwind_two_kernel,
                       (const void*)halo_begin_kernel, (const void*)halo_end_kernel, (const void*)halo_quiesce_kernel,
                       (const void*)extrapolate_kernel, (const void*)wall_face_kernel, (const void*)average_kernel};
  for (const void* f : fns)
    if (cudaFuncGetAttributes(&a, f) != cudaSuccess) return 1;
  return 0;
}

}  // namespace sbte
