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

// ---------------------------------------------------------------- cross-GPU ordering for peer halos
// Each rank owns a few words in IPC-shared device memory -- flags {ready, done, epoch, error, blocks} -- and the
// stencil kernels themselves keep the ranks in step (no separate flag kernels, no host involvement, so a whole time
// step is one CUDA-graph replay):
//   entering pass e = epoch + 1 (epoch counts this rank's completed passes ON THE DEVICE; every rank issues the same
//   sequence of passes): block (0,0) publishes ready = e -- the source array is complete, stream order guarantees it;
//   only the blocks that touch the slab's ends wait: the first block row until the LEFT neighbour is ready for pass e
//   and has finished reading my cells in pass e-1 (its done >= e-1: this pass may overwrite what it read), the last
//   block row likewise on the RIGHT neighbour; interior blocks never wait;
//   leaving: the last block of the grid to finish publishes done = e (my reads of the neighbours' arrays are
//   complete) and epoch = e.
//   quiesce (first order only): wait until the neighbours' done reaches my epoch before the collision update
//   overwrites f, which they read in the same pass.  At second order the update writes f_1 and f_conv, neither of
//   which a neighbour reads before my next pass publishes ready, and the waits above cover every other overwrite.
// Memory ordering: a pass has ONE system-scope fence on the publishing side (before `ready`, by block (0,0); before
// `done`, by the last block, cumulative over the device-scope fence + counter every block leaves through) and
// system-scope acquire loads on the waiting side.  A system-scope fence per block -- the first form of this protocol --
// cost 0.05 ms per step with one neighbour: with peer mappings in place membar.sys takes microseconds
// (profiles/r02_halo_cost_ab.txt).
// Waits are bounded by `timeout` clock cycles (0 = wait for ever; the slab passes SBTE_HALO_TIMEOUT_S, default 120 s).
// A rank may legitimately be late (writing files, instantiating a graph, stopped in a debugger), so running out of
// time is NOT a trap: the waiting rank raises the error word flags[3], stops waiting -- every later wait of this slab
// returns at once -- and the host reports the failure at its next synchronising call (sbte_slab_moments / _download /
// _halo_state).  The CUDA context survives and no GPU hangs.
__device__ __forceinline__ void spin_until(volatile int* my, const volatile int* L, const volatile int* R, int vr, int vd,
                                           long long timeout) {
  if (my[3] == 0) {
    const long long t0 = clock64();
    // {ready, done} of a neighbour are adjacent words, 8-byte aligned: one load -- one NVLink round trip -- per look
    // The load is an acquire at system scope: what the neighbour wrote before it published is visible to this thread's
    // later loads (and, through the block barrier that follows, to the block's) -- no separate fence, which with peer
    // mappings in place costs microseconds on the path every boundary block takes.
    auto behind = [&](const volatile int* nb) {
      if (!nb) return false;
      long long w;
      asm volatile("ld.acquire.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(nb) : "memory");
      return (int)(w & 0xffffffffLL) < vr || (int)(w >> 32) < vd;
    };
    while (behind(L) || behind(R)) {
      if (timeout > 0 && clock64() - t0 > timeout) { my[3] = 1; break; }
      __nanosleep(32);
    }
  }
}
// all threads of the block; returns the pass number.  needL / needR: this block reads the left / right neighbour's
// cells or writes cells that neighbour reads.
__device__ __forceinline__ int halo_enter(const HaloSync& h, bool needL, bool needR) {
  if (!h.my) return 0;
  __shared__ int s_pass;
  if (threadIdx.x == 0) {
    volatile int* m = reinterpret_cast<volatile int*>(h.my);
    const int e = m[2] + 1;
    if (blockIdx.x == 0 && blockIdx.y == 0) {
      __threadfence_system();   // everything earlier kernels wrote (stream order) before the word that announces it
      m[0] = e;
    }
    const volatile int* L = (needL && h.nbL) ? reinterpret_cast<const volatile int*>(h.nbL) : nullptr;
    const volatile int* R = (needR && h.nbR) ? reinterpret_cast<const volatile int*>(h.nbR) : nullptr;
    if (L || R) spin_until(m, L, R, e, e - 1, h.timeout);
    s_pass = e;
  }
  __syncthreads();
  return s_pass;
}
__device__ __forceinline__ void halo_leave(const HaloSync& h, int e) {
  if (!h.my) return;
  __syncthreads();   // every load of this block has been consumed
  if (threadIdx.x == 0) {
    // device scope is enough to count the blocks; the one system-scope fence of the pass is the last block's, and it is
    // cumulative over what the counter ordered before it
    __threadfence();
    const int total = (int)(gridDim.x * gridDim.y);
    if (atomicAdd(h.my + 4, 1) == total - 1) {
      volatile int* m = reinterpret_cast<volatile int*>(h.my);
      m[4] = 0;
      __threadfence_system();
      m[1] = e;
      m[2] = e;
    }
  }
}

__global__ void __launch_bounds__(256)
upwind_one_kernel(const double* __restrict__ f, double* __restrict__ fc, const double* __restrict__ v,
                  const double* __restrict__ dx, int N, int nX, double dt, const double* __restrict__ peerL,
                  const double* __restrict__ peerR, HaloSync hs) {
  const long n3 = (long)N * N * N;
  const int l = blockIdx.y + 1;
  const int pass = halo_enter(hs, l == 1, l == nX);
  const double* fl = f + (long)l * n3;
  const double* fm = cell_src<1>(f, peerL, peerR, n3, nX, l - 1);
  const double* fp = cell_src<1>(f, peerL, peerR, n3, nX, l + 1);
  for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n3; p += (long)gridDim.x * blockDim.x) {
    const int i = (int)(p / (N * N));
    const double cfl = dt * v[i] / dx[l];
    double r;
    if (i < N / 2) r = (1.0 + cfl) * fl[p] - cfl * __ldcg(fp + p);
    else r = (1.0 - cfl) * fl[p] + cfl * __ldcg(fm + p);
    fc[(long)l * n3 + p] = r;
  }
  halo_leave(hs, pass);
}

void launch_upwind_one(cudaStream_t st, const double* f, double* fc, const double* v, const double* dx, int N, int nX,
                       double dt, const double* peerL, const double* peerR, const HaloSync& hs) {
  const long n3 = (long)N * N * N;
  dim3 grid((unsigned)((n3 + 255) / 256), nX);
  upwind_one_kernel<<<grid, 256, 0, st>>>(f, fc, v, dx, N, nX, dt, peerL, peerR, hs);
}

__global__ void halo_quiesce_kernel(int* my, const int* nbL, const int* nbR, long long timeout) {
  volatile int* m = reinterpret_cast<volatile int*>(my);
  spin_until(m, reinterpret_cast<const volatile int*>(nbL), reinterpret_cast<const volatile int*>(nbR), 0, m[2], timeout);
}
void launch_halo_quiesce(cudaStream_t st, int* my, const int* nbL, const int* nbR, long long timeout) {
  halo_quiesce_kernel<<<1, 1, 0, st>>>(my, nbL, nbR, timeout);
}

// ---------------------------------------------------------------- second order (K6b)
// Physical ends of the slab at second order (src/transportroutines.c:271-297 ghosts, :351-404 wall faces), one launch for
// both ends (blockIdx.y: 0 = left, 1 = right, when both are requested):
//   ghost cell by linear extrapolation, f[g] = 2 f[l] - f[in2];
//   wall face from the limited slope s of the wall cell l (2 on the left, nX+1 on the right), which reads the ghost value
//   just formed:  lower half (i < N/2): face = f_l - dx/2 s ; upper half: f_l + dx/2 s.
// fill_* = 1 writes both halves (no wall model: no-flux fill); 0 writes only the outgoing half (the diffuse kernel then
// fills the incoming half).
__global__ void edge_prep_kernel(double* __restrict__ f, double* __restrict__ fl, double* __restrict__ fr,
                                 const double* __restrict__ x, const double* __restrict__ dx, int N, int nX, int do_left,
                                 int do_right, int fill_left, int fill_right) {
  const long n3 = (long)N * N * N;
  const long half = (long)(N / 2) * N * N;
  const int right = (do_left && do_right) ? (int)blockIdx.y : (do_right ? 1 : 0);
  const int l = right ? nX + 1 : 2;                 // the wall cell
  const int g = right ? nX + 2 : 1;                 // the ghost cell next to it
  const int in2 = right ? nX : 3;                   // second cell inward
  double* face = right ? fr : fl;
  const int fill_noflux = right ? fill_right : fill_left;
  for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n3; p += (long)gridDim.x * blockDim.x) {
    const double f0 = f[(long)l * n3 + p], fi = f[(long)in2 * n3 + p];
    const double fg = 2 * f0 - fi;                  // ghost by linear extrapolation: f[g] = 2 f[l] - f[in2]
    f[(long)g * n3 + p] = fg;
    const bool lower = p < half;                    // i < N/2
    const bool outgoing = right ? !lower : lower;
    if (!outgoing && !fill_noflux) continue;
    const double fm = right ? fi : fg, fp = right ? fg : fi;   // cells l-1 and l+1
    const double s = minmod3((f0 - fm) / (x[l] - x[l - 1]), (fp - f0) / (x[l + 1] - x[l]),
                             (fp - fm) / (x[l + 1] - x[l - 1]));
    face[p] = lower ? f0 - 0.5 * dx[l] * s : f0 + 0.5 * dx[l] * s;
  }
}
void launch_edge_prep(cudaStream_t st, double* f, double* fl, double* fr, const double* x, const double* dx, int N, int nX,
                      int do_left, int do_right, int fill_left, int fill_right) {
  const long n3 = (long)N * N * N;
  dim3 grid((unsigned)((n3 + 255) / 256), (do_left && do_right) ? 2 : 1);
  edge_prep_kernel<<<grid, 256, 0, st>>>(f, fl, fr, x, dx, N, nX, do_left, do_right, fill_left, fill_right);
}

// MUSCL / minmod stencil for the owned cells l = 2 .. nX+1. left_wall / right_wall: this rank holds
// the physical boundary: 1 = the wall faces fl / fr (edge_prep_kernel, then the wall model) replace the neighbour
// reconstruction; 2 = an end without wall model (no-flux fill): ghost and face are formed here, with edge_prep_kernel's
// expressions, and that launch is not needed.
//
// Each thread marches one velocity node through UP_CH consecutive cells: the limited slope of cell l is the
// "own" slope of cell l and the upwind-neighbour slope of cell l+1 (v_x > 0) or l-1 (v_x < 0), so carrying it
// along the march computes every slope once (3 FP64 divisions per node and pass instead of 6) with exactly the
// reference's expressions (src/transportroutines.c:411-416,441-446), i.e. the same bits as the per-cell form.
// UP_CH is chosen per launch (launch_upwind_two): 8 for large slabs, fewer (or, to stay within one wave, a few more) for small ones, where the kernel is one
// wave of latency-bound blocks and a shorter march means more blocks and a shorter dependent chain per thread.

__device__ __forceinline__ double slope_at(double fm, double f0, double fp, const double* __restrict__ x, int l) {
  return minmod3((f0 - fm) / (x[l] - x[l - 1]), (fp - f0) / (x[l + 1] - x[l]), (fp - fm) / (x[l + 1] - x[l - 1]));
}

template <int UP_CH>
__global__ void __launch_bounds__(256)
upwind_two_kernel(const double* f, double* __restrict__ fc, const double* __restrict__ fl,
                  const double* __restrict__ fr, const double* __restrict__ v, const double* __restrict__ x,
                  const double* __restrict__ dx, int N, int nX, double dt, int left_wall, int right_wall,
                  const double* peerL, const double* peerR, double force, const double* __restrict__ avg, HaloSync hs) {
  const long n3 = (long)N * N * N;
  const int h = N / 2;
  // Block rows: the first and the last hold just the two cells at each end of the slab -- the only ones whose stencil
  // reaches a neighbour's cells and the only ones a neighbour reads -- so that the rows that may have to wait for a
  // neighbour have the shortest march; the rows between them march UP_CH cells each and never wait.
  int l0, l1;
  if (blockIdx.y == 0) { l0 = 2; l1 = 4; }
  else if (blockIdx.y == gridDim.y - 1) { l0 = nX; l1 = nX + 2; }
  else { l0 = 4 + ((int)blockIdx.y - 1) * UP_CH; l1 = min(l0 + UP_CH, nX); }
  const int pass = halo_enter(hs, blockIdx.y == 0, blockIdx.y == gridDim.y - 1);
  const long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (p < n3) {
  const int i = (int)(p / (N * N));
  const int j = (int)((p / N) % N);
  const double hv = 0.5 * dt * v[i];
  // the whole window this thread marches through, cells l0-2 .. l0+UP_CH+1, is requested before anything consumes it:
  // one memory round trip per thread instead of one per cell (the kernel is latency-bound for small slabs)
  double w[UP_CH + 4], av[UP_CH];
#pragma unroll
  for (int q = 0; q < UP_CH + 4; q++) {
    const int l = l0 - 2 + q;
    w[q] = (l <= l1 + 1) ? __ldcg(cell_src<2>(f, peerL, peerR, n3, nX, l) + p) : 0.0;
  }
  if (avg) {
#pragma unroll
    for (int q = 0; q < UP_CH; q++) av[q] = (l0 + q < l1) ? avg[(long)(l0 + q) * n3 + p] : 0.0;
  }
  // ends without wall model: the ghost by linear extrapolation, f[g] = 2 f[l] - f[in2] (src/transportroutines.c:271-297)
  if (left_wall == 2 && l0 == 2) w[1] = 2 * w[2] - w[3];
  if (right_wall == 2) {
#pragma unroll
    for (int q = 2; q < UP_CH + 4; q++)
      if (l0 - 2 + q == nX + 2) w[q] = 2 * w[q - 1] - w[q - 2];
  }
  auto F = [&](int l) { return w[l - l0 + 2]; };
  auto forced0 = [&](double r, int l) {   // Poiseuille forcing (:428-436,457-465): d/dv_y of the pass input
    if (force == 0.0) return r;
    const double* c0 = f + (long)l * n3;
    if (j == 0) return r - force * c0[p + N];
    if (j == N - 1) return r - force * c0[p - N];
    return r - force * (c0[p + N] - c0[p - N]);
  };
  // avg != null: the second pass of advectTwo ends with f_conv = (f + f_conv) / 2 (src/transportroutines.c:487-491);
  // the pass result goes through the same rounding as when it was stored first and averaged by a second kernel
  auto forced = [&](double r, int l, int q) {
    const double v = forced0(r, l);
    return avg ? 0.5 * (av[q] + v) : v;
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
      if (l == 2 && left_wall) {
        // no wall model: the face edge_prep_kernel would store for this (upper) half, f_l + dx/2 s, s = this cell's slope
        const double face = left_wall == 2 ? f0 + 0.5 * dx[l] * s1 : fl[p];
        r = f0 - cfl * (f0 + 0.5 * dx[l] * s1 - face);
      }
      else r = f0 - cfl * (f0 + 0.5 * dx[l] * s1 - (fm + 0.5 * dx[l - 1] * sprev));
      fc[(long)l * n3 + p] = forced(r, l, q);
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
      if (l == nX + 1 && right_wall) {
        const double face = right_wall == 2 ? f0 - 0.5 * dx[l] * s1 : fr[p];   // lower half: f_l - dx/2 s
        r = f0 - cfl * (face - (f0 - 0.5 * dx[l] * s1));
      }
      else {
        fpp = F(l + 2);
        s2 = slope_at(f0, fp, fpp, x, l + 1);
        r = f0 - cfl * (fp - 0.5 * dx[l + 1] * s2 - (f0 - 0.5 * dx[l] * s1));
      }
      fc[(long)l * n3 + p] = forced(r, l, q);
      fm = f0; f0 = fp; fp = fpp; s1 = s2;
    }
  }
  }
  halo_leave(hs, pass);
}
void launch_upwind_two(cudaStream_t st, const double* f, double* fc, const double* fl, const double* fr,
                       const double* v, const double* x, const double* dx, int N, int nX, double dt, int left_wall,
                       int right_wall, const double* peerL, const double* peerR, double force, const double* avg,
                       const HaloSync& hs) {
  const long n3 = (long)N * N * N;
  const unsigned gx = (unsigned)((n3 + 255) / 256);
  // the shortest march whose grid still fits the device in one wave (two blocks of 256 threads per SM)
  static int slots = 0;
  if (slots == 0) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    slots = 2 * (sms > 0 ? sms : 148);
  }
  // rows: the two cells at either end of the slab, and the nX - 4 cells between them in marches of `ch`
  auto rows = [&](int ch) { return 2 + (nX - 4 + ch - 1) / ch; };
  auto blocks = [&](int ch) { return (long)gx * rows(ch); };
#define SBTE_UPWIND_TWO(CH)                                                                                           \
  upwind_two_kernel<CH><<<dim3(gx, (unsigned)rows(CH)), 256, 0, st>>>(f, fc, fl, fr, v, x, dx, N, nX, dt, left_wall,      \
                                                                     right_wall, peerL, peerR, force, avg, hs)
  if (blocks(2) <= slots) SBTE_UPWIND_TWO(2);
  else if (blocks(3) <= slots) SBTE_UPWIND_TWO(3);
  else if (blocks(4) <= slots) SBTE_UPWIND_TWO(4);
  else if (blocks(5) <= slots) SBTE_UPWIND_TWO(5);
  else if (blocks(6) <= slots) SBTE_UPWIND_TWO(6);
  else if (blocks(8) <= slots) SBTE_UPWIND_TWO(8);
  else if (blocks(10) <= slots) SBTE_UPWIND_TWO(10);   // one wave of longer marches beats two waves of short ones
  else if (blocks(12) <= slots) SBTE_UPWIND_TWO(12);
  else SBTE_UPWIND_TWO(8);
#undef SBTE_UPWIND_TWO
}

// With lazy module loading (the CUDA default) the first launch of a kernel loads it, and the driver cannot do that
// while another kernel is spinning on the device: a pass that waits for a neighbour driven from the same process
// (several slabs on one GPU, or one process driving several GPUs) would then wait for a kernel that cannot be
// loaded.  Loading every transport / halo kernel up front removes that window.
int preload_transport_kernels() {
  cudaFuncAttributes a;
  const void* fns[] = {(const void*)diffuse_bc_kernel, (const void*)upwind_one_kernel,
                       (const void*)upwind_two_kernel<2>, (const void*)upwind_two_kernel<3>,
                       (const void*)upwind_two_kernel<4>, (const void*)upwind_two_kernel<5>,
                       (const void*)upwind_two_kernel<6>, (const void*)upwind_two_kernel<8>,
                       (const void*)upwind_two_kernel<10>, (const void*)upwind_two_kernel<12>,
                       (const void*)halo_quiesce_kernel,
                       (const void*)edge_prep_kernel};
  for (const void* f : fns)
    if (cudaFuncGetAttributes(&a, f) != cudaSuccess) return 1;
  return 0;
}

}  // namespace sbte
