// spectralbte_b200/csrc/fft.cu -- K1 / K3: batched N^3 transforms with the reference's twiddles.
//
// Replaces fft3D of the reference (/root/reference/src/collisions.c:232-283): pre-twiddle x trapezoid
// weight x (2 pi)^-3/2 delta^3, unnormalised 3-D DFT (FFTW at :270), post-twiddle.  The pack of the
// real input (:112-119) is fused into the first pass, the real-part extraction (:186-199, 218-220)
// into the last.  N <= 32 and N = 22, 24 are in scope, so each axis is a dense shared-memory DFT
// (3 N^4 complex MACs: < 0.3 % of the N^6 convolution); the twiddle (cos, sin) tables are computed
// once on the host with the reference's own expressions so the phases carry the same rounding.
//
// Two launches per transform:
//   pass ZY : one CTA per (x-plane, cell): twiddle-in, DFT along z, DFT along y   -> tmp
//   pass X  : one CTA per (y, cell):       DFT along x, twiddle-out, layout write  -> spectra / Re
#include <cooperative_groups.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "internal.h"

namespace sbte {

constexpr int FFT_THREADS = 256;

// dense DFT of the N lines of an N x N tile held in shared memory.
// ALONG_ROW: transform index is the fast one (tile[a][*]); else the slow one (tile[*][b]).
template <bool ALONG_ROW>
__device__ __forceinline__ void dft_tile(const double2* __restrict__ src, double2* __restrict__ dst,
                                         const double2* __restrict__ tw, int N, double sgn) {
  for (int t = threadIdx.x; t < N * N; t += blockDim.x) {
    const int a = t / N, b = t - a * N;          // output element (a, b)
    const int kp = ALONG_ROW ? b : a;            // output frequency along the transformed axis
    double sr = 0.0, si = 0.0;
    int m = 0;
    for (int k = 0; k < N; k++) {
      const double2 x = ALONG_ROW ? src[a * N + k] : src[k * N + b];
      const double2 w = tw[m];
      const double wi = sgn * w.y;
      sr += x.x * w.x - x.y * wi;
      si += x.x * wi + x.y * w.x;
      m += kp;
      if (m >= N) m -= N;
    }
    dst[t] = make_double2(sr, si);
  }
}

// optional stream-K input: the complex input is the fixed-order sum of `tile_np[t]` partial sums
struct PartsIn {
  const double2* parts;          // null: plain input
  size_t stride;                 // double2 elements between parts
  const unsigned char* tile_np;  // partial sums per tile
  int G, cols;                   // cell groups, zeta columns per tile
};

// up to four independent transforms served by one launch of the cluster kernel
struct FftJob {
  const double* in_real;
  const double2* in_cplx;
  int in_parts;              // > 1: the input is the sum of in_parts spectra n3 apart (split convolution), added in order
  double2* out_nat;
  double2* out_lay;
  double* out_real;
  double2* out_layT;         // optional second copy of a parity-layout spectrum with x and y swapped (transposed pairing)
};
struct FftJobs {
  FftJob j[4];
  int cells_per_job, layout, accumulate_real;
};

__global__ void __launch_bounds__(FFT_THREADS)
fft_pass_zy(const double* __restrict__ in_real, const double2* __restrict__ in_cplx, double2* __restrict__ tmp,
            const double2* __restrict__ pre, const double* __restrict__ wt, const double2* __restrict__ dft,
            int N, double prefactor, double sgn, PartsIn pin) {
  extern __shared__ double2 sm[];
  double2* A = sm;
  double2* B = sm + N * N;
  double2* tw = B + N * N;
  const int i = blockIdx.x;
  const long n3 = (long)N * N * N;
  const long cell = blockIdx.y;
  for (int t = threadIdx.x; t < N; t += blockDim.x) tw[t] = dft[t];
  for (int t = threadIdx.x; t < N * N; t += blockDim.x) {
    const int j = t / N, k = t - j * N;
    const long idx = cell * n3 + ((long)i * N + j) * N + k;
    double xr, xi;
    if (in_real) { xr = in_real[idx]; xi = 0.0; }
    else if (pin.parts) {
      const int tile = ((i * N + j) / pin.cols) * pin.G + (int)(cell >> 5);
      const int np = pin.tile_np[tile];
      xr = 0.0; xi = 0.0;
      for (int m = 0; m < np; m++) {
        const double2 z = pin.parts[(size_t)m * pin.stride + idx];
        xr += z.x; xi += z.y;
      }
    }
    else { const double2 z = in_cplx[idx]; xr = z.x; xi = z.y; }
    const double2 cs = pre[i + j + k];
    const double factor = prefactor * wt[i] * wt[j] * wt[k];
    A[t] = make_double2(factor * (cs.x * xr - cs.y * xi), factor * (cs.x * xi + cs.y * xr));
  }
  __syncthreads();
  dft_tile<true>(A, B, tw, N, sgn);   // along z
  __syncthreads();
  dft_tile<false>(B, A, tw, N, sgn);  // along y
  __syncthreads();
  for (int t = threadIdx.x; t < N * N; t += blockDim.x) tmp[cell * n3 + (long)i * N * N + t] = A[t];
}

__global__ void __launch_bounds__(FFT_THREADS)
fft_pass_x(const double2* __restrict__ tmp, const double2* __restrict__ post, const double2* __restrict__ dft,
           int N, double sgn, double2* __restrict__ out_nat, double2* __restrict__ out_lay, int layout,
           double* __restrict__ out_real, int accumulate_real) {
  extern __shared__ double2 sm[];
  double2* A = sm;
  double2* B = sm + N * N;
  double2* tw = B + N * N;
  const int j = blockIdx.x;
  const long n3 = (long)N * N * N;
  const long cell = blockIdx.y;
  for (int t = threadIdx.x; t < N; t += blockDim.x) tw[t] = dft[t];
  for (int t = threadIdx.x; t < N * N; t += blockDim.x) {
    const int i = t / N, k = t - i * N;
    A[t] = tmp[cell * n3 + ((long)i * N + j) * N + k];
  }
  __syncthreads();
  dft_tile<false>(A, B, tw, N, sgn);  // along x (slow index of the [x][z] tile)
  __syncthreads();
  for (int t = threadIdx.x; t < N * N; t += blockDim.x) {
    const int i = t / N, k = t - i * N;
    const long loc = ((long)i * N + j) * N + k;
    const double2 cs = post[loc];
    const double2 z = B[t];
    const double2 o = make_double2(cs.x * z.x - cs.y * z.y, cs.x * z.y + cs.y * z.x);
    if (out_nat) out_nat[cell * n3 + loc] = o;
    if (out_real) {
      if (accumulate_real) out_real[cell * n3 + loc] += o.x;
      else out_real[cell * n3 + loc] = o.x;
    }
    if (out_lay) {
      if (layout == LAY_PARITY) {
        out_lay[cell * n3 + ((long)i * N + j) * N + (k & 1) * (N / 2) + (k >> 1)] = o;
      } else if (layout == LAY_CELLMINOR) {
        out_lay[((cell >> 5) * n3 + loc) * 32 + (cell & 31)] = o;
      } else {
        out_lay[cell * n3 + loc] = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Whole-cell transform for batches (1D): one CTA per cell, the N^3 complex cell stays in shared memory
// (z padded to N+1 so that every axis pass reads conflict-free 16-byte words), each thread transforms whole
// lines held in registers; the DFT matrix entries come from constant memory (c_tw[N][m], compile-time
// offsets), which keeps every DFMA at two register operands.  One launch instead of two, no round trip
// through the intermediate buffer.
// ------------------------------------------------------------------------------------------
__constant__ double2 c_tw[33][32];   // c_tw[N][m] = (cos, sin)(2 pi m / N), filled once per process

void init_fft_constants() {   // constant memory is per device: call with the context's device current
  static std::atomic<unsigned> done_mask{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if ((done_mask.load() >> dev) & 1u) return;
  static double2 h[33][32];
  for (int n = 1; n <= 32; n++)
    for (int m = 0; m < 32; m++) h[n][m] = make_double2(cos(2.0 * M_PI * m / n), sin(2.0 * M_PI * m / n));
  cudaMemcpyToSymbol(c_tw, h, sizeof(h));
  done_mask.fetch_or(1u << dev);
}

constexpr int bitrev(int v, int bits) {
  int r = 0;
  for (int b = 0; b < bits; b++) r |= ((v >> b) & 1) << (bits - 1 - b);
  return r;
}
constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }

// power-of-two line: radix-2 decimation in frequency, entirely in registers (natural-order loads, bit-reversed
// stores so that every register index is a compile-time constant), twiddles from constant memory
template <int N, int M>
__device__ __forceinline__ void dif_stage(double2 (&x)[N], double sgn) {
  constexpr int H = M / 2;
#pragma unroll
  for (int k = 0; k < N; k += M) {
#pragma unroll
    for (int j = 0; j < H; j++) {
      const double wr = c_tw[N][j * (N / M)].x, wi = sgn * c_tw[N][j * (N / M)].y;
      const double2 a = x[k + j], b = x[k + j + H];
      const double dr = a.x - b.x, di = a.y - b.y;
      x[k + j] = make_double2(a.x + b.x, a.y + b.y);
      x[k + j + H] = make_double2(dr * wr - di * wi, dr * wi + di * wr);
    }
  }
  if constexpr (M > 2) dif_stage<N, M / 2>(x, sgn);
}

template <int N>
__device__ __forceinline__ void fft_line_pow2(double2* __restrict__ base, int stride, double sgn) {
  constexpr int LG = ilog2(N);
  double2 x[N];
#pragma unroll
  for (int k = 0; k < N; k++) x[k] = base[k * stride];
  dif_stage<N, N>(x, sgn);
#pragma unroll
  for (int k = 0; k < N; k++) base[bitrev(k, LG) * stride] = x[k];
}

// one output of a dense N-point DFT of a register-resident line; kp may be a run-time (warp-uniform) value
template <int N>
__device__ __forceinline__ double2 dense_output(const double2 (&x)[N], int kp, double sgn) {
  double sr = 0.0, si = 0.0;
  int m = 0;
#pragma unroll
  for (int k = 0; k < N; k++) {
    const double wr = c_tw[N][m].x, wi = sgn * c_tw[N][m].y;
    sr += x[k].x * wr - x[k].y * wi;
    si += x[k].x * wi + x[k].y * wr;
    m += kp;
    if (m >= N) m -= N;
  }
  return make_double2(sr, si);
}

template <int N>
__device__ __forceinline__ void dft_line(double2* __restrict__ base, int stride, double sgn) {
  if constexpr ((N & (N - 1)) == 0) {
    fft_line_pow2<N>(base, stride, sgn);
    return;
  }
  double2 x[N];
#pragma unroll
  for (int k = 0; k < N; k++) x[k] = base[k * stride];
  if constexpr (N > 16) {
    // long dense lines: keep the output loop rolled (the twiddle index is warp-uniform, so the constant-memory
    // reads still broadcast) -- fully unrolled, N = 24 needs more registers than a thread has
#pragma unroll 1
    for (int kp = 0; kp < N; kp++) base[kp * stride] = dense_output<N>(x, kp, sgn);
  } else {
#pragma unroll
    for (int kp = 0; kp < N; kp++) {
      double sr = 0.0, si = 0.0;
#pragma unroll
      for (int k = 0; k < N; k++) {
        const double wr = c_tw[N][(k * kp) % N].x, wi = sgn * c_tw[N][(k * kp) % N].y;
        sr += x[k].x * wr - x[k].y * wi;
        si += x[k].x * wi + x[k].y * wr;
      }
      base[kp * stride] = make_double2(sr, si);
    }
  }
}

// SPEC: 0 = every input / output / epilogue variant decided at run time; 1 = the slab step's forward transform (real
// input, cell-minor spectrum out); 2 = its inverse transform of the convolution's partial sums with the conservation /
// update epilogue (and the chained forward transform).  The specialised instances carry a fraction of the code: with one
// cell per SM (small slabs) every CTA runs through the kernel exactly once and instruction fetch is what it waits for.
template <int N, int SPEC>
__global__ void __launch_bounds__(256, ((N & (N - 1)) == 0) ? 2 : 1)
fft3d_cell_kernel(const double* __restrict__ in_real, const double2* __restrict__ in_cplx, PartsIn pin,
                  const double2* __restrict__ pre, const double2* __restrict__ post, const double* __restrict__ wt,
                  double prefactor, double sgn, double2* __restrict__ out_nat, double2* __restrict__ out_lay, int layout,
                  double* __restrict__ out_real, CellEpi epi) {
  extern __shared__ double2 cellsm[];   // [N][N][N+1]
  constexpr int P = N + 1;
  constexpr long n3 = (long)N * N * N;
  const long cell = blockIdx.x;
  // loads go out in groups of LB independent requests per thread (the CTA is latency-bound otherwise)
  constexpr int LB = 8;
  for (int base = threadIdx.x; base < n3; base += blockDim.x * LB) {
    double xr[LB], xi[LB];
#pragma unroll
    for (int q = 0; q < LB; q++) {
      const int idx = base + q * blockDim.x;
      xr[q] = 0.0; xi[q] = 0.0;
      if (idx < n3) {
        const long g = cell * n3 + idx;
        if (SPEC == 1 || (SPEC == 0 && in_real)) xr[q] = __ldg(in_real + g);
        else if (SPEC == 2 || pin.parts) {
          const int tile = ((idx / N) / pin.cols) * pin.G + (int)(cell >> 5);
          const int np = pin.tile_np[tile];
          for (int m = 0; m < np; m++) {
            const double2 z = __ldg(pin.parts + (size_t)m * pin.stride + g);
            xr[q] += z.x; xi[q] += z.y;
          }
        } else { const double2 z = __ldg(in_cplx + g); xr[q] = z.x; xi[q] = z.y; }
      }
    }
#pragma unroll
    for (int q = 0; q < LB; q++) {
      const int idx = base + q * blockDim.x;
      if (idx < n3) {
        const int i = idx / (N * N), j = (idx / N) % N, k = idx % N;
        const double2 cs = __ldg(pre + i + j + k);
        const double factor = prefactor * __ldg(wt + i) * __ldg(wt + j) * __ldg(wt + k);
        cellsm[(i * N + j) * P + k] =
            make_double2(factor * (cs.x * xr[q] - cs.y * xi[q]), factor * (cs.x * xi[q] + cs.y * xr[q]));
      }
    }
  }
  // The three axis passes are ONE piece of code run once, or -- when the epilogue chains the next stage's forward
  // transform -- twice (rolled loop: the second transform does not double the kernel's instruction footprint).
  int rep = 0;
  double sg = sgn;
#pragma unroll 1
  for (;;) {
  __syncthreads();
  for (int l = threadIdx.x; l < N * N; l += blockDim.x)            // along z: line (i, j)
    dft_line<N>(cellsm + l * P, 1, sg);
  __syncthreads();
  for (int l = threadIdx.x; l < N * N; l += blockDim.x) {          // along y: line (i, k)
    const int i = l / N, k = l % N;
    dft_line<N>(cellsm + (i * N) * P + k, P, sg);
  }
  __syncthreads();
  for (int l = threadIdx.x; l < N * N; l += blockDim.x) {          // along x: line (j, k)
    const int j = l / N, k = l % N;
    dft_line<N>(cellsm + j * P + k, N * P, sg);
  }
  __syncthreads();
  if (rep == 1) break;                                             // that was the chained forward transform
  if (!(SPEC == 2 || (SPEC == 0 && epi.mode == 1))) break;         // plain transform: on to the outputs
  {
    // Q = Re(post * z) stays in shared memory; moments -> multipliers -> corrected Q -> update, all here
    __shared__ double red[5 * 32];
    __shared__ double lam[5];
    double bsum[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    // (both element loops of this epilogue: unrolled so that the global loads of several elements are in flight at once
    // -- a CTA that runs alone on its SM is latency-bound; the sums keep their order)
#pragma unroll 4
    for (int idx = threadIdx.x; idx < n3; idx += blockDim.x) {
      const int i = idx / (N * N), j = (idx / N) % N, k = idx % N;
      const double2 cs = __ldg(post + idx);
      double2& z = cellsm[(i * N + j) * P + k];
      const double q = cs.x * z.x - cs.y * z.y;
      z.x = q;
      const double vi = __ldg(epi.v + i), vj = __ldg(epi.v + j), vk = __ldg(epi.v + k);
      const double pre = __ldg(epi.wt + i) * __ldg(epi.wt + j) * __ldg(epi.wt + k) * epi.dv3;
      bsum[0] += q * pre;
      bsum[1] += q * (pre * vi);
      bsum[2] += q * (pre * vj);
      bsum[3] += q * (pre * vk);
      bsum[4] += q * (pre * 0.5 * (vi * vi + vj * vj + vk * vk));
    }
    block_reduce_sum<5>(bsum, red);
    if (threadIdx.x == 0) {
      const int n = 5;
      for (int k = 0; k < n - 1; k++) {
        const int p = epi.lu.piv[k];
        if (p != k) { const double t = bsum[p]; bsum[p] = bsum[k]; bsum[k] = t; }
        for (int i = k + 1; i < n; i++) bsum[i] -= epi.lu.a[i * n + k] * bsum[k];
      }
      bsum[n - 1] = bsum[n - 1] / epi.lu.a[(n - 1) * n + (n - 1)];
      for (int i = n - 2; i >= 0; i--) {
        double sum = 0.0;
        for (int j = i + 1; j < n; j++) sum += epi.lu.a[i * n + j] * bsum[j];
        bsum[i] = 1.0 / epi.lu.a[i * n + i] * (bsum[i] - sum);
      }
      for (int a = 0; a < n; a++) lam[a] = bsum[a];
    }
    __syncthreads();
    const double l0 = lam[0], l1 = lam[1], l2 = lam[2], l3 = lam[3], l4 = lam[4];
#pragma unroll 4
    for (int idx = threadIdx.x; idx < n3; idx += blockDim.x) {
      const int i = idx / (N * N), j = (idx / N) % N, k = idx % N;
      const double vi = __ldg(epi.v + i), vj = __ldg(epi.v + j), vk = __ldg(epi.v + k);
      const double pre = __ldg(epi.wt + i) * __ldg(epi.wt + j) * __ldg(epi.wt + k) * epi.dv3;
      const double q = cellsm[(i * N + j) * P + k].x -
                       (pre * l0 + (pre * vi) * l1 + (pre * vj) * l2 + (pre * vk) * l3 +
                        (pre * 0.5 * (vi * vi + vj * vj + vk * vk)) * l4);
      const long g = cell * n3 + idx;
      double basev = (epi.a == 1.0) ? epi.x[g] : epi.a * epi.x[g];
      if (epi.y) basev = basev + epi.b * epi.y[g];
      const double fnew = basev + epi.s * q / epi.Kn;
      epi.out[g] = fnew;
      if (epi.next_lay) {
        // chained forward transform of the updated cell (the next Heun stage starts from it): same pre-twiddle and
        // trapezoid factor as the stand-alone forward pass above, imaginary input 0
        const double2 cs = __ldg(epi.next_pre + i + j + k);
        const double factor = epi.next_pref * __ldg(epi.wt + i) * __ldg(epi.wt + j) * __ldg(epi.wt + k);
        const double xr = fnew, xi = 0.0;
        cellsm[(i * N + j) * P + k] = make_double2(factor * (cs.x * xr - cs.y * xi), factor * (cs.x * xi + cs.y * xr));
      }
    }
    if (!epi.next_lay) return;
  }
  rep = 1;
  sg = -1.0;
  }
  if (rep == 1) {
    for (int idx = threadIdx.x; idx < n3; idx += blockDim.x) {       // post-twiddle, cell-minor layout for the convolution
      const int i = idx / (N * N), j = (idx / N) % N, k = idx % N;
      const double2 cs = __ldg(epi.next_post + idx);
      const double2 z = cellsm[(i * N + j) * P + k];
      epi.next_lay[((cell >> 5) * n3 + idx) * 32 + (cell & 31)] =
          make_double2(cs.x * z.x - cs.y * z.y, cs.x * z.y + cs.y * z.x);
    }
    return;
  }
  for (int base = threadIdx.x; base < n3; base += blockDim.x * LB) {
    double2 cs[LB];
#pragma unroll
    for (int q = 0; q < LB; q++) {
      const int idx = base + q * blockDim.x;
      if (idx < n3) cs[q] = __ldg(post + idx);
    }
#pragma unroll
    for (int q = 0; q < LB; q++) {
      const int idx = base + q * blockDim.x;
      if (idx >= n3) continue;
      const int i = idx / (N * N), j = (idx / N) % N, k = idx % N;
      const double2 z = cellsm[(i * N + j) * P + k];
      const double2 o = make_double2(cs[q].x * z.x - cs[q].y * z.y, cs[q].x * z.y + cs[q].y * z.x);
      if (SPEC == 1) {
        out_lay[((cell >> 5) * n3 + idx) * 32 + (cell & 31)] = o;
        continue;
      }
      if (out_nat) out_nat[cell * n3 + idx] = o;
      if (out_real) out_real[cell * n3 + idx] = o.x;
      if (out_lay) {
        if (layout == LAY_PARITY) out_lay[cell * n3 + (idx - k) + (k & 1) * (N / 2) + (k >> 1)] = o;
        else if (layout == LAY_CELLMINOR) out_lay[((cell >> 5) * n3 + idx) * 32 + (cell & 31)] = o;
        else out_lay[cell * n3 + idx] = o;
      }
    }
  }
}

template <int N>
static void launch_cell_n(sbte_ctx* c, const double* in_real, const double2* in_cplx, PartsIn pin, int invert, int batch,
                          double2* out_nat, double2* out_lay, int layout, double* out_real, const CellEpi& epi) {
  const size_t smem = (size_t)N * N * (N + 1) * sizeof(double2);
  static const bool generic_only = getenv("SBTE_CELL_FFT_GENERIC") != nullptr;
  int spec = 0;
  if (!generic_only) {
    if (!invert && in_real && !out_nat && !out_real && out_lay && layout == LAY_CELLMINOR && epi.mode == 0) spec = 1;
    else if (invert && pin.parts && epi.mode == 1) spec = 2;
  }
  auto kern = spec == 1 ? fft3d_cell_kernel<N, 1> : (spec == 2 ? fft3d_cell_kernel<N, 2> : fft3d_cell_kernel<N, 0>);
  static std::atomic<unsigned> configured[3];   // per device: function attributes belong to the device context
  if (!((configured[spec].load() >> c->device) & 1u)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured[spec].fetch_or(1u << c->device);
  }
  const int d = invert ? 1 : 0;
  kern<<<batch, 256, smem, c->stream>>>(in_real, in_cplx, pin, c->d_pre[d], c->d_post[d], c->d_wt, c->pref[d],
                                        invert ? +1.0 : -1.0, out_nat, out_lay, layout, out_real, epi);
  c->launches += 1;
}

// returns false when the whole-cell kernel does not apply (small batches keep the plane-parallel pair)
static bool try_cell_fft(sbte_ctx* c, const double* in_real, const double2* in_cplx, PartsIn pin, int invert, int batch,
                         double2* out_nat, double2* out_lay, int layout, double* out_real, bool accumulate_real,
                         const CellEpi* epi_in = nullptr) {
  if ((batch < 8 && !c->cell_fft_any) || accumulate_real) return false;
  CellEpi epi = {};
  if (epi_in) epi = *epi_in;
  switch (c->N) {
    case 8: launch_cell_n<8>(c, in_real, in_cplx, pin, invert, batch, out_nat, out_lay, layout, out_real, epi); return true;
    case 12: launch_cell_n<12>(c, in_real, in_cplx, pin, invert, batch, out_nat, out_lay, layout, out_real, epi); return true;
    case 16: launch_cell_n<16>(c, in_real, in_cplx, pin, invert, batch, out_nat, out_lay, layout, out_real, epi); return true;
    default: return false;
  }
}

// ------------------------------------------------------------------------------------------
// Cluster transform (0D and any N the whole-cell kernel cannot hold): one thread-block CLUSTER of 8 CTAs per
// cell.  CTA r keeps x-planes [r N/8, (r+1) N/8) in shared memory, transforms them along z and y, then -- after
// a cluster barrier -- owns the x-lines of y in [r N/8, (r+1) N/8) and gathers each line from the eight CTAs'
// shared memory over DSMEM.  One launch per transform (or per group of up to four independent transforms),
// nothing round-trips through global memory, and a single N = 32 cell is spread over 8 SMs instead of
// waiting on two dependent launches of the plane-parallel pair.
// ------------------------------------------------------------------------------------------
constexpr int FFT_CL = 8;

template <int N>
__global__ void __launch_bounds__(256, 1)
fft3d_cluster_kernel(FftJobs jobs, PartsIn pin, const double2* __restrict__ pre, const double2* __restrict__ post,
                     const double* __restrict__ wt, double prefactor, double sgn) {
  namespace cg = cooperative_groups;
  constexpr int PL = N / FFT_CL;        // x-planes (and later y-rows) per CTA
  constexpr int P = N + 1;
  constexpr long n3 = (long)N * N * N;
  extern __shared__ double2 clsm[];     // [PL][N][P]
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const long cell = blockIdx.x / FFT_CL;
  const int job = (int)(cell / jobs.cells_per_job);
  const long coff = (cell - (long)job * jobs.cells_per_job) * n3;   // cell offset inside the job's arrays
  const double* in_real = jobs.j[job].in_real;
  const double2* in_cplx = jobs.j[job].in_cplx;
  const int in_parts = jobs.j[job].in_parts;
  const int i0 = r * PL;
  double2* postsm = clsm + PL * N * P;  // [N][PL][N]: post-twiddles of the outputs this CTA emits (rows j in its slice)
  constexpr int LB = 8;
  static_assert((PL * N * N) % (256 * LB) == 0 || (PL * N * N) < 256 * LB, "load batches must tile");
  pdl_launch_dependents();   // the convolution kernel may start prefetching weights while this grid runs
  for (int base = threadIdx.x; base < N * PL * N; base += blockDim.x * LB) {   // stage the post-twiddles
    double2 pv[LB];
#pragma unroll
    for (int q = 0; q < LB; q++) {
      const int idx = base + q * blockDim.x;
      if (idx < N * PL * N) {
        const int ip = idx / (PL * N), jl = (idx / N) % PL, k = idx % N;
        pv[q] = __ldg(post + ((long)ip * N + i0 + jl) * N + k);
      }
    }
#pragma unroll
    for (int q = 0; q < LB; q++) {
      const int idx = base + q * blockDim.x;
      if (idx < N * PL * N) postsm[idx] = pv[q];
    }
  }
  pdl_wait();                // the twiddles above are static; the input below is the previous kernel's output
  // Loads are issued in groups of LB independent requests per thread before anything consumes them: this
  // kernel is latency-bound (8 CTAs, a few KB each), so exposed round trips are what it costs.
  for (int base = threadIdx.x; base < PL * N * N; base += blockDim.x * LB) {
    double xr[LB], xi[LB];
#pragma unroll
    for (int q = 0; q < LB; q++) {
      const int idx = base + q * blockDim.x;
      xr[q] = 0.0; xi[q] = 0.0;
      if (idx < PL * N * N) {
        const int il = idx / (N * N), j = (idx / N) % N, k = idx % N, i = i0 + il;
        const long loc = ((long)i * N + j) * N + k;
        if (in_real) xr[q] = __ldg(in_real + coff + loc);
        else if (pin.parts) {
          const int tile = ((i * N + j) / pin.cols) * pin.G + (int)(cell >> 5);
          const int np = pin.tile_np[tile];
          for (int m = 0; m < np; m++) {
            const double2 z = pin.parts[(size_t)m * pin.stride + cell * n3 + loc];
            xr[q] += z.x; xi[q] += z.y;
          }
        } else {
          const double2 z = __ldg(in_cplx + coff + loc);
          xr[q] = z.x; xi[q] = z.y;
          for (int m = 1; m < in_parts; m++) {
            const double2 z2 = __ldg(in_cplx + (long)m * n3 + coff + loc);
            xr[q] += z2.x; xi[q] += z2.y;
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < LB; q++) {
      const int idx = base + q * blockDim.x;
      if (idx < PL * N * N) {
        const int il = idx / (N * N), j = (idx / N) % N, k = idx % N, i = i0 + il;
        const double2 cs = __ldg(pre + i + j + k);
        const double factor = prefactor * __ldg(wt + i) * __ldg(wt + j) * __ldg(wt + k);
        clsm[(il * N + j) * P + k] =
            make_double2(factor * (cs.x * xr[q] - cs.y * xi[q]), factor * (cs.x * xi[q] + cs.y * xr[q]));
      }
    }
  }
  __syncthreads();
  for (int l = threadIdx.x; l < PL * N; l += blockDim.x)            // along z: line (il, j)
    dft_line<N>(clsm + l * P, 1, sgn);
  __syncthreads();
  for (int l = threadIdx.x; l < PL * N; l += blockDim.x) {          // along y: line (il, k)
    const int il = l / N, k = l % N;
    dft_line<N>(clsm + (il * N) * P + k, P, sgn);
  }
  cluster.sync();
  double2* out_nat = jobs.j[job].out_nat;
  double2* out_lay = jobs.j[job].out_lay;
  double2* out_layT = jobs.j[job].out_layT;
  double* out_real = jobs.j[job].out_real;
  const int layout = jobs.layout, accumulate = jobs.accumulate_real;
  for (int l = threadIdx.x; l < PL * N; l += blockDim.x) {          // along x: line (j, k), gathered over DSMEM
    const int j = i0 + l / N, k = l % N;
    double2 x[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
      const double2* remote = cluster.map_shared_rank(clsm, i / PL);
      x[i] = remote[((i % PL) * N + j) * P + k];
    }
    auto emit = [&](int ip, double2 z) {   // output frequency ip of this line: post-twiddle and the layout writes
      const long loc = ((long)ip * N + j) * N + k;
      const double2 cs = postsm[(ip * PL + (j - i0)) * N + k];
      const double2 o = make_double2(cs.x * z.x - cs.y * z.y, cs.x * z.y + cs.y * z.x);
      if (out_nat) out_nat[coff + loc] = o;
      if (out_real) {
        if (accumulate) out_real[coff + loc] += o.x;
        else out_real[coff + loc] = o.x;
      }
      if (out_lay) {
        if (layout == LAY_PARITY) out_lay[coff + (loc - k) + (k & 1) * (N / 2) + (k >> 1)] = o;
        else if (layout == LAY_CELLMINOR) out_lay[((cell >> 5) * n3 + loc) * 32 + (cell & 31)] = o;
        else out_lay[coff + loc] = o;
      }
      if (out_layT) out_layT[coff + ((long)j * N + ip) * N + (k & 1) * (N / 2) + (k >> 1)] = o;   // element (ip, j, k) at (j, ip, k)
    };
    if constexpr ((N & (N - 1)) == 0) {
      constexpr int LG = ilog2(N);
      dif_stage<N, N>(x, sgn);
#pragma unroll
      for (int q = 0; q < N; q++) emit(bitrev(q, LG), x[q]);
    } else {
#pragma unroll 1
      for (int ip = 0; ip < N; ip++) emit(ip, dense_output<N>(x, ip, sgn));
    }
  }
  cluster.sync();   // nobody leaves while a neighbour may still read its planes
}

template <int N>
static bool launch_cluster_n(sbte_ctx* c, const FftJobs& jobs, PartsIn pin, int invert, int cells) {
  const size_t smem = (size_t)(N / FFT_CL) * N * (N + 1 + N) * sizeof(double2);   // planes + staged post-twiddles
  auto kern = fft3d_cluster_kernel<N>;
  static std::atomic<unsigned> configured{0};   // per device: function attributes belong to the device context
  if (!((configured.load() >> c->device) & 1u)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured.fetch_or(1u << c->device);
  }
  const int d = invert ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(cells * FFT_CL));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = FFT_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kern, jobs, pin, (const double2*)c->d_pre[d], (const double2*)c->d_post[d],
                     (const double*)c->d_wt, c->pref[d], invert ? +1.0 : -1.0);
  c->launches += 1;
  return true;
}

// N = 16, 32: up to four independent groups of `cells_per_job` cells in one launch.  (N = 24 compiles too, but
// its dense 24-point lines make the cluster kernel slower than the plane-parallel pair: 34.8 vs 18.4 us per
// transform measured on B200, so it keeps the pair.)
bool fft_cluster_supported(int N) { return N == 16 || N == 32; }
static bool launch_cluster(sbte_ctx* c, const FftJobs& jobs, PartsIn pin, int invert, int cells) {
  switch (c->N) {
    case 16: return launch_cluster_n<16>(c, jobs, pin, invert, cells);
    case 24: return launch_cluster_n<24>(c, jobs, pin, invert, cells);
    case 32: return launch_cluster_n<32>(c, jobs, pin, invert, cells);
    default: return false;
  }
}

// several independent forward transforms (real inputs -> spectra in `layout`) in ONE launch; false if N has no
// cluster kernel (the caller then issues them one by one)
bool launch_fft3d_multi(sbte_ctx* c, int njobs, const double* const* in_real, double2* const* out_lay, int layout) {
  if (!fft_cluster_supported(c->N) || njobs > 4 || getenv("SBTE_NO_CLUSTER_FFT")) return false;
  FftJobs jobs = {};
  for (int q = 0; q < njobs; q++) { jobs.j[q].in_real = in_real[q]; jobs.j[q].out_lay = out_lay[q]; }
  const bool want_T = c->fft_layT_multi != nullptr && layout == LAY_PARITY && njobs <= 3;
  if (want_T)
    for (int q = 0; q < njobs; q++) jobs.j[q].out_layT = c->fft_layT_multi[q];
  if (layout == LAY_CELLMINOR) return false;   // the cell-minor interleave belongs to one batched job
  jobs.cells_per_job = 1;
  jobs.layout = layout;
  PartsIn nopart = {nullptr, 0, nullptr, 0, 0};
  const bool ok = launch_cluster(c, jobs, nopart, 0, njobs);
  if (ok && want_T) c->fft_layT_done = true;
  return ok;
}

bool launch_fft3d_inverse_sum(sbte_ctx* c, const double2* parts, int nparts, double* out_real) {
  if (!fft_cluster_supported(c->N)) return false;
  FftJobs jobs = {};
  jobs.j[0].in_cplx = parts; jobs.j[0].in_parts = nparts; jobs.j[0].out_real = out_real;
  jobs.cells_per_job = 1;
  PartsIn nopart = {nullptr, 0, nullptr, 0, 0};
  return launch_cluster(c, jobs, nopart, 1, 1);
}

static bool try_cluster_fft(sbte_ctx* c, const double* in_real, const double2* in_cplx, PartsIn pin, int invert, int batch,
                            double2* out_nat, double2* out_lay, int layout, double* out_real, bool accumulate_real) {
  if (!fft_cluster_supported(c->N) || getenv("SBTE_NO_CLUSTER_FFT")) return false;
  FftJobs jobs = {};
  jobs.j[0].in_real = in_real; jobs.j[0].in_cplx = in_cplx;
  jobs.j[0].out_nat = out_nat; jobs.j[0].out_lay = out_lay; jobs.j[0].out_real = out_real;
  // a caller that wants the x<->y transposed parity-layout copy as well leaves its address in the context
  const bool want_T = c->fft_layT != nullptr && batch == 1 && layout == LAY_PARITY && out_lay != nullptr && !invert;
  jobs.j[0].out_layT = want_T ? c->fft_layT : nullptr;
  jobs.cells_per_job = batch;
  jobs.layout = layout;
  jobs.accumulate_real = accumulate_real ? 1 : 0;
  const bool ok = launch_cluster(c, jobs, pin, invert, batch);
  if (ok && want_T) c->fft_layT_done = true;
  return ok;
}

void launch_fft3d(sbte_ctx* c, const double* in_real, const double2* in_cplx, int invert, int batch,
                  double2* out_nat, double2* out_lay, int layout, double* out_real, bool accumulate_real) {
  PartsIn nopart = {nullptr, 0, nullptr, 0, 0};
  if (try_cell_fft(c, in_real, in_cplx, nopart, invert, batch, out_nat, out_lay, layout, out_real, accumulate_real)) return;
  if (try_cluster_fft(c, in_real, in_cplx, nopart, invert, batch, out_nat, out_lay, layout, out_real, accumulate_real)) return;
  const int N = c->N;
  const size_t smem = (size_t)(2 * N * N + N) * sizeof(double2);
  const int d = invert ? 1 : 0;
  const double sgn = invert ? +1.0 : -1.0;  // FFTW_BACKWARD / FFTW_FORWARD exponent sign
  dim3 grid(N, batch);
  PartsIn none = {nullptr, 0, nullptr, 0, 0};
  fft_pass_zy<<<grid, FFT_THREADS, smem, c->stream>>>(in_real, in_cplx, c->d_tmp, c->d_pre[d], c->d_wt, c->d_dft, N,
                                                      c->pref[d], sgn, none);
  fft_pass_x<<<grid, FFT_THREADS, smem, c->stream>>>(c->d_tmp, c->d_post[d], c->d_dft, N, sgn, out_nat, out_lay,
                                                     layout, out_real, accumulate_real ? 1 : 0);
  c->launches += 2;
}

void launch_fft3d_parts(sbte_ctx* c, const double2* parts, size_t part_stride, const BatchSched& sch, int invert,
                        int batch, double2* out_nat, double* out_real) {
  {
    PartsIn pin0 = {parts, part_stride, sch.tile_np, sch.G, sch.cols};
    if (try_cell_fft(c, nullptr, nullptr, pin0, invert, batch, out_nat, nullptr, 0, out_real, false)) return;
    if (try_cluster_fft(c, nullptr, nullptr, pin0, invert, batch, out_nat, nullptr, 0, out_real, false)) return;
  }
  const int N = c->N;
  const size_t smem = (size_t)(2 * N * N + N) * sizeof(double2);
  const int d = invert ? 1 : 0;
  const double sgn = invert ? +1.0 : -1.0;
  dim3 grid(N, batch);
  PartsIn pin = {parts, part_stride, sch.tile_np, sch.G, sch.cols};
  fft_pass_zy<<<grid, FFT_THREADS, smem, c->stream>>>(nullptr, nullptr, c->d_tmp, c->d_pre[d], c->d_wt, c->d_dft, N,
                                                      c->pref[d], sgn, pin);
  fft_pass_x<<<grid, FFT_THREADS, smem, c->stream>>>(c->d_tmp, c->d_post[d], c->d_dft, N, sgn, out_nat, nullptr, 0,
                                                     out_real, 0);
  c->launches += 2;
}

bool launch_fft3d_parts_update(sbte_ctx* c, const double2* parts, size_t part_stride, const BatchSched& sch, int batch,
                               const CellEpi& epi) {
  PartsIn pin = {parts, part_stride, sch.tile_np, sch.G, sch.cols};
  return try_cell_fft(c, nullptr, nullptr, pin, 1, batch, nullptr, nullptr, 0, nullptr, false, &epi);
}

// plain sum of the partial sums into a natural-layout Q^ (only used when the caller asks for Q^ itself)
__global__ void combine_parts_kernel(double2* __restrict__ out, PartsIn pin, int N, long total) {
  const long n3 = (long)N * N * N;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long cell = e / n3;
    const int loc = (int)(e - cell * n3);
    const int tile = ((loc / N) / pin.cols) * pin.G + (int)(cell >> 5);
    const int np = pin.tile_np[tile];
    double xr = 0.0, xi = 0.0;
    for (int m = 0; m < np; m++) {
      const double2 z = pin.parts[(size_t)m * pin.stride + e];
      xr += z.x; xi += z.y;
    }
    out[e] = make_double2(xr, xi);
  }
}

void launch_combine_parts(sbte_ctx* c, const double2* parts, size_t part_stride, const BatchSched& sch, int batch,
                          double2* out) {
  PartsIn pin = {parts, part_stride, sch.tile_np, sch.G, sch.cols};
  const long total = (long)batch * c->n3;
  combine_parts_kernel<<<148 * 8, 256, 0, c->stream>>>(out, pin, c->N, total);
  c->launches += 1;
}

}  // namespace sbte
