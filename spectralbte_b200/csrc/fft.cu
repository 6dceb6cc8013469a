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

void launch_fft3d(sbte_ctx* c, const double* in_real, const double2* in_cplx, int invert, int batch,
                  double2* out_nat, double2* out_lay, int layout, double* out_real, bool accumulate_real) {
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
