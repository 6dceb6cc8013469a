// spectralbte_b200/csrc/qhat_half.cu -- K2 for one cell (0D), f == g, on HALF of the zeta rows (opt-in: SBTE_HALF0D=1).
//
// The reference keeps Q = Re(fft3D^-1(Q^)) (/root/reference/src/collisions.c:212-221), so the mirror ("B") rows of Q^ may be
// folded into their partner ("A") rows wherever the fold is real (mirror.cuh: mirror_fold_weight; tests/test_half_spectrum_cpu.py).
// With the folded tensor Wh the kernels' formula is unchanged,
//     S[zeta] = sum_xi Wh[zeta][xi] f^[xi] f^[wrap(zeta + N/2 - xi)],      Q = Re(fft3D^-1(S)),
// but the B rows of Wh are zero except
//   (i)  on the steps (xi_x, xi_y) whose x/y phase exponent is not 0 (mirror_exy != 0: about 12 % of the steps at N = 32), and
//   (ii) on the other steps, in the row zeta_z = 0 and at the two entries xi_z = 0, (zeta - xi)_z = 0 of every other row.
// qhat_stream_half_kernel is qhat_stream_kernel (qhat.cu: same register tile, same 128-bit non-allocating weight stream, same
// TMA-staged operand planes) in which the CTAs of B columns skip the steps of (ii); qhat_half_leftover_kernel adds (ii) from
// the few entries it needs.  Weight bytes per evaluation: about 0.5 + 0.5 * 0.12 of the symmetrised stream, plus the sectors
// the leftover entries touch.
// Numerical margin: the reference's spectrum is Hermitian only to round-off (3e-13 of max|f^| at N = 32: its twiddle
// arguments carry the rounding of pi), so this path agrees with the reference's Q to ~3e-14 (N = 16) ... 6e-13 (N = 32, random weights)
// instead of the 1e-14 of qhat_stream_kernel; the tolerance is 1e-12.
// STATUS: runs on the host through tests/emul (tests/test_kernel_emulation_cpu.py: Q to 1e-12); not yet run on a GPU, off by
// default.  The result is NOT the reference's Q^ -- only its real inverse transform is the same -- so this path serves
// ComputeQ, never sbte_qhat.
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "internal.h"
#include "mirror.cuh"

namespace sbte {

bool qhat_half0d_enabled(int N) {
  static const bool on = getenv("SBTE_HALF0D") != nullptr && atoi(getenv("SBTE_HALF0D")) != 0;
  return on && (N == 16 || N == 32);
}

// WT: warps per CTA (8 for one operand pair; 16 for the two pairs of ComputeQ_maxPreserve, whose 128 KB plane ring allows
// only one CTA per SM), as in StreamCfg (qhat.cu)
template <int N, int WT = 8>
struct HalfCfg {
  static constexpr int HALF = N / 2;            // 2-column groups per N-long weight row segment
  static constexpr int RGW = 32 / HALF;         // 4-row groups handled by one warp
  static constexpr int ROWS_W = RGW * 4;        // rows per warp
  static constexpr int RB = N / ROWS_W;         // warps that tile the N rows of one zeta (x,y) column
  static constexpr int PH = WT / RB;            // xi_y phases (warps sharing the same rows)
  static constexpr int NWARP = RB * PH;
  static constexpr int THREADS = NWARP * 32;
  static constexpr int SPC = N / PH;            // steps per xi_x chunk per warp
  static constexpr int PLANE = N * N;           // complex elements per operand plane
  static constexpr int DEPTH = 2;               // weight tiles in flight per thread
  static constexpr int NP = WT / 8;             // operand pairs sharing the weight pass
  static constexpr size_t SMEM = (size_t)2 * 2 * NP * PLANE * sizeof(double2) + 64;
  static_assert(32 % HALF == 0 && N % ROWS_W == 0 && WT % RB == 0 && N % PH == 0 && (WT == 8 || WT == 16), "unsupported N");
};

// the c-th xi_x plane a column visits: the representatives of its own plane (A, unpaired) or the mirror images of its
// partner's representatives (B), cf. mirror_sym_weight
__device__ __forceinline__ int half_chunk_ex(int N, int zx, bool b, int c) {
  return b ? (N - sym_rep(N, (N - zx) % N, c)) % N : sym_rep(N, zx, c);
}

// NP = 2 (ComputeQ_maxPreserve, src/collisions.c:178-210): the summand is g_j^[xi] f^[zeta-xi] + M_j^[xi] g_i^[zeta-xi]
// (xiA/dfA and xiB/dfB); it obeys the same mirror relation, so the same folded tensor serves it.
template <int N, int NP>
__global__ void __launch_bounds__(HalfCfg<N, 8 * NP>::THREADS, (N == 32 && NP == 1) ? 2 : 1)
qhat_stream_half_kernel(const double* __restrict__ Wh, const double2* __restrict__ xiA, const double2* __restrict__ dfA,
                        const double2* __restrict__ xiB, const double2* __restrict__ dfB, double2* __restrict__ qhat,
                        int nsplit) {
  using C = HalfCfg<N, 8 * NP>;
  constexpr int HALF = C::HALF, PLANE = C::PLANE, DEPTH = C::DEPTH;
  constexpr long n3 = (long)N * N * N;
  constexpr uint32_t STAGE_ELEMS = 2 * NP * PLANE;   // per stage: NP x (xi-side plane, dif-side plane)
  SBTE_DYN_SMEM(smraw);
  double2* planes = reinterpret_cast<double2*>(smraw);                       // [2][NP][2][PLANE]
  uint64_t* full = reinterpret_cast<uint64_t*>(smraw + 2 * STAGE_ELEMS * sizeof(double2));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rb = warp % C::RB, ph = warp / C::RB;
  const int cp = lane % HALF, rgw = lane / HALF;
  const bool active = rgw < C::RGW;
  const int part = blockIdx.x / (N * N), colid = blockIdx.x - part * (N * N);
  const int zx = colid / N, zy = colid % N;
  const bool bcol = mirror_paired_column(N, zx, zy) && mirror_is_b_row(N, zx, zy);
  const int nchunk_all = sym_nrep(N, bcol ? (N - zx) % N : zx);
  // the xi_x planes of an A column are shared between nsplit CTAs (shorter tail); a B column keeps so few steps that one
  // CTA takes them all and the others only write their zero partial spectrum
  if (bcol && part > 0) {
    if (tid < N) qhat[(long)part * n3 + (long)colid * N + tid] = make_double2(0.0, 0.0);
    return;
  }
  const int cbeg = bcol ? 0 : part * nchunk_all / nsplit;
  const int nchunk = bcol ? nchunk_all : (part + 1) * nchunk_all / nsplit - cbeg;   // xi_x planes visited by this CTA
  auto chunk_ex = [&](int c) { return half_chunk_ex(N, zx, bcol, cbeg + c); };
  // B columns keep only the steps whose x/y phase exponent is not 0 (the rest is qhat_half_leftover_kernel's)
  auto skipped = [&](int ex, int ey) { return bcol && mirror_exy(N, zx, zy, ex, ey) == 0; };
  const int r0 = rb * C::ROWS_W + (active ? rgw : 0) * 4;   // first of this thread's 4 zeta_z rows
  const int c0 = 2 * cp;                                    // first of its 2 xi_z columns

  // thread-constant operand offsets inside a parity-split line [par][z>>1]
  int offw[5];
#pragma unroll
  for (int k = 0; k < 5; k++) {
    int z = r0 - c0 - 1 + N / 2 + k;
    z = ((z % N) + N) % N;
    offw[k] = (z & 1) * HALF + (z >> 1);
  }
  const int offg0 = cp, offg1 = HALF + cp;

  auto issue_chunk = [&](int chunk) {  // one thread: stage the operand planes of the chunk-th visited xi_x
    const int s = chunk & 1;
    const int ex = chunk_ex(chunk);
    int X = zx + N / 2 - ex;
    if (X < 0) X += N; else if (X > N - 1) X -= N;
    double2* dst = planes + (size_t)s * STAGE_ELEMS;
    mbar_arrive_expect_tx(&full[s], STAGE_ELEMS * (uint32_t)sizeof(double2));
    tma_bulk_g2s(dst, xiA + (size_t)ex * PLANE, PLANE * sizeof(double2), &full[s]);
    tma_bulk_g2s(dst + PLANE, dfA + (size_t)X * PLANE, PLANE * sizeof(double2), &full[s]);
    if (NP > 1) {
      tma_bulk_g2s(dst + 2 * PLANE, xiB + (size_t)ex * PLANE, PLANE * sizeof(double2), &full[s]);
      tma_bulk_g2s(dst + 3 * PLANE, dfB + (size_t)X * PLANE, PLANE * sizeof(double2), &full[s]);
    }
  };

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();

  // weight stream: rows zeta = (zx, zy, r0 + j), j = 0..3; this thread's 16 bytes sit at column c0
  const double* wrow = Wh + ((long)colid * N + r0) * n3 + c0;
  const int NIT = nchunk * C::SPC;  // iterations of this warp over (visited xi_x, xi_y)
  double2 wb[DEPTH][4];
  auto load_w = [&](int it, double2 (&dst)[4]) {
    const int ex = chunk_ex(it / C::SPC), ey = ph + (it % C::SPC) * C::PH;
    if (skipped(ex, ey)) return;
    const double* p = wrow + ((long)ex * N + ey) * N;
#pragma unroll
    for (int j = 0; j < 4; j++) dst[j] = ldg_stream_f64x2(p + (long)j * n3);
  };
  if (active) {
#pragma unroll
    for (int d = 0; d < DEPTH; d++)
      if (d < NIT) load_w(d, wb[d]);
  }
  pdl_wait();
  if (tid == 0) {
    issue_chunk(0);
    if (nchunk > 1) issue_chunk(1);
  }

  double2 acc[4];
#pragma unroll
  for (int j = 0; j < 4; j++) acc[j] = make_double2(0.0, 0.0);

  for (int it0 = 0; it0 < NIT; it0 += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; d++) {
      const int it = it0 + d;
      if (it >= NIT) break;   // block-uniform (NIT need not be a multiple of DEPTH)
      const int chunk = it / C::SPC, step = it % C::SPC;
      const int s = chunk & 1;
      if (step == 0) mbar_wait(&full[s], (chunk >> 1) & 1);
      const int ex = chunk_ex(chunk), ey = ph + step * C::PH;
      int Y = zy + N / 2 - ey;
      if (Y < 0) Y += N; else if (Y > N - 1) Y -= N;
      const double2* st = planes + (size_t)s * STAGE_ELEMS;
      if (active) {
        if (!skipped(ex, ey)) {
          double2 p[4][2];
          {
            const double2* gl = st + ey * N;
            const double2* fl = st + PLANE + Y * N;
            const double2 g0 = gl[offg0], g1 = gl[offg1];
            double2 fw[5];
#pragma unroll
            for (int k = 0; k < 5; k++) fw[k] = fl[offw[k]];
#pragma unroll
            for (int j = 0; j < 4; j++) {
              p[j][0] = cmul(g0, fw[j + 1]);
              p[j][1] = cmul(g1, fw[j]);
            }
          }
          if (NP > 1) {
            const double2* gl = st + 2 * PLANE + ey * N;
            const double2* fl = st + 3 * PLANE + Y * N;
            const double2 g0 = gl[offg0], g1 = gl[offg1];
            double2 fw[5];
#pragma unroll
            for (int k = 0; k < 5; k++) fw[k] = fl[offw[k]];
#pragma unroll
            for (int j = 0; j < 4; j++) {   // p += g * fw as four fused multiply-adds per product
              p[j][0].x = fma(g0.x, fw[j + 1].x, fma(-g0.y, fw[j + 1].y, p[j][0].x));
              p[j][0].y = fma(g0.x, fw[j + 1].y, fma(g0.y, fw[j + 1].x, p[j][0].y));
              p[j][1].x = fma(g1.x, fw[j].x, fma(-g1.y, fw[j].y, p[j][1].x));
              p[j][1].y = fma(g1.x, fw[j].y, fma(g1.y, fw[j].x, p[j][1].y));
            }
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {
            cmac(acc[j], wb[d][j].x, p[j][0]);
            cmac(acc[j], wb[d][j].y, p[j][1]);
          }
        }
        if (it + DEPTH < NIT) load_w(it + DEPTH, wb[d]);
      }
      if (step == C::SPC - 1) {
        // every warp is done with stage s: refill it with the planes of chunk + 2
        __syncthreads();
        if (tid == 0 && chunk + 2 < nchunk) issue_chunk(chunk + 2);
      }
    }
  }

  // deterministic cross-thread reduction: [thread][4] partial sums -> N rows (reuses the plane ring)
  double2* red = planes;
#pragma unroll
  for (int j = 0; j < 4; j++) red[tid * 4 + j] = active ? acc[j] : make_double2(0.0, 0.0);
  __syncthreads();
  if (tid < N) {
    const int r = tid;
    const int rbr = r / C::ROWS_W, rg = (r % C::ROWS_W) / 4, j = r % 4;
    double sr = 0.0, si = 0.0;
    for (int p = 0; p < C::PH; p++) {
      const int w = p * C::RB + rbr;
      for (int q = 0; q < HALF; q++) {
        const double2 v = red[((w * 32) + rg * HALF + q) * 4 + j];
        sr += v.x; si += v.y;
      }
    }
    qhat[(long)part * n3 + (long)colid * N + r] = make_double2(sr, si);
  }
}

// What the B columns skipped above: on every step with mirror_exy == 0 the row zeta_z = 0 (all xi_z) and, for the other
// rows, the entries xi_z = 0 and (zeta - xi)_z = 0.  One CTA per zeta column (A / unpaired columns write zeros), eight
// warps taking the steps round robin, lane = zeta_z row; the zeta_z = 0 row is spread over the lanes (lane = xi_z).
// Operands are read from the parity-split spectrum in global memory (L2-resident: 16 N^3 bytes).
// PACKED: the weights come from the compact tensor written once by half_pack_leftover_kernel,
//   Wl[column][k][0][l] = Wh[row 0][xi_z = l],  [1][l] = Wh[row l][xi_z = 0],  [2][l] = Wh[row l][(zeta - xi)_z = 0]
// (k = ordinal of the folded step, kmax steps per column), read as 3 N contiguous doubles per step instead of two
// 32-byte sectors per row.
__host__ __device__ constexpr int half_kmax(int N) { return (N / 2 + 1) * N; }

template <int N, bool PACKED, int NP>
__global__ void __launch_bounds__(256)
qhat_half_leftover_kernel(const double* __restrict__ Wh, const double2* __restrict__ xiA, const double2* __restrict__ dfA,
                          const double2* __restrict__ xiB, const double2* __restrict__ dfB, double2* __restrict__ out) {
  constexpr int HALF = N / 2;
  constexpr long n3 = (long)N * N * N;
  SBTE_DYN_SMEM(smraw);   // 8 warps x {rows >= 1, row 0} x N partial sums
  double2 (*red)[2][N] = reinterpret_cast<double2 (*)[2][N]>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int colid = blockIdx.x, zx = colid / N, zy = colid % N;
  const bool bcol = mirror_paired_column(N, zx, zy) && mirror_is_b_row(N, zx, zy);
  if (!bcol) {
    if (tid < N) out[(long)colid * N + tid] = make_double2(0.0, 0.0);
    return;
  }
  auto par = [](int z) { return (z & 1) * HALF + (z >> 1); };   // position of element z in a parity-split line
  const int nchunk = sym_nrep(N, (N - zx) % N);
  double2 acc = make_double2(0.0, 0.0);    // lane = row zeta_z (rows >= 1: two entries per step)
  double2 acc0 = make_double2(0.0, 0.0);   // lane = xi_z of the row zeta_z = 0
  if (lane < N) {
    const int r = lane;
    int d = r + N / 2;                     // (zeta - xi)_z at xi_z = 0; also the xi_z with (zeta - xi)_z = 0
    if (d > N - 1) d -= N;
    int d0 = N / 2 - lane;                 // row 0: (zeta - xi)_z at xi_z = lane
    if (d0 < 0) d0 += N;
    const double* wr = Wh + ((long)colid * N + r) * n3;
    const double* w0 = Wh + ((long)colid * N) * n3;
    const double* wl = Wh + (long)colid * half_kmax(N) * 3 * N;   // PACKED: Wh is the compact tensor
    int k = 0;
    for (int c = 0; c < nchunk; c++) {
      const int ex = half_chunk_ex(N, zx, true, c);
      int X = zx + N / 2 - ex;
      if (X < 0) X += N; else if (X > N - 1) X -= N;
      for (int ey = 0; ey < N; ey++) {
        if (mirror_exy(N, zx, zy, ex, ey) != 0) continue;
        const int kk = k++;
        if ((kk & 7) != warp) continue;
        int Y = zy + N / 2 - ey;
        if (Y < 0) Y += N; else if (Y > N - 1) Y -= N;
        const long lg = ((long)ex * N + ey) * N, lf = ((long)X * N + Y) * N;
        // product of the entry (xi_z = a, (zeta - xi)_z = b), summed over the operand pairs
        auto prod = [&](int a, int b) {
          double2 p = cmul(xiA[lg + par(a)], dfA[lf + par(b)]);
          if (NP > 1) {
            const double2 q = cmul(xiB[lg + par(a)], dfB[lf + par(b)]);
            p.x += q.x; p.y += q.y;
          }
          return p;
        };
        const long step = ((long)ex * N + ey) * N;
        const double* ws = wl + (long)kk * 3 * N;
        // row 0, entry xi_z = lane
        cmac(acc0, PACKED ? ws[lane] : w0[step + lane], prod(lane, d0));
        if (r >= 1) {
          cmac(acc, PACKED ? ws[N + lane] : wr[step], prod(0, d));                        // xi_z = 0
          if (d != 0) cmac(acc, PACKED ? ws[2 * N + lane] : wr[step + d], prod(d, 0));    // (zeta - xi)_z = 0
        }
      }
    }
    red[warp][0][lane] = acc;
    red[warp][1][lane] = acc0;
  }
  __syncthreads();
  if (tid < N) {
    double sr = 0.0, si = 0.0;
    if (tid == 0) {
      for (int w = 0; w < 8; w++)
        for (int l = 0; l < N; l++) { sr += red[w][1][l].x; si += red[w][1][l].y; }
    } else {
      for (int w = 0; w < 8; w++) { sr += red[w][0][tid].x; si += red[w][0][tid].y; }
    }
    out[(long)colid * N + tid] = make_double2(sr, si);
  }
}

// writes the compact leftover tensor (same step enumeration as the leftover kernel)
template <int N>
__global__ void __launch_bounds__(256)
half_pack_leftover_kernel(const double* __restrict__ Wh, double* __restrict__ Wl) {
  constexpr long n3 = (long)N * N * N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int colid = blockIdx.x, zx = colid / N, zy = colid % N;
  if (!(mirror_paired_column(N, zx, zy) && mirror_is_b_row(N, zx, zy)) || lane >= N) return;
  int d = lane + N / 2;
  if (d > N - 1) d -= N;
  const double* wr = Wh + ((long)colid * N + lane) * n3;
  const double* w0 = Wh + ((long)colid * N) * n3;
  double* wl = Wl + (long)colid * half_kmax(N) * 3 * N;
  const int nchunk = sym_nrep(N, (N - zx) % N);
  int k = 0;
  for (int c = 0; c < nchunk; c++) {
    const int ex = half_chunk_ex(N, zx, true, c);
    for (int ey = 0; ey < N; ey++) {
      if (mirror_exy(N, zx, zy, ex, ey) != 0) continue;
      const int kk = k++;
      if ((kk & 7) != warp) continue;
      const long step = ((long)ex * N + ey) * N;
      double* ws = wl + (long)kk * 3 * N;
      ws[lane] = w0[step + lane];
      ws[N + lane] = wr[step];
      ws[2 * N + lane] = wr[step + d];
    }
  }
}

#ifndef SBTE_HOST_EMUL   // host launch code: not part of the host emulation
size_t qhat_half_leftover_doubles(int N) { return (size_t)N * N * half_kmax(N) * 3 * N; }

void launch_half_pack_leftover(sbte_ctx* c, const double* Wh, double* Wl) {
  if (c->N == 16) half_pack_leftover_kernel<16><<<16 * 16, 256, 0, c->stream>>>(Wh, Wl);
  else if (c->N == 32) half_pack_leftover_kernel<32><<<32 * 32, 256, 0, c->stream>>>(Wh, Wl);
  else set_error("half_pack_leftover: unsupported N");
  c->launches += 1;
}

// Wl: compact leftover tensor or null (the leftover kernel then gathers from Wh)
template <int N, int NP>
static void launch_half_n(sbte_ctx* c, const double* Wh, const double* Wl, const QhatPair* pr, double2* qhat, int nsplit) {
  using C = HalfCfg<N, 8 * NP>;
  auto kern = qhat_stream_half_kernel<N, NP>;
  static std::atomic<unsigned> configured{0};   // per device: function attributes belong to the device context
  if (!((configured.load() >> c->device) & 1u)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    configured.fetch_or(1u << c->device);
  }
  k2_mark(c);
  const double2* xB = NP > 1 ? pr[1].xi_side : nullptr;
  const double2* dB = NP > 1 ? pr[1].dif_side : nullptr;
  double2* left = qhat + (size_t)nsplit * c->n3;
  kern<<<N * N * nsplit, C::THREADS, C::SMEM, c->stream>>>(Wh, pr[0].xi_side, pr[0].dif_side, xB, dB, qhat, nsplit);
  if (Wl) qhat_half_leftover_kernel<N, true, NP><<<N * N, 256, 8 * 2 * N * sizeof(double2), c->stream>>>(Wl, pr[0].xi_side, pr[0].dif_side, xB, dB, left);
  else qhat_half_leftover_kernel<N, false, NP><<<N * N, 256, 8 * 2 * N * sizeof(double2), c->stream>>>(Wh, pr[0].xi_side, pr[0].dif_side, xB, dB, left);
  k2_mark(c);
  c->launches += 2;
}

// pairs: parity-split operand spectra (one pair: ComputeQ(f, f); two: ComputeQ_maxPreserve); qhat: nsplit + 1 partial
// spectra of n3 elements each (the last one = leftovers)
void launch_qhat_stream_half(sbte_ctx* c, const double* Wh, const double* Wl, int npairs, const QhatPair* pairs, double2* qhat,
                             int nsplit) {
  const int key = c->N * 10 + npairs;
  switch (key) {
    case 161: launch_half_n<16, 1>(c, Wh, Wl, pairs, qhat, nsplit); break;
    case 162: launch_half_n<16, 2>(c, Wh, Wl, pairs, qhat, nsplit); break;
    case 321: launch_half_n<32, 1>(c, Wh, Wl, pairs, qhat, nsplit); break;
    case 322: launch_half_n<32, 2>(c, Wh, Wl, pairs, qhat, nsplit); break;
    default: set_error("qhat_stream_half: unsupported N / number of operand pairs"); break;
  }
}
#endif

}  // namespace sbte
