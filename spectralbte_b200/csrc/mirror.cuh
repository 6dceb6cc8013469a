// spectralbte_b200/csrc/mirror.cuh -- "mirror-paired" batched convolution for f == g with REAL f.
//
// The reference's spectrum of a real distribution function (src/collisions.c:232-283) satisfies, with
// nu(i) = (N - i) mod N per dimension and z(idx) = number of zero components of idx,
//     f^[nu(idx)] = theta^z(idx) conj(f^[idx]),      theta = exp(-2i L_eta L_v)
// (eta_0 = -L_eta has no mirror node on the grid; the transform is quasi-periodic there), and the convolution
// index sigma_zeta(xi) = wrap(zeta + N/2 - xi) (src/collisions.c:141-160) commutes with nu.  Hence row nu(zeta) of
//     Q^[zeta] = sum_xi W[zeta][xi] f^[xi] f^[sigma_zeta(xi)]                         (src/collisions.c:127-165)
// needs, at nu(xi), the complex conjugate of the very product row zeta forms at xi, times theta^(z(xi)+z(sigma)):
//     Q^[nu(zeta)] = sum_xi W[nu(zeta)][nu(xi)] theta^(z(xi) + z(sigma_zeta(xi))) conj(f^[xi] f^[sigma_zeta(xi)]).
// Two weights share one complex product: 8 instead of 12 FP64 instructions wherever the phase is 1.  Valid for
// ARBITRARY real weights (tests/test_hermitian_sharing_cpu.py, tests/test_mirror_emulation_cpu.py).
//
// This header holds what the kernel (qhat_mirror.cu) and its CPU emulation (tests/emul/mirror_emul.cu, test
// infrastructure only) share: the column pairing and the per-step arithmetic of one lane.
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace sbte {

// ---------------------------------------------------------------- column pairing (host)
// Column (zx, zy) pairs with (nu(zx), nu(zy)).  "A" columns own the products: planes zx in [1, N/2) (paired with
// plane N - zx), and in the self-mirrored planes zx in {0, N/2} the columns zy in [1, N/2); the four columns with
// zx, zy in {0, N/2} are their own mirror and run unpaired (zyB = -1).  Tiles hold PAIRS column pairs of one plane.
struct MirrorTile {
  int zx;          // plane of the A columns (the B columns lie in plane nu(zx))
  int zyA[4];      // A columns, -1 = empty slot
  int zyB[4];      // mirror columns nu(zyA) in plane nu(zx), -1 = none (self-mirrored or empty)
};

inline int mirror_nu(int i, int N) { return (N - i) % N; }

inline std::vector<MirrorTile> build_mirror_tiles(int N, int pairs) {
  std::vector<MirrorTile> tiles;
  auto push = [&](int zx, const std::vector<int>& cols, bool paired) {
    for (size_t i = 0; i < cols.size(); i += pairs) {
      MirrorTile t;
      t.zx = zx;
      for (int p = 0; p < 4; p++) { t.zyA[p] = -1; t.zyB[p] = -1; }
      for (int p = 0; p < pairs && i + p < cols.size(); p++) {
        t.zyA[p] = cols[i + p];
        t.zyB[p] = paired ? mirror_nu(cols[i + p], N) : -1;
      }
      tiles.push_back(t);
    }
  };
  for (int zx = 0; zx <= N / 2; zx++) {
    std::vector<int> cols;
    if (zx == 0 || zx == N / 2) {
      for (int zy = 1; zy < N / 2; zy++) cols.push_back(zy);
      push(zx, cols, true);
      push(zx, {0}, false);        // tiles hold CONSECUTIVE A columns (the line-ring variant slides a window over them)
      push(zx, {N / 2}, false);
    } else {
      for (int zy = 0; zy < N; zy++) cols.push_back(zy);
      push(zx, cols, true);
    }
  }
  return tiles;
}

// role of a zeta row in the symmetrised tensor of the mirror kernel: B rows enumerate the mirrored planes
__host__ __device__ __forceinline__ bool mirror_is_b_row(int N, int zx, int zy) {
  if (zx == 0 || zx == N / 2) return zy > N / 2;
  return zx > N / 2;
}

// Symmetrised tensor of the mirror kernel (cf. symmetrize_weights_kernel, qhat.cu): A rows (and the unpaired ones)
// keep the rule of common.cuh -- the sum W + W o sigma on the smaller plane of each pair, W on self-paired planes,
// 0 elsewhere; B rows apply the same rule to the MIRRORED planes, so that the step (xi_x, xi_y) of an A column and
// the step (nu xi_x, nu xi_y) of its B column are representatives together.
__host__ __device__ __forceinline__ double mirror_sym_weight(const double* __restrict__ W, int N, size_t zeta, size_t xi) {
  const size_t n3 = (size_t)N * N * N;
  const int zx = (int)(zeta / ((size_t)N * N)), zy = (int)((zeta / N) % N), zz = (int)(zeta % N);
  const int ex = (int)(xi / ((size_t)N * N)), ey = (int)((xi / N) % N), ez = (int)(xi % N);
  const int X = (zx + N / 2 - ex + N) % N, Y = (zy + N / 2 - ey + N) % N, Z = (zz + N / 2 - ez + N) % N;
  const bool b = mirror_is_b_row(N, zx, zy);
  const int e1 = b ? (N - ex) % N : ex, X1 = b ? (N - X) % N : X;
  if (e1 < X1) return W[zeta * n3 + xi] + W[zeta * n3 + ((size_t)X * N + Y) * N + Z];
  if (e1 == X1) return W[zeta * n3 + xi];
  return 0.0;
}

// ---------------------------------------------------------------- folding mirror rows into their partners
// Q = Re(fft3D^-1(Q^)) (src/collisions.c:212-221) only needs the Hermitian part of Q^: the entry (nu zeta, nu xi) of a
// mirror ("B") row may be added to the entry (zeta, xi) of its "A" row with the factor rho = omega[nu zeta] / omega[zeta]
// (omega = trapezoid weights of the inverse transform) wherever the phase exponent
//     e = z(zeta) - z(xi) - z(zeta - xi)      (z = number of zero index components)
// is 0 -- the folded weight stays real and Q is unchanged (tests/test_half_spectrum_cpu.py).  Folded here: the entries of
// the steps whose x/y part of e is 0 ("foldable steps") with regular z components (zeta_z, xi_z, (zeta - xi)_z != 0);
// every other entry stays in its own row.  What the kernels then form is no longer the reference's Q^ but a spectrum with
// the same real inverse transform, so folding is only used on the way to Q.
__host__ __device__ __forceinline__ int mirror_exy(int N, int zx, int zy, int ex, int ey) {
  const int X = (zx + N / 2 - ex + N) % N, Y = (zy + N / 2 - ey + N) % N;
  return (zx == 0) + (zy == 0) - (ex == 0) - (X == 0) - (ey == 0) - (Y == 0);
}
__host__ __device__ __forceinline__ bool mirror_paired_column(int N, int zx, int zy) {
  return !((zx == 0 || zx == N / 2) && (zy == 0 || zy == N / 2));
}
__host__ __device__ __forceinline__ double mirror_trap(int N, int i) { return (i == 0 || i == N - 1) ? 0.5 : 1.0; }

// Folded tensor of the mirror kernels: sym = 1 starts from the symmetrised weights (mirror_sym_weight), 0 from W.
__host__ __device__ __forceinline__ double mirror_fold_weight(const double* __restrict__ W, int N, size_t zeta, size_t xi, bool sym) {
  const size_t n3 = (size_t)N * N * N;
  const int zx = (int)(zeta / ((size_t)N * N)), zy = (int)((zeta / N) % N), zz = (int)(zeta % N);
  const int ex = (int)(xi / ((size_t)N * N)), ey = (int)((xi / N) % N), ez = (int)(xi % N);
  const double own = sym ? mirror_sym_weight(W, N, zeta, xi) : W[zeta * n3 + xi];
  if (!mirror_paired_column(N, zx, zy)) return own;
  const bool b = mirror_is_b_row(N, zx, zy);
  // the "A" member of this pair of entries decides
  const int azx = b ? (N - zx) % N : zx, azy = b ? (N - zy) % N : zy, azz = b ? (N - zz) % N : zz;
  const int aex = b ? (N - ex) % N : ex, aey = b ? (N - ey) % N : ey, aez = b ? (N - ez) % N : ez;
  const int aZ = (azz + N / 2 - aez + N) % N;
  const bool fold = mirror_exy(N, azx, azy, aex, aey) == 0 && azz != 0 && aez != 0 && aZ != 0;
  if (!fold) return own;
  if (b) return 0.0;
  const size_t zetaB = ((size_t)((N - zx) % N) * N + (N - zy) % N) * N + (N - zz) % N;
  const size_t xiB = ((size_t)((N - ex) % N) * N + (N - ey) % N) * N + (N - ez) % N;
  const double rho = mirror_trap(N, (N - zx) % N) * mirror_trap(N, (N - zy) % N) * mirror_trap(N, (N - zz) % N) /
                     (mirror_trap(N, zx) * mirror_trap(N, zy) * mirror_trap(N, zz));
  return own + rho * (sym ? mirror_sym_weight(W, N, zetaB, xiB) : W[zetaB * n3 + xiB]);
}

// ---------------------------------------------------------------- per-lane arithmetic of one (xi_x, xi_y) step
__host__ __device__ __forceinline__ double2 mir_cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ double2 mir_cmul_conj_b(double2 a, double2 b) {   // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ __forceinline__ double2 mir_conj(double2 a) { return make_double2(a.x, -a.y); }

// Rows R0 .. R0+RH-1 of column A accumulate into accA; the mirrored rows nu(R) of column B into accB (index r).
//   fl/fs : the (zeta - xi)-side line f^[X][Y][.] of column A, element z at fl[z * fs]
//   gl/gs : the xi-side line f^[xi_x][xi_y][.], element c at gl[c * gs]
//   wA    : N x N weights of column A, row zeta_z, column xi_z          (W[zeta_A][(xi_x, xi_y, .)])
//   wB    : N x N weights of column B at the MIRRORED step              (W[zeta_B][(nu xi_x, nu xi_y, .)]); zeros if unpaired
//   R0    : first row of this warp, 0 or N - RH (a RUN-TIME value: both halves of a column execute the same code,
//           which keeps the unrolled body -- and the instruction footprint -- to one copy)
//   theta : exp(-2i L_eta L_v)
// The step-level phase theta^m, m = [xi_x=0] + [xi_y=0] + [X=0] + [Y=0], is NOT applied here: accB is kept in a frame
// rotated by the current step's phase (accB = true value * theta^-m), which the caller changes with
// mirror_frame_update() on the few steps where m changes; that way the mirror rows accumulate directly.
// The main loop treats every entry as phase-free (mirror row += w' conj(p)); the entries that do carry a phase --
// the column xi_z = 0 and the diagonal (zeta - xi)_z = 0, 2 RH of them -- are corrected afterwards by
// w' (theta^k - 1) conj(p), k = 1 or 2.
// LAZY: the (zeta - xi)-side operands are loaded when the sliding window of the rows first needs them (RH + 1 live
// values instead of N: the N >= 20 kernels do not have the registers for the whole line).
// COMBINED: the step is foldable and the tensor is the folded one (mirror_fold_weight): the mirror rows' phase-free
// entries already sit in wA, so the main loop only serves the A rows (6 instead of 8 FP64 instructions per weight pair)
// -- except for the first row of the warp (zeta_z = 0 is never folded; for the other half its wB entries are zeros) --
// and the entries with xi_z = 0 or (zeta - xi)_z = 0 are added in full instead of being corrected.
template <int N, int RH, bool LAZY = false, bool COMBINED = false>
__host__ __device__ __forceinline__ void mirror_step(double2* accA, double2* accB, const double2* fl, int fs,
                                                      const double2* gl, int gs, const double* wA, const double* wB,
                                                      int R0, double2 theta) {
  // slot i of fr holds f^[(i + R0) mod N]: row r, column c reads slot (r + N/2 - c) mod N in either half
  double2 fr[N];
  auto f_at = [&](int slot) {
    int z = slot + R0;
    if (z > N - 1) z -= N;
    return fl[z * fs];
  };
  if (!LAZY) {
#pragma unroll
    for (int z = 0; z < N; z++) fr[z] = f_at(z);
  } else {
    // window of the first column pair: slots N/2 - 1 .. N/2 + RH - 1 (mod N)
#pragma unroll
    for (int j = 0; j <= RH; j++) fr[(N / 2 - 1 + j) % N] = f_at((N / 2 - 1 + j) % N);
  }
  const double* wa_rows = wA + R0 * N;
  // mirrored rows: nu(R0 + r) = (N - R0 - r) mod N, i.e. base - r with base = N (first half; row 0 stays row 0) or N - R0
  const double* wb_top = wB + ((R0 == 0) ? N : N - R0) * N;
  const double* wb_row0 = (R0 == 0) ? wB : wb_top;
#pragma unroll
  for (int c = 0; c < N; c += 2) {
    const double2 g0v = gl[c * gs], g1v = gl[(c + 1) * gs];
    if (LAZY && c > 0) {   // columns c, c + 1 reach two operands further down
      fr[(N / 2 - c + 2 * N) % N] = f_at((N / 2 - c + 2 * N) % N);
      fr[(N / 2 - c - 1 + 2 * N) % N] = f_at((N / 2 - c - 1 + 2 * N) % N);
    }
#pragma unroll
    for (int r = 0; r < RH; r++) {
      const int s0 = (r + N / 2 - c + N) % N;          // operand slot of column c
      const int s1 = (r + N / 2 - c - 1 + N) % N;      // ... of column c + 1
      const int nc0 = (N - c) % N, nc1 = N - c - 1;    // mirrored columns
      const double2 wa = *reinterpret_cast<const double2*>(wa_rows + r * N + c);
      const double* wbr = (r == 0) ? wb_row0 : wb_top - r * N;
      const double wb0 = (!COMBINED || r == 0) ? wbr[nc0] : 0.0, wb1 = (!COMBINED || r == 0) ? wbr[nc1] : 0.0;
      const double2 p0 = mir_cmul(g0v, fr[s0]);
      const double2 p1 = mir_cmul(g1v, fr[s1]);
      accA[r].x = fma(wa.x, p0.x, accA[r].x);
      accA[r].y = fma(wa.x, p0.y, accA[r].y);
      accA[r].x = fma(wa.y, p1.x, accA[r].x);
      accA[r].y = fma(wa.y, p1.y, accA[r].y);
      if (!COMBINED || r == 0) {
        accB[r].x = fma(wb0, p0.x, accB[r].x);
        accB[r].y = fma(-wb0, p0.y, accB[r].y);
        accB[r].x = fma(wb1, p1.x, accB[r].x);
        accB[r].y = fma(-wb1, p1.y, accB[r].y);
      }
    }
  }
  // the entries of the mirror rows with xi_z = 0 and / or (zeta - xi)_z = 0 carry theta^k, k = 1, 2: corrected by
  // (theta^k - 1) where the main loop has added them phase-free, added in full (theta^k) where it has not
  const double2 th2 = mir_cmul(theta, theta);
  const double2 g0 = gl[0], f0 = fl[0];
  const double2 gE1 = mir_cmul_conj_b(make_double2(theta.x - 1.0, theta.y), g0);   // (theta^k - 1) conj(g^[.. 0])
  const double2 gE2 = mir_cmul_conj_b(make_double2(th2.x - 1.0, th2.y), g0);
  const double2 fE1 = mir_cmul_conj_b(make_double2(theta.x - 1.0, theta.y), f0);   // (theta - 1) conj(f^[.. 0])
  const double2 gF1 = mir_cmul_conj_b(theta, g0), gF2 = mir_cmul_conj_b(th2, g0);  // theta^k conj(g^[.. 0])
  const double2 fF1 = mir_cmul_conj_b(theta, f0);
#pragma unroll
  for (int r = 0; r < RH; r++) {
    const bool full = COMBINED && r != 0;       // the main loop skipped this row's mirror entries
    const double* wbr = (r == 0) ? wb_row0 : wb_top - r * N;
    int d = R0 + r + N / 2;                     // (zeta - xi)_z of the entry in column xi_z = 0; also the column of the
    if (d > N - 1) d -= N;                      // row's entry with (zeta - xi)_z = 0
    // column xi_z = 0 (mirrored column 0): w' (theta^k - 1) conj(g_0) conj(f_d), k = 1 + [d = 0]
    {
      const double2 ge = full ? ((d == 0) ? gF2 : gF1) : ((d == 0) ? gE2 : gE1);
      const double2 q = mir_cmul_conj_b(ge, fl[d * fs]);
      const double wb = wbr[0];
      accB[r].x = fma(wb, q.x, accB[r].x);
      accB[r].y = fma(wb, q.y, accB[r].y);
    }
    // diagonal (zeta - xi)_z = 0 at column xi_z = d (d = 0 was handled above): w' (theta - 1) conj(g_d) conj(f_0)
    if (d != 0) {
      const double2 q = mir_cmul_conj_b(full ? fF1 : fE1, gl[d * gs]);
      const double wb = wbr[N - d];
      accB[r].x = fma(wb, q.x, accB[r].x);
      accB[r].y = fma(wb, q.y, accB[r].y);
    }
  }
}

// theta^m for m = 0..4 (the step-level phases), by value in kernel parameters
struct MirrorPhases {
  double2 t[5];
};
__host__ __device__ __forceinline__ double2 mirror_phase(const MirrorPhases& ph, int m) {
  return (m == 0) ? ph.t[0] : (m == 1) ? ph.t[1] : (m == 2) ? ph.t[2] : (m == 3) ? ph.t[3] : ph.t[4];
}
// accB is held as (true value) * theta^-m_cur.  Moves it to the frame of a step with phase theta^m_new (m_new = 0:
// back to the true value, before it is written out).
template <int RH>
__host__ __device__ __forceinline__ void mirror_frame_update(double2* accB, int& m_cur, int m_new, const MirrorPhases& ph) {
  if (m_new == m_cur) return;
  const double2 rot = mir_cmul_conj_b(mirror_phase(ph, m_cur), mirror_phase(ph, m_new));   // theta^(m_cur - m_new)
#pragma unroll
  for (int r = 0; r < RH; r++) accB[r] = mir_cmul(accB[r], rot);
  m_cur = m_new;
}

}  // namespace sbte
