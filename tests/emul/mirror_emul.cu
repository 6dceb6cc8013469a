// tests/emul/mirror_emul.cu -- TEST INFRASTRUCTURE ONLY (never linked into libsbte_b200.so, never on the product
// path): runs the per-lane arithmetic and the column pairing of the mirror-paired batched convolution
// (spectralbte_b200/csrc/mirror.cuh, shared verbatim with the kernel in qhat_mirror.cu) on the HOST for one cell,
// tile by tile and step by step as the kernel does, so that tests/test_mirror_emulation_cpu.py can check the
// indexing and the phase bookkeeping against the oracle without a GPU.  Barriers, TMA and the stream-K split are
// not emulated.
#include <math.h>
#include <string.h>

#include <vector>

#include "../../spectralbte_b200/csrc/common.cuh"
#include "../../spectralbte_b200/csrc/mirror.cuh"

using namespace sbte;

template <int N, bool LAZY>
static void run(const double* W, int sym, int fold, const double2* F, double L_eta, double L_v, double2* qhat) {
  constexpr int PAIRS = (N >= 16) ? 4 : 2, RH = N / 2;   // MirrorCfg / MirrorRingCfg::PAIRS
  const size_t n3 = (size_t)N * N * N;
  const double2 theta = make_double2(cos(-2.0 * L_eta * L_v), sin(-2.0 * L_eta * L_v));
  MirrorPhases ph;
  for (int m = 0; m < 5; m++) ph.t[m] = make_double2(cos(m * -2.0 * L_eta * L_v), sin(m * -2.0 * L_eta * L_v));
  std::vector<double> zero((size_t)N * N, 0.0), wa((size_t)N * N), wb((size_t)N * N);
  for (const MirrorTile& t : build_mirror_tiles(N, PAIRS)) {
    const int zx = t.zx, zxB = mirror_nu(zx, N);
    for (int p = 0; p < PAIRS; p++) {
      if (t.zyA[p] < 0) continue;
      const int zy = t.zyA[p], zyB = t.zyB[p];
      for (int half = 0; half < 2; half++) {
        double2 accA[RH], accB[RH];
        for (int r = 0; r < RH; r++) accA[r] = accB[r] = make_double2(0.0, 0.0);
        int m_cur = 0;   // frame of accB (mirror_frame_update), as the kernels keep it
        const int nrep = sym ? sym_nrep(N, zx) : N;
        for (int cidx = 0; cidx < nrep; cidx++) {
          const int ex = sym ? sym_rep(N, zx, cidx) : cidx;
          int X = zx + N / 2 - ex;
          if (X < 0) X += N; else if (X > N - 1) X -= N;
          for (int ey = 0; ey < N; ey++) {
            int Y = zy + N / 2 - ey;
            if (Y < 0) Y += N; else if (Y > N - 1) Y -= N;
            const double2* fl = F + ((size_t)X * N + Y) * N;
            const double2* gl = F + ((size_t)ex * N + ey) * N;
            // the two TMA boxes of this column pair: N rows x N columns, compact
            for (int r = 0; r < N; r++)
              memcpy(&wa[(size_t)r * N], W + (((size_t)zx * N + zy) * N + r) * n3 + ((size_t)ex * N + ey) * N, N * sizeof(double));
            const double* wbp = zero.data();
            if (zyB >= 0) {
              const int exB = mirror_nu(ex, N), eyB = mirror_nu(ey, N);
              for (int r = 0; r < N; r++)
                memcpy(&wb[(size_t)r * N], W + (((size_t)zxB * N + zyB) * N + r) * n3 + ((size_t)exB * N + eyB) * N,
                       N * sizeof(double));
              wbp = wb.data();
            }
            mirror_frame_update<RH>(accB, m_cur, (ex == 0) + (ey == 0) + (X == 0) + (Y == 0), ph);
            // with the folded tensor the foldable steps of a paired column take the combined body
            if (fold && zyB >= 0 && mirror_exy(N, zx, zy, ex, ey) == 0)
              mirror_step<N, RH, LAZY, true>(accA, accB, fl, 1, gl, 1, wa.data(), wbp, half * (N - RH), theta);
            else
              mirror_step<N, RH, LAZY, false>(accA, accB, fl, 1, gl, 1, wa.data(), wbp, half * (N - RH), theta);
          }
        }
        mirror_frame_update<RH>(accB, m_cur, 0, ph);
        const int R0 = half * (N - RH);
        for (int r = 0; r < RH; r++) {
          qhat[((size_t)zx * N + zy) * N + R0 + r] = accA[r];
          if (zyB >= 0) qhat[((size_t)zxB * N + zyB) * N + mirror_nu(R0 + r, N)] = accB[r];
        }
      }
    }
  }
}

extern "C" {

// Ws2 (mirror rule) from W, both N^3 x N^3 row-major
int mirror_emul_symmetrize(int N, const double* W, double* Ws2) {
  const size_t n3 = (size_t)N * N * N;
  for (size_t zeta = 0; zeta < n3; zeta++)
    for (size_t xi = 0; xi < n3; xi++) Ws2[zeta * n3 + xi] = mirror_sym_weight(W, N, zeta, xi);
  return 0;
}

// folded tensor (mirror_fold_weight) from W, both N^3 x N^3 row-major
int mirror_emul_fold(int N, const double* W, int sym, double* Wh) {
  const size_t n3 = (size_t)N * N * N;
  for (size_t zeta = 0; zeta < n3; zeta++)
    for (size_t xi = 0; xi < n3; xi++) Wh[zeta * n3 + xi] = mirror_fold_weight(W, N, zeta, xi, sym != 0);
  return 0;
}

// Q^ of one cell (fold = 0) or, with the folded tensor (fold = 1), a spectrum with the same real inverse transform;
// W = the tensor the kernel streams (plain / mirror-symmetrised / folded); F, qhat: N^3 complex, natural layout
int mirror_emul_qhat(int N, const double* W, int sym, int fold, const double* F, double L_eta, double L_v, double* qhat) {
  const size_t n3 = (size_t)N * N * N;
  for (size_t i = 0; i < 2 * n3; i++) qhat[i] = NAN;   // every entry must be written exactly by the pairing
  // the operand-loading variant each N uses in the kernels (qhat_mirror.cu)
  if (N == 8) run<8, false>(W, sym, fold, (const double2*)F, L_eta, L_v, (double2*)qhat);
  else if (N == 16) run<16, false>(W, sym, fold, (const double2*)F, L_eta, L_v, (double2*)qhat);
  else if (N == 12) run<12, true>(W, sym, fold, (const double2*)F, L_eta, L_v, (double2*)qhat);   // small stand-in for 20..24
  else if (N == 20) run<20, true>(W, sym, fold, (const double2*)F, L_eta, L_v, (double2*)qhat);
  else if (N == 22) run<22, true>(W, sym, fold, (const double2*)F, L_eta, L_v, (double2*)qhat);
  else if (N == 24) run<24, true>(W, sym, fold, (const double2*)F, L_eta, L_v, (double2*)qhat);
  else return 1;
  return 0;
}

}  // extern "C"
