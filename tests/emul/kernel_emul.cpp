// tests/emul/kernel_emul.cpp -- TEST INFRASTRUCTURE ONLY (built by tests/test_kernel_emulation_cpu.py with g++,
// never linked into libsbte_b200.so).  Compiles the batched convolution kernels of spectralbte_b200/csrc verbatim
// through the host shim cuda_emul.h and runs them CTA by CTA, one OS thread per CUDA thread, on a stream-K
// schedule obtained from the library (sbte_batch_schedule_host): producer/consumer barrier protocol, TMA
// coordinates, shared-memory layout, tile switches and partial-sum flushes of the real kernel source.
#include "cuda_emul.h"

#include <type_traits>

#include "../../spectralbte_b200/csrc/qhat_batch.cu"
#include "../../spectralbte_b200/csrc/qhat_mirror.cu"
#include "../../spectralbte_b200/csrc/qhat_half.cu"

using namespace sbte;

namespace {

CUtensorMap fake_map(const double* W, long long n3, int box_cols, int box_rows) {
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  emul::FakeTensorMap f = {W, n3, n3, box_cols, box_rows};
  static_assert(sizeof(f) <= sizeof(tm), "fake descriptor fits");
  memcpy(&tm, &f, sizeof(f));
  return tm;
}

struct Args {
  int N, cells, P;
  BatchSched sch;
  const double* W;
  const double2* spec;
  double2* parts;
  size_t stride;
  double L_eta, L_v;
};

template <int N>
void run_batch2(const Args& a) {
  using C = Batch2Cfg<N>;
  const CUtensorMap tm = fake_map(a.W, (long long)N * N * N, N, C::COLS * N);
  for (int p = 0; p < a.P; p++)
    emul::run_cta(p, a.P, C::THREADS, C::SMEM,
                  [&](int) { qhat_batch2_kernel<N>(tm, a.spec, a.parts, a.stride, a.cells, a.sch); });
}
template <int N, int ROLL>
void run_batch3(const Args& a) {
  using C = Batch3Cfg<N>;
  const CUtensorMap tm = fake_map(a.W, (long long)N * N * N, N, C::COLS * N);
  for (int p = 0; p < a.P; p++)
    emul::run_cta(p, a.P, C::THREADS, C::SMEM,
                  [&](int) { qhat_batch3_kernel<N, ROLL>(tm, a.spec, a.parts, a.stride, a.cells, a.sch); });
}
MirrorPhases phases(const Args& a) {
  MirrorPhases ph;
  const double ang = -2.0 * a.L_eta * a.L_v;
  for (int m = 0; m < 5; m++) ph.t[m] = make_double2(cos(m * ang), sin(m * ang));
  return ph;
}
template <int N>
void run_mirror(const Args& a, int fold) {
  using C = MirrorCfg<N>;
  const CUtensorMap tm = fake_map(a.W, (long long)N * N * N, N, N);
  const std::vector<MirrorTile> tiles = build_mirror_tiles(N, C::PAIRS);
  const MirrorPhases ph = phases(a);
  for (int p = 0; p < a.P; p++)
    emul::run_cta(p, a.P, C::THREADS, C::SMEM,
                  [&](int) { qhat_mirror_kernel<N>(tm, a.spec, a.parts, a.stride, a.cells, a.sch, tiles.data(), ph, fold); });
}
template <int N>
void run_mirror_ring(const Args& a) {
  using C = MirrorRingCfg<N>;
  const CUtensorMap tm = fake_map(a.W, (long long)N * N * N, N, N);
  const std::vector<MirrorTile> tiles = build_mirror_tiles(N, C::PAIRS);
  const MirrorPhases ph = phases(a);
  for (int p = 0; p < a.P; p++)
    emul::run_cta(p, a.P, C::THREADS, C::SMEM,
                  [&](int) { qhat_mirror_ring_kernel<N>(tm, a.spec, a.parts, a.stride, a.cells, a.sch, tiles.data(), ph); });
}

}  // namespace

extern "C" {

// kind: 0 = default kernels (qhat_batch2 / qhat_batch3), 1 = mirror-paired kernels, 2 = rolled line ring (N = 22, 24),
// 3 = mirror-paired kernel on the folded tensor (the result then only shares Re(fft3D^-1(.)) with Q^).
// Schedule tables as returned by sbte_batch_schedule_host (host pointers); W = the tensor the kernel streams
// (plain, symmetrised or mirror-symmetrised, matching `sym` and `kind`); spec = cell-minor spectra [G][n3][32] complex;
// parts = kmax * stride complex, pre-filled by the caller.
int emul_batched(int kind, int N, int cells, int sym, int P, const long long* cta_begin, const long long* tile_begin,
                 const int* cta_tile, const int* tile_first, const unsigned char* np, int G, int T, int np_cols, int kmax,
                 const double* W, const double* spec, double* parts, double L_eta, double L_v) {
  Args a;
  a.N = N; a.cells = cells; a.P = P;
  a.sch = {cta_begin, tile_begin, cta_tile, tile_first, np, G, T, P, np_cols, kmax, sym};
  a.W = W; a.spec = (const double2*)spec; a.parts = (double2*)parts;
  a.stride = (size_t)G * 32 * (size_t)N * N * N;
  a.L_eta = L_eta; a.L_v = L_v;
  if (kind == 0) {
    if (N == 8) run_batch2<8>(a);
    else if (N == 16) run_batch2<16>(a);
    else if (N == 20) run_batch3<20, 1>(a);
    else if (N == 22) run_batch3<22, 1>(a);
    else if (N == 24) run_batch3<24, 1>(a);
    else return 1;
  } else if (kind == 1) {
    if (N == 8) run_mirror<8>(a, 0);
    else if (N == 16) run_mirror<16>(a, 0);
    else if (N == 20) run_mirror_ring<20>(a);
    else if (N == 22) run_mirror_ring<22>(a);
    else if (N == 24) run_mirror_ring<24>(a);
    else return 1;
  } else if (kind == 2) {
    if (N == 24) run_batch3<24, 3>(a);
    else if (N == 22) run_batch3<22, 11>(a);
    else return 1;
  } else if (kind == 3) {   // folded tensor + combined body on the foldable steps
    if (N == 8) run_mirror<8>(a, 1);
    else if (N == 16) run_mirror<16>(a, 1);
    else return 1;
  } else {
    return 1;
  }
  return 0;
}

// 0D half-spectrum path (qhat_half.cu): Wh = folded tensor (mirror rule, symmetrised); xiA/dfA (and xiB/dfB when npairs = 2:
// ComputeQ_maxPreserve) = parity-split operand spectra of one cell; qhat = (nsplit + 1) partial spectra of n3 complex each
int emul_half0d(int N, int nsplit, int packed, int npairs, const double* Wh, const double* xiA, const double* dfA, const double* xiB,
                const double* dfB, double* qhat) {
  const size_t n3 = (size_t)N * N * N;
  const double2 *xa = (const double2*)xiA, *da = (const double2*)dfA, *xb = (const double2*)xiB, *db = (const double2*)dfB;
  auto run = [&](auto ntag, auto ptag) {
    constexpr int M = decltype(ntag)::value, NP = decltype(ptag)::value;
    using C = HalfCfg<M, 8 * NP>;
    for (int b = 0; b < M * M * nsplit; b++)
      emul::run_cta(b, M * M * nsplit, C::THREADS, C::SMEM,
                    [&](int) { qhat_stream_half_kernel<M, NP>(Wh, xa, da, xb, db, (double2*)qhat, nsplit); });
    std::vector<double> Wl;
    if (packed) {   // compact leftover tensor, poisoned first: every entry the leftover kernel reads must have been packed
      Wl.assign((size_t)M * M * half_kmax(M) * 3 * M, NAN);
      for (int b = 0; b < M * M; b++)
        emul::run_cta(b, M * M, 256, 64, [&](int) { half_pack_leftover_kernel<M>(Wh, Wl.data()); });
    }
    double2* left = (double2*)qhat + (size_t)nsplit * n3;
    for (int b = 0; b < M * M; b++)
      emul::run_cta(b, M * M, 256, 8 * 2 * M * sizeof(double2), [&](int) {
        if (packed) qhat_half_leftover_kernel<M, true, NP>(Wl.data(), xa, da, xb, db, left);
        else qhat_half_leftover_kernel<M, false, NP>(Wh, xa, da, xb, db, left);
      });
  };
  using I16 = std::integral_constant<int, 16>;
  using I1 = std::integral_constant<int, 1>;
  using I2 = std::integral_constant<int, 2>;
  using I32 = std::integral_constant<int, 32>;
  if (N == 16 && npairs == 1) run(I16(), I1());
  else if (N == 16 && npairs == 2) run(I16(), I2());
  else if (N == 32 && npairs == 1) run(I32(), I1());   // 18 GB of tensors: run by hand (tools/emul_half0d_n32.py), not by the suite
  else if (N == 32 && npairs == 2) run(I32(), I2());
  else return 1;
  return 0;
}

}  // extern "C"
