// tests/emul/kernel_emul.cpp -- TEST INFRASTRUCTURE ONLY (built by tests/test_kernel_emulation_cpu.py with g++,
// never linked into libsbte_b200.so).  Compiles the batched convolution kernels of spectralbte_b200/csrc verbatim
// through the host shim cuda_emul.h and runs them CTA by CTA, one OS thread per CUDA thread, on a stream-K
// schedule obtained from the library (sbte_batch_schedule_host): producer/consumer barrier protocol, TMA
// coordinates, shared-memory layout, tile switches and partial-sum flushes of the real kernel source.
#include "cuda_emul.h"

#include <type_traits>

#include "../../spectralbte_b200/csrc/qhat_batch.cu"

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
template <int N>
void run_batch3(const Args& a) {
  using C = Batch3Cfg<N>;
  const CUtensorMap tm = fake_map(a.W, (long long)N * N * N, N, C::COLS * N);
  for (int p = 0; p < a.P; p++)
    emul::run_cta(p, a.P, C::THREADS, C::SMEM,
                  [&](int) { qhat_batch3_kernel<N, false>(tm, a.spec, a.parts, a.stride, a.cells, a.sch, 0); });
}
// the instance for schedules cut at whole chunks only (what the library launches at N = 16 when no range ends inside a chunk)
template <int N>
void run_batch3_whole(const Args& a) {
  using C = Batch3Cfg<N>;
  const CUtensorMap tm = fake_map(a.W, (long long)N * N * N, N, C::COLS * N);
  for (int p = 0; p < a.P; p++)
    emul::run_cta(p, a.P, C::THREADS, C::SMEM,
                  [&](int) { qhat_batch3_kernel<N, false, false>(tm, a.spec, a.parts, a.stride, a.cells, a.sch, 0); });
}
// split tiles: the remainder group `cg_base` (at most 16 live cells), two zeta_y columns per warp
template <int N>
void run_batch3_split(const Args& a, int cg_base) {
  using C = Batch3Cfg<N, true>;
  const CUtensorMap tm = fake_map(a.W, (long long)N * N * N, N, C::COLS * N);
  for (int p = 0; p < a.P; p++)
    emul::run_cta(p, a.P, C::THREADS, C::SMEM,
                  [&](int) { qhat_batch3_kernel<N, true>(tm, a.spec, a.parts, a.stride, a.cells, a.sch, cg_base); });
}
}  // namespace

extern "C" {

// kind: 0 = qhat_batch2 (N = 8, 16) / qhat_batch3 (N = 20, 22, 24); 1 = qhat_batch3<16>; 2 = qhat_batch3<16, whole-chunk cuts only>;
// 100 + g = qhat_batch3<16, split> on group g.
// Schedule tables as returned by sbte_batch_schedule_host (host pointers); W = the tensor the kernel streams
// (plain or symmetrised, matching `sym`); spec = cell-minor spectra [G][n3][32] complex;
// parts = kmax * stride complex, pre-filled by the caller.
int emul_batched(int kind, int N, int cells, int sym, int P, const long long* cta_begin, const long long* tile_begin,
                 const int* cta_tile, const int* tile_first, const unsigned char* np, int G, int T, int np_cols, int kmax,
                 const double* W, const double* spec, double* parts, double L_eta, double L_v) {
  Args a;
  a.N = N; a.cells = cells; a.P = P;
  a.sch = {cta_begin, tile_begin, cta_tile, tile_first, np, G, T, P, np_cols, kmax, sym};
  a.W = W; a.spec = (const double2*)spec; a.parts = (double2*)parts;
  // parts are laid out for ALL cell groups of the batch; a split launch (kind = 100 + g) serves the last group g only
  a.stride = (size_t)(kind >= 100 ? kind - 100 + 1 : G) * 32 * (size_t)N * N * N;
  a.L_eta = L_eta; a.L_v = L_v;
  // hand-over buffers of cut chunks (BatchSched::carry), as ensure_batch_schedule() allocates them
  std::vector<double2> carry((size_t)P * kBatchWarps * N * 32);
  std::vector<int> carry_flag((size_t)P * kBatchWarps, 0);
  a.sch.carry = carry.data();
  a.sch.carry_flag = carry_flag.data();
  if (kind == 0) {
    if (N == 8) run_batch2<8>(a);
    else if (N == 16) run_batch2<16>(a);
    else if (N == 20) run_batch3<20>(a);
    else if (N == 22) run_batch3<22>(a);
    else if (N == 24) run_batch3<24>(a);
    else return 1;
  } else if (kind == 1) {   // N = 16 on the line-ring kernel (the library's default for N = 16)
    if (N == 16) run_batch3<16>(a);
    else return 1;
  } else if (kind == 2) {   // N = 16, the whole-chunk instance of the line-ring kernel (schedule built with whole-chunk cuts)
    for (int p = 0; p <= P; p++)
      if (cta_begin[p] % N != 0) return 3;
    if (N == 16) run_batch3_whole<16>(a);
    else return 1;
  } else if (kind >= 100) {   // N = 16 split tiles; kind - 100 = the cell group they serve (schedule built with split = 1)
    if (N == 16) run_batch3_split<16>(a, kind - 100);
    else return 1;
  } else {
    return 1;
  }
  for (int f : carry_flag)
    if (f != 0) return 2;   // every published running sum must have been taken (and its word re-armed)
  return 0;
}

}  // extern "C"
