// spectralbte_b200/csrc/qhat_mirror.cu -- K2, batched over spatial cells, mirror-paired (opt-in: SBTE_MIRROR=1).
//
// Same contraction as qhat_batch.cu (the N^6 loop of /root/reference/src/collisions.c:127-165 for f == g, one cell
// per lane), but a warp works on a zeta column AND its mirror column (nu(zeta_x), nu(zeta_y)) at once: for a real
// distribution function the complex product f^[xi] f^[zeta - xi] formed for row zeta is, conjugated, the product row
// nu(zeta) needs at nu(xi) (mirror.cuh), so two weights share one product -- 8 instead of 12 FP64 instructions per
// weight pair wherever no index component is zero (about 27 % fewer FP64 instructions at N = 16).
//
// Structure = qhat_batch2_kernel: persistent CTAs, stream-K over (tile, step), producer warpgroup issuing TMA copies
// into a 4-stage ring (xi-side line + one N x N weight box per column), the (zeta - xi)-side plane resident in shared
// memory, decoupled compute warps.  A tile is PAIRS column pairs of one zeta_x plane; two warps share a pair, each
// owning N/2 zeta_z rows of column A and the mirrored rows of column B.
// STATUS: arithmetic, pairing and symmetrisation rule are checked on the CPU against the reference restatement
// (tests/test_mirror_emulation_cpu.py), and this very source -- barriers, TMA coordinates, shared-memory layout, tile
// switches, flushes -- runs on the host through tests/emul/cuda_emul.h (tests/test_kernel_emulation_cpu.py, N = 8, 16,
// 20, 22, 24).  It has not run on a GPU yet (no timing, no sanitizer pass), hence off by default.
#include <math.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "internal.h"
#include "mirror.cuh"

namespace sbte {

// SBTE_MIRROR=1: N = 8, 16 (plane-resident kernel); SBTE_MIRROR=2: also N = 20, 22, 24 (line-ring kernel);
// SBTE_MIRROR=3: additionally, on the way to Q (not Q^), N = 8, 16 stream the FOLDED tensor (mirror.cuh:
// mirror_fold_weight) and run the combined body on the foldable steps
static int mirror_level() {
  static const int level = getenv("SBTE_MIRROR") ? atoi(getenv("SBTE_MIRROR")) : 0;
  return level;
}
bool qhat_mirror_enabled(int N) {
  if (mirror_level() >= 1 && (N == 8 || N == 16)) return true;
  return mirror_level() >= 2 && (N == 20 || N == 22 || N == 24);
}
bool qhat_mirror_fold_enabled(int N) { return mirror_level() >= 3 && (N == 8 || N == 16); }
int qhat_mirror_pairs(int N) { return (N >= 16) ? 4 : 2; }
int qhat_mirror_align(int N) { return (N >= 20) ? N : 1; }   // the line-ring kernel works on whole xi_x chunks

__global__ void symmetrize_weights_mirror_kernel(const double* __restrict__ W, double* __restrict__ Ws2, int N) {
  const size_t n3 = (size_t)N * N * N, total = n3 * n3;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t zeta = e / n3;
    Ws2[e] = mirror_sym_weight(W, N, zeta, e - zeta * n3);
  }
}

__global__ void fold_weights_mirror_kernel(const double* __restrict__ W, double* __restrict__ Wh, int N, int sym) {
  const size_t n3 = (size_t)N * N * N, total = n3 * n3;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t zeta = e / n3;
    Wh[e] = mirror_fold_weight(W, N, zeta, e - zeta * n3, sym != 0);
  }
}

#ifndef SBTE_HOST_EMUL   // host launch code: not part of the host emulation
void launch_symmetrize_weights_mirror(sbte_ctx* c, const double* W, double* Ws2) {
  symmetrize_weights_mirror_kernel<<<148 * 16, 256, 0, c->stream>>>(W, Ws2, c->N);
  c->launches += 1;
}
void launch_fold_weights_mirror(sbte_ctx* c, const double* W, double* Wh, bool sym) {
  fold_weights_mirror_kernel<<<148 * 16, 256, 0, c->stream>>>(W, Wh, c->N, sym ? 1 : 0);
  c->launches += 1;
}
#endif

template <int N>
struct MirrorCfg {
  static constexpr int PAIRS = (N >= 16) ? 4 : 2;
  static constexpr int CWARPS = 2 * PAIRS;          // two warps per column pair: rows [0, N/2) and [N/2, N)
  static constexpr int RH = N / 2;
  static constexpr int CONSUMERS = CWARPS * 32;
  static constexpr int THREADS = CONSUMERS + 128;   // + one producer warpgroup
  static constexpr bool REG_SPLIT = THREADS > 256;  // 384 threads: 40 (producer) / 232 (compute) registers
  static constexpr int LINE = N * 32;
  static constexpr int PLANE = N * LINE;
  static constexpr int STAGES = 4;
  static constexpr int WTILE = N * N;               // doubles per column box
  static constexpr size_t STAGE_BYTES = (size_t)LINE * 16 + (size_t)2 * PAIRS * WTILE * 8;
  static constexpr size_t SMEM = (size_t)PLANE * 16 + STAGES * STAGE_BYTES + (size_t)WTILE * 8 + 256;
  static_assert(SMEM <= 227 * 1024, "plane + stages + zero box must fit in shared memory");
};

template <int N>
__global__ void __launch_bounds__(MirrorCfg<N>::THREADS, 1)
qhat_mirror_kernel(const __grid_constant__ CUtensorMap tmapW, const double2* __restrict__ spec,
                   double2* __restrict__ parts, size_t part_stride, int cells, BatchSched sch,
                   const MirrorTile* __restrict__ tiles, MirrorPhases ph, int fold) {
  using C = MirrorCfg<N>;
  constexpr long n3 = (long)N * N * N;
  constexpr int S = C::STAGES, RH = C::RH;
  SBTE_DYN_SMEM(smraw);
  double2* plane = reinterpret_cast<double2*>(smraw);
  unsigned char* stage0 = smraw + (size_t)C::PLANE * 16;
  double* zero_box = reinterpret_cast<double*>(stage0 + S * C::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + S * C::STAGE_BYTES + (size_t)C::WTILE * 8);
  uint64_t* full = bars;            // [S]  TMA transaction barriers
  uint64_t* empty = bars + S;       // [S]  one arrival per compute warp
  uint64_t* fullPlane = bars + 2 * S;
  uint64_t* emptyPlane = bars + 2 * S + 1;
  auto stage_line = [&](int s) { return reinterpret_cast<double2*>(stage0 + (size_t)s * C::STAGE_BYTES); };
  auto stage_w = [&](int s) {
    return reinterpret_cast<double*>(stage0 + (size_t)s * C::STAGE_BYTES + (size_t)C::LINE * 16);
  };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long g0 = sch.cta_begin[blockIdx.x];
  const int n = (int)(sch.cta_begin[blockIdx.x + 1] - g0);
  if (n <= 0) return;
  const int G = sch.G;
  const bool sym = sch.sym != 0;

  for (int i = tid; i < C::WTILE; i += blockDim.x) zero_box[i] = 0.0;
  if (tid == 0) {
    for (int b = 0; b < S; b++) { mbar_init(&full[b], 1); mbar_init(&empty[b], C::CWARPS); }
    mbar_init(fullPlane, 1);
    mbar_init(emptyPlane, C::CWARPS);
    mbar_fence_init();
  }
  __syncthreads();

  // step s of tile (rb, cg): chunk ordinal s / N -> xi_x (all planes, or the representatives of plane zx), xi_y = s % N
  auto decode = [&](int t, int s, int& rb, int& cg, int& zx, int& ex, int& ey, int& X) {
    rb = t / G;
    cg = t - rb * G;
    zx = tiles[rb].zx;
    const int c = s / N;
    ey = s - c * N;
    ex = sym ? sym_rep(N, zx, c) : c;
    X = zx + N / 2 - ex;
    if (X < 0) X += N; else if (X > N - 1) X -= N;
  };

  if (warp >= C::CWARPS) {
    // ===== producer warpgroup =====
    if (C::REG_SPLIT) SBTE_SETMAXNREG_DEC(40);   // 40*128 + 232*256 = 64512 = 384*168
    if (warp == C::CWARPS && lane == 0) {
      int cur_cg = -1, cur_X = -1, epoch = -1;
      int t = sch.cta_tile[blockIdx.x];
      long long te = sch.tile_begin[t + 1];
      int sl = (int)(g0 - sch.tile_begin[t]);
      for (int k = 0; k < n; k++, sl++) {
        if (g0 + k == te) { t++; te = sch.tile_begin[t + 1]; sl = 0; }
        int rb, cg, zx, ex, ey, X;
        decode(t, sl, rb, cg, zx, ex, ey, X);
        if (cg != cur_cg || X != cur_X) {
          if (epoch >= 0) mbar_wait(emptyPlane, epoch & 1);   // every warp released the previous plane
          epoch++;
          cur_cg = cg; cur_X = X;
          mbar_arrive_expect_tx(fullPlane, (uint32_t)(C::PLANE * 16));
          const double2* src = spec + (size_t)cg * n3 * 32 + (size_t)X * N * C::LINE;
          for (int y = 0; y < N; y++)
            tma_bulk_g2s(plane + (size_t)y * C::LINE, src + (size_t)y * C::LINE, C::LINE * 16, fullPlane);
        }
        const int st = k % S;
        if (k >= S) mbar_wait(&empty[st], ((k / S) - 1) & 1);
        const MirrorTile mt = tiles[rb];
        int boxes = 0;
#pragma unroll
        for (int p = 0; p < C::PAIRS; p++) boxes += (mt.zyA[p] >= 0) + (mt.zyB[p] >= 0);
        mbar_arrive_expect_tx(&full[st], (uint32_t)(C::LINE * 16 + boxes * C::WTILE * 8));
        const int s = ex * N + ey;
        tma_bulk_g2s(stage_line(st), spec + (size_t)cg * n3 * 32 + (size_t)s * C::LINE, C::LINE * 16, &full[st]);
        const int zxB = (N - zx) % N, sB = ((N - ex) % N) * N + (N - ey) % N;   // mirrored plane and step
#pragma unroll
        for (int p = 0; p < C::PAIRS; p++) {
          if (mt.zyA[p] >= 0)
            tma_tensor2d_g2s(stage_w(st) + p * C::WTILE, &tmapW, s * N, (zx * N + mt.zyA[p]) * N, &full[st]);
          if (mt.zyB[p] >= 0)
            tma_tensor2d_g2s(stage_w(st) + (C::PAIRS + p) * C::WTILE, &tmapW, sB * N, (zxB * N + mt.zyB[p]) * N, &full[st]);
        }
      }
    }
    return;
  }

  // ===== compute warps =====
  if (C::REG_SPLIT) SBTE_SETMAXNREG_INC(232);
  const int pair = warp % C::PAIRS, half = warp / C::PAIRS;
  double2 accA[RH], accB[RH];
#pragma unroll
  for (int r = 0; r < RH; r++) { accA[r] = make_double2(0.0, 0.0); accB[r] = make_double2(0.0, 0.0); }

  int cur_t = -1, cur_cg = -1, cur_X = -1, epoch = -1;
  int zx = 0, zyA = -1, zyB = -1;
  int m_cur = 0;   // accB is held rotated by the current step-level phase (mirror_frame_update)

  auto flush = [&]() {
    const int rb = cur_t / G, cg = cur_t - rb * G;
    const long cell = (long)cg * 32 + lane;
    mirror_frame_update<RH>(accB, m_cur, 0, ph);
    if (cell < cells && zyA >= 0) {
      const int part = (int)blockIdx.x - sch.tile_first[cur_t];
      double2* base = parts + (size_t)part * part_stride + cell * n3;
      double2* outA = base + ((long)zx * N + zyA) * N;
      if (half == 0) {
#pragma unroll
        for (int r = 0; r < RH; r++) outA[r] = accA[r];
      } else {
#pragma unroll
        for (int r = 0; r < RH; r++) outA[RH + r] = accA[r];
      }
      if (zyB >= 0) {
        double2* outB = base + ((long)((N - zx) % N) * N + zyB) * N;
        if (half == 0) {
#pragma unroll
          for (int r = 0; r < RH; r++) outB[(N - r) % N] = accB[r];
        } else {
#pragma unroll
          for (int r = 0; r < RH; r++) outB[(N - RH - r) % N] = accB[r];
        }
      }
    }
  };

  int t = sch.cta_tile[blockIdx.x];
  long long te = sch.tile_begin[t + 1];
  int sl = (int)(g0 - sch.tile_begin[t]);
  for (int k = 0; k < n; k++, sl++) {
    if (g0 + k == te) { t++; te = sch.tile_begin[t + 1]; sl = 0; }
    int rb, cg, zxx, ex, ey, X;
    decode(t, sl, rb, cg, zxx, ex, ey, X);
    if (t != cur_t) {
      if (cur_t >= 0) {
        flush();
#pragma unroll
        for (int r = 0; r < RH; r++) { accA[r] = make_double2(0.0, 0.0); accB[r] = make_double2(0.0, 0.0); }
      }
      cur_t = t;
      zx = zxx;
      zyA = tiles[rb].zyA[pair];
      zyB = tiles[rb].zyB[pair];
    }
    if (cg != cur_cg || X != cur_X) {
      epoch++;
      cur_cg = cg; cur_X = X;
      mbar_wait(fullPlane, epoch & 1);
    }
    const int st = k % S;
    mbar_wait(&full[st], (k / S) & 1);

    if (zyA >= 0) {
      int Y = zyA + N / 2 - ey;
      if (Y < 0) Y += N; else if (Y > N - 1) Y -= N;
      const double2* fl = plane + (size_t)Y * C::LINE + lane;
      const double2* gl = stage_line(st) + lane;
      const double* wA = stage_w(st) + pair * C::WTILE;
      const double* wB = (zyB >= 0) ? stage_w(st) + (C::PAIRS + pair) * C::WTILE : zero_box;
      mirror_frame_update<RH>(accB, m_cur, (ex == 0) + (ey == 0) + (X == 0) + (Y == 0), ph);
      // folded tensor (fold != 0): the foldable steps of a paired column take the combined body
      if (fold && zyB >= 0 && mirror_exy(N, zx, zyA, ex, ey) == 0)
        mirror_step<N, RH, false, true>(accA, accB, fl, 32, gl, 32, wA, wB, half * RH, ph.t[1]);
      else
        mirror_step<N, RH, false, false>(accA, accB, fl, 32, gl, 32, wA, wB, half * RH, ph.t[1]);
    }

    // release the stage (and the plane when the next step needs another one)
    bool plane_done = (k == n - 1);
    if (!plane_done) {
      int rb2, cg2, zx2, ex2, ey2, X2;
      if (g0 + k + 1 == te) decode(t + 1, 0, rb2, cg2, zx2, ex2, ey2, X2);
      else decode(t, sl + 1, rb2, cg2, zx2, ex2, ey2, X2);
      plane_done = (cg2 != cur_cg) || (X2 != cur_X);
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&empty[st]);
      if (plane_done) mbar_arrive(emptyPlane);
    }
  }
  flush();
}

#ifndef SBTE_HOST_EMUL   // host launch code: not part of the host emulation
template <int N>
static void launch_mirror_n(sbte_ctx* c, const double2* spec, double2* parts, size_t part_stride, int cells,
                            const BatchSched& sch, bool fold) {
  using C = MirrorCfg<N>;
  auto kern = qhat_mirror_kernel<N>;
  static std::atomic<unsigned> configured{0};   // per device: function attributes belong to the device context
  if (!((configured.load() >> c->device) & 1u)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    configured.fetch_or(1u << c->device);
  }
  MirrorPhases ph;
  const double ang = -2.0 * c->L_eta * c->L_v;
  for (int m = 0; m < 5; m++) ph.t[m] = make_double2(cos(m * ang), sin(m * ang));
  k2_mark(c);
  kern<<<sch.P, C::THREADS, C::SMEM, c->stream>>>(fold ? c->tmapMh : (sch.sym ? c->tmapMs : c->tmapM), spec, parts, part_stride,
                                                   cells, sch, c->d_mtiles, ph, fold ? 1 : 0);
  k2_mark(c);
  c->launches += 1;
}
#endif


// ------------------------------------------------------------------------------------------
// line-ring variant (N = 20, 22, 24: the (zeta - xi)-side plane does not fit in shared memory), cf. qhat_batch3_kernel:
// the PAIRS consecutive A columns of a tile need, at step xi_y, the lines Y = zeta_y0 + w + N/2 - xi_y (w = slot):
// a window that slides by one line per step through an R-slot ring.  Both warps of a slot read the same line; the
// slot that reads a line last also arrives for the slots that never read it.  Stream-K ranges are whole chunks.
template <int N>
struct MirrorRingCfg {
  static constexpr int PAIRS = 4;
  static constexpr int CWARPS = 2 * PAIRS;
  static constexpr int RH = N / 2;                 // zeta_z rows per warp
  static constexpr int THREADS = CWARPS * 32 + 128;
  static constexpr int LINE = N * 32;
  static constexpr int RING = 10;
  static constexpr int STAGES = 2;
  static constexpr int LPC = N + PAIRS - 1;          // lines streamed per chunk
  static constexpr int WTILE = N * N;                    // doubles per column box
  static constexpr int WSLOT = (WTILE + 15) / 16 * 16;   // box slots start on 128-byte boundaries (TMA tensor copies)
  static constexpr size_t LINE_BYTES = (size_t)LINE * 16;
  static constexpr size_t STAGE_BYTES = LINE_BYTES + (size_t)2 * PAIRS * WSLOT * 8;
  static constexpr size_t SMEM = RING * LINE_BYTES + STAGES * STAGE_BYTES + (size_t)WTILE * 8 + 512;
  static_assert(SMEM <= 227 * 1024, "line ring + stages + zero box must fit in shared memory");
};

template <int R>
__device__ __forceinline__ void ring_clear(double2* a, double2* b) {
#pragma unroll
  for (int r = 0; r < R; r++) { a[r] = make_double2(0.0, 0.0); b[r] = make_double2(0.0, 0.0); }
}
template <int N>
__global__ void __launch_bounds__(MirrorRingCfg<N>::THREADS, 1)
qhat_mirror_ring_kernel(const __grid_constant__ CUtensorMap tmapW, const double2* __restrict__ spec,
                        double2* __restrict__ parts, size_t part_stride, int cells, BatchSched sch,
                        const MirrorTile* __restrict__ tiles, MirrorPhases ph) {
  using C = MirrorRingCfg<N>;
  constexpr long n3 = (long)N * N * N;
  constexpr int S = C::STAGES, R = C::RING, L = C::LPC, RH = C::RH;
  SBTE_DYN_SMEM(smraw);
  double2* ring = reinterpret_cast<double2*>(smraw);
  unsigned char* stage0 = smraw + R * C::LINE_BYTES;
  double* zero_box = reinterpret_cast<double*>(stage0 + S * C::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + S * C::STAGE_BYTES + (size_t)C::WTILE * 8);
  uint64_t* fullS = bars;              // [S]
  uint64_t* emptyS = bars + S;         // [S]
  uint64_t* fullL = bars + 2 * S;      // [R]
  uint64_t* emptyL = bars + 2 * S + R; // [R]
  auto stage_line = [&](int s) { return reinterpret_cast<double2*>(stage0 + (size_t)s * C::STAGE_BYTES); };
  auto stage_w = [&](int s) { return reinterpret_cast<double*>(stage0 + (size_t)s * C::STAGE_BYTES + C::LINE_BYTES); };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long g0 = sch.cta_begin[blockIdx.x];
  const int n = (int)(sch.cta_begin[blockIdx.x + 1] - g0);   // multiple of N (whole chunks)
  if (n <= 0) return;
  const int G = sch.G;
  const int nchunk = n / N;
  const bool sym = sch.sym != 0;

  for (int i = tid; i < C::WTILE; i += blockDim.x) zero_box[i] = 0.0;
  if (tid == 0) {
    for (int b = 0; b < S; b++) { mbar_init(&fullS[b], 1); mbar_init(&emptyS[b], C::CWARPS); }
    for (int b = 0; b < R; b++) { mbar_init(&fullL[b], 1); mbar_init(&emptyL[b], C::CWARPS); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp >= C::CWARPS) {
    SBTE_SETMAXNREG_DEC(40);
    if (warp == C::CWARPS && lane == 0) {
      int k = 0;          // local step counter (stage ring)
      long q = 0;         // line sequence number (line ring)
      int t = sch.cta_tile[blockIdx.x];
      long long te = sch.tile_begin[t + 1];
      int cl = (int)((g0 - sch.tile_begin[t]) / N);   // chunk ordinal inside the tile
      for (int ch = 0; ch < nchunk; ch++, cl++) {
        if (g0 + (long long)ch * N == te) { t++; te = sch.tile_begin[t + 1]; cl = 0; }
        const int rb = t / G, cg = t - rb * G;
        const MirrorTile mt = tiles[rb];
        const int zx = mt.zx, zy0 = mt.zyA[0], zxB = (N - zx) % N;
        const int ex = sym ? sym_rep(N, zx, cl) : cl;
        int X = zx + N / 2 - ex;
        if (X < 0) X += N; else if (X > N - 1) X -= N;
        int boxes = 0;
#pragma unroll
        for (int p = 0; p < C::PAIRS; p++) boxes += (mt.zyA[p] >= 0) + (mt.zyB[p] >= 0);
        const double2* gs = spec + (size_t)cg * n3 * 32;
        int issued = 0;   // lines of this chunk issued so far
        for (int ey = 0; ey < N; ey++, k++) {
          const int need = C::PAIRS + ey;   // lines needed by step ey: j' <= PAIRS-1 + ey
          for (; issued < need && issued < L; issued++, q++) {
            const int slot = (int)(q % R);
            if (q >= R) mbar_wait(&emptyL[slot], (uint32_t)(((q / R) - 1) & 1));
            int Y = (zy0 + N / 2 + C::PAIRS - 1 - issued) % N;
            if (Y < 0) Y += N;
            mbar_arrive_expect_tx(&fullL[slot], (uint32_t)C::LINE_BYTES);
            tma_bulk_g2s(ring + (size_t)slot * C::LINE, gs + ((size_t)X * N + Y) * C::LINE, (uint32_t)C::LINE_BYTES,
                         &fullL[slot]);
          }
          const int st = k % S;
          if (k >= S) mbar_wait(&emptyS[st], (uint32_t)(((k / S) - 1) & 1));
          const int s = ex * N + ey;
          const int sB = ((N - ex) % N) * N + (N - ey) % N;   // mirrored step
          mbar_arrive_expect_tx(&fullS[st], (uint32_t)(C::LINE_BYTES + (size_t)boxes * C::WTILE * 8));
          tma_bulk_g2s(stage_line(st), gs + (size_t)s * C::LINE, (uint32_t)C::LINE_BYTES, &fullS[st]);
#pragma unroll
          for (int p = 0; p < C::PAIRS; p++) {
            if (mt.zyA[p] >= 0)
              tma_tensor2d_g2s(stage_w(st) + p * C::WSLOT, &tmapW, s * N, (zx * N + mt.zyA[p]) * N, &fullS[st]);
            if (mt.zyB[p] >= 0)
              tma_tensor2d_g2s(stage_w(st) + (C::PAIRS + p) * C::WSLOT, &tmapW, sB * N, (zxB * N + mt.zyB[p]) * N, &fullS[st]);
          }
        }
      }
    }
    return;
  }

  SBTE_SETMAXNREG_INC(232);
  const int pair = warp % C::PAIRS, half = warp / C::PAIRS;
  double2 accA[RH], accB[RH];
  ring_clear<RH>(accA, accB);
  int cur_t = -1, zx = 0, zyA = -1, zyB = -1;
  int k = 0;
  long qbase = 0;
  int m_cur = 0;   // accB is held rotated by the current step-level phase (mirror_frame_update)

  auto flush = [&]() {
    const int rb = cur_t / G, cg = cur_t - rb * G;
    const long cell = (long)cg * 32 + lane;
    mirror_frame_update<RH>(accB, m_cur, 0, ph);
    if (cell < cells && zyA >= 0) {
      const int part = (int)blockIdx.x - sch.tile_first[cur_t];
      double2* base = parts + (size_t)part * part_stride + cell * n3;
      double2* outA = base + ((long)zx * N + zyA) * N + half * RH;
#pragma unroll
      for (int r = 0; r < RH; r++) outA[r] = accA[r];
      if (zyB >= 0) {
        double2* outB = base + ((long)((N - zx) % N) * N + zyB) * N;
        if (half == 0) {
#pragma unroll
          for (int r = 0; r < RH; r++) outB[(N - r) % N] = accB[r];
        } else {
#pragma unroll
          for (int r = 0; r < RH; r++) outB[(N - RH - r) % N] = accB[r];
        }
      }
    }
  };

  int t = sch.cta_tile[blockIdx.x];
  long long te = sch.tile_begin[t + 1];
  int cl = (int)((g0 - sch.tile_begin[t]) / N);
  for (int ch = 0; ch < nchunk; ch++, cl++, qbase += L) {
    if (g0 + (long long)ch * N == te) { t++; te = sch.tile_begin[t + 1]; cl = 0; }
    if (t != cur_t) {
      if (cur_t >= 0) {
        flush();
        ring_clear<RH>(accA, accB);
      }
      cur_t = t;
      const int rb = t / G;
      zx = tiles[rb].zx;
      zyA = tiles[rb].zyA[pair];
      zyB = tiles[rb].zyB[pair];
    }
    const int ex = sym ? sym_rep(N, zx, cl) : cl;
    int X = zx + N / 2 - ex;
    if (X < 0) X += N; else if (X > N - 1) X -= N;
    for (int ey = 0; ey < N; ey++, k++) {
      const int st = k % S;
      const int jl = C::PAIRS - 1 + ey - pair;          // this slot's line within the chunk
      const long q = qbase + jl;
      const int slot = (int)(q % R);
      mbar_wait(&fullS[st], (uint32_t)((k / S) & 1));
      mbar_wait(&fullL[slot], (uint32_t)((q / R) & 1));

      if (zyA >= 0) {
        int Y = zyA + N / 2 - ey;
        if (Y < 0) Y += N; else if (Y > N - 1) Y -= N;
        const double2* fl = ring + (size_t)slot * C::LINE + lane;
        const double2* gl = stage_line(st) + lane;
        const double* wA = stage_w(st) + pair * C::WSLOT;
        const double* wB = (zyB >= 0) ? stage_w(st) + (C::PAIRS + pair) * C::WSLOT : zero_box;
        mirror_frame_update<RH>(accB, m_cur, (ex == 0) + (ey == 0) + (X == 0) + (Y == 0), ph);
        // one pass over all RH rows: splitting the rows into several passes lets ptxas hoist the later passes' loads
        // and spills far more (3.3 KB against 0.5 KB at N = 24, none at N = 20)
        mirror_step<N, RH, true>(accA, accB, fl, 32, gl, 32, wA, wB, half * RH, ph.t[1]);
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&emptyS[st]);
        // readers of line jl are the slots with 0 <= jl - (PAIRS-1) + w' < N (both warps of a slot); the slot that
        // reads it last in step order also arrives for the slots that never read it
        uint32_t cnt = 1;
        if (jl < C::PAIRS - 1 && pair == C::PAIRS - 1) cnt = (uint32_t)(C::PAIRS - jl);
        if (jl > N - 1 && pair == N + C::PAIRS - 2 - jl) cnt = (uint32_t)(jl - N + 2);
        mbar_arrive_cnt(&emptyL[slot], cnt);
      }
    }
  }
  flush();
}

#ifndef SBTE_HOST_EMUL   // host launch code: not part of the host emulation
template <int N>
static void launch_mirror_ring_n(sbte_ctx* c, const double2* spec, double2* parts, size_t part_stride, int cells,
                                 const BatchSched& sch) {
  using C = MirrorRingCfg<N>;
  auto kern = qhat_mirror_ring_kernel<N>;
  static std::atomic<unsigned> configured{0};
  if (!((configured.load() >> c->device) & 1u)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    configured.fetch_or(1u << c->device);
  }
  MirrorPhases ph;
  const double ang = -2.0 * c->L_eta * c->L_v;
  for (int m = 0; m < 5; m++) ph.t[m] = make_double2(cos(m * ang), sin(m * ang));
  k2_mark(c);
  kern<<<sch.P, C::THREADS, C::SMEM, c->stream>>>(sch.sym ? c->tmapMs : c->tmapM, spec, parts, part_stride, cells, sch,
                                                   c->d_mtiles, ph);
  k2_mark(c);
  c->launches += 1;
}
#endif

#ifndef SBTE_HOST_EMUL   // host launch code: not part of the host emulation
// fold: stream the folded tensor (c->tmapMh, built for sch.sym) -- the result is then not Q^ but a spectrum with the same Q
void launch_qhat_mirror(sbte_ctx* c, const double2* spec, double2* parts, size_t part_stride, int cells,
                        const BatchSched& sch, bool fold) {
  if (!c->mirror_ok) { set_error("qhat_mirror: tensor maps / tile table not initialised"); return; }
  if (fold && !qhat_mirror_fold_enabled(c->N)) { set_error("qhat_mirror: no folded variant for this N"); return; }
  switch (c->N) {
    case 8: launch_mirror_n<8>(c, spec, parts, part_stride, cells, sch, fold); break;
    case 16: launch_mirror_n<16>(c, spec, parts, part_stride, cells, sch, fold); break;
    case 20: launch_mirror_ring_n<20>(c, spec, parts, part_stride, cells, sch); break;
    case 22: launch_mirror_ring_n<22>(c, spec, parts, part_stride, cells, sch); break;
    case 24: launch_mirror_ring_n<24>(c, spec, parts, part_stride, cells, sch); break;
    default: set_error("qhat_mirror: unsupported N"); break;
  }
}
#endif

}  // namespace sbte
