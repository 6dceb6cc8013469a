// spectralbte_b200/csrc/qhat_batch.cu -- K2, batched over spatial cells (the 1D-3V case).
//
//     Q^_b[zeta] = sum_xi W[zeta][xi] * f^_b[xi] * f^_b[wrap(zeta + N/2 - xi)]      b = cell
// (reference: the N^6 loop of /root/reference/src/collisions.c:127-165, run once per cell per RK
// stage by /root/reference/exec/boltz.c:285-345 with f == g).
//
// Every weight is shared by all cells, so the kernel is FP64-bound (6 FP64-pipe instructions per
// weight and cell), not HBM-bound.  Mapping:
//   lane  = cell (32 cells per warp; spectra are stored cell-minor so a warp reads 512 contiguous bytes)
//   warp  = one zeta (x,y) column: all N zeta_z rows
//   CTA   = COLS consecutive zeta_y columns of one zeta_x  x  one group of 32 cells
//   step  = one (xi_x, xi_y): each thread multiplies the N x N Toeplitz tile
//           p[r][c] = f^[xi_z = c] * f^[wrap(r + N/2 - c)] held entirely in registers (N + N complex
//           operands for N^2 products) against the weight tile, which every lane reads as a
//           shared-memory broadcast.
// Data movement is all TMA: the weight tile (COLS*N rows x N columns of W) is a 2-D tensor-map
// copy, the xi-side line and the (zeta - xi)-side plane are 1-D bulk copies, each completing on an
// mbarrier; a 3-stage ring keeps two steps of weights/lines in flight behind the FP64 pipe.
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include "common.cuh"
#include "internal.h"

namespace sbte {

// ------------------------------------------------------------------------------------------
// any even N (12, 20, 22, 28, ...): lanes = cells on the cell-minor layout, operands read straight from L1/L2
// (no TMA staging), the xi loop with incrementally wrapped indices; visits only the representative xi_x planes
// when handed the symmetrised tensor.
// Each warp owns ANY_R consecutive zeta_z rows of one zeta (x,y) column: the xi-side operand is shared by the rows
// and the (zeta - xi)-side operands form a sliding window (Toeplitz in z), so one step costs two 512-byte operand
// reads for ANY_R products instead of two per product.
// ------------------------------------------------------------------------------------------
template <int ANY_R>
__global__ void __launch_bounds__(256)
qhat_batch_any_kernel(const double* __restrict__ W, const double2* __restrict__ spec, double2* __restrict__ qhat,
                      int N, int cells, int sym) {
  const long n3 = (long)N * N * N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = blockIdx.x;
  const int groups_z = (N + ANY_R - 1) / ANY_R;
  const int task = blockIdx.y * 8 + warp;               // (zeta_x, zeta_y, group of ANY_R zeta_z rows)
  if (task >= N * N * groups_z) return;
  const int col = task / groups_z, zz0 = (task % groups_z) * ANY_R;
  const int zx = col / N, zy = col % N;
  const int n2 = N / 2;
  const double2* S = spec + (size_t)cg * n3 * 32 + lane;
  const double* w0 = W + ((size_t)col * N + zz0) * n3;   // row zeta = (zx, zy, zz0)
  bool live[ANY_R];
#pragma unroll
  for (int r = 0; r < ANY_R; r++) live[r] = zz0 + r < N;
  const int nrep = sym ? sym_nrep(N, zx) : N;
  double ar[ANY_R], ai[ANY_R];
#pragma unroll
  for (int r = 0; r < ANY_R; r++) { ar[r] = 0.0; ai[r] = 0.0; }
  for (int c = 0; c < nrep; c++) {
    const int ex = sym ? sym_rep(N, zx, c) : c;
    int x = zx + n2 - ex;
    if (x < 0) x += N; else if (x > N - 1) x -= N;
    for (int ey = 0; ey < N; ey++) {
      int y = zy + n2 - ey;
      if (y < 0) y += N; else if (y > N - 1) y -= N;
      const double2* gl = S + (size_t)((ex * N + ey) * N) * 32;
      const double2* fl = S + (size_t)((x * N + y) * N) * 32;
      const double* wl = w0 + (ex * N + ey) * N;
      // window fw[r] = f^[wrap(zz0 + r + N/2 - xi_z)], xi_z = 0
      double2 fw[ANY_R];
      int z = zz0 + n2;
      if (z > N - 1) z -= N;
#pragma unroll
      for (int r = 0; r < ANY_R; r++) {
        int zr = z + r;
        if (zr > N - 1) zr -= N;
        fw[r] = fl[zr * 32];
      }
      for (int ez = 0; ez < N; ez++) {
        const double2 g = gl[ez * 32];
#pragma unroll
        for (int r = 0; r < ANY_R; r++) {
          const double wv = live[r] ? wl[(size_t)r * n3 + ez] : 0.0;
          const double pr = g.x * fw[r].x - g.y * fw[r].y, pi = g.x * fw[r].y + g.y * fw[r].x;
          ar[r] = fma(wv, pr, ar[r]);
          ai[r] = fma(wv, pi, ai[r]);
        }
        // xi_z + 1: every row's operand index drops by one
#pragma unroll
        for (int r = ANY_R - 1; r > 0; r--) fw[r] = fw[r - 1];
        z = (z == 0) ? N - 1 : z - 1;
        fw[0] = fl[z * 32];
      }
    }
  }
  const long cell = (long)cg * 32 + lane;
  if (cell < cells) {
#pragma unroll
    for (int r = 0; r < ANY_R; r++)
      if (live[r]) qhat[cell * n3 + (long)col * N + zz0 + r] = make_double2(ar[r], ai[r]);
  }
}

#ifndef SBTE_HOST_EMUL   // host launch code: not part of the host emulation
void launch_qhat_batch_any(sbte_ctx* c, const double2* spec, double2* qhat, int cells, bool sym) {
  const int groups = (cells + 31) / 32;
  static const int R0 = getenv("SBTE_ANY_R") ? atoi(getenv("SBTE_ANY_R")) : 4;
  const int R = (R0 == 8 || R0 == 2) ? R0 : 4;
  const int tasks = c->N * c->N * ((c->N + R - 1) / R);
  dim3 grid(groups, (unsigned)((tasks + 7) / 8));
  k2_mark(c);
  if (R == 8)
    qhat_batch_any_kernel<8><<<grid, 256, 0, c->stream>>>(sym ? c->d_Ws : c->d_W, spec, qhat, c->N, cells, sym ? 1 : 0);
  else if (R == 2)
    qhat_batch_any_kernel<2><<<grid, 256, 0, c->stream>>>(sym ? c->d_Ws : c->d_W, spec, qhat, c->N, cells, sym ? 1 : 0);
  else
    qhat_batch_any_kernel<4><<<grid, 256, 0, c->stream>>>(sym ? c->d_Ws : c->d_W, spec, qhat, c->N, cells, sym ? 1 : 0);
  k2_mark(c);
  c->launches += 1;
}
#endif

// N = 20, 22: the line-ring kernel with partly empty row-blocks (SBTE_NO_BATCH3G=1: back to the any-N kernel)
static bool batch3_general(int N) {
  static const bool on = getenv("SBTE_NO_BATCH3G") == nullptr;
  return on && (N == 20 || N == 22);
}
bool qhat_batch_supported(int N) { return N == 8 || N == 16 || N == 24 || batch3_general(N); }
// Where a stream-K range may end.  The line-ring kernels take a cut at any step (a cut xi_x chunk's running sum is handed
// from one CTA to the next, so the canonical summation order is kept); the resident-plane kernel (N = 8, and N = 16 under
// SBTE_N16_PLANE) needs whole chunks.  0 = whole chunks only, 1 = the schedule builder chooses (see qhat_batch_cut_cost),
// 2 = any step.  SBTE_CHUNK_CUTS=1 forces whole chunks, =0 forces cuts at any step (A/B runs).
int qhat_batch_cut_mode(int N) {
  static const char* env = getenv("SBTE_CHUNK_CUTS");
  static const bool plane16 = getenv("SBTE_N16_PLANE") != nullptr;
  if (N == 8 || (N == 16 && plane16)) return 0;
  if (env) return atoi(env) != 0 ? 0 : 2;
  return 1;
}
// Measured price of cutting anywhere, as a fraction of the launch: with whole-chunk cuts every CTA reaches its chunk ends
// at the same moments and the SMs run the (45-75 KB, fully unrolled) step body in lockstep, which the instruction caches
// they share reward; ranges of unequal phase lose that (ncu, N = 24, 250 cells: stall_no_instruction 0.29 -> 0.45 per
// issue, 12.11 -> 12.46 ms).  N = 16's body fits the per-SM cache; there the price is that of the step loop with run-time
// bounds (the instance for whole-chunk schedules, CUTS = false, has them at compile time: 2.487 against 2.505 ms at 640
// cells).  The builder cuts anywhere only where the better balance is worth more than this.
double qhat_batch_cut_cost(int N) { return N <= 16 ? 0.012 : N <= 20 ? 0.015 : N <= 22 ? 0.03 : 0.035; }
int qhat_batch_cols(int N) { return (N >= 16) ? 8 : 4; }

// ------------------------------------------------------------------------------------------
// v2: persistent CTAs, stream-K split of the (tile, step) iteration space, decoupled warps
// ------------------------------------------------------------------------------------------
// The work is T = groups * row-blocks tiles of N^2 steps each.  One CTA per SM takes an equal,
// contiguous share of the T*N^2 global steps (tile-major, cell group fastest so that concurrently
// running CTAs read the same weight rows from L2).  A tile cut by a CTA boundary is written as two
// (or more) partial sums into `parts[k]`; the inverse transform adds the parts in fixed order, so the
// result is deterministic and independent of the SM count only through that order.
// Inside a CTA the warps are decoupled: a 4-stage ring of (weight tile, xi-side line) with full
// (TMA transaction) and empty (one arrival per warp) mbarriers; lane 0 of warp 0 issues the TMA
// copies two to three steps ahead.  No __syncthreads in the main loop.
template <int N>
struct Batch2Cfg {
  static constexpr int COLS = (N >= 16) ? 8 : 4;
  static constexpr int CONSUMERS = COLS * 32;      // compute threads: one warp per zeta (x,y) column
  static constexpr int THREADS = CONSUMERS + 128;  // + one producer warpgroup (one lane issues the TMA copies)
  // register split (setmaxnreg works on warpgroups): 384 threads compile to <= 168 registers; the
  // producer warpgroup drops to 24 and the two compute warpgroups grow to 240 (240*256 + 24*128 <= 64512)
  static constexpr bool REG_SPLIT = THREADS > 256;
  static constexpr int ROWS = COLS * N;
  static constexpr int LINE = N * 32;
  static constexpr int PLANE = N * LINE;
  static constexpr int STAGES = 4;
  static constexpr size_t STAGE_BYTES = (size_t)LINE * 16 + (size_t)ROWS * N * 8;
  static constexpr size_t SMEM = (size_t)PLANE * 16 + STAGES * STAGE_BYTES + 256;
};

template <int N>
__global__ void __launch_bounds__(Batch2Cfg<N>::THREADS, 1)
qhat_batch2_kernel(const __grid_constant__ CUtensorMap tmapW, const double2* __restrict__ spec,
                   double2* __restrict__ parts, size_t part_stride, int cells, BatchSched sch) {
  using C = Batch2Cfg<N>;
  constexpr long n3 = (long)N * N * N;
  constexpr int S = C::STAGES;
  SBTE_DYN_SMEM(smraw);
  double2* plane = reinterpret_cast<double2*>(smraw);
  unsigned char* stage0 = smraw + (size_t)C::PLANE * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + S * C::STAGE_BYTES);
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

  if (tid == 0) {
    for (int b = 0; b < S; b++) { mbar_init(&full[b], 1); mbar_init(&empty[b], C::COLS); }
    mbar_init(fullPlane, 1);
    mbar_init(emptyPlane, C::COLS);
    mbar_fence_init();
  }
  __syncthreads();

  // A tile (row-block rb, cell group cg) has len = (visited xi_x planes) * N steps; tile_begin[] holds
  // the global step offsets.  Step s of a tile -> chunk ordinal s / N -> xi_x (all planes, or the
  // representatives of the symmetrised tensor), xi_y = s % N, dif-side plane X = wrap(zeta_x + N/2 - xi_x).
  auto decode = [&](int t, int s, int& rb, int& cg, int& ex, int& ey, int& X) {
    rb = t / G;
    cg = t - rb * G;
    const int zxx = (rb * C::COLS) / N;
    const int c = s / N;
    ey = s - c * N;
    ex = sym ? sym_rep(N, zxx, c) : c;
    X = zxx + N / 2 - ex;
    if (X < 0) X += N; else if (X > N - 1) X -= N;
  };

  if (warp >= C::COLS) {
    // ===== producer warpgroup: one lane issues every TMA copy, up to S steps ahead of the consumers =====
    if (C::REG_SPLIT) SBTE_SETMAXNREG_DEC(24);
    if (warp == C::COLS && lane == 0) {
      int cur_cg = -1, cur_X = -1, epoch = -1;
      int t = sch.cta_tile[blockIdx.x];
      long long te = sch.tile_begin[t + 1];
      int sl = (int)(g0 - sch.tile_begin[t]);
      for (int k = 0; k < n; k++, sl++) {
        if (g0 + k == te) { t++; te = sch.tile_begin[t + 1]; sl = 0; }
        int rb, cg, ex, ey, X;
        decode(t, sl, rb, cg, ex, ey, X);
        const int s = ex * N + ey;
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
        if (k >= S) mbar_wait(&empty[st], ((k / S) - 1) & 1);  // consumers finished the previous use of this stage
        mbar_arrive_expect_tx(&full[st], (uint32_t)C::STAGE_BYTES);
        tma_bulk_g2s(stage_line(st), spec + (size_t)cg * n3 * 32 + (size_t)s * C::LINE, C::LINE * 16, &full[st]);
        tma_tensor2d_g2s(stage_w(st), &tmapW, s * N, rb * C::ROWS, &full[st]);
      }
    }
    return;
  }

  // ===== compute warps =====
  if (C::REG_SPLIT) SBTE_SETMAXNREG_INC(240);
  double2 acc[N];
#pragma unroll
  for (int r = 0; r < N; r++) acc[r] = make_double2(0.0, 0.0);

  int cur_t = -1, cur_cg = -1, cur_X = -1, epoch = -1;
  int zx = 0, zy = 0;

  // Canonical summation order (independent of the schedule, hence of the SM count and of how many cells this rank
  // holds): Q^ = ((C_0 + C_1) + C_2) + ..., C_c = the sum over the N steps of the c-th visited xi_x chunk started from
  // zero.  CTA ranges are whole chunks.  The CTA that owns a tile's first chunk keeps the running fold in part 0 (a
  // reduction into its own 16 N bytes per thread at every chunk end); a CTA that
  // enters the tile later writes every chunk as its own part 1 + (c - e0), e0 = chunks held by the first CTA, and the
  // inverse transform continues the same left fold over the parts.  The fold into part 0 is a fire-and-forget
  // red.global.add.f64 per component (one writer per location, program order): a load/add/store there cost 8 % at
  // N = 24 (profiles/r02_canonical_fold_ab.txt), the reduction in L2 costs nothing measurable.
  auto chunk_end = [&](int c) {
    const int cg = cur_t - (cur_t / G) * G;
    const long cell = (long)cg * 32 + lane;
    if (cell < cells) {
      const int first = sch.tile_first[cur_t];
      double2* out = parts + cell * n3 + ((long)zx * N + zy) * N;
      bool fold = false;
      if ((int)blockIdx.x == first) {
        fold = c > 0;
      } else {
        const int e0 = (int)((sch.cta_begin[first + 1] - sch.tile_begin[cur_t]) / N);
        out += (size_t)(1 + c - e0) * part_stride;
      }
      if (fold) {   // part 0 += C_c in L2, no round trip: this thread is the only writer of these locations
        double* o = reinterpret_cast<double*>(out);
#pragma unroll
        for (int r = 0; r < N; r++) { red_add_f64(o + 2 * r, acc[r].x); red_add_f64(o + 2 * r + 1, acc[r].y); }
      } else {
#pragma unroll
        for (int r = 0; r < N; r++) out[r] = acc[r];
      }
    }
#pragma unroll
    for (int r = 0; r < N; r++) acc[r] = make_double2(0.0, 0.0);
  };

  int t = sch.cta_tile[blockIdx.x];
  long long te = sch.tile_begin[t + 1];
  int sl = (int)(g0 - sch.tile_begin[t]);
  for (int k = 0; k < n; k++, sl++) {
    if (g0 + k == te) { t++; te = sch.tile_begin[t + 1]; sl = 0; }
    int rb, cg, ex, ey, X;
    decode(t, sl, rb, cg, ex, ey, X);
    if (t != cur_t) {
      cur_t = t;
      const int q0 = rb * C::COLS;
      zx = q0 / N;
      zy = (q0 % N) + warp;
    }
    if (cg != cur_cg || X != cur_X) {
      epoch++;
      cur_cg = cg; cur_X = X;
      mbar_wait(fullPlane, epoch & 1);
    }
    const int st = k % S;
    mbar_wait(&full[st], (k / S) & 1);

    int Y = zy + N / 2 - ey;
    if (Y < 0) Y += N; else if (Y > N - 1) Y -= N;
    const double2* fl = plane + (size_t)Y * C::LINE + lane;
    const double2* gl = stage_line(st) + lane;
    const double* wt = stage_w(st) + warp * N * N;

    double2 fr[N];
#pragma unroll
    for (int z = 0; z < N; z++) fr[z] = fl[z * 32];
#pragma unroll
    for (int c = 0; c < N; c += 2) {
      const double2 g0v = gl[c * 32], g1v = gl[(c + 1) * 32];
#pragma unroll
      for (int r = 0; r < N; r++) {
        const double2 w2 = *reinterpret_cast<const double2*>(wt + r * N + c);
        const double2 p0 = cmul(g0v, fr[(r + N / 2 - c + N) % N]);
        const double2 p1 = cmul(g1v, fr[(r + N / 2 - c - 1 + N) % N]);
        cmac(acc[r], w2.x, p0);
        cmac(acc[r], w2.y, p1);
      }
    }

    // release the stage (and the plane when the next step needs another one)
    bool plane_done = (k == n - 1);
    if (!plane_done) {
      int rb2, cg2, ex2, ey2, X2;
      if (g0 + k + 1 == te) decode(t + 1, 0, rb2, cg2, ex2, ey2, X2);
      else decode(t, sl + 1, rb2, cg2, ex2, ey2, X2);
      plane_done = (cg2 != cur_cg) || (X2 != cur_X);
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&empty[st]);
      if (plane_done) mbar_arrive(emptyPlane);
    }
    if (ey == N - 1) chunk_end(sl / N);
  }
}

#ifndef SBTE_HOST_EMUL   // host launch code: not part of the host emulation
template <int N>
static void launch_batch2_n(sbte_ctx* c, const double2* spec, double2* parts, size_t part_stride, int cells,
                            const BatchSched& sch) {
  using C = Batch2Cfg<N>;
  auto kern = qhat_batch2_kernel<N>;
  static std::atomic<unsigned> configured{0};   // per device: function attributes belong to the device context
  if (!((configured.load() >> c->device) & 1u)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    configured.fetch_or(1u << c->device);
  }
  k2_mark(c);
  kern<<<sch.P, C::THREADS, C::SMEM, c->stream>>>(sch.sym ? c->tmapWs : c->tmapW, spec, parts, part_stride, cells, sch);
  k2_mark(c);
  c->launches += 1;
}
#endif

// ------------------------------------------------------------------------------------------
// v3 ("line ring"): same mapping as v2 for N whose (zeta - xi)-side plane does not fit in shared
// memory (N = 24: 288 KB).  The COLS columns of a CTA need, at step xi_y, the COLS consecutive lines
// Y = zeta_y0 + w + N/2 - xi_y (w = warp): a window that slides by one line per step.  Per xi_x chunk the
// producer therefore streams L = N + COLS - 1 lines through an R-slot ring (line j' of the chunk is
// Y = zeta_y0 + N/2 + COLS-1 - j'; warp w reads line j' = COLS-1 + xi_y - w at step xi_y).  Every slot
// has a full (TMA) and an empty mbarrier of COLS arrivals; lines at the chunk edges have fewer than COLS
// readers, so their last reader arrives for the absent ones (mbarrier.arrive with a count).
// Stream-K ranges may be cut anywhere (qhat_batch_cut_mode): a cut chunk's running sum passes from one CTA to the next.
// When COLS does not divide N (N = 20, 22) the last row-block of every zeta_x plane is partly empty: its surplus
// warps keep the barrier protocol going (waits and arrivals) but neither multiply nor write, and the weight
// tile's surplus rows are the next plane's (or, past the end of the tensor, TMA zero fill).
// SPLIT (N = 16): the tile of a cell group with at most 16 live cells.  A warp then serves TWO zeta_y columns with 16
// cells each (lanes 0-15: column 2w, lanes 16-31: column 2w+1), the CTA all 16 columns of a zeta_x plane: half the work
// of the two ordinary tiles the padded group would need.  Same per-lane arithmetic in the same order, hence the same bits.
template <int N, bool SPLIT = false>
struct Batch3Cfg {
  static constexpr int WARPS = 8;                    // compute warps
  static constexpr int COLS = SPLIT ? 16 : 8;        // zeta_y columns per CTA
  static constexpr int BPX = (N + COLS - 1) / COLS;  // row-blocks per zeta_x plane
  static constexpr bool PARTIAL = (N % COLS) != 0;
  static constexpr int CONSUMERS = WARPS * 32;
  static constexpr int THREADS = CONSUMERS + 128;
  static constexpr int ROWS = COLS * N;
  static constexpr int LINE = N * 32;
  // stages of (xi-side line, weight tile) the producer may run ahead, and a line ring deep enough for that lookahead
  // (the COLS lines being read + one per stage): N = 16: 4 / 12 (split: 2 / 18), N = 20: 3 / 11, N = 22, 24: 2 / 10
  static constexpr int STAGES = SPLIT ? 2 : (N <= 16) ? 4 : (N <= 20) ? 3 : 2;
  static constexpr int RING = COLS + STAGES;
  static constexpr int LPC = N + COLS - 1;          // lines streamed per chunk
  static constexpr size_t LINE_BYTES = (size_t)LINE * 16;
  static constexpr size_t STAGE_BYTES = LINE_BYTES + (size_t)ROWS * N * 8;
  static constexpr size_t SMEM = RING * LINE_BYTES + STAGES * STAGE_BYTES + 512;
  static_assert(SMEM <= 227 * 1024, "line ring + stages must fit in shared memory");
  static_assert(WARPS == kBatchWarps, "carry slots are sized for kBatchWarps warps per CTA");
  static_assert(!SPLIT || (N % 16 == 0 && ROWS <= 256), "split tiles: whole zeta_x planes of 16 columns, TMA box <= 256 rows");
};

// cg_base: cell group the (single-group) schedule of a SPLIT launch refers to; 0 otherwise
// CUTS = false: the instance for schedules cut at whole chunks only (no hand-over code, step loop with compile-time bounds)
// cta: this CTA's index in the schedule (the block index, except in the paired launch below)
template <int N, bool SPLIT, bool CUTS>
__device__ __forceinline__ void qhat_batch3_body(const CUtensorMap& tmapW, const double2* __restrict__ spec,
                                                 double2* __restrict__ parts, size_t part_stride, int cells,
                                                 const BatchSched& sch, int cg_base, const unsigned cta) {
  using C = Batch3Cfg<N, SPLIT>;
  constexpr long n3 = (long)N * N * N;
  constexpr int S = C::STAGES, R = C::RING;
  SBTE_DYN_SMEM(smraw);
  double2* ring = reinterpret_cast<double2*>(smraw);
  unsigned char* stage0 = smraw + R * C::LINE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + S * C::STAGE_BYTES);
  uint64_t* fullS = bars;              // [S]
  uint64_t* emptyS = bars + S;         // [S]
  uint64_t* fullL = bars + 2 * S;      // [R]
  uint64_t* emptyL = bars + 2 * S + R; // [R]
  auto stage_line = [&](int s) { return reinterpret_cast<double2*>(stage0 + (size_t)s * C::STAGE_BYTES); };
  auto stage_w = [&](int s) { return reinterpret_cast<double*>(stage0 + (size_t)s * C::STAGE_BYTES + C::LINE_BYTES); };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = sch.G;
  const bool sym = sch.sym != 0;
  // The range [g0, g1) in segments of one xi_x chunk each (chunk boundaries are the multiples of N): a cut may fall
  // inside a chunk, whose running sum then passes from this CTA's predecessor to it (or from it to its successor)
  // through sch.carry.  Order: the head of the chunk cut by g1 FIRST (published at once: the successor needs it only at
  // the very end of its own range, so nobody waits), the whole chunks, the tail of the chunk cut by g0 LAST.
  // The host schedule gives every CTA at least N steps, so the two cut chunks are different chunks.
  // Positions are kept relative to g0 (32 bits): h0 = steps of the cut chunk at the start, n1 = steps of the one at the end.
  int h0, n1, nwhole, tb, te;
  const int t0 = sch.cta_tile[cta];
  {
    const long long g0 = sch.cta_begin[cta], g1 = sch.cta_begin[cta + 1];
    if (g1 <= g0) return;
    const long long w0 = ((g0 + N - 1) / N) * N, w1 = (g1 / N) * N;
    h0 = (int)(w0 - g0);
    n1 = (int)(g1 - w1);
    nwhole = (int)((w1 - w0) / N);
    tb = (int)(sch.tile_begin[t0] - g0);       // bounds of the tile the whole chunks are in, cached
    te = (int)(sch.tile_begin[t0 + 1] - g0);
  }
  const int seg_pre = n1 > 0 ? 1 : 0, seg_post = h0 > 0 ? 1 : 0;
  const int nseg = seg_pre + nwhole + seg_post;
  // segment i -> its tile, its chunk's ordinal in the tile and the xi_y range [ey0, ey1) this CTA computes of it
  int tw = t0;
  auto segment = [&](int i, int& t, int& cl, int& ey0, int& ey1) {
    if (i >= seg_pre && i - seg_pre < nwhole) {
      const int rc = h0 + (i - seg_pre) * N;
      if (rc >= te) {   // tiles are whole chunks: one switch at most
        tw++; tb = te;
        te = (int)(sch.tile_begin[tw + 1] - sch.cta_begin[cta]);
      }
      t = tw; cl = (rc - tb) / N; ey0 = 0; ey1 = N;
    } else if (i < seg_pre) {   // at most once, before any whole chunk
      const long long w1 = sch.cta_begin[cta + 1] - n1;
      t = t0;
      while (w1 >= sch.tile_begin[t + 1]) t++;
      cl = (int)((w1 - sch.tile_begin[t]) / N); ey0 = 0; ey1 = n1;
    } else {                    // at most once, after the whole chunks
      t = t0;
      cl = (int)((sch.cta_begin[cta] + h0 - N - sch.tile_begin[t0]) / N); ey0 = N - h0; ey1 = N;
    }
  };

  if (tid == 0) {
    for (int b = 0; b < S; b++) { mbar_init(&fullS[b], 1); mbar_init(&emptyS[b], C::WARPS); }
    for (int b = 0; b < R; b++) { mbar_init(&fullL[b], 1); mbar_init(&emptyL[b], C::COLS); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp >= C::WARPS) {
    SBTE_SETMAXNREG_DEC(24);
    if (warp == C::WARPS && lane == 0) {
      int k = 0;          // local step counter (stage ring)
      long q = 0;         // line sequence number (line ring)
      for (int i = 0; i < nseg; i++) {
        int t, cl, ey0, ey1;
        segment(i, t, cl, ey0, ey1);
        const int rb = t / G, cg = t - rb * G;
        const int zx = rb / C::BPX, zy0 = (rb % C::BPX) * C::COLS;
        const int ex = sym ? sym_rep(N, zx, cl) : cl;
        int X = zx + N / 2 - ex;
        if (X < 0) X += N; else if (X > N - 1) X -= N;
        const double2* gs = spec + (size_t)(cg_base + cg) * n3 * 32;
        int issued = ey0;                      // next line j' of this chunk to issue: the segment needs [ey0, lend)
        const int lend = C::COLS - 1 + ey1;
        for (int ey = ey0; ey < ey1; ey++, k++) {
          // lines needed by step ey: j' <= COLS-1 + ey
          const int need = C::COLS + ey;
          for (; issued < need && issued < lend; issued++, q++) {
            const int slot = (int)(q % R);
            if (q >= R) mbar_wait(&emptyL[slot], (uint32_t)(((q / R) - 1) & 1));
            int Y = (zy0 + N / 2 + C::COLS - 1 - issued) % N;
            if (Y < 0) Y += N;
            mbar_arrive_expect_tx(&fullL[slot], (uint32_t)C::LINE_BYTES);
            tma_bulk_g2s(ring + (size_t)slot * C::LINE, gs + ((size_t)X * N + Y) * C::LINE, (uint32_t)C::LINE_BYTES,
                         &fullL[slot]);
          }
          const int st = k % S;
          if (k >= S) mbar_wait(&emptyS[st], (uint32_t)(((k / S) - 1) & 1));
          const int s = ex * N + ey;
          mbar_arrive_expect_tx(&fullS[st], (uint32_t)C::STAGE_BYTES);
          tma_bulk_g2s(stage_line(st), gs + (size_t)s * C::LINE, (uint32_t)C::LINE_BYTES, &fullS[st]);
          tma_tensor2d_g2s(stage_w(st), &tmapW, s * N, (zx * N + zy0) * N, &fullS[st]);
        }
      }
    }
    return;
  }

  SBTE_SETMAXNREG_INC(240);
  // this lane's zeta_y column inside the CTA's row-block and its cell inside the group
  const int col = SPLIT ? 2 * warp + (lane >> 4) : warp;
  const int clane = SPLIT ? (lane & 15) : lane;
  double2 acc[N];
#pragma unroll
  for (int r = 0; r < N; r++) acc[r] = make_double2(0.0, 0.0);
  int cur_t = -1, zx = 0, zy = 0;
  int k = 0;
  using QT = std::conditional_t<(N <= 16), int, long>;   // line sequence numbers (32 bits are plenty; N = 16 runs faster with them)
  QT qbase = 0;

  // canonical summation order, as in qhat_batch2_kernel: left fold over the chunk sums, part 0 kept by the CTA that owns
  // the tile's first step (it folds the chunks it completes), one part per chunk from every later CTA.  A chunk cut by
  // a range boundary is completed -- and written -- by the later of its two CTAs.
  auto chunk_end = [&](int c) {
    const int cg = cur_t - (cur_t / G) * G;
    const long cell = (long)(cg_base + cg) * 32 + clane;
    if (cell < cells && (!C::PARTIAL || zy < N)) {
      const int first = sch.tile_first[cur_t];
      double2* out = parts + cell * n3 + ((long)zx * N + zy) * N;
      bool fold = false;
      if ((int)cta == first) {
        fold = c > 0;
      } else {
        const int e0 = (int)((sch.cta_begin[first + 1] - sch.tile_begin[cur_t]) / N);   // chunks the first CTA completes
        out += (size_t)((e0 > 0 ? 1 : 0) + c - e0) * part_stride;
      }
      if (fold) {   // part 0 += C_c in L2, no round trip: this thread is the only writer of these locations
        double* o = reinterpret_cast<double*>(out);
#pragma unroll
        for (int r = 0; r < N; r++) { red_add_f64(o + 2 * r, acc[r].x); red_add_f64(o + 2 * r + 1, acc[r].y); }
      } else {
#pragma unroll
        for (int r = 0; r < N; r++) out[r] = acc[r];
      }
    }
#pragma unroll
    for (int r = 0; r < N; r++) acc[r] = make_double2(0.0, 0.0);
  };

  // The steps [ey0, ey1) of the current chunk: the general form, and (used for N > 16, see below) the form for whole
  // chunks, ey0 = 0 and ey1 = N at compile time.
  auto run_steps = [&](auto whole_tag, int ey0_, int ey1_) {
    constexpr bool WHOLE = decltype(whole_tag)::value;
    const int ey0 = WHOLE ? 0 : ey0_, ey1 = WHOLE ? N : ey1_;
    const bool live = !C::PARTIAL || zy < N;
    for (int ey = ey0; ey < ey1; ey++, k++) {
      const int st = k % S;
      const int jl = C::COLS - 1 + ey - col;           // this column's line within the chunk
      const QT q = qbase + (jl - ey0);                 // the segment streams lines ey0 .. COLS-2 + ey1
      const int slot = (int)(q % R);
      mbar_wait(&fullS[st], (uint32_t)((k / S) & 1));
      mbar_wait(&fullL[slot], (uint32_t)((q / R) & 1));

      if (live) {
        const double2* fl = ring + (size_t)slot * C::LINE + clane;
        const double2* gl = stage_line(st) + clane;
        const double* wt = stage_w(st) + col * N * N;
        double2 fr[N];
#pragma unroll
        for (int z = 0; z < N; z++) fr[z] = fl[z * 32];
#pragma unroll
        for (int c = 0; c < N; c += 2) {
          const double2 g0v = gl[c * 32], g1v = gl[(c + 1) * 32];
#pragma unroll
          for (int r = 0; r < N; r++) {
            const double2 w2 = *reinterpret_cast<const double2*>(wt + r * N + c);
            const double2 p0 = cmul(g0v, fr[(r + N / 2 - c + N) % N]);
            const double2 p1 = cmul(g1v, fr[(r + N / 2 - c - 1 + N) % N]);
            cmac(acc[r], w2.x, p0);
            cmac(acc[r], w2.y, p1);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&emptyS[st]);
      if (clane == 0) {
        // readers of line jl in this segment are the columns col' with ey0 <= jl - (COLS-1) + col' < ey1; the one that
        // comes last in step order (the largest) also arrives for the columns that never read the line
        uint32_t cnt = 1;
        if (WHOLE) {
          if (jl < C::COLS - 1 && col == C::COLS - 1) cnt = (uint32_t)(C::COLS - jl);
          if (jl > N - 1 && col == N + C::COLS - 2 - jl) cnt = (uint32_t)(jl - N + 2);
        } else {
          const int hi = ey1 - 1 - jl + C::COLS - 1, lo = ey0 - jl + C::COLS - 1;
          const int cmax = hi < C::COLS - 1 ? hi : C::COLS - 1;
          const int cmin = lo > 0 ? lo : 0;
          if (col == cmax) cnt = (uint32_t)(C::COLS - (cmax - cmin));
        }
        mbar_arrive_cnt(&emptyL[slot], cnt);
      }
    }
    qbase += C::COLS - 1 + ey1 - ey0;
  };
  auto enter_tile = [&](int t) {
    if (t != cur_t) {
      cur_t = t;
      const int rb = t / G;
      zx = rb / C::BPX;
      zy = (rb % C::BPX) * C::COLS + col;
    }
  };

  // a chunk begun by the previous CTA: continue from its running sum (same operations in the same order as an uncut
  // chunk, hence the same bits)
  auto carry_in = [&]() {
    int* flag = sch.carry_flag + ((size_t)cta - 1) * C::WARPS + warp;
    if (lane == 0) carry_await(flag);
    __syncwarp();
    const double2* in = sch.carry + (((size_t)cta - 1) * C::WARPS + warp) * C::LINE + lane;
#pragma unroll
    for (int r = 0; r < N; r++) acc[r] = carry_load(in + r * 32);
    __syncwarp();
    if (lane == 0) carry_rearm(flag);
  };
  // a chunk the next CTA completes: hand the running sum over
  auto carry_out = [&]() {
    double2* out = sch.carry + ((size_t)cta * C::WARPS + warp) * C::LINE + lane;
#pragma unroll
    for (int r = 0; r < N; r++) { out[r * 32] = acc[r]; acc[r] = make_double2(0.0, 0.0); }
    __threadfence();
    __syncwarp();
    if (lane == 0) carry_publish(sch.carry_flag + (size_t)cta * C::WARPS + warp);
  };

  if constexpr (!CUTS) {
    // whole-chunk schedule: nothing but whole chunks, one instance of the step body, bounds known at compile time
#pragma unroll 1
    for (int i = 0; i < nwhole; i++) {
      int t, cl, ey0, ey1;
      segment(i, t, cl, ey0, ey1);
      enter_tile(t);
      run_steps(std::true_type{}, 0, N);
      chunk_end(cl);
    }
  } else if constexpr (N <= 16) {
    // One instance of the step body for every segment: two of them (2 x 27 KB) would not share the 32 KB instruction
    // cache, and a small slab's CTA changes between them every few chunks (80 cells: 0.183 -> 0.173 ms per launch).
#pragma unroll 1
    for (int i = 0; i < nseg; i++) {
      int t, cl, ey0, ey1;
      segment(i, t, cl, ey0, ey1);
      enter_tile(t);
      if (ey0 > 0) carry_in();
      run_steps(std::false_type{}, ey0, ey1);
      if (ey1 == N) chunk_end(cl);
      else carry_out();
    }
  } else {
    // phase 0: the head of the chunk cut by g1 (handed to the next CTA); 1: the whole chunks; 2: the tail of the chunk
    // cut by g0 (continued from the previous CTA's running sum).  The general form of the step loop exists once, for 0
    // and 2; the whole chunks have the loop with compile-time bounds to themselves (N = 24 loses 2 % without).
    int i = 0;
#pragma unroll 1
    for (int phase = 0; phase < 3; phase++) {
      if (phase == 1) {
#pragma unroll 1
        for (int w = 0; w < nwhole; w++, i++) {
          int t, cl, ey0, ey1;
          segment(i, t, cl, ey0, ey1);
          enter_tile(t);
          run_steps(std::true_type{}, 0, N);
          chunk_end(cl);
        }
        continue;
      }
      if (phase == 0 ? seg_pre == 0 : seg_post == 0) continue;
      int t, cl, ey0, ey1;
      segment(i, t, cl, ey0, ey1);
      i++;
      enter_tile(t);
      if (phase == 2) carry_in();
      run_steps(std::false_type{}, ey0, ey1);
      if (phase == 2) chunk_end(cl);
      else carry_out();
    }
  }
}

template <int N, bool SPLIT, bool CUTS = true>
__global__ void __launch_bounds__(Batch3Cfg<N, SPLIT>::THREADS, 1)
qhat_batch3_kernel(const __grid_constant__ CUtensorMap tmapW, const double2* __restrict__ spec,
                   double2* __restrict__ parts, size_t part_stride, int cells, BatchSched sch, int cg_base) {
  qhat_batch3_body<N, SPLIT, CUTS>(tmapW, spec, parts, part_stride, cells, sch, cg_base, blockIdx.x);
}

// The main tiles and the split tiles of the remainder group in ONE launch: CTAs [0, schM.P) run the ordinary-tile code on
// schM, the rest the split-tile code on schS (group cg_split).  One CTA per SM either way, so an SM that is through with
// its ordinary share takes a split CTA at once -- no drain and refill of the whole device between two launches.
template <int N>
__global__ void __launch_bounds__(Batch3Cfg<N, false>::THREADS, 1)
qhat_batch3_pair_kernel(const __grid_constant__ CUtensorMap tmapM, const __grid_constant__ CUtensorMap tmapS,
                        const double2* __restrict__ spec, double2* __restrict__ parts, size_t part_stride, int cells_main,
                        int cells_all, BatchSched schM, BatchSched schS, int cg_split) {
  static_assert(Batch3Cfg<N, false>::THREADS == Batch3Cfg<N, true>::THREADS, "one block shape for both tile kinds");
  if (blockIdx.x < (unsigned)schM.P)
    qhat_batch3_body<N, false, true>(tmapM, spec, parts, part_stride, cells_main, schM, 0, blockIdx.x);
  else
    qhat_batch3_body<N, true, true>(tmapS, spec, parts, part_stride, cells_all, schS, cg_split, blockIdx.x - (unsigned)schM.P);
}

#ifndef SBTE_HOST_EMUL   // host launch code: not part of the host emulation
template <int N, bool SPLIT = false, bool CUTS = true>
static void launch_batch3_n(sbte_ctx* c, const CUtensorMap& tmap, const double2* spec, double2* parts, size_t part_stride, int cells,
                            const BatchSched& sch, int cg_base = 0) {
  using C = Batch3Cfg<N, SPLIT>;
  if (!CUTS && sch.cuts) { set_error("qhat_batch: this kernel instance takes whole-chunk schedules only"); return; }
  auto kern = qhat_batch3_kernel<N, SPLIT, CUTS>;
  static std::atomic<unsigned> configured{0};   // per device: function attributes belong to the device context
  if (!((configured.load() >> c->device) & 1u)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    configured.fetch_or(1u << c->device);
  }
  k2_mark(c);
  kern<<<sch.P, C::THREADS, C::SMEM, c->stream>>>(tmap, spec, parts, part_stride, cells, sch, cg_base);
  k2_mark(c);
  c->launches += 1;
}
#endif

#ifndef SBTE_HOST_EMUL   // host launch code: not part of the host emulation
// main + split tiles of an N = 16 slab in one launch (both schedules of the general, cut-anywhere kind)
bool qhat_batch_pair_supported(int N) {
  static const bool off = getenv("SBTE_NO_PAIR_LAUNCH") != nullptr;
  return N == 16 && !off;
}
void launch_qhat_batch_pair(sbte_ctx* c, const double2* spec, double2* parts, size_t part_stride, int cells_main, int cells_all,
                            const BatchSched& schM, const BatchSched& schS, int cg_split) {
  if (!c->tmap_ok) { set_error("qhat_batch: weight tensor map not initialised"); return; }
  if (c->N != 16) { set_error("qhat_batch: the paired launch exists for N = 16 only"); return; }
  constexpr size_t smem = Batch3Cfg<16, false>::SMEM > Batch3Cfg<16, true>::SMEM ? Batch3Cfg<16, false>::SMEM : Batch3Cfg<16, true>::SMEM;
  auto kern = qhat_batch3_pair_kernel<16>;
  static std::atomic<unsigned> configured{0};   // per device: function attributes belong to the device context
  if (!((configured.load() >> c->device) & 1u)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured.fetch_or(1u << c->device);
  }
  k2_mark(c);
  kern<<<schM.P + schS.P, Batch3Cfg<16, false>::THREADS, smem, c->stream>>>(schM.sym ? c->tmapWs : c->tmapW, schS.sym ? c->tmapWs16 : c->tmapW16,
                                                                           spec, parts, part_stride, cells_main, cells_all, schM, schS, cg_split);
  k2_mark(c);
  c->launches += 1;
}

void launch_qhat_batch2(sbte_ctx* c, const double2* spec, double2* parts, size_t part_stride, int cells,
                        const BatchSched& sch) {
  if (!c->tmap_ok) { set_error("qhat_batch: weight tensor map not initialised"); return; }
  const CUtensorMap& tm = sch.sym ? c->tmapWs : c->tmapW;
  switch (c->N) {
    case 8: launch_batch2_n<8>(c, spec, parts, part_stride, cells, sch); break;
    case 16: {
      // The line-ring kernel has no plane switch (no bubble at chunk ends) although it streams N + 7 lines per chunk;
      // SBTE_N16_PLANE=1 keeps the resident-plane kernel for A/B runs
      static const bool ring16 = getenv("SBTE_N16_PLANE") == nullptr;   // default: line ring (2.47 vs 2.69 ms at 640 cells)
      if (!ring16) launch_batch2_n<16>(c, spec, parts, part_stride, cells, sch);
      else if (sch.cuts) launch_batch3_n<16>(c, tm, spec, parts, part_stride, cells, sch);
      else launch_batch3_n<16, false, false>(c, tm, spec, parts, part_stride, cells, sch);
      break;
    }
    case 20: launch_batch3_n<20>(c, tm, spec, parts, part_stride, cells, sch); break;
    case 22: launch_batch3_n<22>(c, tm, spec, parts, part_stride, cells, sch); break;
    case 24: launch_batch3_n<24>(c, tm, spec, parts, part_stride, cells, sch); break;
    default: set_error("qhat_batch: unsupported N"); break;
  }
}

// the remainder group (at most 16 live cells) of an N = 16 slab: two columns per warp (Batch3Cfg<16, true>)
bool qhat_batch_split_supported(int N) {
  static const bool off = getenv("SBTE_NO_SPLIT16") != nullptr;
  return N == 16 && !off;
}
void launch_qhat_batch_split(sbte_ctx* c, const double2* spec, double2* parts, size_t part_stride, int cells,
                             const BatchSched& sch, int cg_base) {
  if (!c->tmap_ok) { set_error("qhat_batch: weight tensor map not initialised"); return; }
  if (c->N != 16) { set_error("qhat_batch: split tiles exist for N = 16 only"); return; }
  launch_batch3_n<16, true>(c, sch.sym ? c->tmapWs16 : c->tmapW16, spec, parts, part_stride, cells, sch, cg_base);
}
#endif

}  // namespace sbte
