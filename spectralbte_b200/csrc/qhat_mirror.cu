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
// (tests/test_mirror_emulation_cpu.py); the kernel itself has not run on a GPU yet, hence off by default.
#include <math.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "internal.h"
#include "mirror.cuh"

namespace sbte {

bool qhat_mirror_enabled(int N) {
  static const bool on = getenv("SBTE_MIRROR") != nullptr && atoi(getenv("SBTE_MIRROR")) != 0;
  return on && (N == 8 || N == 16);
}
int qhat_mirror_pairs(int N) { return (N >= 16) ? 4 : 2; }

__global__ void symmetrize_weights_mirror_kernel(const double* __restrict__ W, double* __restrict__ Ws2, int N) {
  const size_t n3 = (size_t)N * N * N, total = n3 * n3;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t zeta = e / n3;
    Ws2[e] = mirror_sym_weight(W, N, zeta, e - zeta * n3);
  }
}

void launch_symmetrize_weights_mirror(sbte_ctx* c, const double* W, double* Ws2) {
  symmetrize_weights_mirror_kernel<<<148 * 16, 256, 0, c->stream>>>(W, Ws2, c->N);
  c->launches += 1;
}

struct MirrorPhases {
  double2 t[5];   // theta^m, m = 0..4
};

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
                   const MirrorTile* __restrict__ tiles, MirrorPhases ph) {
  using C = MirrorCfg<N>;
  constexpr long n3 = (long)N * N * N;
  constexpr int S = C::STAGES, RH = C::RH;
  extern __shared__ __align__(128) unsigned char smraw[];
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
    if (C::REG_SPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // 40*128 + 232*256 = 64512 = 384*168
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
  if (C::REG_SPLIT) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int pair = warp % C::PAIRS, half = warp / C::PAIRS;
  double2 accA[RH], accB[RH];
#pragma unroll
  for (int r = 0; r < RH; r++) { accA[r] = make_double2(0.0, 0.0); accB[r] = make_double2(0.0, 0.0); }

  int cur_t = -1, cur_cg = -1, cur_X = -1, epoch = -1;
  int zx = 0, zyA = -1, zyB = -1;

  auto flush = [&]() {
    const int rb = cur_t / G, cg = cur_t - rb * G;
    const long cell = (long)cg * 32 + lane;
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
      const int m = (ex == 0) + (ey == 0) + (X == 0) + (Y == 0);
      const double2 phi = (m == 0) ? ph.t[0] : (m == 1) ? ph.t[1] : (m == 2) ? ph.t[2] : (m == 3) ? ph.t[3] : ph.t[4];
      if (half == 0) mirror_step<N, 0, RH>(accA, accB, fl, 32, gl, 32, wA, wB, ph.t[1], phi);
      else mirror_step<N, RH, RH>(accA, accB, fl, 32, gl, 32, wA, wB, ph.t[1], phi);
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

template <int N>
static void launch_mirror_n(sbte_ctx* c, const double2* spec, double2* parts, size_t part_stride, int cells,
                            const BatchSched& sch) {
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
  kern<<<sch.P, C::THREADS, C::SMEM, c->stream>>>(sch.sym ? c->tmapMs : c->tmapM, spec, parts, part_stride, cells, sch,
                                                   c->d_mtiles, ph);
  k2_mark(c);
  c->launches += 1;
}

void launch_qhat_mirror(sbte_ctx* c, const double2* spec, double2* parts, size_t part_stride, int cells,
                        const BatchSched& sch) {
  if (!c->mirror_ok) { set_error("qhat_mirror: tensor maps / tile table not initialised"); return; }
  switch (c->N) {
    case 8: launch_mirror_n<8>(c, spec, parts, part_stride, cells, sch); break;
    case 16: launch_mirror_n<16>(c, spec, parts, part_stride, cells, sch); break;
    default: set_error("qhat_mirror: unsupported N"); break;
  }
}

}  // namespace sbte
