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
#include "common.cuh"
#include "internal.h"

namespace sbte {

template <int N>
struct BatchCfg {
  static constexpr int COLS = (N >= 16) ? 8 : 4;     // zeta_y columns (= warps) per CTA; must be < N
  static constexpr int THREADS = COLS * 32;
  static constexpr int ROWS = COLS * N;              // weight rows per CTA
  static constexpr int LINE = N * 32;                // double2 per operand line (N modes x 32 cells)
  static constexpr int PLANE = N * LINE;             // double2 per operand plane
  static constexpr int STAGES = 3;
  static constexpr size_t STAGE_BYTES = (size_t)LINE * 16 + (size_t)ROWS * N * 8;
  static constexpr size_t SMEM = (size_t)PLANE * 16 + STAGES * STAGE_BYTES + 128;
  static_assert(COLS < N && N % COLS == 0, "columns must tile zeta_y");
};

template <int N>
__global__ void __launch_bounds__(BatchCfg<N>::THREADS, 1)
qhat_batch_kernel(const __grid_constant__ CUtensorMap tmapW, const double2* __restrict__ spec,
                  double2* __restrict__ qhat, int cells) {
  using C = BatchCfg<N>;
  constexpr long n3 = (long)N * N * N;
  constexpr int NSTEP = N * N;
  extern __shared__ __align__(128) unsigned char smraw[];
  double2* plane = reinterpret_cast<double2*>(smraw);                                  // [N][N][32]
  unsigned char* stage0 = smraw + (size_t)C::PLANE * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + C::STAGES * C::STAGE_BYTES);  // [0..2] stage, [3] plane
  auto stage_line = [&](int s) { return reinterpret_cast<double2*>(stage0 + (size_t)s * C::STAGE_BYTES); };
  auto stage_w = [&](int s) {
    return reinterpret_cast<double*>(stage0 + (size_t)s * C::STAGE_BYTES + (size_t)C::LINE * 16);
  };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cg = blockIdx.x;                       // cell group
  const int q0 = blockIdx.y * C::COLS;             // first zeta (x,y) column of this CTA
  const int zx = q0 / N, zy = (q0 % N) + warp;     // this warp's column (COLS divides N: same zeta_x)
  const double2* gspec = spec + (size_t)cg * n3 * 32;

  auto issue_stage = [&](int step) {
    const int s = step % C::STAGES;
    mbar_arrive_expect_tx(&bars[s], (uint32_t)C::STAGE_BYTES);
    tma_bulk_g2s(stage_line(s), gspec + (size_t)step * C::LINE, C::LINE * 16, &bars[s]);
    tma_tensor2d_g2s(stage_w(s), &tmapW, step * N, q0 * N, &bars[s]);
  };
  auto issue_plane = [&](int chunk) {
    int X = zx + N / 2 - chunk;
    if (X < 0) X += N; else if (X > N - 1) X -= N;
    mbar_arrive_expect_tx(&bars[3], (uint32_t)(C::PLANE * 16));
    for (int y = 0; y < N; y++)
      tma_bulk_g2s(plane + (size_t)y * C::LINE, gspec + ((size_t)X * N + y) * C::LINE, C::LINE * 16, &bars[3]);
  };

  if (tid == 0) {
    for (int b = 0; b < 4; b++) mbar_init(&bars[b], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    issue_plane(0);
    for (int s = 0; s < C::STAGES; s++) issue_stage(s);
  }

  double2 acc[N];
#pragma unroll
  for (int r = 0; r < N; r++) acc[r] = make_double2(0.0, 0.0);

  for (int step = 0; step < NSTEP; step++) {
    const int chunk = step / N, ey = step - chunk * N;
    const int s = step % C::STAGES;
    if (ey == 0) mbar_wait(&bars[3], chunk & 1);
    mbar_wait(&bars[s], (step / C::STAGES) & 1);

    int Y = zy + N / 2 - ey;
    if (Y < 0) Y += N; else if (Y > N - 1) Y -= N;
    const double2* fl = plane + (size_t)Y * C::LINE + lane;
    const double2* gl = stage_line(s) + lane;
    const double* wt = stage_w(s) + warp * N * N;     // [row r][col c], warp-uniform addresses

    double2 fr[N];
#pragma unroll
    for (int z = 0; z < N; z++) fr[z] = fl[z * 32];
#pragma unroll
    for (int c = 0; c < N; c += 2) {
      const double2 g0 = gl[c * 32], g1 = gl[(c + 1) * 32];
#pragma unroll
      for (int r = 0; r < N; r++) {
        const double2 w2 = *reinterpret_cast<const double2*>(wt + r * N + c);
        const double2 p0 = cmul(g0, fr[(r + N / 2 - c + N) % N]);
        const double2 p1 = cmul(g1, fr[(r + N / 2 - c - 1 + N) % N]);
        cmac(acc[r], w2.x, p0);
        cmac(acc[r], w2.y, p1);
      }
    }

    __syncthreads();  // all warps done with stage s (and, at ey == N-1, with the plane)
    if (tid == 0) {
      if (step + C::STAGES < NSTEP) issue_stage(step + C::STAGES);
      if (ey == N - 1 && chunk + 1 < N) issue_plane(chunk + 1);
    }
  }

  const long cell = (long)cg * 32 + lane;
  if (cell < cells) {
    double2* out = qhat + cell * n3 + ((long)zx * N + zy) * N;
#pragma unroll
    for (int r = 0; r < N; r++) out[r] = acc[r];
  }
}

bool qhat_batch_supported(int N) { return N == 8 || N == 16; }

template <int N>
static void launch_batch_n(sbte_ctx* c, const double2* spec, double2* qhat, int cells) {
  using C = BatchCfg<N>;
  auto kern = qhat_batch_kernel<N>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    configured = true;
  }
  const int groups = (cells + 31) / 32;
  dim3 grid(groups, N * N / C::COLS);
  k2_mark(c);
  kern<<<grid, C::THREADS, C::SMEM, c->stream>>>(c->tmapW, spec, qhat, cells);
  k2_mark(c);
  c->launches += 1;
}

void launch_qhat_batch(sbte_ctx* c, const double2* spec, double2* qhat, int cells) {
  if (!c->tmap_ok) { set_error("qhat_batch: weight tensor map not initialised"); return; }
  switch (c->N) {
    case 8: launch_batch_n<8>(c, spec, qhat, cells); break;
    case 16: launch_batch_n<16>(c, spec, qhat, cells); break;
    default: set_error("qhat_batch: unsupported N"); break;
  }
}

}  // namespace sbte
