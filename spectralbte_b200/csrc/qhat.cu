// spectralbte_b200/csrc/qhat.cu -- K2: the weighted spectral convolution
//     Q^[zeta] = sum_xi W[zeta][xi] * g^[xi] * f^[wrap(zeta + N/2 - xi)]
// (reference hot loop: /root/reference/src/collisions.c:127-165; wrap once per dimension :141-158).
//
// Three kernels:
//   qhat_generic : any N, any batch.  One CTA per (zeta row, cell); correctness path for sizes the
//                  tuned kernels do not cover (N = 6, 8, 12, 22, ...).
//   qhat_stream  : batch 1 (0D).  HBM-bound: every weight is read exactly once with 128-bit
//                  non-allocating loads, software-pipelined in registers; the operand planes
//                  g^[xi_x][.][.] and f^[X][.][.] are staged by TMA bulk copies (cp.async.bulk +
//                  mbarrier) into a double-buffered shared-memory ring; each thread owns a
//                  4-row x 2-column register tile so the Toeplitz structure in z gives 8 weights per
//                  7 operand reads.  NP operand pairs can share one weight pass (maxPreserve).
//   qhat_batch   : many cells share each weight (1D).  FP64-bound: lanes = cells, one warp per
//                  zeta (x,y) column, the whole N x N (zeta_z, xi_z) Toeplitz tile in registers.
//                  (qhat_batch.cu)
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"
#include "internal.h"

namespace sbte {

// ------------------------------------------------------------------------------------------
// generic
// ------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(256)
qhat_generic_kernel(const double* __restrict__ W, const double2* __restrict__ xi0, const double2* __restrict__ df0,
                    const double2* __restrict__ xi1, const double2* __restrict__ df1, double2* __restrict__ qhat,
                    int N) {
  __shared__ double red[2 * 32];
  const long n3 = (long)N * N * N;
  const int zeta = blockIdx.x;
  const long cell = blockIdx.y;
  const int n2 = N / 2;
  const int zx = zeta / (N * N), zy = (zeta / N) % N, zz = zeta % N;
  const double* w = W + (long)zeta * n3;
  const double2* g0 = xi0 + cell * n3;
  const double2* f0 = df0 + cell * n3;
  const double2* g1 = NP > 1 ? xi1 + cell * n3 : nullptr;
  const double2* f1 = NP > 1 ? df1 + cell * n3 : nullptr;
  double acc[2] = {0.0, 0.0};
  for (int xi = threadIdx.x; xi < n3; xi += blockDim.x) {
    const int ex = xi / (N * N), ey = (xi / N) % N, ez = xi % N;
    int x = zx + n2 - ex, y = zy + n2 - ey, z = zz + n2 - ez;
    if (x < 0) x += N; else if (x > N - 1) x -= N;
    if (y < 0) y += N; else if (y > N - 1) y -= N;
    if (z < 0) z += N; else if (z > N - 1) z -= N;
    const int idx = z + N * (y + N * x);
    double2 p = cmul(g0[xi], f0[idx]);
    if (NP > 1) {
      const double2 q = cmul(g1[xi], f1[idx]);
      p.x += q.x; p.y += q.y;
    }
    const double wv = w[xi];
    acc[0] = fma(wv, p.x, acc[0]);
    acc[1] = fma(wv, p.y, acc[1]);
  }
  block_reduce_sum<2>(acc, red);
  if (threadIdx.x == 0) qhat[cell * n3 + zeta] = make_double2(acc[0], acc[1]);
}

void launch_qhat_generic(sbte_ctx* c, int npairs, const QhatPair* pairs, double2* qhat, int batch) {
  dim3 grid((unsigned)c->n3, batch);
  k2_mark(c);
  if (npairs == 1)
    qhat_generic_kernel<1><<<grid, 256, 0, c->stream>>>(c->d_W, pairs[0].xi_side, pairs[0].dif_side, nullptr, nullptr,
                                                        qhat, c->N);
  else
    qhat_generic_kernel<2><<<grid, 256, 0, c->stream>>>(c->d_W, pairs[0].xi_side, pairs[0].dif_side, pairs[1].xi_side,
                                                        pairs[1].dif_side, qhat, c->N);
  k2_mark(c);
  c->launches += 1;
}

// ------------------------------------------------------------------------------------------
// stream kernel
// ------------------------------------------------------------------------------------------
constexpr int largest_divisor_le(int n, int cap) {
  int best = 1;
  for (int d = 1; d <= cap; d++)
    if (n % d == 0) best = d;
  return best;
}

// WT: target warps per CTA (8 for one operand pair; 16 for two pairs, whose 128 KB plane ring allows only one
// CTA per SM so the CTA itself has to bring the latency-hiding warps)
template <int N, int WT = 8>
struct StreamCfg {
  static constexpr int HALF = N / 2;            // 2-column groups per N-long weight row segment
  static constexpr int RGW = 32 / HALF;         // 4-row groups handled by one warp
  static constexpr int ROWS_W = RGW * 4;        // rows per warp
  static constexpr int RB = N / ROWS_W;         // warps that tile the N rows of one zeta (x,y) column
  static constexpr int PH = largest_divisor_le(N, (WT / RB) < 1 ? 1 : (WT / RB));  // xi_y phases (warps sharing the same rows)
  static constexpr int NWARP = RB * PH;
  static constexpr int THREADS = NWARP * 32;
  static constexpr int SPC = N / PH;            // steps per xi_x chunk per warp
  static constexpr int PLANE = N * N;           // complex elements per operand plane
  static_assert(N % 2 == 0 && 32 % HALF == 0 || N == 24, "unsupported N");
  static_assert(N % ROWS_W == 0, "rows must tile");
  static_assert(N % PH == 0, "phases must tile");
};

// operand pairs of one weight pass: xi-side spectrum g^[xi] and dif-side spectrum f^[zeta - xi] of up to four products
struct StreamOperands {
  const double2* xi[4];
  const double2* df[4];
};

// SYM: W is the symmetrised tensor Ws and only the representative xi_x planes are visited (f == g only).
// TP ("transposed pairing", NP = 2, f == g): for a tensor that is invariant under swapping the x and y axes of both
// indices -- W[(zy,zx,zz)][(ey,ex,ez)] = W[(zx,zy,zz)][(ex,ey,ez)], which isotropic weights are (src/weights.c:265-281:
// they depend on |zeta|, |xi|, |xi - zeta/2| and a product of per-axis trapezoid factors) -- the rows of column (zy, zx)
// are the rows of column (zx, zy) read against the x<->y transposed spectrum:
//     Q^(zy,zx,zz) = sum_xi W[(zx,zy,zz)][xi] F(xi) F(sigma(xi)),   F(ex,ey,ez) = f^(ey,ex,ez).
// Only the columns zx >= zy are streamed (N(N+1)/2 of N^2: 51.6 % of the bytes at N = 32); pair A is the spectrum itself
// and accumulates column (zx,zy), pair B is the transposed spectrum and accumulates column (zy,zx): two outputs from one
// weight, 12 FP64 instructions per weight.  The library verifies the invariance of the bound tensor before using it.
template <int N, int NP, int DEPTH, bool SYM, int WT, bool TP>
__global__ void __launch_bounds__(StreamCfg<N, WT>::THREADS, (N == 32 && NP == 1 && DEPTH <= 2) ? 2 : 1)
qhat_stream_kernel(const double* __restrict__ W, StreamOperands ops, double2* __restrict__ qhat, int nsplit) {
  // NP = 2 with TP: one pair per orientation (ComputeQ).  NP = 4 with TP: two summed pairs per orientation
  // (ComputeQ_maxPreserve): pairs 0,1 against the spectra, pairs 2,3 against their x<->y transposes.
  static_assert(!TP || NP == 2 || NP == 4, "transposed pairing streams two or four operand pairs");
  static_assert(NP != 4 || TP, "four pairs only as two orientations of two");
  using C = StreamCfg<N, WT>;
  constexpr int HALF = C::HALF, PLANE = C::PLANE;
  constexpr long n3 = (long)N * N * N;
  // Eight planes of 16 KB do not fit twice: with four pairs a stage holds HALF a chunk -- the xi_y lines
  // [sub N/2, (sub+1) N/2) of every xi-side plane and the N/2 dif-side lines they meet (contiguous modulo N).
  constexpr int SUB = (NP == 4 && (size_t)4 * NP * PLANE * sizeof(double2) > 200 * 1024) ? 2 : 1;   // staged sub-chunks per xi_x chunk
  constexpr int LPS = N / SUB;                      // lines per staged plane
  constexpr int SPLANE = LPS * N;                   // complex elements per staged plane
  constexpr int SPS = C::SPC / SUB;                 // steps per sub-chunk
  static_assert(C::SPC % SUB == 0 && (SUB == 1 || C::PH * SPS == LPS), "a sub-chunk covers a contiguous range of xi_y");
  constexpr uint32_t STAGE_ELEMS = 2 * NP * SPLANE;  // per stage: NP x (xi-side plane, dif-side plane)
  extern __shared__ __align__(128) unsigned char smraw[];
  double2* planes = reinterpret_cast<double2*>(smraw);                       // [2][NP][2][SPLANE]
  uint64_t* full = reinterpret_cast<uint64_t*>(smraw + 2 * STAGE_ELEMS * sizeof(double2));
  int* done_cnt = reinterpret_cast<int*>(full + 2);   // [2] warps that have finished reading a stage

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rb = warp % C::RB, ph = warp / C::RB;
  const int cp = lane % HALF, rgw = lane / HALF;
  const bool active = rgw < C::RGW;
  // nsplit > 1: the xi_x planes of a column are shared between nsplit CTAs, each writing its own partial
  // spectrum qhat[part]; the inverse transform adds the parts in order.  More, shorter CTAs = a shorter tail.
  constexpr int NCOL = TP ? N * (N + 1) / 2 : N * N;   // columns streamed
  const int part = blockIdx.x / NCOL, cidx = blockIdx.x - part * NCOL;
  int zx, zy;
  if (TP) {   // cidx enumerates the columns zx >= zy: cidx = zx (zx + 1) / 2 + zy
    zx = (int)((sqrtf(8.0f * (float)cidx + 1.0f) - 1.0f) * 0.5f);
    while ((zx + 1) * (zx + 2) / 2 <= cidx) zx++;
    while (zx * (zx + 1) / 2 > cidx) zx--;
    zy = cidx - zx * (zx + 1) / 2;
  } else {
    zx = cidx / N; zy = cidx % N;
  }
  const int colid = zx * N + zy;
  const int nchunk_all = SYM ? sym_nrep(N, zx) : N;
  const int cbeg = part * nchunk_all / nsplit;
  const int nchunk = (part + 1) * nchunk_all / nsplit - cbeg;   // xi_x planes visited by this CTA
  auto chunk_ex = [&](int c) { return SYM ? sym_rep(N, zx, cbeg + c) : cbeg + c; };
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

  // first dif-side line of a sub-chunk in ascending order: the lines Y = zy + N/2 - ey, ey in [sub LPS, (sub+1) LPS)
  auto sub_ylo = [&](int sub) {
    if (SUB == 1) return 0;
    int y = zy + N / 2 - (sub * LPS + LPS - 1);
    y %= N;
    return y < 0 ? y + N : y;
  };
  auto issue_chunk = [&](int q) {  // one thread: stage the operand planes of sub-chunk q = chunk * SUB + sub
    const int s = q & 1;
    const int chunk = q / SUB, sub = q - chunk * SUB;
    const int ex = chunk_ex(chunk);
    int X = zx + N / 2 - ex;
    if (X < 0) X += N; else if (X > N - 1) X -= N;
    double2* dst = planes + (size_t)s * STAGE_ELEMS;
    mbar_arrive_expect_tx(&full[s], STAGE_ELEMS * (uint32_t)sizeof(double2));
    const int ylo = sub_ylo(sub);
    const int n1 = (SUB == 1) ? N : min(LPS, N - ylo);      // lines before the wrap
#pragma unroll
    for (int p = 0; p < NP; p++) {
      tma_bulk_g2s(dst + (2 * p) * SPLANE, ops.xi[p] + (size_t)ex * PLANE + (size_t)sub * SPLANE, SPLANE * sizeof(double2),
                   &full[s]);
      const double2* dsrc = ops.df[p] + (size_t)X * PLANE;
      tma_bulk_g2s(dst + (2 * p + 1) * SPLANE, dsrc + (size_t)ylo * N, (uint32_t)(n1 * N * sizeof(double2)), &full[s]);
      if (SUB > 1 && n1 < LPS)
        tma_bulk_g2s(dst + (2 * p + 1) * SPLANE + n1 * N, dsrc, (uint32_t)((LPS - n1) * N * sizeof(double2)), &full[s]);
    }
  };

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    done_cnt[0] = 0;
    done_cnt[1] = 0;
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();   // the inverse transform may stage its twiddles while this grid drains

  // weight stream: rows zeta = (zx, zy, r0 + j), j = 0..3; this thread's 16 bytes sit at column c0
  const double* wrow = W + ((long)colid * N + r0) * n3 + c0;
  const int NIT = nchunk * C::SPC;  // iterations of this warp over (visited xi_x, xi_y)
  double2 wb[DEPTH][4];
  auto load_w = [&](int it, double2 (&dst)[4]) {
    const int ex = chunk_ex(it / C::SPC), ey = ph + (it % C::SPC) * C::PH;
    const double* p = wrow + ((long)ex * N + ey) * N;
#pragma unroll
    for (int j = 0; j < 4; j++) dst[j] = ldg_stream_f64x2(p + (long)j * n3);
  };
  if (active) {
#pragma unroll
    for (int d = 0; d < DEPTH; d++)
      if (d < NIT) load_w(d, wb[d]);
  }
  // the weights above do not depend on the forward transform; the operand planes below do
  pdl_wait();
  if (tid == 0 && nchunk > 0) {   // a CTA with no xi_x planes (nsplit > planes of this column) writes a zero partial sum
    issue_chunk(0);
    if (nchunk * SUB > 1) issue_chunk(1);
  }

  double2 acc[4], accB[TP ? 4 : 1];
#pragma unroll
  for (int j = 0; j < 4; j++) acc[j] = make_double2(0.0, 0.0);
#pragma unroll
  for (int j = 0; j < (TP ? 4 : 1); j++) accB[j] = make_double2(0.0, 0.0);

  for (int it0 = 0; it0 < NIT; it0 += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; d++) {
      const int it = it0 + d;
      if (it >= NIT) break;   // block-uniform (NIT need not be a multiple of DEPTH)
      const int chunk = it / C::SPC, step = it % C::SPC;
      const int sub = step / SPS, q = chunk * SUB + sub;      // staged sub-chunk this step reads
      const int s = q & 1;
      if (step % SPS == 0) mbar_wait(&full[s], (q >> 1) & 1);
      const int ey = ph + step * C::PH;
      int Y = zy + N / 2 - ey;
      if (Y < 0) Y += N; else if (Y > N - 1) Y -= N;
      int yl = Y - sub_ylo(sub);                               // line of the staged dif-side plane
      if (yl < 0) yl += N;
      const int el = ey - sub * LPS;                           // line of the staged xi-side plane
      const double2* st = planes + (size_t)s * STAGE_ELEMS;
      // sum of the products of operand pairs p0 and (if p1 >= 0) p1 for this thread's 4 x 2 tile
      auto products = [&](int p0, int p1, double2 (&p)[4][2]) {
        {
          const double2* gl = st + (2 * p0) * SPLANE + el * N;
          const double2* fl = st + (2 * p0 + 1) * SPLANE + yl * N;
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
        if (p1 >= 0) {
          const double2* gl = st + (2 * p1) * SPLANE + el * N;
          const double2* fl = st + (2 * p1 + 1) * SPLANE + yl * N;
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
      };
      if (active) {
        double2 p[4][2];
        // orientation e: one pair (NP = 1, or NP = 2 with TP) or two summed pairs (maxPreserve: NP = 2 without TP, NP = 4)
        products(0, (NP == 4 || (NP == 2 && !TP)) ? 1 : -1, p);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          cmac(acc[j], wb[d][j].x, p[j][0]);
          cmac(acc[j], wb[d][j].y, p[j][1]);
        }
        if (TP && zx != zy) {   // the same weights against the transposed spectra: the rows of column (zy, zx)
          products(NP == 4 ? 2 : 1, NP == 4 ? 3 : -1, p);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            cmac(accB[j], wb[d][j].x, p[j][0]);
            cmac(accB[j], wb[d][j].y, p[j][1]);
          }
        }
        if (it + DEPTH < NIT) load_w(it + DEPTH, wb[d]);
      }
      if (step % SPS == SPS - 1) {
        // No CTA-wide barrier at the end of a (sub-)chunk: the LAST warp to finish reading stage s refills it with the
        // planes of sub-chunk q + 2 (its full barrier cannot complete that phase before every warp has passed this
        // point, so a fast warp simply waits there).  With one CTA per SM a __syncthreads here idled every warp.
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          if (atomicAdd(&done_cnt[s], 1) == C::NWARP - 1) {
            __threadfence_block();
            done_cnt[s] = 0;
            if (q + 2 < nchunk * SUB) issue_chunk(q + 2);
          }
        }
      }
    }
  }

  // deterministic cross-thread reduction: [thread][4] partial sums -> N rows (reuses the plane ring)
  double2* red = planes;
  __syncthreads();   // every warp has left the main loop: the plane ring is free
#pragma unroll
  for (int round = 0; round < (TP ? 2 : 1); round++) {
    if (round == 1) {
      if (zx == zy) break;   // a diagonal column is its own transpose: pair B would rewrite the rows pair A wrote
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
      red[tid * 4 + j] = active ? (round == 0 ? acc[j] : accB[TP ? j : 0]) : make_double2(0.0, 0.0);
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
      const int out_col = round == 0 ? colid : zy * N + zx;
      qhat[(long)part * n3 + (long)out_col * N + r] = make_double2(sr, si);
    }
  }
}

bool qhat_stream_supported(int N) { return N == 16 || N == 24 || N == 32; }

template <int N, int NP, int DEPTH, bool SYM, int WT = 8, bool TP = false>
static void launch_stream_inst(sbte_ctx* c, const double* W, const QhatPair* pairs, double2* qhat, int nsplit = 1) {
  using C = StreamCfg<N, WT>;
  const size_t full2 = (size_t)2 * 2 * NP * C::PLANE * sizeof(double2);   // two stages of whole planes
  const size_t smem = ((NP == 4 && full2 > 200 * 1024) ? full2 / 2 : full2) + 64;
  auto kern = qhat_stream_kernel<N, NP, DEPTH, SYM, WT, TP>;
  StreamOperands ops = {};
  for (int q = 0; q < NP; q++) { ops.xi[q] = pairs[q].xi_side; ops.df[q] = pairs[q].dif_side; }
  static std::atomic<unsigned> configured{0};   // per device: function attributes belong to the device context
  if (!((configured.load() >> c->device) & 1u)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured.fetch_or(1u << c->device);
  }
  k2_mark(c);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((TP ? N * (N + 1) / 2 : N * N) * nsplit);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // overlap the prologue with the transform before it
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, W, ops, qhat, nsplit);
  k2_mark(c);
  c->launches += 1;
}

template <int N, bool SYM>
static void launch_stream_n(sbte_ctx* c, const double* W, int npairs, const QhatPair* pairs, double2* qhat, int depth,
                            int nsplit) {
  if (npairs == 1) {
    if (depth >= 4) launch_stream_inst<N, 1, 4, SYM>(c, W, pairs, qhat, nsplit);
    else launch_stream_inst<N, 1, 2, SYM>(c, W, pairs, qhat, nsplit);
  } else {
    // two operand pairs need 128 KB of plane ring => one CTA per SM.  Either 8 warps with four weight tiles per
    // thread in flight, or (default) 16 warps with two: the same 64 KB per SM in flight, twice the warps to
    // overlap the shared-memory operand reads, the FP64 pipe and the weight stream
    if (getenv("SBTE_MP_NARROW")) launch_stream_inst<N, 2, 4, SYM>(c, W, pairs, qhat, nsplit);
    else launch_stream_inst<N, 2, 2, SYM, 16>(c, W, pairs, qhat, nsplit);
  }
}

// sym: stream the symmetrised tensor (caller guarantees xi-side and dif-side operands describe f == g)
// nsplit: partial spectra qhat[0..nsplit) of n3 elements each, to be added by the caller
void launch_qhat_stream(sbte_ctx* c, int npairs, const QhatPair* pairs, double2* qhat, int depth, bool sym, int nsplit) {
  const double* W = sym ? c->d_Ws : c->d_W;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > c->N / 2) nsplit = c->N / 2;   // every column has at least N/2 representative planes: no CTA stays empty
  switch (c->N) {
    case 16: sym ? launch_stream_n<16, true>(c, W, npairs, pairs, qhat, depth, nsplit) : launch_stream_n<16, false>(c, W, npairs, pairs, qhat, depth, nsplit); break;
    case 24: sym ? launch_stream_n<24, true>(c, W, npairs, pairs, qhat, depth, nsplit) : launch_stream_n<24, false>(c, W, npairs, pairs, qhat, depth, nsplit); break;
    case 32: sym ? launch_stream_n<32, true>(c, W, npairs, pairs, qhat, depth, nsplit) : launch_stream_n<32, false>(c, W, npairs, pairs, qhat, depth, nsplit); break;
    default: set_error("qhat_stream: unsupported N"); break;
  }
}

// transposed pairing (TP): pairs[0] = the spectrum (both sides), pairs[1] = its x<->y transpose (both sides); f == g only
bool qhat_stream_tp_supported(int N) { return N == 32 || N == 16; }
// npairs = 2: ComputeQ (pairs[0] spectrum, pairs[1] transposed spectrum); npairs = 4: ComputeQ_maxPreserve (pairs[0..1]
// its two summed products, pairs[2..3] the same against the transposed spectra)
void launch_qhat_stream_tp(sbte_ctx* c, int npairs, const QhatPair* pairs, double2* qhat, bool sym, int nsplit) {
  const double* W = sym ? c->d_Ws : c->d_W;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > c->N / 2) nsplit = c->N / 2;
  if (npairs == 4) {
    switch (c->N) {
      case 16: sym ? launch_stream_inst<16, 4, 2, true, 16, true>(c, W, pairs, qhat, nsplit) : launch_stream_inst<16, 4, 2, false, 16, true>(c, W, pairs, qhat, nsplit); break;
      case 32: sym ? launch_stream_inst<32, 4, 2, true, 16, true>(c, W, pairs, qhat, nsplit) : launch_stream_inst<32, 4, 2, false, 16, true>(c, W, pairs, qhat, nsplit); break;
      default: set_error("qhat_stream (transposed pairing): unsupported N"); break;
    }
    return;
  }
  // 16 warps with two weight tiles in flight per thread; 8 warps with four were 15 % slower (profiles/r02_tp_tune.txt)
  switch (c->N) {
    case 16: sym ? launch_stream_inst<16, 2, 2, true, 16, true>(c, W, pairs, qhat, nsplit) : launch_stream_inst<16, 2, 2, false, 16, true>(c, W, pairs, qhat, nsplit); break;
    case 32: sym ? launch_stream_inst<32, 2, 2, true, 16, true>(c, W, pairs, qhat, nsplit) : launch_stream_inst<32, 2, 2, false, 16, true>(c, W, pairs, qhat, nsplit); break;
    default: set_error("qhat_stream (transposed pairing): unsupported N"); break;
  }
}

// dst(x, y, .) = src(y, x, .) for a parity-split (or natural) spectrum: lines of N complex values move as a whole
__global__ void transpose_xy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, int N) {
  const int line = blockIdx.x;                 // destination line (x, y)
  const int x = line / N, y = line - x * N;
  const double2* s = src + ((size_t)y * N + x) * N;
  double2* d = dst + (size_t)line * N;
  for (int k = threadIdx.x; k < N; k += blockDim.x) d[k] = s[k];
}
void launch_transpose_xy(sbte_ctx* c, const double2* src, double2* dst) {
  transpose_xy_kernel<<<c->N * c->N, 32, 0, c->stream>>>(src, dst, c->N);
  c->launches += 1;
}

// max |W[(zx,zy,zz)][(ex,ey,ez)] - W[(zy,zx,zz)][(ey,ex,ez)]| and max |W| over the tensor, as bit patterns of
// non-negative doubles (atomicMax on unsigned long long orders them correctly); out[0] = difference, out[1] = magnitude
__global__ void xy_symmetry_kernel(const double* __restrict__ W, unsigned long long* __restrict__ out, int N) {
  const size_t n3 = (size_t)N * N * N, total = n3 * n3;
  double dmax = 0.0, amax = 0.0;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int zeta = (int)(e / n3), xi = (int)(e - (size_t)zeta * n3);
    const int zx = zeta / (N * N), zy = (zeta / N) % N, zz = zeta % N;
    if (zx < zy) continue;                     // every unordered pair once
    const int ex = xi / (N * N), ey = (xi / N) % N, ez = xi % N;
    const double a = W[e];
    const double b = W[(((size_t)zy * N + zx) * N + zz) * n3 + ((size_t)ey * N + ex) * N + ez];
    dmax = fmax(dmax, fabs(a - b));
    amax = fmax(amax, fmax(fabs(a), fabs(b)));
  }
  for (int o = 16; o > 0; o >>= 1) {
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, (unsigned long long)__double_as_longlong(dmax));
    atomicMax(out + 1, (unsigned long long)__double_as_longlong(amax));
  }
}
// returns 0 and the two maxima; synchronises the context's stream
int weights_xy_symmetry(sbte_ctx* c, const double* W, double* max_diff, double* max_abs) {
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, 16) != cudaSuccess) return 1;
  cudaMemsetAsync(d, 0, 16, c->stream);
  xy_symmetry_kernel<<<148 * 16, 256, 0, c->stream>>>(W, d, c->N);
  c->launches += 1;
  unsigned long long h[2] = {0, 0};
  cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, c->stream);
  const cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  if (e != cudaSuccess) return 1;
  memcpy(max_diff, &h[0], 8);
  memcpy(max_abs, &h[1], 8);
  return 0;
}

// Ws[zeta][xi] = W[zeta][xi] + W[zeta][sigma(xi)] on representative planes, W on self-paired planes, 0 elsewhere
__global__ void symmetrize_weights_kernel(const double* __restrict__ W, double* __restrict__ Ws, int N) {
  const size_t n3 = (size_t)N * N * N, total = n3 * n3;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int zeta = (int)(e / n3), xi = (int)(e - (size_t)zeta * n3);
    const int zx = zeta / (N * N), zy = (zeta / N) % N, zz = zeta % N;
    const int ex = xi / (N * N), ey = (xi / N) % N, ez = xi % N;
    const int X = (zx + N / 2 - ex + N) % N, Y = (zy + N / 2 - ey + N) % N, Z = (zz + N / 2 - ez + N) % N;
    double v;
    if (ex < X) v = W[e] + W[(size_t)zeta * n3 + ((size_t)X * N + Y) * N + Z];
    else if (ex == X) v = W[e];
    else v = 0.0;
    Ws[e] = v;
  }
}

void launch_symmetrize_weights(sbte_ctx* c, const double* W, double* Ws) {
  symmetrize_weights_kernel<<<148 * 16, 256, 0, c->stream>>>(W, Ws, c->N);
  c->launches += 1;
}

}  // namespace sbte
