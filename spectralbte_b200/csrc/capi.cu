// spectralbte_b200/csrc/capi.cu -- the sbte_* C ABI (include/sbte_b200.h, section 2): context,
// weights, device-pointer operations and the 0D step.  The drop-in symbols live in dropin.cu, the
// 1D slab in slab.cu.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>

#include "../../include/sbte_b200.h"
#include "common.cuh"
#include "internal.h"

namespace sbte {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

void k2_mark(sbte_ctx* c) {
  if (!c->k2_prof) return;
  if (c->k2_ev_used == c->k2_ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->k2_ev.push_back(e);
  }
  cudaEventRecord(c->k2_ev[c->k2_ev_used++], c->stream);
}

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      sbte::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                            \
      return 1;                                                                                       \
    }                                                                                                 \
  } while (0)

static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// Gram matrix of the moment functionals and its scaled-pivot LU (src/conserve.c:268-317, 89-168).
// The pivot row is the first row that improves on the diagonal entry (:115-126); row swaps touch
// only columns >= k (:140-145), which the solve in conserve.cu mirrors.
static int build_lu(const sbte_ctx* c, ConsLU* out) {
  const int N = c->N, n = 5;
  const double* v = c->v.data();
  const double* wt = c->wt.data();
  double* A = out->a;
  double s[5];
  for (int a = 0; a < n; a++)
    for (int b = 0; b < n; b++) {
      double acc = 0.0;
      for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
          for (int k = 0; k < N; k++) {
            const double pre = wt[i] * wt[j] * wt[k] * c->dv * c->dv * c->dv * 1.0;
            const double row[5] = {pre, pre * v[i], pre * v[j], pre * v[k],
                                   pre * 0.5 * (v[i] * v[i] + v[j] * v[j] + v[k] * v[k])};
            acc += row[a] * row[b];
          }
      A[a * n + b] = acc;
    }
  for (int i = 0; i < n; i++) {
    s[i] = fabs(A[i * n]);
    for (int j = 0; j < n; j++)
      if (s[i] < fabs(A[i * n + j])) s[i] = fabs(A[i * n + j]);
  }
  for (int k = 0; k < n - 1; k++) {
    double best = fabs(A[k * n + k] / s[k]);
    int prow = k;
    bool taken = false;
    for (int i = k; i < n; i++) {
      const double cand = fabs(A[i * n + k] / s[i]);
      if (best < cand) {
        best = cand;
        if (!taken) { prow = i; taken = true; }
      }
    }
    out->piv[k] = prow;
    if (best == 0.0) { set_error("conservation matrix is singular"); return 1; }
    if (prow != k) {
      for (int j = k; j < n; j++) { const double t = A[k * n + j]; A[k * n + j] = A[prow * n + j]; A[prow * n + j] = t; }
      const double t = s[k]; s[k] = s[prow]; s[prow] = t;
    }
    for (int i = k + 1; i < n; i++) {
      const double m = A[i * n + k] / A[k * n + k];
      A[i * n + k] = m;
      for (int j = k + 1; j < n; j++) A[i * n + j] -= m * A[k * n + j];
    }
  }
  out->piv[n - 1] = n - 1;
  return 0;
}

static void invalidate_graphs(sbte_ctx* c) {
  c->graph_gen++;
  for (auto& g : c->step_graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  c->step_graphs.clear();
}

static void free_scratch(sbte_ctx* c) {
  invalidate_graphs(c);
  cudaFree(c->d_tmp); cudaFree(c->d_specA); cudaFree(c->d_specB); cudaFree(c->d_specC);
  for (int i = 0; i < 3; i++) { cudaFree(c->d_lay[i]); cudaFree(c->d_layT[i]); c->d_layT[i] = nullptr; }
  cudaFree(c->d_qhat); cudaFree(c->d_Q); cudaFree(c->d_f); cudaFree(c->d_g); cudaFree(c->d_M); cudaFree(c->d_mom);
  c->d_tmp = c->d_specA = c->d_specB = c->d_specC = c->d_qhat = nullptr;
  c->d_lay[0] = c->d_lay[1] = c->d_lay[2] = nullptr;
  c->d_Q = c->d_f = c->d_g = c->d_M = c->d_mom = nullptr;
  c->cap = 0;
}

int ensure_capacity(sbte_ctx* c, int cells) {
  if (cells <= c->cap) return 0;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  free_scratch(c);
  const size_t padded = (size_t)((cells + 31) / 32) * 32;  // cell-minor layout works in groups of 32
  const size_t cb = padded * (size_t)c->n3 * sizeof(double2);
  const size_t rb = padded * (size_t)c->n3 * sizeof(double);
  CK(cudaMalloc(&c->d_tmp, cb));
  CK(cudaMalloc(&c->d_specA, cb));
  CK(cudaMalloc(&c->d_lay[0], cb));
  CK(cudaMemsetAsync(c->d_lay[0], 0, cb, c->stream));  // padding cells of the last group read as zero
  CK(cudaMalloc(&c->d_qhat, cb));
  CK(cudaMalloc(&c->d_Q, rb));
  CK(cudaMalloc(&c->d_f, rb));
  CK(cudaMalloc(&c->d_g, rb));
  CK(cudaMalloc(&c->d_mom, padded * 8 * sizeof(double)));
  // small-batch paths (0D, two-species pairs): second/third spectra and Maxwellian scratch, 32 cells
  const size_t cb32 = (size_t)32 * (size_t)c->n3 * sizeof(double2);
  CK(cudaMalloc(&c->d_specB, cb32));
  CK(cudaMalloc(&c->d_specC, cb32));
  CK(cudaMalloc(&c->d_lay[1], cb32));
  CK(cudaMalloc(&c->d_lay[2], cb32));
  CK(cudaMalloc(&c->d_M, 4 * (size_t)c->n3 * sizeof(double)));
  for (int i = 0; i < 3; i++) CK(cudaMalloc(&c->d_layT[i], (size_t)c->n3 * sizeof(double2)));
  c->cap = cells;
  return 0;
}

static int encode_weight_map(sbte_ctx* c, const double* W, CUtensorMap* out, int box_columns = 0) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available"); return 1; }
  const int N = c->N;
  const int cols_per_cta = box_columns > 0 ? box_columns : ((N >= 16) ? 8 : 4);  // Batch2Cfg<N>::COLS unless given
  cuuint64_t gdim[2] = {(cuuint64_t)c->n3, (cuuint64_t)c->n3};
  cuuint64_t gstride[1] = {(cuuint64_t)c->n3 * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)N, (cuuint32_t)(cols_per_cta * N)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((encode_fn)fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)W, gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return 1; }
  return 0;
}

static int make_tensor_map(sbte_ctx* c) {
  c->tmap_ok = false;
  if (c->d_Ws) { cudaFree(c->d_Ws); c->d_Ws = nullptr; }   // a new tensor invalidates the symmetrised copy
  c->xy_sym = -1;
  c->sched_cells = 0;
  if (!qhat_batch_supported(c->N) || !c->d_W) return 0;
  if (encode_weight_map(c, c->d_W, &c->tmapW)) return 1;
  if (qhat_batch_split_supported(c->N) && encode_weight_map(c, c->d_W, &c->tmapW16, 16)) return 1;
  c->tmap_ok = true;
  return 0;
}

static int release_weights(sbte_ctx* c) {
  invalidate_graphs(c);
  if (c->d_Ws) { cudaFree(c->d_Ws); c->d_Ws = nullptr; }
  if (c->owns_W && c->d_W) cudaFree((void*)c->d_W);
  c->d_W = nullptr; c->owns_W = false; c->host_key = nullptr; c->tmap_ok = false;
  return 0;
}

static int alloc_weights(sbte_ctx* c) {
  release_weights(c);
  double* p = nullptr;
  CK(cudaMalloc(&p, (size_t)c->n3 * c->n3 * sizeof(double)));
  c->d_W = p; c->owns_W = true;
  return 0;
}

__global__ void synth_weights_kernel(double* __restrict__ W, size_t n, unsigned long long seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long z = (unsigned long long)(i + 1) * 0x9E3779B97F4A7C15ULL + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    W[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
}

// ---- stream-K schedule of the batched convolution ---------------------------------------------
// T = groups * row-blocks tiles of N^2 steps; P persistent CTAs take equal contiguous shares of the
// T*N^2 global steps. tile_first / tile_np tell kernels which CTA writes which partial sum.
struct HostSchedule {
  std::vector<long long> begin, tbegin;   // [P+1], [T+1]
  std::vector<int> ctile, first;          // [P], [T]
  std::vector<unsigned char> np;          // [T], or [N*N*G] when kept per zeta column
  int G = 0, T = 0, P = 0, np_cols = 0, kmax = 1;
  bool cuts = false;   // ranges may end inside a chunk
};

// Pure host arithmetic (no CUDA call): also exported as sbte_batch_schedule_host for the CPU tests.
// split: the schedule of one remainder group on split tiles (qhat_batch.cu Batch3Cfg<N, true>: 16 columns per tile); its
// CTAs may be as short as one chunk, so that the launch adds about one chunk time after the main launch.
static void build_batch_schedule(int N, int cells, bool sym, int ctas, HostSchedule* out, bool split = false,
                                 int cuts = -1) {   // cuts: -1 = qhat_batch_cut_mode(N), else that mode
  const int cols = split ? 16 : qhat_batch_cols(N);
  // row-blocks never straddle a zeta_x plane: bpx blocks of `cols` zeta_y columns per plane, the last one
  // partly empty when cols does not divide N (N = 20, 22)
  const int bpx = (N + cols - 1) / cols;
  const int G = (cells + 31) / 32, RB = N * bpx, T = G * RB;
  // tile t = (row-block rb = t / G, cell group cg = t % G); its length is (visited xi_x planes) * N steps
  std::vector<long long>& tbegin = out->tbegin;
  tbegin.assign(T + 1, 0);
  for (int t = 0; t < T; t++) {
    const int zx = (t / G) / bpx;
    tbegin[t + 1] = tbegin[t] + (long long)(sym ? sym_nrep(N, zx) : N) * N;
  }
  const long long total = tbegin[T];
  // Cut granularity.  Whole xi_x chunks: the resident-plane kernel cannot do otherwise, and the busiest CTA then sets the
  // duration (ceil(chunks / P) chunk times).  Any step (line-ring kernels: a cut chunk's running sum passes from CTA p to
  // p + 1): equal shares, at the price qhat_batch_cut_cost(N).  Every CTA needs at least one chunk's worth of steps, so
  // that the chunks it shares with its two neighbours are different chunks.
  int mode = cuts >= 0 ? cuts : qhat_batch_cut_mode(N);
  if (mode != 0 && qhat_batch_cut_mode(N) == 0) mode = 0;
  const long long chunks = total / N;
  const int P_whole = (int)std::max<long long>(1, std::min<long long>(ctas, chunks / (split ? 1 : 4)));   // a few chunks per CTA
  const int P_any = (int)std::max<long long>(1, std::min<long long>(ctas, chunks));
  if (mode == 1) {
    const double t_whole = (double)((chunks + P_whole - 1) / P_whole);
    const double t_any = (double)chunks / P_any * (1.0 + qhat_batch_cut_cost(N));
    mode = t_any < t_whole ? 2 : 0;
  }
  const long long align = mode == 2 ? 1 : N;
  const int P = mode == 2 ? P_any : P_whole;
  std::vector<long long>& begin = out->begin;
  begin.assign(P + 1, 0);
  for (int p = 0; p <= P; p++) begin[p] = (long long)(((__int128)p * (total / align)) / P) * align;
  auto owner = [&](long long g) {   // last CTA whose range starts at or before g
    int lo = 0, hi = P - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) / 2; if (begin[mid] <= g) lo = mid; else hi = mid - 1; }
    return lo;
  };
  auto tile_of = [&](long long g) {
    int lo = 0, hi = T - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) / 2; if (tbegin[mid] <= g) lo = mid; else hi = mid - 1; }
    return lo;
  };
  std::vector<int>& first = out->first;
  std::vector<int>& ctile = out->ctile;
  std::vector<unsigned char>& np = out->np;
  first.assign(T, 0); ctile.assign(P, 0); np.assign(T, 0);
  int kmax = 1;
  for (int t = 0; t < T; t++) {
    // CTAs with an empty range never write: pick owners among non-empty ranges
    const int a = owner(tbegin[t]);
    first[t] = a;
    // canonical summation (qhat_batch.cu chunk_end): the first CTA folds the e0 chunks it completes into part 0, every
    // chunk a later CTA completes is a part of its own (a cut chunk is completed by the later of its two CTAs)
    const long long e0 = (std::min(begin[a + 1], tbegin[t + 1]) - tbegin[t]) / N;
    const long long nchunks = (tbegin[t + 1] - tbegin[t]) / N;
    const int parts_t = (e0 > 0 ? 1 : 0) + (int)(nchunks - e0);
    np[t] = (unsigned char)parts_t;
    kmax = std::max(kmax, parts_t);
  }
  for (int p = 0; p < P; p++) ctile[p] = tile_of(std::min(begin[p], total - 1));
  // the inverse transform looks the part count up as np[(column / np_cols) * G + cell group]; with partly
  // empty row-blocks that table is kept per zeta column
  const bool per_column = (N % cols) != 0;
  if (per_column) {
    std::vector<int> col_rb((size_t)N * N, 0);   // row-block that writes zeta column q
    for (int q = 0; q < N * N; q++) col_rb[q] = (q / N) * bpx + (q % N) / cols;
    std::vector<unsigned char> npc((size_t)N * N * G);
    for (int q = 0; q < N * N; q++)
      for (int g = 0; g < G; g++) npc[(size_t)q * G + g] = np[(size_t)col_rb[q] * G + g];
    np.swap(npc);
  }
  out->G = G; out->T = T; out->P = P; out->np_cols = per_column ? 1 : cols; out->kmax = kmax;
  out->cuts = mode == 2;
}

// uploads one host schedule; returns the device view through *dev (tables live in *mem)
static int upload_schedule(const HostSchedule& h, bool sym, void** mem, BatchSched* dev) {
  const int T = h.T, P = h.P;
  const size_t o1 = (size_t)(P + 1) * sizeof(long long);
  const size_t o2 = o1 + (size_t)(T + 1) * sizeof(long long);
  const size_t o3 = o2 + (size_t)P * sizeof(int);
  const size_t o4 = o3 + (size_t)T * sizeof(int);
  const size_t bytes = o4 + h.np.size();
  CK(cudaMalloc(mem, bytes));
  std::vector<unsigned char> blob(bytes);
  memcpy(blob.data(), h.begin.data(), o1);
  memcpy(blob.data() + o1, h.tbegin.data(), o2 - o1);
  memcpy(blob.data() + o2, h.ctile.data(), o3 - o2);
  memcpy(blob.data() + o3, h.first.data(), o4 - o3);
  memcpy(blob.data() + o4, h.np.data(), h.np.size());
  CK(cudaMemcpy(*mem, blob.data(), bytes, cudaMemcpyHostToDevice));
  unsigned char* base = (unsigned char*)*mem;
  *dev = {(const long long*)base, (const long long*)(base + o1), (const int*)(base + o2), (const int*)(base + o3),
          base + o4, h.G, T, P, h.np_cols, h.kmax, sym ? 1 : 0};
  dev->cuts = h.cuts ? 1 : 0;
  return 0;
}

// Schedules of a batched ComputeQ over `cells` cells.  Ordinarily one launch (c->sched_main == c->sched).  At N = 16 a
// last cell group with at most 16 live cells runs on split tiles in a second launch (c->sched_split, group c->split_cg):
// c->sched is then only the view the inverse transform needs -- part counts per zeta column for every group.
static int ensure_batch_schedule(sbte_ctx* c, int cells, bool sym) {
  if (c->sched_cells == cells && c->sched_sym == (int)sym && (c->d_sched_mem || c->d_sched_mem2)) return 0;
  c->graph_gen++;   // the schedule tables move: captured slab steps must be rebuilt
  CK(cudaStreamSynchronize(c->stream));
  for (void** m : {&c->d_sched_mem, &c->d_sched_mem2, &c->d_sched_mem3})
    if (*m) { cudaFree(*m); *m = nullptr; }
  if (c->sm_count == 0) CK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
  int ctas = c->sm_count;
  const char* pe = getenv("SBTE_BATCH_CTAS");
  if (pe && atoi(pe) > 0) ctas = atoi(pe);
  const int N = c->N, rem = cells % 32, G = (cells + 31) / 32;
  static const bool plane16 = getenv("SBTE_N16_PLANE") != nullptr;
  auto longest = [&](const HostSchedule& hs) {   // chunks of the busiest CTA = duration of the launch in chunk times
    long long m = 0;
    for (int p = 0; p < hs.P; p++) m = std::max(m, hs.begin[p + 1] - hs.begin[p]);
    return (double)m / N;
  };
  // The split launch follows the main one, so it pays only where dropping the padded group shortens the main launch
  // by more than the split launch lasts (80 cells: 6 -> 4 + 1 chunk times; 44 cells: 4 -> 4 + 1, not used).
  bool split = qhat_batch_split_supported(N) && !plane16 && rem >= 1 && rem <= 16;
  HostSchedule h, h2;
  if (split) {
    HostSchedule h0;
    build_batch_schedule(N, cells, sym, ctas, &h0);
    build_batch_schedule(N, 16, sym, ctas, &h2, true);
    double t_split = longest(h2) + 0.5;   // + launch gap and pipeline fill
    if (cells - rem > 0) {
      build_batch_schedule(N, cells - rem, sym, ctas, &h);
      t_split += longest(h);
    }
    if (t_split >= longest(h0)) { split = false; h = h0; }
  } else {
    build_batch_schedule(N, cells, sym, ctas, &h);
  }
  c->cells_main = split ? cells - rem : cells;
  c->split_on = split;
  c->split_cg = G - 1;
  int kmax = 1;
  if (c->cells_main > 0) {
    if (upload_schedule(h, sym, &c->d_sched_mem, &c->sched_main)) return 1;
    kmax = h.kmax;
  }
  c->sched = c->sched_main;
  if (split) {
    if (upload_schedule(h2, sym, &c->d_sched_mem2, &c->sched_split)) return 1;
    kmax = std::max(kmax, h2.kmax);
    // part counts per zeta column and cell group, main groups first, the split group last
    std::vector<unsigned char> npc((size_t)N * N * G);
    for (int q = 0; q < N * N; q++) {
      for (int g = 0; g < G - 1; g++) npc[(size_t)q * G + g] = h.np[(size_t)(q / h.np_cols) * h.G + g];
      npc[(size_t)q * G + (G - 1)] = h2.np[(size_t)(q / h2.np_cols) * h2.G];
    }
    CK(cudaMalloc(&c->d_sched_mem3, npc.size()));
    CK(cudaMemcpy(c->d_sched_mem3, npc.data(), npc.size(), cudaMemcpyHostToDevice));
    c->sched = {nullptr, nullptr, nullptr, nullptr, (const unsigned char*)c->d_sched_mem3, G, 0, 0, 1, kmax, sym ? 1 : 0};
  }
  {
    // hand-over buffers for cut chunks (BatchSched::carry): one accumulator set per compute warp and CTA and a flag word each
    // (the main and the split schedule each have their own slots: the two may share one launch)
    const size_t Pm = c->cells_main > 0 ? (size_t)h.P : 0, Ps = split ? (size_t)h2.P : 0;
    const size_t Pmax = Pm + Ps;
    const size_t acc_bytes = Pmax * kBatchWarps * (size_t)N * 32 * sizeof(double2);
    const size_t flag_bytes = Pmax * kBatchWarps * sizeof(int);
    if (c->carry_bytes < acc_bytes + flag_bytes) {
      if (c->d_carry) cudaFree(c->d_carry);
      c->d_carry = nullptr;
      c->carry_bytes = 0;
      CK(cudaMalloc(&c->d_carry, acc_bytes + flag_bytes));
      c->carry_bytes = acc_bytes + flag_bytes;
    }
    int* flags = (int*)((unsigned char*)c->d_carry + acc_bytes);
    CK(cudaMemsetAsync(flags, 0, flag_bytes, c->stream));   // on the stream the kernels run on (it does not wait for the null stream)
    c->sched_main.carry = (double2*)c->d_carry;
    c->sched_main.carry_flag = flags;
    c->sched_split.carry = (double2*)c->d_carry + Pm * kBatchWarps * (size_t)N * 32;
    c->sched_split.carry_flag = flags + Pm * kBatchWarps;
  }
  c->sched_cells = cells;
  c->sched_sym = (int)sym;
  // partial-sum workspace: kmax parts of (padded cells) x n3 complex
  const size_t stride = (size_t)G * 32 * (size_t)c->n3;
  if (!c->d_parts || c->parts_stride < stride || c->parts_cap < kmax) {
    if (c->d_parts) cudaFree(c->d_parts);
    c->d_parts = nullptr;
    CK(cudaMalloc(&c->d_parts, (size_t)kmax * stride * sizeof(double2)));
    c->parts_stride = stride;
    c->parts_cap = kmax;
  }
  return 0;
}

// the convolution launches of one batched ComputeQ: the main tiles, then the split tiles of the remainder group
static void launch_batched_conv(sbte_ctx* c, const double2* spec, int batch) {
  if (c->cells_main > 0 && c->split_on && c->sched_main.cuts && qhat_batch_pair_supported(c->N)) {
    launch_qhat_batch_pair(c, spec, c->d_parts, c->parts_stride, c->cells_main, batch, c->sched_main, c->sched_split, c->split_cg);
    return;
  }
  if (c->cells_main > 0) launch_qhat_batch2(c, spec, c->d_parts, c->parts_stride, c->cells_main, c->sched_main);
  if (c->split_on) launch_qhat_batch_split(c, spec, c->d_parts, c->parts_stride, batch, c->sched_split, c->split_cg);
}

// Symmetrised weights for f == g (see common.cuh): built lazily, once per bound tensor.
static int ensure_sym(sbte_ctx* c) {
  if (c->d_Ws) return 0;
  c->graph_gen++;
  CK(cudaMalloc(&c->d_Ws, (size_t)c->n3 * c->n3 * sizeof(double)));
  launch_symmetrize_weights(c, c->d_W, c->d_Ws);
  if (qhat_batch_supported(c->N)) {
    if (encode_weight_map(c, c->d_Ws, &c->tmapWs, 0)) return 1;
    if (qhat_batch_split_supported(c->N) && encode_weight_map(c, c->d_Ws, &c->tmapWs16, 16)) return 1;
  }
  return 0;
}
// Transposed pairing of the 0D stream (qhat.cu, TP): only for f == g, for the N it is built for, and only after the bound
// tensor has been checked -- once, one pass over it -- to be invariant under x <-> y of both indices to 1e-14 of its
// largest entry (isotropic weights are, to round-off; an arbitrary file need not be and then keeps the full stream).
static bool want_xy(sbte_ctx* c, bool same) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("SBTE_NO_XYSYM"); env = (e && atoi(e) != 0) ? 0 : 1; }
  if (!same || !c->xy_enabled || env != 1 || !qhat_stream_tp_supported(c->N) || !c->d_W) return false;
  if (c->xy_sym < 0) {
    double d = 0.0, a = 0.0;
    if (weights_xy_symmetry(c, c->d_W, &d, &a)) { c->xy_sym = 0; return false; }
    c->xy_sym_dev = a > 0.0 ? d / a : 0.0;
    c->xy_sym = (a > 0.0 && d <= 1e-14 * a) ? 1 : 0;
  }
  return c->xy_sym == 1;
}
static bool want_sym(sbte_ctx* c, bool same) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("SBTE_NO_SYM"); env = (e && atoi(e) != 0) ? 0 : 1; }
  return same && c->sym_enabled && env == 1;
}

// ---- convolution dispatch -------------------------------------------------------------------
// spectra of f (dif side) and g (xi side) -> qhat (natural layout)
bool use_pdl() {
  static const bool on = getenv("SBTE_NO_PDL") == nullptr;
  return on;
}

static int resolve_k2(sbte_ctx* c, int batch, int k2, bool same = true) {
  if (k2 == SBTE_K2_AUTO) {
    if (batch == 1) return qhat_stream_supported(c->N) ? SBTE_K2_STREAM : SBTE_K2_GENERIC;
    if (!same) return SBTE_K2_GENERIC;   // the batched kernels compute Q(f, f)
    return (batch >= 8 || qhat_batch_supported(c->N)) ? SBTE_K2_BATCH : SBTE_K2_GENERIC;
  }
  return k2;
}

// nsplit: in = the caller can add this many partial spectra (d_qhat + p * n3), out = how many were written
int qhat_from_real(sbte_ctx* c, const double* d_f, const double* d_g, double2* d_qhat, int batch, int k2,
                   int* nsplit = nullptr) {
  const int max_split = nsplit ? *nsplit : 1;
  if (nsplit) *nsplit = 1;
  if (!c->d_W) { set_error("no weights bound"); return 1; }
  if (ensure_capacity(c, batch)) return 1;
  const bool same = (d_f == d_g);
  k2 = resolve_k2(c, batch, k2, same);
  if (k2 == SBTE_K2_BATCH && same && !qhat_batch_supported(c->N)) {
    // any even N: lanes = cells, one warp per zeta row
    const bool sym = want_sym(c, true);
    if (sym && ensure_sym(c)) return 1;
    launch_fft3d(c, d_f, nullptr, 0, batch, nullptr, c->d_lay[0], LAY_CELLMINOR, nullptr, false);
    launch_qhat_batch_any(c, c->d_lay[0], d_qhat, batch, sym);
    return check_launch("qhat");
  }
  if (k2 == SBTE_K2_BATCH) {
    if (!same) { set_error("batched convolution requires f == g (single species)"); return 1; }
    const bool sym = want_sym(c, true);
    if (sym && ensure_sym(c)) return 1;
    if (ensure_batch_schedule(c, batch, sym)) return 1;
    launch_fft3d(c, d_f, nullptr, 0, batch, nullptr, c->d_lay[0], LAY_CELLMINOR, nullptr, false);
    launch_batched_conv(c, c->d_lay[0], batch);
    launch_combine_parts(c, c->d_parts, c->parts_stride, c->sched, batch, d_qhat);
  } else if (k2 == SBTE_K2_STREAM || k2 == SBTE_K2_STREAM_DEEP) {
    if (batch != 1 || !qhat_stream_supported(c->N)) { set_error("stream convolution: batch must be 1, N in {16,24,32}"); return 1; }
    const bool tp = want_xy(c, same) && k2 == SBTE_K2_STREAM;
    c->fft_layT = tp ? c->d_lay[1] : nullptr;   // the cluster transform writes the transposed copy along with the spectrum
    c->fft_layT_done = false;
    launch_fft3d(c, d_f, nullptr, 0, 1, nullptr, c->d_lay[0], LAY_PARITY, nullptr, false);
    c->fft_layT = nullptr;
    const double2* gl = c->d_lay[0];
    if (!same) {
      launch_fft3d(c, d_g, nullptr, 0, 1, nullptr, c->d_lay[1], LAY_PARITY, nullptr, false);
      gl = c->d_lay[1];
    }
    QhatPair p = {gl, c->d_lay[0]};
    const bool sym = want_sym(c, same);
    if (sym && ensure_sym(c)) return 1;
    // one cell keeps 1024 CTAs busy for only 3.5 waves: splitting each column's xi_x planes between two CTAs
    // shortens the under-filled last wave (the two partial spectra are added by the inverse transform)
    static const int want = getenv("SBTE_NO_SPLIT") ? 1 : (getenv("SBTE_SPLIT") ? atoi(getenv("SBTE_SPLIT")) : 2);
    static const bool all_n = getenv("SBTE_SPLIT_ALL") != nullptr;   // testing: split the smaller grids too
    const int ns = ((c->N == 32 || all_n) && want >= 1 && want <= max_split) ? want : 1;
    if (tp) {
      // half of the zeta columns, each weight against the spectrum and against its x <-> y transpose
      if (!c->fft_layT_done) launch_transpose_xy(c, c->d_lay[0], c->d_lay[1]);
      const QhatPair tp[2] = {{c->d_lay[0], c->d_lay[0]}, {c->d_lay[1], c->d_lay[1]}};
      // one CTA per streamed column: splitting the columns (as the one-pair kernel does) only loses here -- 0.449 ms
      // against 0.479 / 0.467 / 0.490 ms with 2 / 3 / 4 parts, sustained (profiles/r02_tp_tune.txt)
      launch_qhat_stream_tp(c, 2, tp, d_qhat, sym, 1);
      if (nsplit) *nsplit = 1;
      return check_launch("qhat");
    } else {
      launch_qhat_stream(c, 1, &p, d_qhat, k2 == SBTE_K2_STREAM_DEEP ? 4 : 2, sym, ns);
    }
    if (nsplit) *nsplit = ns;
  } else {
    if (!same && batch > 4) { set_error("generic convolution with f != g is limited to 4 cells"); return 1; }
    launch_fft3d(c, d_f, nullptr, 0, batch, c->d_specA, nullptr, 0, nullptr, false);
    const double2* gs = c->d_specA;
    if (!same) {
      launch_fft3d(c, d_g, nullptr, 0, batch, c->d_specB, nullptr, 0, nullptr, false);
      gs = c->d_specB;
    }
    QhatPair p = {gs, c->d_specA};
    launch_qhat_generic(c, 1, &p, d_qhat, batch);
  }
  return check_launch("qhat");
}

int compute_q_dev(sbte_ctx* c, const double* d_f, const double* d_g, double* d_Q, int batch, int k2) {
  if (ensure_capacity(c, batch)) return 1;  // before c->d_qhat is read: growth reallocates the scratch
  if (resolve_k2(c, batch, k2) == SBTE_K2_BATCH && d_f == d_g && qhat_batch_supported(c->N)) {
    // fast path: forward transform -> stream-K convolution -> inverse transform summing the partial sums
    if (!c->d_W) { set_error("no weights bound"); return 1; }
    const bool sym = want_sym(c, true);
    if (sym && ensure_sym(c)) return 1;
    if (ensure_batch_schedule(c, batch, sym)) return 1;
    launch_fft3d(c, d_f, nullptr, 0, batch, nullptr, c->d_lay[0], LAY_CELLMINOR, nullptr, false);
    launch_batched_conv(c, c->d_lay[0], batch);
    launch_fft3d_parts(c, c->d_parts, c->parts_stride, c->sched, 1, batch, nullptr, d_Q);
    return check_launch("batched compute_q");
  }
  int nsplit = (batch == 1 && fft_cluster_supported(c->N)) ? 8 : 1;   // d_qhat holds 32 cells' worth of spectrum
  if (qhat_from_real(c, d_f, d_g, c->d_qhat, batch, k2, &nsplit)) return 1;
  if (nsplit > 1) launch_fft3d_inverse_sum(c, c->d_qhat, nsplit, d_Q);
  else launch_fft3d(c, nullptr, c->d_qhat, 1, batch, nullptr, nullptr, 0, d_Q, false);
  return check_launch("inverse fft");
}

// One collision stage of the 1D step for a slab: out = a x + b y + s conserve(Q(src, src)) / Kn.
// Fast path (batched kernels with a whole-cell inverse transform): conservation and update run as the
// epilogue of the inverse transform; otherwise the three separate kernels.
// chain: bit 0 = the cell-minor spectrum of d_src is already in d_lay[0] (the previous stage left it there),
//        bit 1 = leave the spectrum of `out` there for the next stage (the inverse transform's kernel goes on with the
//        forward transform of the updated cell).  Both only on the fused path; *chained tells the caller what happened.
int collide_stage_dev(sbte_ctx* c, const double* d_src, double* d_Q, int batch, int k2, double* out, double a,
                      const double* x, double b, const double* y, double s, double Kn, int chain, int* chained) {
  static const bool no_fuse = getenv("SBTE_NO_FUSE") != nullptr;
  static const bool no_chain = getenv("SBTE_NO_CHAIN") != nullptr;
  if (chained) *chained = 0;
  struct Scope {   // the same transform kernel for every slab size (see sbte_ctx::cell_fft_any)
    sbte_ctx* c; bool old;
    explicit Scope(sbte_ctx* c_) : c(c_), old(c_->cell_fft_any) { c->cell_fft_any = true; }
    ~Scope() { c->cell_fft_any = old; }
  } scope(c);
  if (!no_fuse && resolve_k2(c, batch, k2) == SBTE_K2_BATCH && qhat_batch_supported(c->N) &&
      batch >= 2 && (c->N == 8 || c->N == 16)) {
    if (ensure_capacity(c, batch)) return 1;
    if (!c->d_W) { set_error("no weights bound"); return 1; }
    const bool sym = want_sym(c, true);
    if (sym && ensure_sym(c)) return 1;
    if (ensure_batch_schedule(c, batch, sym)) return 1;
    if (!(chain & 1)) launch_fft3d(c, d_src, nullptr, 0, batch, nullptr, c->d_lay[0], LAY_CELLMINOR, nullptr, false);
    launch_batched_conv(c, c->d_lay[0], batch);
    CellEpi epi = {};
    epi.mode = 1; epi.v = c->d_v; epi.wt = c->d_wt; epi.dv3 = c->dv * c->dv * c->dv; epi.lu = c->lu;
    epi.a = a; epi.x = x; epi.b = b; epi.y = y; epi.s = s; epi.Kn = Kn; epi.out = out;
    if ((chain & 2) && !no_chain) {
      epi.next_lay = c->d_lay[0]; epi.next_pre = c->d_pre[0]; epi.next_post = c->d_post[0]; epi.next_pref = c->pref[0];
      if (chained) *chained = 1;
    }
    if (launch_fft3d_parts_update(c, c->d_parts, c->parts_stride, c->sched, batch, epi)) return check_launch("fused collision stage");
    set_error("fused collision stage: no whole-cell transform for this N");
    return 1;
  }
  if (compute_q_dev(c, d_src, d_src, d_Q, batch, k2)) return 1;
  launch_conserve(c, d_Q, batch);
  launch_update(c, out, a, x, b, y, s, Kn, d_Q, (long)batch * c->n3);
  return check_launch("collision stage");
}

// ComputeQ_maxPreserve (src/collisions.c:178-210) with the three products folded into one weight pass:
//   Q^ = sum W ( g_j^[xi] (M_i + g_i)^[zeta-xi] + M_j^[xi] g_i^[zeta-xi] ),  (M_i + g_i)^ = f^.
int compute_q_maxpreserve_dev(sbte_ctx* c, const double* d_f, const double* d_g, double* d_Q, int k2) {
  if (!c->d_W) { set_error("no weights bound"); return 1; }
  if (ensure_capacity(c, 1)) return 1;
  k2 = resolve_k2(c, 1, k2);
  const long n3 = c->n3;
  double* Mi = c->d_M; double* gi = c->d_M + n3; double* Mj = c->d_M + 2 * n3; double* gj = c->d_M + 3 * n3;
  const bool same = (d_f == d_g);
  launch_maxwellian_split(c, d_f, nullptr, Mi, gi);
  if (!same) launch_maxwellian_split(c, d_g, Mi, Mj, gj);
  else { Mj = Mi; gj = gi; }
  const bool stream = (k2 == SBTE_K2_STREAM || k2 == SBTE_K2_STREAM_DEEP);
  if (stream && !qhat_stream_supported(c->N)) { set_error("stream convolution: N must be in {16,24,32}"); return 1; }
  const int lay = stream ? LAY_PARITY : LAY_NATURAL;
  // three spectra: f^ (A), g_i^ (B), M_j^ (C); g_j^ == g_i^ for one species (f == g)
  const double* ins[4] = {d_f, gi, Mj, gj};
  double2* outs[4] = {c->d_lay[0], c->d_lay[1], c->d_lay[2], c->d_specB};
  const bool tp = stream && k2 == SBTE_K2_STREAM && want_xy(c, same);
  c->fft_layT_multi = tp ? c->d_layT : nullptr;   // the cluster transform writes the transposed copies along with the spectra
  c->fft_layT_done = false;
  const bool multi = launch_fft3d_multi(c, same ? 3 : 4, ins, outs, lay);
  c->fft_layT_multi = nullptr;
  if (!multi) {   // one launch where the cluster kernel exists
    for (int q = 0; q < (same ? 3 : 4); q++) launch_fft3d(c, ins[q], nullptr, 0, 1, nullptr, outs[q], lay, nullptr, false);
  }
  const double2* gjhat = same ? c->d_lay[1] : c->d_specB;
  QhatPair pairs[2] = {{gjhat, c->d_lay[0]}, {c->d_lay[2], c->d_lay[1]}};
  if (tp) {
    // transposed pairing: half of the zeta columns; the two summed products against the spectra give column (zx, zy),
    // against the transposed spectra column (zy, zx)
    if (!c->fft_layT_done)
      for (int q = 0; q < 3; q++) launch_transpose_xy(c, c->d_lay[q], c->d_layT[q]);
    const QhatPair p4[4] = {{c->d_lay[1], c->d_lay[0]}, {c->d_lay[2], c->d_lay[1]},
                            {c->d_layT[1], c->d_layT[0]}, {c->d_layT[2], c->d_layT[1]}};
    const bool symt = want_sym(c, true);
    if (symt && ensure_sym(c)) return 1;
    launch_qhat_stream_tp(c, 4, p4, c->d_qhat, symt, 1);
    launch_fft3d(c, nullptr, c->d_qhat, 1, 1, nullptr, nullptr, 0, d_Q, false);
    return check_launch("maxpreserve (transposed pairing)");
  }
  const bool sym = stream && want_sym(c, same);   // for f == g the three-product summand is symmetric as a whole
  if (sym && ensure_sym(c)) return 1;
  // splitting the columns between CTAs (as the one-pair kernel does at N = 32) does not pay here: N = 32 already runs
  // 6.9 waves, and at N = 16 the kernel is 27 us and two parts cost more than they save (measured 27 -> 31 us)
  static const int mp_split = getenv("SBTE_MP_SPLIT") ? atoi(getenv("SBTE_MP_SPLIT")) : 1;
  const int ns = (stream && c->N == 16 && fft_cluster_supported(c->N) && mp_split >= 1 && mp_split <= 8) ? mp_split : 1;
  if (stream) launch_qhat_stream(c, 2, pairs, c->d_qhat, k2 == SBTE_K2_STREAM_DEEP ? 4 : 2, sym, ns);
  else launch_qhat_generic(c, 2, pairs, c->d_qhat, 1);
  if (ns > 1) launch_fft3d_inverse_sum(c, c->d_qhat, ns, d_Q);
  else launch_fft3d(c, nullptr, c->d_qhat, 1, 1, nullptr, nullptr, 0, d_Q, false);
  return check_launch("maxpreserve");
}

}  // namespace sbte

using namespace sbte;

extern "C" {

const char* sbte_last_error(void) { return g_err.c_str(); }

int sbte_create(sbte_ctx** out, int N, double L_v, const double* v, const double* eta, int device) {
  *out = nullptr;
  if (N < 2 || N > 32 || (N % 2) != 0) { set_error("N must be even and in [2, 32]"); return 1; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available: libsbte_b200 has no CPU fallback");
    return 1;
  }
  CK(cudaSetDevice(device));
  sbte_ctx* c = new sbte_ctx();
  c->N = N; c->n3 = (long)N * N * N; c->device = device; c->L_v = L_v;
  c->v.assign(v, v + N); c->eta.assign(eta, eta + N);
  c->wt.assign(N, 1.0); c->wt[0] = 0.5; c->wt[N - 1] = 0.5;   // src/collisions.c:48-54
  c->dv = v[1] - v[0];                                        // src/collisions.c:39-43
  c->deta = eta[1] - eta[0];
  c->L_eta = -eta[0];
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  init_fft_constants();

  std::vector<double2> tab(N);
  for (int m = 0; m < N; m++) tab[m] = make_double2(cos(2.0 * M_PI * m / N), sin(2.0 * M_PI * m / N));
  CK(cudaMalloc(&c->d_dft, N * sizeof(double2)));
  CK(cudaMemcpy(c->d_dft, tab.data(), N * sizeof(double2), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&c->d_v, N * sizeof(double)));
  CK(cudaMalloc(&c->d_wt, N * sizeof(double)));
  CK(cudaMemcpy(c->d_v, v, N * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_wt, c->wt.data(), N * sizeof(double), cudaMemcpyHostToDevice));

  // twiddle tables with the reference's own expressions (src/collisions.c:238-281), in host libm
  const double scale3 = pow(1.0 / sqrt(2.0 * M_PI), 3.0);
  for (int d = 0; d < 2; d++) {
    const double delta = d ? c->deta : c->dv;
    const double L_start = d ? c->L_v : c->L_eta;
    const double L_end = d ? c->L_eta : c->L_v;
    const double* arr = d ? c->v.data() : c->eta.data();
    const double sign = d ? -1.0 : 1.0;
    c->pref[d] = scale3 * delta * delta * delta;
    std::vector<double2> pre(3 * N - 2), post((size_t)c->n3);
    for (int s = 0; s < 3 * N - 2; s++) {
      const double sum = sign * (double)s * L_start * delta;
      pre[s] = make_double2(cos(sum), sin(sum));
    }
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++)
        for (int k = 0; k < N; k++) {
          const double sum = sign * L_end * (arr[i] + arr[j] + arr[k]);
          post[k + (size_t)N * (j + (size_t)N * i)] = make_double2(cos(sum), sin(sum));
        }
    CK(cudaMalloc(&c->d_pre[d], pre.size() * sizeof(double2)));
    CK(cudaMalloc(&c->d_post[d], post.size() * sizeof(double2)));
    CK(cudaMemcpy(c->d_pre[d], pre.data(), pre.size() * sizeof(double2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_post[d], post.data(), post.size() * sizeof(double2), cudaMemcpyHostToDevice));
  }
  if (build_lu(c, &c->lu)) { delete c; return 1; }
  *out = c;
  return 0;
}

int sbte_destroy(sbte_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_scratch(c);
  release_weights(c);
  cudaFree(c->d_v); cudaFree(c->d_wt); cudaFree(c->d_dft);
  for (int d = 0; d < 2; d++) { cudaFree(c->d_pre[d]); cudaFree(c->d_post[d]); }
  if (c->h_pin) cudaFreeHost(c->h_pin);
  for (cudaEvent_t e : c->k2_ev) cudaEventDestroy(e);
  if (c->d_sched_mem) cudaFree(c->d_sched_mem);
  if (c->d_sched_mem2) cudaFree(c->d_sched_mem2);
  if (c->d_sched_mem3) cudaFree(c->d_sched_mem3);
  if (c->d_carry) cudaFree(c->d_carry);
  if (c->d_parts) cudaFree(c->d_parts);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

// The stream-K schedule ensure_batch_schedule() would upload for (N, cells, sym) on a device with `ctas` SMs.
// Pure host arithmetic: works without a GPU (CPU tests, sizing).  split bit 0: the schedule of the split-tile launch that
// serves a remainder group of at most 16 cells at N = 16; bit 1: cuts at whole xi_x chunks only (what SBTE_CHUNK_CUTS=1
// makes the library use); bit 2: cuts at any step wherever the kernel takes them (SBTE_CHUNK_CUTS=0); neither: the library's choice.  dims = {G, T, P, np_cols, kmax, np_len};
// any array pointer may be null (query dims first, then call again with arrays of P+1, T+1, P, T, np_len entries).
int sbte_batch_schedule_host(int N, int cells, int sym, int ctas, int split, long long* cta_begin, long long* tile_begin,
                             int* cta_tile, int* tile_first, unsigned char* np, int* dims) {
  if (N < 2 || N > 32 || (N % 2) != 0 || cells < 1 || ctas < 1) { set_error("batch schedule: bad arguments"); return 1; }
  if (!qhat_batch_supported(N)) { set_error("batch schedule: this N runs the any-N kernel (no schedule)"); return 1; }
  if ((split & 1) && (N != 16 || cells > 16)) { set_error("batch schedule: split tiles serve one group of at most 16 cells at N = 16"); return 1; }
  HostSchedule h;
  build_batch_schedule(N, (split & 1) ? 16 : cells, sym != 0, ctas, &h, (split & 1) != 0, (split & 2) ? 0 : (split & 4) ? 2 : -1);
  if (dims) { dims[0] = h.G; dims[1] = h.T; dims[2] = h.P; dims[3] = h.np_cols; dims[4] = h.kmax; dims[5] = (int)h.np.size(); }
  if (cta_begin) memcpy(cta_begin, h.begin.data(), h.begin.size() * sizeof(long long));
  if (tile_begin) memcpy(tile_begin, h.tbegin.data(), h.tbegin.size() * sizeof(long long));
  if (cta_tile) memcpy(cta_tile, h.ctile.data(), h.ctile.size() * sizeof(int));
  if (tile_first) memcpy(tile_first, h.first.data(), h.first.size() * sizeof(int));
  if (np) memcpy(np, h.np.data(), h.np.size());
  return 0;
}

int sbte_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// lets kernels of either context read the other's memory (same-process multi-GPU: sbte_slab_peer_attach)
int sbte_enable_peer_access(sbte_ctx* a, sbte_ctx* b) {
  if (a->device == b->device) return 0;
  for (int dir = 0; dir < 2; dir++) {
    sbte_ctx* from = dir ? b : a;
    sbte_ctx* to = dir ? a : b;
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, from->device, to->device));
    if (!can) { set_error("no peer access between GPU " + std::to_string(from->device) + " and " + std::to_string(to->device)); return 1; }
    CK(cudaSetDevice(from->device));
    cudaError_t e = cudaDeviceEnablePeerAccess(to->device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
    else if (e != cudaSuccess) { set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return 1; }
  }
  return 0;
}

int sbte_sync(sbte_ctx* c) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  CK(cudaStreamSynchronize(c->stream)); return check_launch("sync");
}
void* sbte_stream(sbte_ctx* c) { return (void*)c->stream; }
unsigned long long sbte_launch_count(sbte_ctx* c) { return c->launches; }
int sbte_reserve(sbte_ctx* c, int cells) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  return ensure_capacity(c, cells);
}

int sbte_set_symmetrize(sbte_ctx* c, int enable) {
  if (c->sym_enabled != (enable != 0)) c->graph_gen++;
  c->sym_enabled = enable != 0;
  return 0;
}

// transposed pairing of the 0D stream: switch, and what the check of the bound tensor found
int sbte_set_xy_pairing(sbte_ctx* c, int enable) {
  c->xy_enabled = enable != 0;
  return 0;
}
int sbte_xy_pairing_state(sbte_ctx* c, int* state, double* deviation) {
  cudaSetDevice(c->device);
  if (c->xy_sym < 0 && c->d_W && qhat_stream_tp_supported(c->N)) (void)want_xy(c, true);   // examine the tensor now
  if (state) *state = c->xy_sym;
  if (deviation) *deviation = c->xy_sym_dev;
  return 0;
}

int sbte_k2_profile(sbte_ctx* c, int enable) {
  c->k2_prof = enable != 0;
  c->k2_ev_used = 0;
  return 0;
}

int sbte_k2_profile_read(sbte_ctx* c, double* total_ms, int* launches) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  CK(cudaStreamSynchronize(c->stream));
  double sum = 0.0;
  const size_t pairs = c->k2_ev_used / 2;
  for (size_t i = 0; i < pairs; i++) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->k2_ev[2 * i], c->k2_ev[2 * i + 1]));
    sum += ms;
  }
  *total_ms = sum;
  *launches = (int)pairs;
  c->k2_ev_used = 0;
  return 0;
}

int sbte_dev_alloc(void** d_ptr, size_t bytes) { CK(cudaMalloc(d_ptr, bytes)); return 0; }
int sbte_dev_free(void* d_ptr) { CK(cudaFree(d_ptr)); return 0; }
int sbte_h2d(sbte_ctx* c, void* d_dst, const void* src, size_t bytes) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  CK(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
int sbte_d2h(sbte_ctx* c, void* dst, const void* d_src, size_t bytes) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  CK(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int sbte_d2d(sbte_ctx* c, void* d_dst, const void* d_src, size_t bytes) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  CK(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

// ---- weights
int sbte_weights_upload_rows(sbte_ctx* c, double* const* rows) {
  CK(cudaSetDevice(c->device));
  if (alloc_weights(c)) return 1;
  // rows are separate host allocations (src/weights.c:61-63): gather through a pinned bounce buffer
  const size_t row_bytes = (size_t)c->n3 * sizeof(double);
  const size_t rows_per_chunk = std::max<size_t>(1, (64u << 20) / row_bytes);
  double* pin[2] = {nullptr, nullptr};
  CK(cudaMallocHost(&pin[0], rows_per_chunk * row_bytes));
  CK(cudaMallocHost(&pin[1], rows_per_chunk * row_bytes));
  cudaEvent_t ev[2];
  CK(cudaEventCreate(&ev[0])); CK(cudaEventCreate(&ev[1]));
  int b = 0;
  for (size_t r0 = 0; r0 < (size_t)c->n3; r0 += rows_per_chunk, b ^= 1) {
    const size_t nr = std::min(rows_per_chunk, (size_t)c->n3 - r0);
    CK(cudaEventSynchronize(ev[b]));
    for (size_t r = 0; r < nr; r++) memcpy(pin[b] + r * c->n3, rows[r0 + r], row_bytes);
    CK(cudaMemcpyAsync((double*)c->d_W + r0 * c->n3, pin[b], nr * row_bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(ev[b], c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  cudaFreeHost(pin[0]); cudaFreeHost(pin[1]);
  cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
  c->host_key = (const void*)rows;
  return make_tensor_map(c);
}

int sbte_weights_upload(sbte_ctx* c, const double* W) {
  CK(cudaSetDevice(c->device));
  if (alloc_weights(c)) return 1;
  CK(cudaMemcpy((void*)c->d_W, W, (size_t)c->n3 * c->n3 * sizeof(double), cudaMemcpyHostToDevice));
  return make_tensor_map(c);
}

int sbte_weights_load_file(sbte_ctx* c, const char* path) {
  CK(cudaSetDevice(c->device));
  FILE* fp = fopen(path, "rb");
  if (!fp) { set_error(std::string("cannot open weight file ") + path); return 1; }
  if (alloc_weights(c)) { fclose(fp); return 1; }
  const size_t total = (size_t)c->n3 * c->n3;
  const size_t chunk = (size_t)(64u << 20) / sizeof(double);
  // two pinned buffers: the read of one chunk overlaps the upload of the previous one; every failure path releases the
  // file, the buffers and the half-filled tensor
  double* pin[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  auto cleanup = [&](bool failed) {
    fclose(fp);
    for (int b = 0; b < 2; b++) {
      if (done[b]) { cudaEventSynchronize(done[b]); cudaEventDestroy(done[b]); }
      if (pin[b]) cudaFreeHost(pin[b]);
    }
    if (failed) release_weights(c);
  };
  cudaError_t e = cudaSuccess;
  for (int b = 0; b < 2 && e == cudaSuccess; b++) {
    e = cudaMallocHost(&pin[b], chunk * sizeof(double));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) { cleanup(true); set_error(std::string("weight file staging buffers: ") + cudaGetErrorString(e)); return 1; }
  int b = 0;
  for (size_t off = 0; off < total; off += chunk, b ^= 1) {
    const size_t n = std::min(chunk, total - off);
    cudaEventSynchronize(done[b]);                  // the previous upload from this buffer has finished
    if (fread(pin[b], sizeof(double), n, fp) != n) {   // src/weights.c:82-86
      cleanup(true);
      set_error("Error reading weight file");
      return 1;
    }
    e = cudaMemcpyAsync((double*)c->d_W + off, pin[b], n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaEventRecord(done[b], c->stream);
    if (e != cudaSuccess) { cleanup(true); set_error(std::string("weight upload: ") + cudaGetErrorString(e)); return 1; }
  }
  e = cudaStreamSynchronize(c->stream);
  cleanup(e != cudaSuccess);
  if (e != cudaSuccess) { set_error(std::string("weight upload: ") + cudaGetErrorString(e)); return 1; }
  return make_tensor_map(c);
}

int sbte_weights_bind_device(sbte_ctx* c, const double* d_W) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  release_weights(c);
  c->d_W = d_W; c->owns_W = false;
  return make_tensor_map(c);
}

int sbte_weights_fill_synthetic(sbte_ctx* c, unsigned long long seed) {
  CK(cudaSetDevice(c->device));
  if (alloc_weights(c)) return 1;
  synth_weights_kernel<<<148 * 8, 256, 0, c->stream>>>((double*)c->d_W, (size_t)c->n3 * c->n3, seed);
  c->launches += 1;
  CK(cudaStreamSynchronize(c->stream));
  return make_tensor_map(c);
}

const double* sbte_weights_device(sbte_ctx* c) { return c->d_W; }

int sbte_weights_generate_iso(sbte_ctx* c, double lambda) {
  CK(cudaSetDevice(c->device));
  if (alloc_weights(c)) return 1;
  int used = 0;
  if (generate_weights_iso(c, (double*)c->d_W, lambda, &used)) { release_weights(c); return 1; }
  return make_tensor_map(c);
}

int sbte_weights_save_file(sbte_ctx* c, const char* path) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  if (!c->d_W) { set_error("no weights bound"); return 1; }
  FILE* fp = fopen(path, "wb");
  if (!fp) { set_error(std::string("cannot create weight file ") + path); return 1; }
  const size_t total = (size_t)c->n3 * c->n3;
  const size_t chunk = (size_t)(64u << 20) / sizeof(double);
  double* pin = nullptr;
  cudaError_t e = cudaMallocHost(&pin, chunk * sizeof(double));
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  for (size_t off = 0; off < total && e == cudaSuccess; off += chunk) {
    const size_t n = std::min(chunk, total - off);
    e = cudaMemcpy(pin, c->d_W + off, n * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && fwrite(pin, sizeof(double), n, fp) != n) {
      fclose(fp);
      cudaFreeHost(pin);
      set_error("Something is wrong with storing the weights");   // src/weights.c:104-107
      return 1;
    }
  }
  const bool closed = fclose(fp) == 0;
  if (pin) cudaFreeHost(pin);
  if (e != cudaSuccess) { set_error(std::string("weight file: ") + cudaGetErrorString(e)); return 1; }
  if (!closed) { set_error("Something is wrong with storing the weights"); return 1; }
  return 0;
}

// ---- device-pointer operations
int sbte_fft3d(sbte_ctx* c, const double* d_in, double* d_out, int invert, int batch) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  if (ensure_capacity(c, batch)) return 1;
  launch_fft3d(c, nullptr, (const double2*)d_in, invert, batch, (double2*)d_out, nullptr, 0, nullptr, false);
  return check_launch("fft3d");
}

int sbte_qhat(sbte_ctx* c, const double* d_f, const double* d_g, double* d_qhat, int batch, int k2) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  return qhat_from_real(c, d_f, d_g, (double2*)d_qhat, batch, k2);
}

int sbte_compute_q(sbte_ctx* c, const double* d_f, const double* d_g, double* d_Q, int batch, int k2) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  return compute_q_dev(c, d_f, d_g, d_Q, batch, k2);
}

int sbte_compute_q_maxpreserve(sbte_ctx* c, const double* d_f, const double* d_g, double* d_Q, int k2) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  return compute_q_maxpreserve_dev(c, d_f, d_g, d_Q, k2);
}

int sbte_conserve(sbte_ctx* c, double* d_Q, int batch) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  launch_conserve(c, d_Q, batch);
  return check_launch("conserve");
}

int sbte_moment_functionals(sbte_ctx* c, const double* d_Q, double* d_b5, int batch) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  launch_moment_functionals(c, d_Q, d_b5, batch);
  return check_launch("moment_functionals");
}

int sbte_moments(sbte_ctx* c, const double* d_f, double* d_mom8, int batch) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  launch_moments(c, d_f, d_mom8, batch);
  return check_launch("moments");
}

static int step_0d_direct(sbte_ctx* c, double* d_f, double dt, double Kn, int order, int k2);

// exec/boltz.c:189-241.  The step is launch-bound at small N (24 launches, ~0.2 ms at N=16), so after one
// direct execution (allocations, attributes, symmetrised weights) it is captured into a CUDA graph and
// replayed; profiling, SBTE_NO_GRAPH=1 or a change of arguments fall back to direct launches.
int sbte_step_0d(sbte_ctx* c, double* d_f, double dt, double Kn, int order, int k2) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  if (ensure_capacity(c, 1)) return 1;
  static int no_graph = -1;
  if (no_graph < 0) { const char* e = getenv("SBTE_NO_GRAPH"); no_graph = (e && atoi(e) != 0) ? 1 : 0; }
  if (no_graph || c->k2_prof) return step_0d_direct(c, d_f, dt, Kn, order, k2);
  const int sym = c->sym_enabled ? 1 : 0;
  sbte_ctx::StepGraph* g = nullptr;
  for (auto& s : c->step_graphs)
    if (s.d_f == d_f && s.dt == dt && s.Kn == Kn && s.order == order && s.k2 == k2 && s.sym == sym) g = &s;
  if (!g) {
    if (c->step_graphs.size() > 8) invalidate_graphs(c);
    c->step_graphs.push_back({nullptr, d_f, dt, Kn, order, k2, sym, 0, 0});
    g = &c->step_graphs.back();
  }
  if (g->exec) {
    CK(cudaGraphLaunch(g->exec, c->stream));
    c->launches += g->launches;
    return 0;
  }
  if (g->seen++ == 0) return step_0d_direct(c, d_f, dt, Kn, order, k2);   // warm path first
  const unsigned long long before = c->launches;
  CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = step_0d_direct(c, d_f, dt, Kn, order, k2);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
  if (rc != 0 || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    if (rc == 0) set_error(std::string("graph capture of the 0D step failed: ") + cudaGetErrorString(e));
    return 1;
  }
  g->launches = c->launches - before;
  c->launches = before;
  e = cudaGraphInstantiate(&g->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { g->exec = nullptr; set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); return 1; }
  CK(cudaGraphLaunch(g->exec, c->stream));
  c->launches += g->launches;
  return 0;
}

static int step_0d_direct(sbte_ctx* c, double* d_f, double dt, double Kn, int order, int k2) {
  const long n3 = c->n3;
  double* Q = c->d_Q;
  if (compute_q_maxpreserve_dev(c, d_f, d_f, Q, k2)) return 1;
  launch_conserve(c, Q, 1);
  if (order == 1) {
    launch_update(c, d_f, 1.0, d_f, 0.0, nullptr, dt, Kn, Q, n3);
  } else {
    double* f1 = c->d_g;
    launch_update(c, f1, 1.0, d_f, 0.0, nullptr, dt, Kn, Q, n3);
    if (compute_q_maxpreserve_dev(c, f1, f1, Q, k2)) return 1;
    launch_conserve(c, Q, 1);
    launch_update(c, d_f, 0.5, d_f, 0.5, f1, 0.5 * dt, Kn, Q, n3);
  }
  return check_launch("step_0d");
}

// ---- host-pointer forms
int sbte_compute_q_host(sbte_ctx* c, const double* f, const double* g, double* Q, int k2) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  if (ensure_capacity(c, 1)) return 1;
  const size_t bytes = (size_t)c->n3 * sizeof(double);
  CK(cudaMemcpyAsync(c->d_f, f, bytes, cudaMemcpyHostToDevice, c->stream));
  const double* dg = c->d_f;
  if (g != f) {
    CK(cudaMemcpyAsync(c->d_g, g, bytes, cudaMemcpyHostToDevice, c->stream));
    dg = c->d_g;
  }
  if (compute_q_dev(c, c->d_f, dg, c->d_Q, 1, k2)) return 1;
  CK(cudaMemcpyAsync(Q, c->d_Q, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return check_launch("compute_q_host");
}

int sbte_compute_q_maxpreserve_host(sbte_ctx* c, const double* f, const double* g, double* Q, int k2) {
  cudaSetDevice(c->device);   // one process may drive several GPUs
  if (ensure_capacity(c, 1)) return 1;
  const size_t bytes = (size_t)c->n3 * sizeof(double);
  CK(cudaMemcpyAsync(c->d_f, f, bytes, cudaMemcpyHostToDevice, c->stream));
  const double* dg = c->d_f;
  if (g != f) {
    CK(cudaMemcpyAsync(c->d_g, g, bytes, cudaMemcpyHostToDevice, c->stream));
    dg = c->d_g;
  }
  if (compute_q_maxpreserve_dev(c, c->d_f, dg, c->d_Q, k2)) return 1;
  CK(cudaMemcpyAsync(Q, c->d_Q, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return check_launch("compute_q_maxpreserve_host");
}

}  // extern "C"
