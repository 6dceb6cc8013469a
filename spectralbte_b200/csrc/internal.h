// spectralbte_b200/csrc/internal.h -- library-internal state and kernel launchers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <string>
#include <vector>

struct sbte_slab;

// Conservation data handed to kernels by value: the factored Gram matrix of the five moment
// functionals and its pivots (reference: src/conserve.c:89-168,268-317).
struct ConsLU {
  double a[25];
  int piv[5];
};

// Epilogue of the batched inverse transform: with the cell's Q still in shared memory, project it onto the
// conserved moments (src/conserve.c:207-264) and apply the Euler / Heun update out = a x + b y + (s Q)/Kn
// (exec/boltz.c:296-343), instead of three more passes over global memory.
struct CellEpi {
  int mode;               // 0: none (Q is written as is); 1: conserve + update
  const double* v;        // velocity grid
  const double* wt;       // trapezoid weights
  double dv3;
  ConsLU lu;
  double a; const double* x;
  double b; const double* y;   // y may be null
  double s, Kn;
  double* out;
  // chained forward transform of `out` (null: none): cell-minor spectrum for the next stage's convolution, written by the
  // same kernel while the updated cell is still in shared memory -- one launch and one trip through global memory less
  double2* next_lay;
  const double2* next_pre;    // forward pre-twiddles [3N-2]
  const double2* next_post;   // forward post-twiddles [N^3]
  double next_pref;           // forward prefactor (2 pi)^-3/2 dv^3
};

// stream-K schedule of the batched convolution (device tables, see qhat_batch.cu)
constexpr int kBatchWarps = 8;   // compute warps per CTA of the batched convolution kernels (one carry slot each)

struct BatchSched {
  const long long* cta_begin;   // [P+1] first global step of every CTA
  const long long* tile_begin;  // [T+1] first global step of every tile (tile lengths vary when symmetrised)
  const int* cta_tile;          // [P] tile that contains cta_begin[p]
  const int* tile_first;        // [T] first CTA that touches tile t
  const unsigned char* tile_np; // number of partial sums, indexed [(zeta column / cols) * G + cell group]
  int G, T, P, cols, kmax, sym; // cols: zeta columns per tile_np entry (the tile width, or 1: table kept per column)
  // hand-over of the running sum of a xi_x chunk that a stream-K cut divides between CTA p and p + 1 (line-ring kernels):
  // carry[p] = one accumulator set per compute warp, carry_flag[p * warps + w] = published
  double2* carry = nullptr;
  int* carry_flag = nullptr;
  int cuts = 0;                 // 1: some range ends inside a chunk (the kernel instance that hands running sums over)
};

struct sbte_ctx {
  int N = 0;
  long n3 = 0;
  int device = 0;
  double L_v = 0, L_eta = 0, dv = 0, deta = 0;
  cudaStream_t stream = nullptr;
  std::vector<double> v, eta, wt;

  // device tables
  double* d_v = nullptr;         // [N]
  double* d_wt = nullptr;        // [N] trapezoid weights
  double2* d_dft = nullptr;      // [N] (cos, sin)(2 pi m / N)
  double2* d_pre[2] = {nullptr, nullptr};   // [3N-2] pre-twiddle (cos,sin) by i+j+k; 0 = forward, 1 = inverse
  double2* d_post[2] = {nullptr, nullptr};  // [N^3] post-twiddle (cos,sin)
  double pref[2] = {0, 0};       // (2 pi)^-3/2 delta^3
  ConsLU lu;

  // weights
  const double* d_W = nullptr;   // N^3 x N^3, row-major [zeta][xi]
  bool owns_W = false;
  const void* host_key = nullptr;  // identity of the host row-pointer array the cached copy came from
  CUtensorMap tmapW;
  bool tmap_ok = false;
  // symmetrised copy for f == g (Ws[zeta][xi] = W[zeta][xi] + W[zeta][sigma_zeta(xi)], see common.cuh)
  bool sym_enabled = true;
  // x <-> y invariance of the bound tensor: -1 not examined yet, 0 no, 1 yes (|W[t zeta][t xi] - W[zeta][xi]| <= 1e-14 max|W|)
  int xy_sym = -1;
  double xy_sym_dev = 0.0;            // the measured relative deviation
  bool xy_enabled = true;
  double2* fft_layT = nullptr;        // set around a forward transform: also write the x<->y transposed parity-layout copy here
  double2* const* fft_layT_multi = nullptr;   // the same for launch_fft3d_multi: one destination per job
  bool fft_layT_done = false;         // ... and whether the kernel that ran could do it (else launch_transpose_xy)
  double* d_Ws = nullptr;
  CUtensorMap tmapWs;
  CUtensorMap tmapW16, tmapWs16;   // boxes of 16 zeta_y columns: split tiles of the N = 16 remainder group

  // scratch, sized for `cap` cells
  int cap = 0;
  double2* d_tmp = nullptr;      // FFT intermediate           [cap][n3]
  double2* d_specA = nullptr;    // spectra, natural layout     [cap][n3]  (operand "f": zeta - xi side)
  double2* d_specB = nullptr;    //                             [cap][n3]  (operand "g": xi side)
  double2* d_specC = nullptr;    // third spectrum (maxPreserve: Maxwellian)
  double2* d_lay[3] = {nullptr, nullptr, nullptr};  // kernel-specific operand layouts of A, B, C
  double2* d_layT[3] = {nullptr, nullptr, nullptr}; // their x<->y transposes (one cell each; transposed pairing of maxPreserve)
  double2* d_qhat = nullptr;     // [cap][n3]
  double* d_Q = nullptr;         // [cap][n3]
  double* d_f = nullptr;         // staging for host-pointer entry points [cap][n3]
  double* d_g = nullptr;
  double* d_M = nullptr;         // Maxwellian / perturbation scratch (0D)
  double* d_mom = nullptr;       // [cap][8] moments
  double* h_pin = nullptr;       // pinned host staging, 3 * n3 doubles
  // batched-convolution schedule + partial-sum workspace (valid for sched_cells cells)
  int sched_cells = 0;
  int sched_sym = -1;
  BatchSched sched = {nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, 0};        // what the inverse transform reads
  BatchSched sched_main = {nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, 0};   // main convolution launch
  BatchSched sched_split = {nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, 0};  // split tiles of the remainder group
  bool split_on = false;
  int split_cg = 0, cells_main = 0;
  void* d_sched_mem = nullptr;
  void* d_sched_mem2 = nullptr;
  void* d_sched_mem3 = nullptr;
  void* d_carry = nullptr;          // carry + flags of the batched convolution (see BatchSched)
  size_t carry_bytes = 0;
  double2* d_parts = nullptr;
  size_t parts_stride = 0;          // double2 elements per part
  int parts_cap = 0;                // parts allocated
  int sm_count = 0;
  // CUDA graphs of the launch-bound 0D time step (24 launches per RK2 step), keyed on its arguments
  struct StepGraph {
    cudaGraphExec_t exec;
    double* d_f;
    double dt, Kn;
    int order, k2, sym;
    unsigned long long launches;
    int seen;
  };
  std::vector<StepGraph> step_graphs;
  unsigned long long graph_gen = 1;   // bumped whenever a buffer a captured graph may reference is replaced
  unsigned long long launches = 0;  // kernels launched through this context
  // slab collisions take the whole-cell transform at ANY batch size: which kernel transforms a cell must not depend on
  // how many cells the rank holds (rank-count invariance, bit for bit); other callers keep it for batches >= 8
  bool cell_fft_any = false;
  int wg_nonconverged = 0, wg_classes = 0;   // last device weight generation: integrals that ended abnormally (QUADPACK)
  bool k2_prof = false;             // bracket every K2 launch with CUDA events
  std::vector<cudaEvent_t> k2_ev;   // [2*i], [2*i+1] = start/stop of launch i
  size_t k2_ev_used = 0;
};

namespace sbte {

void set_error(const std::string& msg);
void k2_mark(sbte_ctx* c);   // records a profiling event on c->stream when K2 profiling is on
int ensure_capacity(sbte_ctx* c, int cells);

// operand layouts produced by the forward transform's last pass
enum SpecLayout {
  LAY_NATURAL = 0,   // [cell][x][y][z]
  LAY_PARITY = 1,    // [cell][x][y][z&1][z>>1]            (stream kernel: conflict-free 2-column reads)
  LAY_CELLMINOR = 2  // [cell/32][x][y][z][cell%32]        (batched kernel: lanes = cells)
};

void init_fft_constants();
// programmatic dependent launch between the 0D transforms and the stream convolution (SBTE_NO_PDL=1 disables)
bool use_pdl();
// fft.cu -- K1 / K3. in_real: N^3 doubles per cell; in_cplx: N^3 double2 per cell (exactly one non-null).
// out_nat / out_lay / out_real may each be null. batch = number of cells.
void launch_fft3d(sbte_ctx* c, const double* in_real, const double2* in_cplx, int invert, int batch,
                  double2* out_nat, double2* out_lay, int layout, double* out_real, bool accumulate_real);
// same, the complex input being the sum of the stream-K partial sums described by `sch`
// several single-cell forward transforms in one launch (cluster kernel); false: not available for this N
bool launch_fft3d_multi(sbte_ctx* c, int njobs, const double* const* in_real, double2* const* out_lay, int layout);
void launch_fft3d_parts(sbte_ctx* c, const double2* parts, size_t part_stride, const BatchSched& sch, int invert,
                        int batch, double2* out_nat, double* out_real);
// same with the conserve + update epilogue; false when this N has no whole-cell kernel (nothing was launched)
bool launch_fft3d_parts_update(sbte_ctx* c, const double2* parts, size_t part_stride, const BatchSched& sch, int batch,
                               const CellEpi& epi);
void launch_combine_parts(sbte_ctx* c, const double2* parts, size_t part_stride, const BatchSched& sch, int batch,
                          double2* out);

// qhat.cu -- K2
struct QhatPair {
  const double2* xi_side;    // g^[xi]
  const double2* dif_side;   // f^[zeta - xi]
};
// generic: natural-layout operands, any N, any batch (one CTA per (zeta, cell))
void launch_qhat_generic(sbte_ctx* c, int npairs, const QhatPair* pairs, double2* qhat, int batch);
// stream kernel (N in {16,24,32}, batch 1): parity-layout operands
bool qhat_stream_supported(int N);
void launch_qhat_stream(sbte_ctx* c, int npairs, const QhatPair* pairs, double2* qhat, int depth, bool sym,
                        int nsplit = 1);
// inverse transform of the sum of `nparts` partial spectra (n3 apart, added in order) of one cell into a real field;
// cluster kernel only
bool fft_cluster_supported(int N);
bool launch_fft3d_inverse_sum(sbte_ctx* c, const double2* parts, int nparts, double* out_real);
void launch_symmetrize_weights(sbte_ctx* c, const double* W, double* Ws);
// transposed pairing (f == g, tensor invariant under x <-> y of both indices): half of the zeta columns are streamed
bool qhat_stream_tp_supported(int N);
void launch_qhat_stream_tp(sbte_ctx* c, int npairs, const QhatPair* pairs, double2* qhat, bool sym, int nsplit);
void launch_transpose_xy(sbte_ctx* c, const double2* src, double2* dst);
int weights_xy_symmetry(sbte_ctx* c, const double* W, double* max_diff, double* max_abs);
// batched kernel (N in {8,16}): cell-minor operand layout, cells padded to a multiple of 32
bool qhat_batch_supported(int N);
int qhat_batch_cols(int N);
void launch_qhat_batch_any(sbte_ctx* c, const double2* spec_cellminor, double2* qhat, int cells, bool sym);
bool qhat_batch_pair_supported(int N);   // main + split tiles in one launch
void launch_qhat_batch_pair(sbte_ctx* c, const double2* spec, double2* parts, size_t part_stride, int cells_main, int cells_all,
                            const BatchSched& schM, const BatchSched& schS, int cg_split);
int qhat_batch_cut_mode(int N);      // 0 = stream-K cuts at whole xi_x chunks only, 1 = builder's choice, 2 = at any step
double qhat_batch_cut_cost(int N);   // measured price of cutting anywhere (fraction of the launch)
void launch_qhat_batch2(sbte_ctx* c, const double2* spec_cellminor, double2* parts, size_t part_stride, int cells,
                        const BatchSched& sch);
bool qhat_batch_split_supported(int N);
void launch_qhat_batch_split(sbte_ctx* c, const double2* spec_cellminor, double2* parts, size_t part_stride, int cells,
                             const BatchSched& sch, int cg_base);

// conserve.cu -- K4 / K5 / moments
void launch_conserve(sbte_ctx* c, double* Q, int batch);
// out = a*x + b*y + s*Q/Kn   (x, y, out may alias; y may be null)   -- time-integration glue
void launch_update(sbte_ctx* c, double* out, double a, const double* x, double b, const double* y, double s,
                   double Kn, const double* Q, long n);
void launch_moments(sbte_ctx* c, const double* f, double* mom8, int batch);   // rho,ux,uy,uz,T,Epos,Eneg,-
void launch_maxwellian_split(sbte_ctx* c, const double* f, const double* Msub, double* M, double* g);
void launch_moment_functionals(sbte_ctx* c, const double* Q, double* b5, int batch);

// weightgen.cu -- isotropic weights by adaptive GK21 on the device
int generate_weights_iso(sbte_ctx* c, double* d_W, double lambda, int* max_intervals);

}  // namespace sbte
