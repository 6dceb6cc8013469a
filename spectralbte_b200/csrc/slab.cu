// spectralbte_b200/csrc/slab.cu -- device-resident 1D-3V state: a slab of cells_local + 2*order
// cells per rank, the transport half-steps and the batched collision half-step.
//
// Mirrors the 1D branch of the reference driver (/root/reference/exec/boltz.c:264-353) and the
// ghost-cell logic of src/transportroutines.c:107-172 (order 1) and :261-404 (order 2), with the
// blocking MPI halo replaced either by peer memory (the stencil kernels read the neighbour GPU's boundary cells and
// order themselves with device-side counters: sbte_slab_ipc_* / peer_attach, transport.cu halo_enter / halo_leave)
// or by "ghost cells are filled by the caller" (NCCL between the regions reported by sbte_slab_halo_regions);
// either way f never leaves the device.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/sbte_b200.h"
#include "internal.h"
#include "transport.h"

namespace sbte {
int compute_q_dev(sbte_ctx* c, const double* d_f, const double* d_g, double* d_Q, int batch, int k2);
int collide_stage_dev(sbte_ctx* c, const double* d_src, double* d_Q, int batch, int k2, double* out, double a,
                      const double* x, double b, const double* y, double s, double Kn, int chain, int* chained);
}

struct sbte_slab {
  sbte_ctx* c = nullptr;
  int nX = 0, order = 1, ic = 0, rank = 0, nranks = 1, ncell = 0;
  double dt = 0;
  double TWall_in = 1.0;   // Init_field 1 wall parameter (src/initializer.c:275, src/transportroutines.c:57)
  double *d_x = nullptr, *d_dx = nullptr;
  double *d_f = nullptr, *d_fc = nullptr, *d_f1 = nullptr, *d_ft = nullptr;  // f, f_conv, f_1, f_tmp
  double *d_fl = nullptr, *d_fr = nullptr;                                   // wall faces
  double* d_Q = nullptr;                                                     // nX cells
  double* d_mom = nullptr;
  // peer-memory halo (one process per GPU, CUDA IPC): the neighbours' slabs mapped into this process
  struct Peer {
    bool on = false, mapped = false;                // mapped: opened through CUDA IPC (closed on destroy)
    double* arr[3] = {nullptr, nullptr, nullptr};   // their f, f_conv, f_tmp
    int* flags = nullptr;                           // their {ready, done} counters
    int cells = 0;
  } nb[2];
  int* d_flags = nullptr;   // my {ready, done, epoch}; the pass counter lives on the device (graph replay)
  int p2p = 0;              // 1: stencils read the neighbours' boundary cells over NVLink
  long long halo_timeout = 0;   // bound of every device-side wait for a neighbour, in SM clock cycles (0 = none)
  double halo_timeout_s = 0;
  // CUDA graph of one whole time step (single rank or peer halos: nothing but launches on one stream)
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    double Kn = 0;
    int k2 = 0, seen = 0;
    unsigned long long gen = 0, launches = 0;
  } graph;
};

using namespace sbte;

#define CKS(call)                                                                 \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      sbte::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));        \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

static int launch_ok(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error(std::string(what) + ": " + cudaGetErrorString(e)); return 1; }
  return 0;
}

static const double T0_WALL = 1.0, T1_WALL = 2.0;  // src/transportroutines.c:46-47

static inline double* cell(double* base, long n3, int l) { return base + (long)l * n3; }

// seconds -> SM clock cycles at the device's maximum clock (clock64 never runs faster, so the wait lasts at least
// `seconds`); seconds <= 0: unbounded
static long long timeout_cycles(int device, double seconds) {
  if (!(seconds > 0)) return 0;
  int khz = 0;
  if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device) != cudaSuccess || khz <= 0) khz = 2000000;
  return (long long)(seconds * 1e3 * (double)khz);
}

// the error word of the peer halo (a wait for a neighbour ran out of time): checked where the host synchronises anyway
static int peer_halo_failed(sbte_slab* s) {
  if (!s->p2p || !s->d_flags) return 0;
  int err = 0;
  if (cudaMemcpy(&err, s->d_flags + 3, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  if (!err) return 0;
  char msg[256];
  snprintf(msg, sizeof msg, "peer halo: a neighbouring rank did not arrive within %g s (SBTE_HALO_TIMEOUT_S / "
           "sbte_slab_set_halo_timeout; 0 = wait for ever); the slab contents are invalid from that pass on", s->halo_timeout_s);
  set_error(msg);
  return 1;
}

// ghost cells owned by the physical boundaries, order 1 (src/transportroutines.c:116-137,156-172)
static void fill_ghosts_one(sbte_slab* s, double* f) {
  sbte_ctx* c = s->c;
  const long n3 = c->n3;
  const size_t cb = (size_t)n3 * sizeof(double);
  const int nX = s->nX, ic = s->ic;
  cudaStream_t st = c->stream;
  const bool first = s->rank == 0, last = s->rank == s->nranks - 1;
  if (first) {
    if (ic == 3 || ic == 5) { launch_diffuse_bc(st, cell(f, n3, 1), cell(f, n3, 0), c->d_v, c->d_wt, c->N, c->dv, T0_WALL, 0); c->launches++; }
    else if (ic == 1) { launch_diffuse_bc(st, cell(f, n3, 1), cell(f, n3, 0), c->d_v, c->d_wt, c->N, c->dv, 2.0 * s->TWall_in, 0); c->launches++; }
    else if (ic != 6) cudaMemcpyAsync(cell(f, n3, 0), cell(f, n3, 1), cb, cudaMemcpyDeviceToDevice, st);
  }
  if (last) {
    if (ic == 3 || ic == 5) { launch_diffuse_bc(st, cell(f, n3, nX), cell(f, n3, nX + 1), c->d_v, c->d_wt, c->N, c->dv, T1_WALL, 1); c->launches++; }
    else if (ic != 6) cudaMemcpyAsync(cell(f, n3, nX + 1), cell(f, n3, nX), cb, cudaMemcpyDeviceToDevice, st);
  }
  if (ic == 6 && s->nranks == 1) {  // periodic wrap inside one rank (:167-170)
    cudaMemcpyAsync(cell(f, n3, 0), cell(f, n3, nX), cb, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(cell(f, n3, nX + 1), cell(f, n3, 1), cb, cudaMemcpyDeviceToDevice, st);
  }
}

// which of the three exchangeable arrays is `p` (index into Peer::arr), -1 if none
static int array_id(const sbte_slab* s, const double* p) {
  if (p == s->d_f) return 0;
  if (p == s->d_fc) return 1;
  if (p == s->d_ft) return 2;
  return -1;
}

// peer-memory halo of an upwind pass reading `src`: the mapped boundary cells of the neighbours (null where there is
// no peer) and the flag words the stencil kernel orders itself with
static HaloSync peer_halo(sbte_slab* s, const double* src, const double** peerL, const double** peerR) {
  *peerL = *peerR = nullptr;
  HaloSync hs = {nullptr, nullptr, nullptr, 0};
  if (!s->p2p) return hs;
  const long n3 = s->c->n3;
  const int id = array_id(s, src);
  if (s->nb[0].on) *peerL = s->nb[0].arr[id] + (long)s->nb[0].cells * n3;   // their last `order` owned cells
  if (s->nb[1].on) *peerR = s->nb[1].arr[id] + (long)s->order * n3;         // their first `order` owned cells
  hs.my = s->d_flags;
  hs.nbL = s->nb[0].on ? s->nb[0].flags : nullptr;
  hs.nbR = s->nb[1].on ? s->nb[1].flags : nullptr;
  hs.timeout = s->halo_timeout;
  return hs;
}

// one upwindTwo pass src -> dst (src/transportroutines.c:241-470); ghosts from neighbours must be in place.
// avg != null: dst = (avg + pass result) / 2, the closing average of advectTwo (:487-491) folded into its second pass.
static void upwind_two_pass(sbte_slab* s, double* src, double* dst, const double* avg) {
  sbte_ctx* c = s->c;
  const int nX = s->nX, ic = s->ic, N = c->N;
  cudaStream_t st = c->stream;
  const bool first = s->rank == 0, last = s->rank == s->nranks - 1;
  const bool wallL = (ic == 3 || ic == 5 || ic == 1), wallR = (ic == 3 || ic == 5);
  // Physical ends this rank holds.  With a wall model: extrapolated ghost + outgoing half of the wall face (one launch
  // for both ends), then the diffuse kernel fills the incoming half.  Without one (no-flux fill) the stencil kernel forms
  // ghost and face itself (SBTE_NO_EDGE_FUSE=1: the separate launch, for A/B runs).
  static const bool fuse = getenv("SBTE_NO_EDGE_FUSE") == nullptr;
  const bool prepL = first && (wallL || !fuse), prepR = last && (wallR || !fuse);
  if (prepL || prepR) {
    launch_edge_prep(st, src, s->d_fl, s->d_fr, s->d_x, s->d_dx, N, nX, prepL ? 1 : 0, prepR ? 1 : 0, wallL ? 0 : 1, wallR ? 0 : 1);
    c->launches++;
  }
  if (first && wallL) {
    launch_diffuse_bc(st, s->d_fl, s->d_fl, c->d_v, c->d_wt, N, c->dv, (ic == 1) ? 2.0 * s->TWall_in : T0_WALL, 0);
    c->launches++;
  }
  if (last && wallR) { launch_diffuse_bc(st, s->d_fr, s->d_fr, c->d_v, c->d_wt, N, c->dv, T1_WALL, 1); c->launches++; }
  const double *peerL, *peerR;
  const HaloSync hs = peer_halo(s, src, &peerL, &peerR);
  // Poiseuille forcing coefficient Ma*0.5*dt/(2*h_v), Ma = 1, h_v = 2 L_v/(N-1) (src/transportroutines.c:37,255,431)
  const double force = (ic == 5) ? 1.0 * 0.5 * s->dt / (2 * (2 * c->L_v / (N - 1))) : 0.0;
  launch_upwind_two(st, src, dst, s->d_fl, s->d_fr, c->d_v, s->d_x, s->d_dx, N, nX, s->dt, first ? (prepL ? 1 : 2) : 0,
                    last ? (prepR ? 1 : 2) : 0, peerL, peerR, force, avg, hs);
  c->launches++;
}

static void pick(sbte_slab* s, int which, double*& A, double*& B) {
  if (which == 0) { A = s->d_f; B = s->d_fc; } else { A = s->d_fc; B = s->d_f; }
}

extern "C" {

int sbte_slab_create(sbte_ctx* c, sbte_slab** out, int cells_local, int order, const double* x, const double* dx,
                     int init_field, double dt, int rank, int nranks) {
  *out = nullptr;
  if (order != 1 && order != 2) { set_error("Space_order must be 1 or 2"); return 1; }
  // Init_field 5 (Poiseuille): forcing at order 2; at order 1 the reference's forcing block (:217-225) has no
  // observable effect (it runs after the k loop closed and every write is overwritten), reproduced as such
  if (cells_local < 2 * order) { set_error("too few cells per rank"); return 1; }
  CKS(cudaSetDevice(c->device));
  sbte_slab* s = new sbte_slab();
  s->c = c; s->nX = cells_local; s->order = order; s->ic = init_field; s->dt = dt; s->rank = rank; s->nranks = nranks;
  s->ncell = cells_local + 2 * order;
  {
    const char* e = getenv("SBTE_HALO_TIMEOUT_S");
    s->halo_timeout_s = (e && *e) ? atof(e) : 120.0;
    s->halo_timeout = timeout_cycles(c->device, s->halo_timeout_s);
  }
  const long n3 = c->n3;
  const size_t sb = (size_t)s->ncell * n3 * sizeof(double);
  std::vector<double> hx(x, x + s->ncell), hdx(dx, dx + s->ncell);
  if (rank == nranks - 1) {
    // right ghost coordinates by extension, as the last-rank branch of src/mesh_setup.c:165-175 does
    // (the reference's single-rank branch leaves them uninitialised, :121-146)
    for (int g = cells_local + order; g < s->ncell; g++) { hdx[g] = hdx[g - 1]; hx[g] = hx[g - 1] + hdx[g - 1]; }
  }
  CKS(cudaMalloc(&s->d_x, s->ncell * sizeof(double)));
  CKS(cudaMalloc(&s->d_dx, s->ncell * sizeof(double)));
  CKS(cudaMemcpy(s->d_x, hx.data(), s->ncell * sizeof(double), cudaMemcpyHostToDevice));
  CKS(cudaMemcpy(s->d_dx, hdx.data(), s->ncell * sizeof(double), cudaMemcpyHostToDevice));
  CKS(cudaMalloc(&s->d_f, sb));
  CKS(cudaMalloc(&s->d_fc, sb));
  CKS(cudaMemset(s->d_f, 0, sb));
  CKS(cudaMemset(s->d_fc, 0, sb));
  if (order == 2) {
    CKS(cudaMalloc(&s->d_f1, sb));
    CKS(cudaMalloc(&s->d_ft, sb));
    CKS(cudaMemset(s->d_f1, 0, sb));
    CKS(cudaMemset(s->d_ft, 0, sb));
    CKS(cudaMalloc(&s->d_fl, n3 * sizeof(double)));
    CKS(cudaMalloc(&s->d_fr, n3 * sizeof(double)));
  }
  CKS(cudaMalloc(&s->d_Q, (size_t)cells_local * n3 * sizeof(double)));
  CKS(cudaMalloc(&s->d_mom, (size_t)cells_local * 8 * sizeof(double)));
  if (ensure_capacity(c, cells_local)) { sbte_slab_destroy(s); return 1; }
  *out = s;
  return 0;
}

// Unmaps the neighbours' slabs (CUDA IPC) and switches the peer halo off.  Exported memory must not be freed while
// another process still maps it, so multi-process callers detach on every rank, synchronise the ranks, and only then
// destroy their slabs (spectralbte_b200/halo.py SlabHalo.close).
int sbte_slab_peer_detach(sbte_slab* s) {
  cudaSetDevice(s->c->device);
  cudaStreamSynchronize(s->c->stream);
  if (s->graph.exec) { cudaGraphExecDestroy(s->graph.exec); }   // the captured step references the mapped pointers
  s->graph = sbte_slab::StepGraph();
  for (int side = 0; side < 2; side++) {
    sbte_slab::Peer& p = s->nb[side];
    if (p.on && p.mapped) {
      // (a closed ring of two ranks maps the same neighbour on both sides: mappings are reference-counted, one close each)
      for (int a = 0; a < 3; a++)
        if (p.arr[a]) cudaIpcCloseMemHandle(p.arr[a]);
      if (p.flags) cudaIpcCloseMemHandle(p.flags);
    }
  }
  for (int side = 0; side < 2; side++) s->nb[side] = sbte_slab::Peer();
  s->p2p = 0;
  cudaGetLastError();
  return 0;
}

int sbte_slab_destroy(sbte_slab* s) {
  if (!s) return 0;
  cudaSetDevice(s->c->device);
  sbte_slab_peer_detach(s);
  cudaFree(s->d_flags);
  cudaFree(s->d_x); cudaFree(s->d_dx); cudaFree(s->d_f); cudaFree(s->d_fc); cudaFree(s->d_f1); cudaFree(s->d_ft);
  cudaFree(s->d_fl); cudaFree(s->d_fr); cudaFree(s->d_Q); cudaFree(s->d_mom);
  delete s;
  return 0;
}

// ---- peer-memory halo set-up (one process per GPU): export my slabs, import the neighbours'
int sbte_slab_ipc_export(sbte_slab* s, unsigned char* handles256) {
  cudaSetDevice(s->c->device);
  if (!s->d_flags) {
    CKS(cudaMalloc(&s->d_flags, 256));
    CKS(cudaMemset(s->d_flags, 0, 256));
  }
  memset(handles256, 0, 256);
  cudaIpcMemHandle_t h;
  double* arrs[3] = {s->d_f, s->d_fc, s->d_ft};
  for (int a = 0; a < 3; a++) {
    if (!arrs[a]) continue;
    CKS(cudaIpcGetMemHandle(&h, arrs[a]));
    memcpy(handles256 + 64 * a, &h, sizeof(h));
  }
  CKS(cudaIpcGetMemHandle(&h, s->d_flags));
  memcpy(handles256 + 192, &h, sizeof(h));
  return 0;
}

int sbte_slab_ipc_import(sbte_slab* s, int side, const unsigned char* handles256, int neighbour_cells) {
  cudaSetDevice(s->c->device);
  if (side < 0 || side > 1) { set_error("side must be 0 (left) or 1 (right)"); return 1; }
  sbte_slab::Peer& p = s->nb[side];
  cudaIpcMemHandle_t h;
  double* mine[3] = {s->d_f, s->d_fc, s->d_ft};
  for (int a = 0; a < 3; a++) {
    if (!mine[a]) continue;
    memcpy(&h, handles256 + 64 * a, sizeof(h));
    CKS(cudaIpcOpenMemHandle((void**)&p.arr[a], h, cudaIpcMemLazyEnablePeerAccess));
  }
  memcpy(&h, handles256 + 192, sizeof(h));
  CKS(cudaIpcOpenMemHandle((void**)&p.flags, h, cudaIpcMemLazyEnablePeerAccess));
  p.cells = neighbour_cells;
  p.on = true;
  p.mapped = true;
  return 0;
}

// same-process form (one process driving several slabs / GPUs with peer access already enabled): no IPC mapping
int sbte_slab_peer_attach(sbte_slab* s, int side, sbte_slab* other) {
  if (side < 0 || side > 1) { set_error("side must be 0 (left) or 1 (right)"); return 1; }
  if (other->order != s->order || other->c->N != s->c->N) { set_error("neighbour slab has a different shape"); return 1; }
  for (sbte_slab* t : {s, other})
    if (!t->d_flags) {
      CKS(cudaSetDevice(t->c->device));
      CKS(cudaMalloc(&t->d_flags, 256));
      CKS(cudaMemset(t->d_flags, 0, 256));
    }
  sbte_slab::Peer& p = s->nb[side];
  p.arr[0] = other->d_f; p.arr[1] = other->d_fc; p.arr[2] = other->d_ft;
  p.flags = other->d_flags;
  p.cells = other->nX;
  p.on = true;
  p.mapped = false;
  return 0;
}

// diagnostic: this rank's {ready, done, epoch, error} words (synchronous copy)
int sbte_slab_halo_state(sbte_slab* s, int* state4) {
  cudaSetDevice(s->c->device);
  state4[0] = state4[1] = state4[2] = state4[3] = 0;
  if (!s->d_flags) return 0;
  CKS(cudaMemcpy(state4, s->d_flags, 4 * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

// bound of the device-side waits for a neighbour; seconds <= 0: wait for ever.  A captured step graph carries the
// old bound as a kernel argument, so it is dropped and re-captured.
int sbte_slab_set_halo_timeout(sbte_slab* s, double seconds) {
  cudaSetDevice(s->c->device);
  s->halo_timeout_s = seconds > 0 ? seconds : 0;
  s->halo_timeout = timeout_cycles(s->c->device, seconds);
  if (s->graph.exec) {
    cudaStreamSynchronize(s->c->stream);
    cudaGraphExecDestroy(s->graph.exec);
  }
  s->graph = sbte_slab::StepGraph();
  return 0;
}

int sbte_slab_set_peer_halo(sbte_slab* s, int enable) {
  cudaSetDevice(s->c->device);
  if (enable && !s->d_flags) { set_error("export/import the IPC handles before enabling peer halos"); return 1; }
  if (enable && preload_transport_kernels()) { set_error("could not load the transport kernels"); return 1; }
  s->p2p = enable ? 1 : 0;
  return 0;
}

// TWall_in of initialize_transport (src/transportroutines.c:57): the Init_field 1 left wall is a diffuse wall at
// 2 * TWall_in (:120,:365).  Baked into a captured step graph as a kernel argument, so the graph is dropped.
int sbte_slab_set_twall_in(sbte_slab* s, double TWall_in) {
  cudaSetDevice(s->c->device);
  if (!(TWall_in > 0)) { set_error("TWall_in must be positive"); return 1; }
  s->TWall_in = TWall_in;
  if (s->graph.exec) {
    cudaStreamSynchronize(s->c->stream);
    cudaGraphExecDestroy(s->graph.exec);
  }
  s->graph = sbte_slab::StepGraph();
  return 0;
}

double* sbte_slab_f(sbte_slab* s) { return s->d_f; }
double* sbte_slab_fconv(sbte_slab* s) { return s->d_fc; }

int sbte_slab_upload(sbte_slab* s, const double* f_host) {
  cudaSetDevice(s->c->device);
  return sbte_h2d(s->c, s->d_f, f_host, (size_t)s->ncell * s->c->n3 * sizeof(double));
}
int sbte_slab_download(sbte_slab* s, double* f_host) {
  cudaSetDevice(s->c->device);
  if (sbte_d2h(s->c, f_host, s->d_f, (size_t)s->ncell * s->c->n3 * sizeof(double))) return 1;
  return peer_halo_failed(s);
}

int sbte_slab_halo_regions(sbte_slab* s, int which, int stage, int side, double** d_send, double** d_recv,
                           size_t* count) {
  double *A, *B;
  pick(s, which, A, B);
  double* arr = (s->order == 2 && stage == 1) ? s->d_ft : A;
  const long n3 = s->c->n3;
  const int o = s->order, nX = s->nX;
  *count = (size_t)o * n3;
  if (side == 0) {  // left neighbour: send my first `o` owned cells, receive into my left ghosts
    *d_send = cell(arr, n3, o);
    *d_recv = cell(arr, n3, 0);
  } else {          // right neighbour: send my last `o` owned cells, receive into my right ghosts
    *d_send = cell(arr, n3, nX);
    *d_recv = cell(arr, n3, nX + o);
  }
  return 0;
}

int sbte_slab_upwind_stage(sbte_slab* s, int which, int stage) {
  cudaSetDevice(s->c->device);
  double *A, *B;
  pick(s, which, A, B);
  sbte_ctx* c = s->c;
  if (s->order == 1) {
    fill_ghosts_one(s, A);
    const double *peerL, *peerR;
    const HaloSync hs = peer_halo(s, A, &peerL, &peerR);
    launch_upwind_one(c->stream, A, B, c->d_v, s->d_dx, c->N, s->nX, s->dt, peerL, peerR, hs);
    c->launches++;
  } else {
    if (stage == 0) upwind_two_pass(s, A, s->d_ft, nullptr);
    else upwind_two_pass(s, s->d_ft, B, A);   // with the closing average (A + pass) / 2
  }
  return launch_ok("upwind");
}

// kept for callers of the staged interface: the average of advectTwo (src/transportroutines.c:487-491) is now part of
// the second upwind stage, so there is nothing left to do here
int sbte_slab_advect_finish(sbte_slab* s, int which) {
  (void)s; (void)which;
  return 0;
}

int sbte_slab_advect(sbte_slab* s, int which) {
  cudaSetDevice(s->c->device);
  if (s->nranks != 1 && !s->p2p) { set_error("sbte_slab_advect needs ghost cells: exchange halos with the staged calls or enable peer halos"); return 1; }
  if (sbte_slab_upwind_stage(s, which, 0)) return 1;
  if (s->order == 2) {
    if (sbte_slab_upwind_stage(s, which, 1)) return 1;
    if (sbte_slab_advect_finish(s, which)) return 1;
  }
  return 0;
}

// exec/boltz.c:285-345 for all owned cells at once
int sbte_slab_collide(sbte_slab* s, double Kn, int k2) {
  cudaSetDevice(s->c->device);
  sbte_ctx* c = s->c;
  const long n3 = c->n3;
  const int o = s->order, nX = s->nX;
  double* fc = cell(s->d_fc, n3, o);
  double* f = cell(s->d_f, n3, o);
  if (s->p2p && o == 1) {   // first order: the update below overwrites f, which the neighbours read in the same pass
    launch_halo_quiesce(c->stream, s->d_flags, s->nb[0].on ? s->nb[0].flags : nullptr, s->nb[1].on ? s->nb[1].flags : nullptr,
                        s->halo_timeout);
    c->launches++;
  }
  if (o == 1) {
    // f = f_conv + dt Q / Kn
    if (collide_stage_dev(c, fc, s->d_Q, nX, k2, f, 1.0, fc, 0.0, nullptr, s->dt, Kn, 0, nullptr)) return 1;
  } else {
    double* f1 = cell(s->d_f1, n3, o);
    // f_1 = f_conv + dt Q(f_conv) / Kn ;  f_conv = (f_conv + f_1)/2 + dt/2 Q(f_1) / Kn  (Heun)
    // the first stage's closing kernel goes on with the forward transform of f_1, which the second stage starts from
    int chained = 0;
    if (collide_stage_dev(c, fc, s->d_Q, nX, k2, f1, 1.0, fc, 0.0, nullptr, s->dt, Kn, 2, &chained)) return 1;
    if (collide_stage_dev(c, f1, s->d_Q, nX, k2, fc, 0.5, fc, 0.5, f1, 0.5 * s->dt, Kn, chained ? 1 : 0, nullptr)) return 1;
  }
  return launch_ok("collide");
}

static int slab_step_direct(sbte_slab* s, double Kn, int k2) {
  if (sbte_slab_advect(s, 0)) return 1;
  if (sbte_slab_collide(s, Kn, k2)) return 1;
  if (s->order == 2 && sbte_slab_advect(s, 1)) return 1;
  return 0;
}

// One time step.  After a first direct step (which also sizes every buffer) the launch sequence is captured once
// and replayed: ~20-40 launches per step become one graph launch, which is what a small slab per GPU needs.
int sbte_slab_step(sbte_slab* s, double Kn, int k2) {
  cudaSetDevice(s->c->device);
  sbte_ctx* c = s->c;
  static const bool no_graph = getenv("SBTE_NO_GRAPH") != nullptr;
  if (no_graph || c->k2_prof) return slab_step_direct(s, Kn, k2);
  sbte_slab::StepGraph& g = s->graph;
  if (g.exec && (g.Kn != Kn || g.k2 != k2 || g.gen != c->graph_gen)) {   // arguments or the context's buffers changed
    cudaGraphExecDestroy(g.exec);
    g = sbte_slab::StepGraph();
  }
  if (g.exec) {
    CKS(cudaGraphLaunch(g.exec, c->stream));
    c->launches += g.launches;
    return 0;
  }
  if (g.seen == 0 || g.Kn != Kn || g.k2 != k2 || g.gen != c->graph_gen) {
    if (slab_step_direct(s, Kn, k2)) return 1;
    g.seen = 1; g.Kn = Kn; g.k2 = k2; g.gen = c->graph_gen;
    return 0;
  }
  const unsigned long long before = c->launches;
  CKS(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = slab_step_direct(s, Kn, k2);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
  if (rc != 0 || e != cudaSuccess || !graph || g.gen != c->graph_gen) {
    if (graph) cudaGraphDestroy(graph);
    if (rc == 0) set_error(std::string("graph capture of the slab step failed: ") + cudaGetErrorString(e));
    g.seen = 0;
    return 1;
  }
  g.launches = c->launches - before;
  c->launches = before;
  e = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { g.exec = nullptr; set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); return 1; }
  CKS(cudaGraphLaunch(g.exec, c->stream));
  c->launches += g.launches;
  return 0;
}

int sbte_slab_moments(sbte_slab* s, double* mom_host) {
  cudaSetDevice(s->c->device);
  sbte_ctx* c = s->c;
  launch_moments(c, cell(s->d_f, c->n3, s->order), s->d_mom, s->nX);
  if (launch_ok("moments")) return 1;
  if (sbte_d2h(c, mom_host, s->d_mom, (size_t)s->nX * 8 * sizeof(double))) return 1;
  return peer_halo_failed(s);
}

}  // extern "C"
