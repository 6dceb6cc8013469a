// spectralbte_b200/csrc/transport.h -- launchers of the transport kernels (transport.cu)
#pragma once
#include <cuda_runtime.h>
namespace sbte {
// peer-memory halo of a stencil pass: this rank's flag words and the neighbours' (null: no neighbour on that side);
// my == null: no cross-GPU ordering (one rank, or ghost cells exchanged by messages)
struct HaloSync {
  int* my;
  const int* nbL;
  const int* nbR;
  long long timeout;
};
void launch_diffuse_bc(cudaStream_t st, const double* in, double* out, const double* v, const double* wt, int N,
                       double hv, double TW, int bdry);
void launch_upwind_one(cudaStream_t st, const double* f, double* fc, const double* v, const double* dx, int N, int nX,
                       double dt, const double* peerL, const double* peerR, const HaloSync& hs);
int preload_transport_kernels();
void launch_halo_quiesce(cudaStream_t st, int* my, const int* nbL, const int* nbR, long long timeout_cycles);
void launch_upwind_two(cudaStream_t st, const double* f, double* fc, const double* fl, const double* fr,
                       const double* v, const double* x, const double* dx, int N, int nX, double dt, int left_wall,
                       int right_wall, const double* peerL, const double* peerR, double force, const double* avg,
                       const HaloSync& hs);
void launch_edge_prep(cudaStream_t st, double* f, double* fl, double* fr, const double* x, const double* dx, int N, int nX,
                      int do_left, int do_right, int fill_left, int fill_right);
}  // namespace sbte
