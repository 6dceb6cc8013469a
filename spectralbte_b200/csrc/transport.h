// spectralbte_b200/csrc/transport.h -- launchers of the transport kernels (transport.cu)
#pragma once
#include <cuda_runtime.h>
namespace sbte {
void launch_diffuse_bc(cudaStream_t st, const double* in, double* out, const double* v, const double* wt, int N,
                       double hv, double TW, int bdry);
void launch_upwind_one(cudaStream_t st, const double* f, double* fc, const double* v, const double* dx, int N, int nX,
                       double dt, const double* peerL, const double* peerR);
int preload_transport_kernels();
void launch_halo_begin(cudaStream_t st, int* my, const int* nbL, const int* nbR, long long timeout_cycles);
void launch_halo_end(cudaStream_t st, int* my);
void launch_halo_quiesce(cudaStream_t st, int* my, const int* nbL, const int* nbR, long long timeout_cycles);
void launch_extrapolate(cudaStream_t st, double* f, long n3, int dst, int a, int b);
void launch_wall_face(cudaStream_t st, const double* f, double* face, const double* x, const double* dx, int N, int l,
                      int right, int fill_noflux);
void launch_upwind_two(cudaStream_t st, const double* f, double* fc, const double* fl, const double* fr,
                       const double* v, const double* x, const double* dx, int N, int nX, double dt, int left_wall,
                       int right_wall, const double* peerL, const double* peerR, double force);
void launch_average(cudaStream_t st, const double* f, double* fc, long n);
}  // namespace sbte
