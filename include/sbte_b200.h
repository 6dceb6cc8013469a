/*
 * include/sbte_b200.h -- C ABI of libsbte_b200.so, the B200 (sm_100a) collision hot path of SpectralBTE.
 *
 * Two groups of entry points, all with C linkage, plain pointers and sizes:
 *
 *  (1) DROP-IN symbols: byte-for-byte the link interface of the reference's collision, conservation
 *      and transport modules, so exec/boltz.c and src/initializer.c link against this library in
 *      place of src/collisions.c, src/conserve.c, src/transportroutines.c and
 *      src/boundaryConditions.c with no source change.  They take HOST pointers,
 *      return void and, like the reference, report failure by printing and exit(1).
 *
 *  (2) sbte_* extension surface: the device-resident fast path (weights uploaded once, spectra,
 *      slabs and moments stay in HBM).  They return 0 on success, non-zero on failure with the
 *      message available from sbte_last_error().  Pointers named d_* are DEVICE pointers.
 *
 * There is no CPU fallback: every entry point fails loudly when no CUDA device is usable.
 * Citations are to the reference tree (/root/reference).
 */
#ifndef SBTE_B200_H
#define SBTE_B200_H
#include <stddef.h>
#if defined(__GNUC__)
#define SBTE_API __attribute__((visibility("default")))
#else
#define SBTE_API
#endif
#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ (1) drop-in link interface */

/* ABI mirror of `species` (src/species.h:11-26); only .mass and .name are read by this library. */
typedef struct sbte_species {
  size_t id;
  size_t num_levels;
  size_t *lev_id;
  double Rgas, mass, mm, d_ref, T_ref, mu_ref, omega, E0;
  double *Ei, *gi;
  char name[80];
} sbte_species;

/* src/collisions.h:27 (src/collisions.c:33-74). Retains nothing from vel/zeta after returning
 * (the reference keeps the pointers; the grids are copied to the device here). */
SBTE_API void initialize_coll(int nodes, double length, double *vel, double *zeta);
/* src/collisions.h:33 (src/collisions.c:81-89) */
SBTE_API void dealloc_coll(void);
/* src/collisions.h:38 (src/collisions.c:212-221). f, g, Q: N^3 host doubles; conv_weights: N^3 row
 * pointers of N^3 doubles (src/weights.c:61-63). The rows are uploaded once and cached, keyed on
 * the conv_weights pointer value; they are treated as immutable (BASELINE/SURVEY 8b). */
SBTE_API void ComputeQ(double *f, double *g, double *Q, double **conv_weights);
/* src/collisions.h:40 (src/collisions.c:178-210) */
SBTE_API void ComputeQ_maxPreserve(double *f, double *g, double *Q, double **conv_weights);
/* src/collisions.h:42 (src/collisions.c:232-283); in/out: N^3 interleaved complex (fftw_complex) */
SBTE_API void fft3D(double (*in)[2], double (*out)[2], int invert);

/* src/conserve.h:6 (src/conserve.c:17-43); only num_spec == 1 is supported */
SBTE_API void initialize_conservation(int nodes, double h_v, double *vel, sbte_species *mix, int num_spec);
/* src/conserve.h:8 (src/conserve.c:45-73) */
SBTE_API void initialize_conservation_fast(int nodes, double h_v, double *vel);
/* src/conserve.h:16 (src/conserve.c:207-264): Q[0] -> N^3 host doubles, corrected in place */
SBTE_API void conserveAllMoments(double **Q);
/* src/conserve.h:10 (src/conserve.c:76-84) */
SBTE_API void dealloc_conservation(void);

/* src/transportroutines.h:3 (src/transportroutines.c:27-59) */
SBTE_API void initialize_transport(int numV, int numX, double lv, double *xnodes, double *dxnodes, double *vel, int IC,
                          double timestep, double TWall_in, sbte_species *mix);
/* src/transportroutines.h:20,22 (src/transportroutines.c:473-492): f, f_conv are arrays of
 * numX + 2*order cell pointers. Single rank per process (the reference's MPI halo is replaced by
 * sbte_slab_* + the caller's exchange, see section (2)). */
SBTE_API void advectOne(double **f, double **f_conv, int id);
SBTE_API void advectTwo(double **f, double **f_conv, int id);
/* src/transportroutines.h:24 (src/transportroutines.c:494-501) */
SBTE_API void dealloc_trans(void);

/* ------------------------------------------------------------------ (2) device-resident extension */

typedef struct sbte_ctx sbte_ctx;
typedef struct sbte_slab sbte_slab;

SBTE_API const char *sbte_last_error(void);

/* One context per (N, L_v) grid and device. v, eta: N host doubles each (src/initializer.c:66-82). */
SBTE_API int sbte_create(sbte_ctx **out, int N, double L_v, const double *v, const double *eta, int device);
SBTE_API int sbte_destroy(sbte_ctx *c);
/* number of visible GPUs; mutual peer access between the GPUs of two contexts (one process driving several GPUs) */
SBTE_API int sbte_device_count(void);
SBTE_API int sbte_enable_peer_access(sbte_ctx *a, sbte_ctx *b);
SBTE_API int sbte_sync(sbte_ctx *c);
SBTE_API void *sbte_stream(sbte_ctx *c);                      /* the cudaStream_t every kernel is launched on */
SBTE_API unsigned long long sbte_launch_count(sbte_ctx *c);   /* kernels launched so far through c */
SBTE_API int sbte_reserve(sbte_ctx *c, int cells);            /* pre-size scratch for a batch of cells */
/* CUDA-event timing of the convolution kernel (K2) alone, on the context stream: enable, run, then
 * read the summed device time and the number of K2 launches since the last read. */
SBTE_API int sbte_k2_profile(sbte_ctx *c, int enable);
SBTE_API int sbte_k2_profile_read(sbte_ctx *c, double *total_ms, int *launches);

/* raw device memory for C hosts without CUDA headers */
SBTE_API int sbte_dev_alloc(void **d_ptr, size_t bytes);
SBTE_API int sbte_dev_free(void *d_ptr);
SBTE_API int sbte_h2d(sbte_ctx *c, void *d_dst, const void *src, size_t bytes);
SBTE_API int sbte_d2h(sbte_ctx *c, void *dst, const void *d_src, size_t bytes);
SBTE_API int sbte_d2d(sbte_ctx *c, void *d_dst, const void *d_src, size_t bytes);   /* async on the context stream */

/* Weights: N^3 x N^3 doubles, row-major [zeta][xi]; file format = src/weights.c:78-88,101-103
 * (headerless native doubles, name Weights/N%d_isotropic_L_v%g_lambda%g.wts, :68). */
SBTE_API int sbte_weights_upload_rows(sbte_ctx *c, double *const *rows);
SBTE_API int sbte_weights_upload(sbte_ctx *c, const double *W);
SBTE_API int sbte_weights_load_file(sbte_ctx *c, const char *path);
SBTE_API int sbte_weights_bind_device(sbte_ctx *c, const double *d_W);
SBTE_API int sbte_weights_fill_synthetic(sbte_ctx *c, unsigned long long seed); /* splitmix64 -> [-0.5,0.5) */
SBTE_API const double *sbte_weights_device(sbte_ctx *c);
/* isotropic weights generated on the device (src/weights.c:156-281: adaptive GK21 per (zeta, xi) pair) */
SBTE_API int sbte_weights_generate_iso(sbte_ctx *c, double lambda);
/* write the bound weights in the reference's .wts format (src/weights.c:100-103) */
SBTE_API int sbte_weights_save_file(sbte_ctx *c, const char *path);

/* For f == g the summand W[zeta][xi] f^[xi] f^[zeta-xi] is symmetric under xi <-> zeta-xi, so the library
 * streams a symmetrised copy Ws = W + W o sigma over half of the xi_x planes (built lazily on the device;
 * costs one extra N^6 tensor). On by default; disable to stream the tensor exactly as the reference does. */
SBTE_API int sbte_set_symmetrize(sbte_ctx *c, int enable);
/* Isotropic weights (src/weights.c:265-281) are invariant under swapping the x and y axes of both indices, so for f == g
 * the rows of zeta column (zy, zx) are the rows of column (zx, zy) read against the x<->y transposed spectrum: the 0D
 * stream then reads only the columns zx >= zy (51.6 % of them at N = 32) and forms two outputs per weight.  Used only
 * after the bound tensor has been CHECKED (one pass, once) to have that invariance to 1e-14 of its largest entry; any other
 * tensor keeps the full stream.  On by default; sbte_xy_pairing_state reports -1 not examined / 0 not invariant / 1 in
 * use, and the measured relative deviation. */
SBTE_API int sbte_set_xy_pairing(sbte_ctx *c, int enable);
SBTE_API int sbte_xy_pairing_state(sbte_ctx *c, int *state, double *deviation);

/* The stream-K schedule of the batched convolution for `cells` cells on a device with `ctas` SMs (what the
 * library uploads before a batched ComputeQ; exec/boltz.c:285-345 has no counterpart -- its cells are a plain loop).
 * Pure host arithmetic, no device needed.  split bit 0: the second launch that serves a last cell group with at most
 * 16 live cells at N = 16 on tiles of 16 zeta_y columns x 16 cells (two columns per warp); bit 1: ranges cut at whole
 * xi_x chunks only (the library's schedule under SBTE_CHUNK_CUTS=1); bit 2: ranges cut at any step wherever the kernel
 * takes that (SBTE_CHUNK_CUTS=0); neither: the library's own choice between the two (the better predicted duration).
 * dims = {G, T, P, np_cols, kmax, np_len}; array pointers may be null:
 * query dims first, then pass arrays of P+1, T+1, P, T and np_len entries.  Fails for N without a scheduled kernel. */
SBTE_API int sbte_batch_schedule_host(int N, int cells, int sym, int ctas, int split, long long *cta_begin, long long *tile_begin,
                                      int *cta_tile, int *tile_first, unsigned char *np, int *dims);

/* kernel selection for the convolution */
enum { SBTE_K2_AUTO = 0, SBTE_K2_GENERIC = 1, SBTE_K2_STREAM = 2, SBTE_K2_BATCH = 3, SBTE_K2_STREAM_DEEP = 4 };

/* fft3D on device data (src/collisions.c:232-283); batch cells of N^3 interleaved complex */
SBTE_API int sbte_fft3d(sbte_ctx *c, const double *d_in, double *d_out, int invert, int batch);
/* Q^ = compute_Qhat before the inverse transform (src/collisions.c:108-165); d_qhat: batch x N^3 complex */
SBTE_API int sbte_qhat(sbte_ctx *c, const double *d_f, const double *d_g, double *d_qhat, int batch, int k2);
/* ComputeQ (src/collisions.c:212-221) for `batch` cells laid out [cell][N^3] */
SBTE_API int sbte_compute_q(sbte_ctx *c, const double *d_f, const double *d_g, double *d_Q, int batch, int k2);
/* ComputeQ_maxPreserve (src/collisions.c:178-210), one cell; one weight pass for the three products */
SBTE_API int sbte_compute_q_maxpreserve(sbte_ctx *c, const double *d_f, const double *d_g, double *d_Q, int k2);
/* conserveAllMoments (src/conserve.c:207-264) on `batch` cells in place */
SBTE_API int sbte_conserve(sbte_ctx *c, double *d_Q, int batch);
/* b = C Q, the five conserved functionals per cell (src/conserve.c:217-240); d_b5: batch x 5 */
SBTE_API int sbte_moment_functionals(sbte_ctx *c, const double *d_Q, double *d_b5, int batch);
/* per cell: rho, u_x, u_y, u_z, T, E_pos, E_neg, p (src/momentRoutines.c:58-183); d_mom8: batch x 8 */
SBTE_API int sbte_moments(sbte_ctx *c, const double *d_f, double *d_mom8, int batch);
/* one 0D time step, order 1 (Euler) or 2 (Heun) (exec/boltz.c:189-241); f stays on the device */
SBTE_API int sbte_step_0d(sbte_ctx *c, double *d_f, double dt, double Kn, int order, int k2);

/* host-pointer forms used by the drop-in symbols and by end-to-end timing */
SBTE_API int sbte_compute_q_host(sbte_ctx *c, const double *f, const double *g, double *Q, int k2);
SBTE_API int sbte_compute_q_maxpreserve_host(sbte_ctx *c, const double *f, const double *g, double *Q, int k2);

/* ---- 1D slabs: cells_local + 2*order cells of N^3 doubles, contiguous, ghosts at both ends ----
 * rank/nranks describe the block partition of the global mesh (src/mesh_setup.c:46-53); walls and
 * extrapolation apply only on rank 0 / rank nranks-1 (src/transportroutines.c:107-172,261-404). */
SBTE_API int sbte_slab_create(sbte_ctx *c, sbte_slab **out, int cells_local, int order, const double *x, const double *dx,
                     int init_field, double dt, int rank, int nranks);
SBTE_API int sbte_slab_destroy(sbte_slab *s);
/* TWall_in of initialize_transport (src/transportroutines.c:57; default 1 = src/initializer.c:275): Init_field 1's
 * left wall is a diffuse wall at 2 * TWall_in */
SBTE_API int sbte_slab_set_twall_in(sbte_slab *s, double TWall_in);
SBTE_API double *sbte_slab_f(sbte_slab *s);        /* device pointer of the f slab   (exec/boltz.c f_inhom) */
SBTE_API double *sbte_slab_fconv(sbte_slab *s);    /* device pointer of f_conv */
SBTE_API int sbte_slab_upload(sbte_slab *s, const double *f_host);      /* (cells_local + 2*order) x N^3 */
SBTE_API int sbte_slab_download(sbte_slab *s, double *f_host);
/* transport half of the step: advectOne / advectTwo on the slab (ghost cells must already hold the
 * neighbours' data when nranks > 1; sbte_slab_halo_* give the send/recv regions). which: 0 f->f_conv, 1 f_conv->f */
SBTE_API int sbte_slab_advect(sbte_slab *s, int which);
/* stage = 0,1 for the two upwindTwo passes inside advectTwo (halo needed before each) */
SBTE_API int sbte_slab_upwind_stage(sbte_slab *s, int which, int stage);
/* kept for callers of the staged interface; the closing average of advectTwo (src/transportroutines.c:487-491) is part
 * of the second upwind stage, so this does nothing */
SBTE_API int sbte_slab_advect_finish(sbte_slab *s, int which);
/* device pointers + element counts of the boundary cells to send / ghost cells to receive for the
 * array the next upwind pass reads. side 0 = left neighbour, 1 = right neighbour */
SBTE_API int sbte_slab_halo_regions(sbte_slab *s, int which, int stage, int side, double **d_send, double **d_recv,
                           size_t *count);
/* Peer-memory halo (replaces the MPI_Send/Recv of src/transportroutines.c:107-172,261-344 on one NVLink node,
 * one process per GPU): every rank exports CUDA IPC handles of its slabs (256 bytes), imports its left (side 0)
 * and right (side 1) neighbours' (for Init_field 6 at order 1 the ring closes: rank 0's left is the last rank),
 * then enables the mode. The upwind kernels then read the neighbours' boundary cells directly over NVLink and
 * ranks are ordered by two device-side counters per rank; no ghost-cell messages, no host synchronisation. */
SBTE_API int sbte_slab_ipc_export(sbte_slab *s, unsigned char *handles256);
SBTE_API int sbte_slab_ipc_import(sbte_slab *s, int side, const unsigned char *handles256, int neighbour_cells);
/* same-process form: `other` is a slab of another context (another stream or GPU with peer access enabled) */
SBTE_API int sbte_slab_peer_attach(sbte_slab *s, int side, sbte_slab *other);
SBTE_API int sbte_slab_set_peer_halo(sbte_slab *s, int enable);
/* unmap the neighbours' slabs and leave the peer mode; with one process per GPU every rank detaches, the ranks
 * synchronise, and only then are the slabs destroyed (exported memory must outlive its remote mappings) */
SBTE_API int sbte_slab_peer_detach(sbte_slab *s);
/* this rank's {ready, done, epoch, error} words.  error != 0: a device-side wait for a neighbour ran out of time
 * (the waiting kernels do not trap: they raise this word, stop waiting, and sbte_slab_moments / sbte_slab_download
 * then fail with a message). */
SBTE_API int sbte_slab_halo_state(sbte_slab *s, int *state4);
/* bound of those waits in seconds (default 120, or SBTE_HALO_TIMEOUT_S); <= 0: wait for ever */
SBTE_API int sbte_slab_set_halo_timeout(sbte_slab *s, double seconds);
/* collision half: per-cell ComputeQ + conserve + Euler / Heun (exec/boltz.c:285-345) */
SBTE_API int sbte_slab_collide(sbte_slab *s, double Kn, int k2);
/* whole single-rank step (exec/boltz.c:264-353) */
SBTE_API int sbte_slab_step(sbte_slab *s, double Kn, int k2);
/* rho, u_x, u_y, u_z, T, Epos, Eneg, p of every owned cell -> host (cells_local x 8) */
SBTE_API int sbte_slab_moments(sbte_slab *s, double *mom_host);

#ifdef __cplusplus
}
#endif
#endif
