/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99) of the SpectralBTE collision hot path, used as the parity checker
 * for the CUDA library.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product (spectralbte_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_*.py check this restatement against
 *   (1) the reference's golden vectors tests/BKW8/target/ and tests/heat_transport/target/
 *       (committed under tests/golden/), and
 *   (2) the reference's own C sources compiled unmodified into oracle/_ref/libref.so
 *       (oracle/build_ref.sh), function by function on seeded inputs.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Layout differs from the reference on purpose: one context struct instead of file-static
 * globals, one contiguous weight matrix W[zeta][xi] instead of N^3 row pointers, contiguous
 * slabs f[cell][N^3] instead of arrays of cell pointers.  Index convention everywhere:
 * flat = k + N*(j + N*i), i <-> v_x (slowest).  Complex arrays are interleaved (re,im).
 */
#ifndef ORC_ORACLE_H
#define ORC_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

/* grid_rule 0: 0D rule (src/initializer.c:66-82); 1: 1D rule (src/initializer.c:256-266) */
orc_ctx *orc_create(int N, double L_v, int grid_rule);
void orc_destroy(orc_ctx *c);
void orc_get_grid(const orc_ctx *c, double *v, double *eta);
int orc_n(const orc_ctx *c);

/* src/collisions.c:232-283 */
void orc_fft3d(orc_ctx *c, const double *in, double *out, int invert);
/* bare unnormalised 3-D DFT (what fftw_execute does at src/collisions.c:270); sign -1 fwd, +1 bwd */
void orc_dft3(orc_ctx *c, double *io, int sign);
/* src/collisions.c:127-165 ; W is N^3 x N^3 row-major [zeta][xi] */
void orc_qhat(orc_ctx *c, const double *W, const double *fhat, const double *ghat, double *qhat);
/* src/collisions.c:108-169 + 212-221 ; qhat_out (2*N^3) may be NULL */
void orc_compute_q(orc_ctx *c, const double *W, const double *f, const double *g, double *Q,
                   double *qhat_out);
void orc_compute_q_rowmod(orc_ctx *c, const double *W, long distinct_rows, const double *f, const double *g,
                          double *Q);
/* src/collisions.c:91-106,178-210 */
void orc_compute_q_maxpreserve(orc_ctx *c, const double *W, const double *f, const double *g,
                               double *Q);
void orc_find_maxwellian(orc_ctx *c, const double *f, double *M, double *rho_u_T /*5*/);

/* src/momentRoutines.c:58-72,116-142,146-165,168-183 (mass = 1, KB = 1) */
double orc_density(const orc_ctx *c, const double *f);
void orc_bulk_velocity(const orc_ctx *c, const double *f, double rho, double *u3);
double orc_temperature(const orc_ctx *c, const double *f, const double *u3, double rho);
void orc_energy(const orc_ctx *c, const double *f, double *pos_neg);

/* src/conserve.c:17-43,89-168,173-202,207-264,268-317 (Ns = 1, mass = 1) */
void orc_conserve(orc_ctx *c, double *Q);
void orc_moment_functionals(const orc_ctx *c, const double *Q, double *b5); /* b = C Q */
void orc_conserve_lu(const orc_ctx *c, double *lu25, int *piv5);

/* src/boundaryConditions.c:39-84 (mass = 1, KB = 1); bdry 0 = left wall, else right */
void orc_diffuse_bc(const orc_ctx *c, const double *in, double *out, double TW, int bdry);

/* single-rank transport on contiguous slabs.
 * order 1: slab has nX+2 cells (ghosts 0 and nX+1), src/transportroutines.c:94-238
 * order 2: slab has nX+4 cells (ghosts 0,1,nX+2,nX+3), src/transportroutines.c:241-470,477-492
 * x, dx have the same cell count as the slab.  ic = Init_field. */
void orc_upwind_one(const orc_ctx *c, int nX, const double *x, const double *dx, double dt,
                    int ic, double *f, double *f_conv);
void orc_upwind_two(const orc_ctx *c, int nX, const double *x, const double *dx, double dt,
                    int ic, double *f, double *f_conv);
void orc_advect_two(const orc_ctx *c, int nX, const double *x, const double *dx, double dt,
                    int ic, double *f, double *f_conv, double *f_tmp);

/* exec/boltz.c:189-249 : one 0D time step, order 1 (Euler) or 2 (RK2), maxPreserve + conserve */
void orc_step_0d(orc_ctx *c, const double *W, double *f, double dt, double Kn, int order);
/* exec/boltz.c:264-353 : one 1D time step; f, f_conv, f_1 are slabs of nX+2*order cells */
void orc_step_1d(orc_ctx *c, const double *W, int nX, const double *x, const double *dx,
                 double dt, double Kn, int order, int ic, double *f, double *f_conv, double *f_1,
                 double *f_tmp);

/* src/initializer.c:91-198 (0D, cases 0..5) and :297-421 (1D, cases 0,1,2,3,5,6) */
void orc_init_hom(const orc_ctx *c, int init_flag, double *f);
void orc_init_inhom(const orc_ctx *c, int init_flag, int nX, int order, double *f_slab);
/* single-rank mesh, src/mesh_setup.c:121-146 with the right ghosts filled as the last-rank
 * branch does (:165-175); zones given as counts/lengths */
void orc_make_mesh(int nzones, const int *zone_n, const double *zone_len, int order, double *x,
                   double *dx);

/* src/output.c:159-213 : row = time-less [rho, u_x, T, p, Eneg/Epos, slice(N)] */
void orc_row_0d(const orc_ctx *c, const double *f, double *row);
/* src/output.c:301-343 : [rho, u_x, T, p] */
void orc_row_1d(const orc_ctx *c, const double *f, double *row4);

/* src/weights.c:33-56,156-160,181-206,261,265-281 (d_ref = 2, equal masses) */
void orc_weights_iso(const orc_ctx *c, double lambda, double *W);
double orc_weight_one(const orc_ctx *c, double lambda, int zeta_flat, int xi_flat);

#ifdef __cplusplus
}
#endif
#endif
