/* oracle/shims/shim.c -- TEST INFRASTRUCTURE ONLY.
 * Implementations behind the shim headers (fftw3.h, mpi.h, gsl/gsl_integration.h) so that the
 * reference's own C sources compile UNMODIFIED from /root/reference (see oracle/build_ref.sh). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "fftw3.h"
#include "mpi.h"
#include "gsl/gsl_integration.h"
#include "../qag21.h"

/* ---------------- FFTW stand-in: exact separable dense DFT ---------------- */
struct orc_shim_plan {
  int n[3];
  int sign;
  fftw_complex *io;
  double *tw; /* cos/sin of 2*pi*m/n per axis length (all axes equal here, kept general) */
};

fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in, fftw_complex *out, int sign,
                           unsigned flags) {
  struct orc_shim_plan *p = malloc(sizeof(*p));
  (void)flags;
  if (in != out) { fprintf(stderr, "fftw shim: only in-place plans supported\n"); abort(); }
  if (n0 > 64 || n1 > 64 || n2 > 64) { fprintf(stderr, "fftw shim: N <= 64 only\n"); abort(); }
  p->n[0] = n0; p->n[1] = n1; p->n[2] = n2; p->sign = sign; p->io = in; p->tw = NULL;
  return p;
}

static void dft_axis(fftw_complex *x, int n, long stride, long nlines, const long *starts, int sign) {
  double *c = malloc(sizeof(double) * n), *s = malloc(sizeof(double) * n);
  int m; long l;
  for (m = 0; m < n; m++) { /* exact-angle table, m in [0,n) */
    c[m] = cos(2.0 * M_PI * (double)m / (double)n);
    s[m] = (double)sign * sin(2.0 * M_PI * (double)m / (double)n);
  }
  /* lines are independent: thread them so the stand-in does not handicap the reference's timing
   * (FFTW itself would spend well under a millisecond here) */
#pragma omp parallel for schedule(static)
  for (l = 0; l < nlines; l++) {
    double tr[64], ti[64];
    fftw_complex *ln = x + starts[l];
    int j, k;
    for (k = 0; k < n; k++) {
      double ar = 0.0, ai = 0.0;
      for (j = 0; j < n; j++) {
        const int mm = (int)(((long)j * k) % n);
        const double xr = ln[j * stride][0], xi = ln[j * stride][1];
        ar += xr * c[mm] - xi * s[mm];
        ai += xr * s[mm] + xi * c[mm];
      }
      tr[k] = ar; ti[k] = ai;
    }
    for (k = 0; k < n; k++) { ln[k * stride][0] = tr[k]; ln[k * stride][1] = ti[k]; }
  }
  free(c); free(s);
}

void fftw_execute(const fftw_plan p) {
  const int n0 = p->n[0], n1 = p->n[1], n2 = p->n[2];
  long *st = malloc(sizeof(long) * (size_t)((long)n0 * n1 + (long)n0 * n2 + (long)n1 * n2));
  long cnt; int a, b;
  /* axis 2 (fastest) */
  cnt = 0; for (a = 0; a < n0; a++) for (b = 0; b < n1; b++) st[cnt++] = ((long)a * n1 + b) * n2;
  dft_axis(p->io, n2, 1, cnt, st, p->sign);
  /* axis 1 */
  cnt = 0; for (a = 0; a < n0; a++) for (b = 0; b < n2; b++) st[cnt++] = (long)a * n1 * n2 + b;
  dft_axis(p->io, n1, n2, cnt, st, p->sign);
  /* axis 0 */
  cnt = 0; for (a = 0; a < n1; a++) for (b = 0; b < n2; b++) st[cnt++] = (long)a * n2 + b;
  dft_axis(p->io, n0, (long)n1 * n2, cnt, st, p->sign);
  free(st);
}
void fftw_destroy_plan(fftw_plan p) { free(p); }
void *fftw_malloc(size_t n) { return malloc(n); }
void fftw_free(void *p) { free(p); }

/* ---------------- MPI stand-in: exactly one rank ---------------- */
int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = 1; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
int MPI_Send(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c) {
  (void)buf; (void)count; (void)t; (void)dest; (void)tag; (void)c;
  fprintf(stderr, "mpi shim: MPI_Send reached with a single rank\n"); abort();
}
int MPI_Recv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *s) {
  (void)buf; (void)count; (void)t; (void)src; (void)tag; (void)c; (void)s;
  fprintf(stderr, "mpi shim: MPI_Recv reached with a single rank\n"); abort();
}
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm c) {
  (void)buf; (void)count; (void)t; (void)root; (void)c; return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
double MPI_Wtime(void) {
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---------------- GSL stand-in ---------------- */
gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n) {
  gsl_integration_workspace *w = malloc(sizeof(*w)); w->limit = n; return w;
}
void gsl_integration_workspace_free(gsl_integration_workspace *w) { free(w); }
int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *w, double *result,
                        double *abserr) {
  (void)w;
  if (key != 2) { fprintf(stderr, "gsl shim: only key=2 (GK21) is provided\n"); abort(); }
  return orc_qag21(f->function, f->params, a, b, epsabs, epsrel, limit, result, abserr);
}

/* exec/boltz.c:154 references the anisotropic loader; every config in scope has Anisotropic 0
 * (SURVEY.md section 2 row 14), so aniso_weights.c (GSL cquad/glfixed/j0) is not compiled. */
void initialize_weights_AnIso(int nodes, double *zeta, double lam, double Lv, int weightFlag,
                              double **conv_weights, double glance) {
  (void)nodes; (void)zeta; (void)lam; (void)Lv; (void)weightFlag; (void)conv_weights; (void)glance;
  fprintf(stderr, "shim: anisotropic weights are out of scope\n"); exit(1);
}
