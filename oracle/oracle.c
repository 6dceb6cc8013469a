/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into or loaded by the product).
 * CPU restatement of the SpectralBTE collision hot path; see oracle/oracle.h for the contract and
 * the pinning status.  All citations are relative to /root/reference.
 */
#include "oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "qag21.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct orc_ctx {
  int N;
  long n3;
  double L_v, L_eta, dv, deta;
  double *v, *eta, *wt; /* grids and trapezoid weights {1/2,1,...,1,1/2} */
  double *dftc, *dfts;  /* cos/sin(2 pi m / N), m in [0,N) */
  double *fh, *gh, *qh, *tmp, *out; /* complex scratch, 2*n3 doubles each */
  double *Mi, *Mj, *gi, *gj, *Qa;   /* real scratch */
  double lu[25];
  int piv[5];
};

/* ------------------------------------------------------------------ set-up */

/* Grids: src/initializer.c:66-82 (0D) and :256-266 (1D); trapezoid weights src/collisions.c:48-54 */
orc_ctx *orc_create(int N, double L_v, int grid_rule) {
  orc_ctx *c = calloc(1, sizeof(*c));
  int i;
  c->N = N;
  c->n3 = (long)N * N * N;
  c->L_v = L_v;
  c->v = malloc(sizeof(double) * N);
  c->eta = malloc(sizeof(double) * N);
  c->wt = malloc(sizeof(double) * N);
  {
    const double dv = 2 * L_v / (N - 1);
    double deta, L_eta;
    for (i = 0; i < N; i++) c->v[i] = -L_v + i * dv;
    if (grid_rule == 0) {
      deta = (2 * M_PI / N) / dv;
      L_eta = ((N % 2) == 0) ? 0.5 * N * deta : 0.5 * (N - 1) * deta;
    } else {
      L_eta = 0.5 * (N - 1) * M_PI / L_v;
      deta = M_PI * (N - 1) / (N * L_v);
    }
    for (i = 0; i < N; i++) c->eta[i] = -L_eta + i * deta;
  }
  /* what initialize_coll derives from the arrays it is handed: src/collisions.c:39-43 */
  c->dv = c->v[1] - c->v[0];
  c->deta = c->eta[1] - c->eta[0];
  c->L_eta = -c->eta[0];
  for (i = 0; i < N; i++) c->wt[i] = 1.0;
  c->wt[0] = 0.5;
  c->wt[N - 1] = 0.5;
  c->dftc = malloc(sizeof(double) * N);
  c->dfts = malloc(sizeof(double) * N);
  for (i = 0; i < N; i++) {
    c->dftc[i] = cos(2.0 * M_PI * i / N);
    c->dfts[i] = sin(2.0 * M_PI * i / N);
  }
  c->fh = malloc(sizeof(double) * 2 * c->n3);
  c->gh = malloc(sizeof(double) * 2 * c->n3);
  c->qh = malloc(sizeof(double) * 2 * c->n3);
  c->tmp = malloc(sizeof(double) * 2 * c->n3);
  c->out = malloc(sizeof(double) * 2 * c->n3);
  c->Mi = malloc(sizeof(double) * c->n3);
  c->Mj = malloc(sizeof(double) * c->n3);
  c->gi = malloc(sizeof(double) * c->n3);
  c->gj = malloc(sizeof(double) * c->n3);
  c->Qa = malloc(sizeof(double) * c->n3);
  orc_conserve_lu(c, c->lu, c->piv);
  return c;
}

void orc_destroy(orc_ctx *c) {
  if (!c) return;
  free(c->v); free(c->eta); free(c->wt); free(c->dftc); free(c->dfts);
  free(c->fh); free(c->gh); free(c->qh); free(c->tmp); free(c->out);
  free(c->Mi); free(c->Mj); free(c->gi); free(c->gj); free(c->Qa);
  free(c);
}

void orc_get_grid(const orc_ctx *c, double *v, double *eta) {
  memcpy(v, c->v, sizeof(double) * c->N);
  memcpy(eta, c->eta, sizeof(double) * c->N);
}
int orc_n(const orc_ctx *c) { return c->N; }

/* ------------------------------------------------------------------ transforms */

/* Unnormalised 3-D DFT, kernel exp(sign * 2 pi i jk/N): the operation FFTW performs at
 * src/collisions.c:270 for the plans made at :67-68. Axis-by-axis dense sums. */
void orc_dft3(orc_ctx *c, double *io, int sign) {
  const int N = c->N;
  const long strides[3] = {1, N, (long)N * N};
  double *lr = malloc(sizeof(double) * N), *li = malloc(sizeof(double) * N);
  int ax;
  for (ax = 0; ax < 3; ax++) {
    const long st = strides[ax];
    const long sa = strides[(ax + 1) % 3], sb = strides[(ax + 2) % 3];
    int a, b, j, k;
    for (a = 0; a < N; a++)
      for (b = 0; b < N; b++) {
        double *base = io + 2 * (a * sa + b * sb);
        for (k = 0; k < N; k++) {
          double sr = 0.0, si = 0.0;
          for (j = 0; j < N; j++) {
            const int m = (j * k) % N;
            const double wr = c->dftc[m], wi = sign * c->dfts[m];
            const double xr = base[2 * j * st], xi = base[2 * j * st + 1];
            sr += xr * wr - xi * wi;
            si += xr * wi + xi * wr;
          }
          lr[k] = sr; li[k] = si;
        }
        for (k = 0; k < N; k++) { base[2 * k * st] = lr[k]; base[2 * k * st + 1] = li[k]; }
      }
  }
  free(lr); free(li);
}

/* src/collisions.c:232-283: pre-twiddle x trapezoid weight x (2 pi)^-3/2 delta^3, DFT, post-twiddle */
void orc_fft3d(orc_ctx *c, const double *in, double *out, int invert) {
  const int N = c->N;
  const double scale3 = pow(1.0 / sqrt(2.0 * M_PI), 3.0); /* src/collisions.c:46 */
  double delta, L_start, L_end, sign;
  const double *arr;
  int i, j, k;
  if (!invert) { delta = c->dv; L_start = c->L_eta; L_end = c->L_v; arr = c->eta; sign = 1.0; }
  else { delta = c->deta; L_start = c->L_v; L_end = c->L_eta; arr = c->v; sign = -1.0; }
  {
    const double prefactor = scale3 * delta * delta * delta;
    for (i = 0; i < N; i++)
      for (j = 0; j < N; j++)
        for (k = 0; k < N; k++) {
          const long idx = k + (long)N * (j + (long)N * i);
          const double sum = sign * (double)(i + j + k) * L_start * delta;
          const double factor = prefactor * c->wt[i] * c->wt[j] * c->wt[k];
          const double cs = cos(sum), sn = sin(sum);
          c->tmp[2 * idx] = factor * (cs * in[2 * idx] - sn * in[2 * idx + 1]);
          c->tmp[2 * idx + 1] = factor * (cs * in[2 * idx + 1] + sn * in[2 * idx]);
        }
  }
  orc_dft3(c, c->tmp, invert ? +1 : -1);
  for (i = 0; i < N; i++)
    for (j = 0; j < N; j++)
      for (k = 0; k < N; k++) {
        const long idx = k + (long)N * (j + (long)N * i);
        const double sum = sign * L_end * (arr[i] + arr[j] + arr[k]);
        const double cs = cos(sum), sn = sin(sum);
        const double tr = c->tmp[2 * idx], ti = c->tmp[2 * idx + 1];
        out[2 * idx] = cs * tr - sn * ti;
        out[2 * idx + 1] = cs * ti + sn * tr;
      }
}

/* ------------------------------------------------------------------ the N^6 convolution */

/* src/collisions.c:127-165.  Q^[zeta] = sum_xi W[zeta][xi] g^[xi] f^[wrap(zeta + N/2 - xi)], the
 * wrap applied once per dimension (:141-158).  The xi-sum runs in flat xi order, sequentially, as
 * in the reference; the per-dimension wrapped index tables replace its div/mod arithmetic. */
static long g_row_mod = 0; /* > 0: weight row zeta is read from row (zeta % g_row_mod) (bounded CPU baselines) */

void orc_qhat(orc_ctx *c, const double *W, const double *fhat, const double *ghat, double *qhat) {
  const int N = c->N, n2 = N / 2;
  const long n3 = c->n3;
  long zeta;
#pragma omp parallel for schedule(static)
  for (zeta = 0; zeta < n3; zeta++) {
    const int zx = (int)(zeta / ((long)N * N));
    const int zy = (int)((zeta - (long)zx * N * N) / N);
    const int zz = (int)(zeta - (long)N * (zy + (long)zx * N));
    const double *w = W + (g_row_mod > 0 ? zeta % g_row_mod : zeta) * n3;
    double ar = 0.0, ai = 0.0;
    int ex, ey, ez;
    long xi = 0;
    for (ex = 0; ex < N; ex++) {
      int x = zx + n2 - ex;
      if (x < 0) x += N; else if (x > N - 1) x -= N;
      for (ey = 0; ey < N; ey++) {
        int y = zy + n2 - ey;
        if (y < 0) y += N; else if (y > N - 1) y -= N;
        {
          const double *fl = fhat + 2 * ((long)N * (y + (long)N * x));
          for (ez = 0; ez < N; ez++, xi++) {
            int z = zz + n2 - ez;
            if (z < 0) z += N; else if (z > N - 1) z -= N;
            {
              const double gr = ghat[2 * xi], gim = ghat[2 * xi + 1];
              const double fr = fl[2 * z], fim = fl[2 * z + 1];
              ar += w[xi] * (gr * fr - gim * fim);
              ai += w[xi] * (gr * fim + gim * fr);
            }
          }
        }
      }
    }
    qhat[2 * zeta] = ar;
    qhat[2 * zeta + 1] = ai;
  }
}

/* compute_Qhat, src/collisions.c:108-169: pack, two forward transforms, convolution, inverse */
static void qhat_pipeline(orc_ctx *c, const double *W, const double *f_mat, const double *g_mat) {
  long i;
  for (i = 0; i < c->n3; i++) {
    c->out[2 * i] = f_mat[i]; c->out[2 * i + 1] = 0.0;
  }
  orc_fft3d(c, c->out, c->fh, 0);
  for (i = 0; i < c->n3; i++) {
    c->out[2 * i] = g_mat[i]; c->out[2 * i + 1] = 0.0;
  }
  orc_fft3d(c, c->out, c->gh, 0);
  orc_qhat(c, W, c->fh, c->gh, c->qh);
  orc_fft3d(c, c->qh, c->out, 1);
}

/* ComputeQ with the N^3 weight rows aliased onto `distinct` stored rows: same loop, same memory
 * stream per row, bounded host memory (bench.py cpu_baseline when oracle/_ref is unavailable) */
void orc_compute_q_rowmod(orc_ctx *c, const double *W, long distinct, const double *f, const double *g, double *Q) {
  g_row_mod = distinct;
  orc_compute_q(c, W, f, g, Q, NULL);
  g_row_mod = 0;
}

/* src/collisions.c:212-221 */
void orc_compute_q(orc_ctx *c, const double *W, const double *f, const double *g, double *Q,
                   double *qhat_out) {
  long i;
  qhat_pipeline(c, W, f, g);
  if (qhat_out) memcpy(qhat_out, c->qh, sizeof(double) * 2 * c->n3);
  for (i = 0; i < c->n3; i++) Q[i] = c->out[2 * i];
}

/* ------------------------------------------------------------------ moments */

/* src/momentRoutines.c:58-72 */
double orc_density(const orc_ctx *c, const double *f) {
  const int N = c->N;
  const double dv3 = c->dv * c->dv * c->dv;
  double r = 0.0;
  int i, j, k;
  for (i = 0; i < N; i++)
    for (j = 0; j < N; j++)
      for (k = 0; k < N; k++) r += dv3 * c->wt[i] * c->wt[j] * c->wt[k] * f[k + N * (j + N * i)];
  return 1.0 * r;
}

/* src/momentRoutines.c:116-142 */
void orc_bulk_velocity(const orc_ctx *c, const double *f, double rho, double *u) {
  const int N = c->N;
  const double dv3 = c->dv * c->dv * c->dv;
  int i, j, k;
  u[0] = u[1] = u[2] = 0.0;
  for (i = 0; i < N; i++)
    for (j = 0; j < N; j++)
      for (k = 0; k < N; k++) {
        const double t1 = c->v[i] * dv3 * c->wt[i] * c->wt[j] * c->wt[k] / rho;
        const double t2 = c->v[j] * dv3 * c->wt[i] * c->wt[j] * c->wt[k] / rho;
        const double t3 = c->v[k] * dv3 * c->wt[i] * c->wt[j] * c->wt[k] / rho;
        const double fv = f[k + N * (j + N * i)];
        u[0] += t1 * fv; u[1] += t2 * fv; u[2] += t3 * fv;
      }
  u[0] = 1.0 * u[0]; u[1] = 1.0 * u[1]; u[2] = 1.0 * u[2];
}

/* src/momentRoutines.c:168-183 */
double orc_temperature(const orc_ctx *c, const double *f, const double *u, double rho) {
  const int N = c->N;
  const double dv3 = c->dv * c->dv * c->dv;
  double r = 0.0;
  int i, j, k;
  for (i = 0; i < N; i++)
    for (j = 0; j < N; j++)
      for (k = 0; k < N; k++) {
        const double t = (c->v[i] - u[0]) * (c->v[i] - u[0]) + (c->v[j] - u[1]) * (c->v[j] - u[1]) +
                         (c->v[k] - u[2]) * (c->v[k] - u[2]);
        r += t * dv3 * c->wt[i] * c->wt[j] * c->wt[k] * f[k + N * (j + N * i)] / (3.0 * rho);
      }
  return (1.0 * 1.0 / 1.0) * r;
}

/* src/momentRoutines.c:146-165 */
void orc_energy(const orc_ctx *c, const double *f, double *pn) {
  const int N = c->N;
  const double dv3 = c->dv * c->dv * c->dv;
  double pos = 0.0, neg = 0.0;
  int i, j, k;
  for (i = 0; i < N; i++)
    for (j = 0; j < N; j++)
      for (k = 0; k < N; k++) {
        const double e = dv3 * c->wt[i] * c->wt[j] * c->wt[k] * f[k + N * (j + N * i)] *
                         (c->v[i] * c->v[i] + c->v[j] * c->v[j] + c->v[k] * c->v[k]);
        if (e > 0) pos += e; else neg -= e;
      }
  pn[0] = pos; pn[1] = neg;
}

/* ------------------------------------------------------------------ Maxwellian split */

/* src/collisions.c:91-106 (first half: the Maxwellian with the moments of f) */
void orc_find_maxwellian(orc_ctx *c, const double *f, double *M, double *ruT) {
  const int N = c->N;
  double u[3];
  const double rho = orc_density(c, f);
  double T, pre;
  int i, j, k;
  orc_bulk_velocity(c, f, rho, u);
  T = orc_temperature(c, f, u, rho);
  pre = rho * pow(0.5 / (M_PI * T), 1.5);
  for (i = 0; i < N; i++)
    for (j = 0; j < N; j++)
      for (k = 0; k < N; k++)
        M[k + N * (j + N * i)] =
            pre * exp(-(0.5 / T) * ((c->v[i] - u[0]) * (c->v[i] - u[0]) + (c->v[j] - u[1]) * (c->v[j] - u[1]) +
                                    (c->v[k] - u[2]) * (c->v[k] - u[2])));
  if (ruT) { ruT[0] = rho; ruT[1] = u[0]; ruT[2] = u[1]; ruT[3] = u[2]; ruT[4] = T; }
}

/* src/collisions.c:178-210.  Quirk kept: both perturbations subtract M_i (:104). */
void orc_compute_q_maxpreserve(orc_ctx *c, const double *W, const double *f, const double *g,
                               double *Q) {
  long i;
  orc_find_maxwellian(c, f, c->Mi, NULL);
  for (i = 0; i < c->n3; i++) c->gi[i] = f[i] - c->Mi[i];
  orc_find_maxwellian(c, g, c->Mj, NULL);
  for (i = 0; i < c->n3; i++) c->gj[i] = g[i] - c->Mi[i];
  qhat_pipeline(c, W, c->Mi, c->gj);
  for (i = 0; i < c->n3; i++) Q[i] = c->out[2 * i];
  qhat_pipeline(c, W, c->gi, c->Mj);
  for (i = 0; i < c->n3; i++) Q[i] += c->out[2 * i];
  qhat_pipeline(c, W, c->gi, c->gj);
  for (i = 0; i < c->n3; i++) Q[i] += c->out[2 * i];
}

/* ------------------------------------------------------------------ conservation */

static void cons_row(const orc_ctx *c, int i, int j, int k, double pre, double *t) {
  t[0] = pre;
  t[1] = pre * c->v[i];
  t[2] = pre * c->v[j];
  t[3] = pre * c->v[k];
  t[4] = pre * 0.5 * (c->v[i] * c->v[i] + c->v[j] * c->v[j] + c->v[k] * c->v[k]);
}

/* src/conserve.c:268-317 (Gram matrix of the moment functionals) + :89-168 (scaled-pivot LU; the
 * pivot row is the FIRST row that improves on the diagonal, :115-126; swaps touch columns >= k) */
void orc_conserve_lu(const orc_ctx *c, double *A, int *piv) {
  const int N = c->N, n = 5;
  double s[5], t[5];
  int i, j, k, a, b;
  for (a = 0; a < 25; a++) A[a] = 0.0;
  for (a = 0; a < n; a++)
    for (b = 0; b < n; b++) {
      double acc = 0.0;
      for (i = 0; i < N; i++)
        for (j = 0; j < N; j++)
          for (k = 0; k < N; k++) {
            cons_row(c, i, j, k, c->wt[i] * c->wt[j] * c->wt[k] * c->dv * c->dv * c->dv * 1.0, t);
            acc += t[a] * t[b];
          }
      A[a * n + b] = acc;
    }
  for (i = 0; i < n; i++) {
    s[i] = fabs(A[i * n]);
    for (j = 0; j < n; j++)
      if (s[i] < fabs(A[i * n + j])) s[i] = fabs(A[i * n + j]);
  }
  for (k = 0; k < n - 1; k++) {
    double ck = fabs(A[k * n + k] / s[k]);
    int i0 = k, found = 0;
    for (i = k; i < n; i++)
      if (ck < fabs(A[i * n + k] / s[i])) {
        ck = fabs(A[i * n + k] / s[i]);
        if (!found) { i0 = i; found = 1; }
      }
    piv[k] = i0;
    if (ck == 0.0) { fprintf(stderr, "orc_conserve_lu: singular\n"); exit(1); }
    if (i0 != k) {
      double tmp;
      for (j = k; j < n; j++) { tmp = A[k * n + j]; A[k * n + j] = A[i0 * n + j]; A[i0 * n + j] = tmp; }
      tmp = s[k]; s[k] = s[i0]; s[i0] = tmp;
    }
    for (i = k + 1; i < n; i++) {
      const double m = A[i * n + k] / A[k * n + k];
      A[i * n + k] = m;
      for (j = k + 1; j < n; j++) A[i * n + j] -= m * A[k * n + j];
    }
  }
  piv[n - 1] = n - 1;
}

/* b = sum C Q: src/conserve.c:217-240 */
void orc_moment_functionals(const orc_ctx *c, const double *Q, double *b) {
  const int N = c->N;
  double t[5];
  int i, j, k, a;
  for (a = 0; a < 5; a++) b[a] = 0.0;
  for (i = 0; i < N; i++)
    for (j = 0; j < N; j++)
      for (k = 0; k < N; k++) {
        const double q = Q[k + N * (j + N * i)];
        cons_row(c, i, j, k, c->wt[i] * c->wt[j] * c->wt[k] * c->dv * c->dv * c->dv * 1.0, t);
        for (a = 0; a < 5; a++) b[a] += q * t[a];
      }
}

/* src/conserve.c:207-264 with the solve of :173-202 */
void orc_conserve(orc_ctx *c, double *Q) {
  const int N = c->N, n = 5;
  const double *A = c->lu;
  double b[5], t[5];
  int i, j, k;
  orc_moment_functionals(c, Q, b);
  for (k = 0; k < n - 1; k++) {
    i = c->piv[k];
    if (i != k) { const double tmp = b[i]; b[i] = b[k]; b[k] = tmp; }
    for (i = k + 1; i < n; i++) b[i] -= A[i * n + k] * b[k];
  }
  b[n - 1] = b[n - 1] / A[(n - 1) * n + (n - 1)];
  for (i = n - 2; i >= 0; i--) {
    double sum = 0.0;
    for (j = i + 1; j < n; j++) sum += A[i * n + j] * b[j];
    b[i] = 1.0 / A[i * n + i] * (b[i] - sum);
  }
  for (i = 0; i < N; i++)
    for (j = 0; j < N; j++)
      for (k = 0; k < N; k++) {
        cons_row(c, i, j, k, c->wt[i] * c->wt[j] * c->wt[k] * c->dv * c->dv * c->dv * 1.0 / 1, t);
        Q[k + N * (j + N * i)] -= (t[0] * b[0] + t[1] * b[1] + t[2] * b[2] + t[3] * b[3] + t[4] * b[4]);
      }
}

/* ------------------------------------------------------------------ walls and transport */

/* src/boundaryConditions.c:39-84 */
void orc_diffuse_bc(const orc_ctx *c, const double *in, double *out, double TW, int bdry) {
  const int N = c->N;
  const double h = c->dv;
  double sig = 0.0;
  int i, j, k;
  const int o0 = (bdry == 0) ? 0 : N / 2, o1 = (bdry == 0) ? N / 2 : N; /* outgoing half */
  const int n0 = (bdry == 0) ? N / 2 : 0, n1 = (bdry == 0) ? N : N / 2; /* incoming half */
  for (i = o0; i < o1; i++)
    for (j = 0; j < N; j++)
      for (k = 0; k < N; k++)
        sig += c->v[i] * c->wt[i] * c->wt[j] * c->wt[k] * h * h * h * in[k + N * (j + N * i)];
  if (bdry == 0) sig *= -sqrt(2.0 * M_PI * 1.0 / (1.0 * TW));
  else sig *= sqrt(2.0 * M_PI * 1.0 / (1.0 * TW));
  for (i = n0; i < n1; i++)
    for (j = 0; j < N; j++)
      for (k = 0; k < N; k++)
        out[k + N * (j + N * i)] =
            sig * pow(0.5 * 1.0 / (M_PI * 1.0 * TW), 1.5) *
            exp(-0.5 * 1.0 / (1.0 * TW) * (c->v[i] * c->v[i] + c->v[j] * c->v[j] + c->v[k] * c->v[k]));
}

/* src/transportroutines.c:80-90 */
static double mm3(double a, double b, double d) {
  if (a > 0 && b > 0 && d > 0) { double m = a < b ? a : b; return m < d ? m : d; }
  if (a < 0 && b < 0 && d < 0) { double m = a > b ? a : b; return m > d ? m : d; }
  return 0;
}

#define T0_WALL 1.0 /* src/transportroutines.c:46-47 */
#define T1_WALL 2.0
#define TWALL_IN 1.0 /* src/initializer.c:275 */

/* src/transportroutines.c:94-238, one rank (rank 0 == last rank). */
void orc_upwind_one(const orc_ctx *c, int nX, const double *x, const double *dx, double dt, int ic,
                    double *f, double *fc) {
  const int N = c->N;
  const long n3 = c->n3;
  int i, l;
  long jk;
  (void)x;
  /* Init_field 5 (Poiseuille): the forcing block of :217-225 sits AFTER the k loop has closed, so with k == N
   * every write lands on element (i, j+1, 0) -- which the next j (or i) iteration overwrites with its plain
   * upwind value -- and the very last one falls one double past the cell (into the next cell's allocation in
   * this process image, never read back).  The observable result at order 1 is therefore the unforced scheme
   * with the diffuse walls of :117,134; that is what is restated (checked against oracle/_ref in
   * tests/golden/make_golden.py, vectors tr_ic5_o1_*). */
  /* ghost cells, :107-172 */
  if (ic == 3 || ic == 5) orc_diffuse_bc(c, f + 1 * n3, f + 0 * n3, T0_WALL, 0);
  else if (ic == 1) orc_diffuse_bc(c, f + 1 * n3, f + 0 * n3, 2.0 * TWALL_IN, 0);
  else if (ic != 6) memcpy(f, f + n3, sizeof(double) * n3);
  if (ic == 3 || ic == 5) orc_diffuse_bc(c, f + (long)nX * n3, f + (long)(nX + 1) * n3, T1_WALL, 1);
  else if (ic != 6) memcpy(f + (long)(nX + 1) * n3, f + (long)nX * n3, sizeof(double) * n3);
  if (ic == 6) {
    memcpy(f, f + (long)nX * n3, sizeof(double) * n3);
    memcpy(f + (long)(nX + 1) * n3, f + n3, sizeof(double) * n3);
  }
  /* stencil, :203-216 */
  for (l = 1; l < nX + 1; l++)
    for (i = 0; i < N; i++) {
      const double cfl = dt * c->v[i] / dx[l];
      const double *fl = f + l * n3 + (long)i * N * N;
      double *o = fc + l * n3 + (long)i * N * N;
      if (i < N / 2) {
        const double *fr = fl + n3;
        for (jk = 0; jk < (long)N * N; jk++) o[jk] = (1.0 + cfl) * fl[jk] - cfl * fr[jk];
      } else {
        const double *fm = fl - n3;
        for (jk = 0; jk < (long)N * N; jk++) o[jk] = (1.0 - cfl) * fl[jk] + cfl * fm[jk];
      }
    }
}

/* src/transportroutines.c:241-470, one rank. Ghosts: f[1], f[nX+2] by linear extrapolation
 * (:271-275, :293-297); f[0], f[nX+3] are never read on a single rank (the wall branches replace
 * the only stencils that would reach them). */
void orc_upwind_two(const orc_ctx *c, int nX, const double *x, const double *dx, double dt, int ic,
                    double *f, double *fc) {
  const int N = c->N, h = N / 2;
  const long n3 = c->n3, nn = (long)N * N;
  double *fl = malloc(sizeof(double) * n3), *fr = malloc(sizeof(double) * n3);
  long p;
  int i, l;
  /* Poiseuille forcing, :428-436,457-465: Ma = 1 (:255), h_v = 2 L_v / (N-1) (:37) */
  const double h_v = 2 * c->L_v / (N - 1), Ma = 1.0;
#define CELL(m) (f + (long)(m) * n3)
  for (p = 0; p < n3; p++) CELL(1)[p] = 2 * CELL(2)[p] - CELL(3)[p];
  for (p = 0; p < n3; p++) CELL(nX + 2)[p] = 2 * CELL(nX + 1)[p] - CELL(nX)[p];
  /* left wall face, :351-378 */
  for (p = 0; p < h * nn; p++) {
    const double s1 = mm3((CELL(2)[p] - CELL(1)[p]) / (x[2] - x[1]), (CELL(3)[p] - CELL(2)[p]) / (x[3] - x[2]),
                          (CELL(3)[p] - CELL(1)[p]) / (x[3] - x[1]));
    fl[p] = CELL(2)[p] - 0.5 * dx[2] * s1;
  }
  if (ic == 3 || ic == 5) orc_diffuse_bc(c, fl, fl, T0_WALL, 0);
  else if (ic == 1) orc_diffuse_bc(c, fl, fl, 2.0 * TWALL_IN, 0);
  else
    for (p = h * nn; p < n3; p++) {
      const double s1 = mm3((CELL(2)[p] - CELL(1)[p]) / (x[2] - x[1]), (CELL(3)[p] - CELL(2)[p]) / (x[3] - x[2]),
                            (CELL(3)[p] - CELL(1)[p]) / (x[3] - x[1]));
      fl[p] = CELL(2)[p] + 0.5 * dx[2] * s1;
    }
  /* right wall face, :380-404 */
  for (p = h * nn; p < n3; p++) {
    const double s1 = mm3((CELL(nX + 1)[p] - CELL(nX)[p]) / (x[nX + 1] - x[nX]),
                          (CELL(nX + 2)[p] - CELL(nX + 1)[p]) / (x[nX + 2] - x[nX + 1]),
                          (CELL(nX + 2)[p] - CELL(nX)[p]) / (x[nX + 2] - x[nX]));
    fr[p] = CELL(nX + 1)[p] + 0.5 * dx[nX + 1] * s1;
  }
  if (ic == 3 || ic == 5) orc_diffuse_bc(c, fr, fr, T1_WALL, 1);
  else
    for (p = 0; p < h * nn; p++) {
      const double s1 = mm3((CELL(nX + 1)[p] - CELL(nX)[p]) / (x[nX + 1] - x[nX]),
                            (CELL(nX + 2)[p] - CELL(nX + 1)[p]) / (x[nX + 2] - x[nX + 1]),
                            (CELL(nX + 2)[p] - CELL(nX)[p]) / (x[nX + 2] - x[nX]));
      fr[p] = CELL(nX + 1)[p] - 0.5 * dx[nX + 1] * s1;
    }
  /* MUSCL stencil, :406-468 */
  for (l = 2; l < nX + 2; l++)
    for (i = 0; i < N; i++) {
      const double cfl = 0.5 * dt * c->v[i] / dx[l];
      for (p = (long)i * nn; p < (long)(i + 1) * nn; p++) {
        const double s1 = mm3((CELL(l)[p] - CELL(l - 1)[p]) / (x[l] - x[l - 1]),
                              (CELL(l + 1)[p] - CELL(l)[p]) / (x[l + 1] - x[l]),
                              (CELL(l + 1)[p] - CELL(l - 1)[p]) / (x[l + 1] - x[l - 1]));
        double r;
        if (i >= h) {
          if (l == 2) r = CELL(l)[p] - cfl * (CELL(l)[p] + 0.5 * dx[l] * s1 - fl[p]);
          else {
            const double s0 = mm3((CELL(l - 1)[p] - CELL(l - 2)[p]) / (x[l - 1] - x[l - 2]),
                                  (CELL(l)[p] - CELL(l - 1)[p]) / (x[l] - x[l - 1]),
                                  (CELL(l)[p] - CELL(l - 2)[p]) / (x[l] - x[l - 2]));
            r = CELL(l)[p] - cfl * (CELL(l)[p] + 0.5 * dx[l] * s1 - (CELL(l - 1)[p] + 0.5 * dx[l - 1] * s0));
          }
        } else {
          if (l == nX + 1) r = CELL(l)[p] - cfl * (fr[p] - (CELL(l)[p] - 0.5 * dx[l] * s1));
          else {
            const double s2 = mm3((CELL(l + 1)[p] - CELL(l)[p]) / (x[l + 1] - x[l]),
                                  (CELL(l + 2)[p] - CELL(l + 1)[p]) / (x[l + 2] - x[l + 1]),
                                  (CELL(l + 2)[p] - CELL(l)[p]) / (x[l + 2] - x[l]));
            r = CELL(l)[p] - cfl * (CELL(l + 1)[p] - 0.5 * dx[l + 1] * s2 - (CELL(l)[p] - 0.5 * dx[l] * s1));
          }
        }
        if (ic == 5) { /* central difference in v_y of the pass input, one-sided (sign as written) at the ends */
          const int j = (int)((p / N) % N);
          if (j == 0) r = r - Ma * 0.5 * dt / (2 * h_v) * CELL(l)[p + N];
          else if (j == N - 1) r = r - Ma * 0.5 * dt / (2 * h_v) * CELL(l)[p - N];
          else r = r - Ma * 0.5 * dt / (2 * h_v) * (CELL(l)[p + N] - CELL(l)[p - N]);
        }
        fc[(long)l * n3 + p] = r;
      }
    }
#undef CELL
  free(fl); free(fr);
}

/* src/transportroutines.c:477-492 */
void orc_advect_two(const orc_ctx *c, int nX, const double *x, const double *dx, double dt, int ic,
                    double *f, double *fc, double *ft) {
  const long n3 = c->n3;
  long p;
  int l;
  orc_upwind_two(c, nX, x, dx, dt, ic, f, ft);
  orc_upwind_two(c, nX, x, dx, dt, ic, ft, fc);
  for (l = 2; l < nX + 2; l++)
    for (p = 0; p < n3; p++) fc[l * n3 + p] = 0.5 * (f[l * n3 + p] + fc[l * n3 + p]);
}

/* ------------------------------------------------------------------ time steps */

/* exec/boltz.c:189-249, one species */
void orc_step_0d(orc_ctx *c, const double *W, double *f, double dt, double Kn, int order) {
  const long n3 = c->n3;
  double *Q = c->Qa;
  long i;
  orc_compute_q_maxpreserve(c, W, f, f, Q);
  orc_conserve(c, Q);
  if (order == 1) {
    for (i = 0; i < n3; i++) f[i] += dt * Q[i] / Kn;
  } else {
    double *f1 = malloc(sizeof(double) * n3);
    for (i = 0; i < n3; i++) { f1[i] = f[i]; f1[i] += dt * Q[i] / Kn; }
    orc_compute_q_maxpreserve(c, W, f1, f1, Q);
    orc_conserve(c, Q);
    for (i = 0; i < n3; i++) { f[i] = 0.5 * (f[i] + f1[i]); f[i] += 0.5 * dt * Q[i] / Kn; }
    free(f1);
  }
}

/* exec/boltz.c:264-353, one species, one rank */
void orc_step_1d(orc_ctx *c, const double *W, int nX, const double *x, const double *dx, double dt,
                 double Kn, int order, int ic, double *f, double *fc, double *f1, double *ft) {
  const long n3 = c->n3;
  double *Q = c->Qa;
  long p;
  int l;
  if (order == 1) orc_upwind_one(c, nX, x, dx, dt, ic, f, fc);
  else orc_advect_two(c, nX, x, dx, dt, ic, f, fc, ft);
  for (l = order; l < nX + order; l++) {
    double *cf = f + l * n3, *cc = fc + l * n3, *c1 = f1 + l * n3;
    orc_compute_q(c, W, cc, cc, Q, NULL);
    orc_conserve(c, Q);
    if (order == 1) {
      for (p = 0; p < n3; p++) { cf[p] = cc[p]; cf[p] += dt * Q[p] / Kn; }
    } else {
      for (p = 0; p < n3; p++) { c1[p] = cc[p]; c1[p] += dt * Q[p] / Kn; }
      orc_compute_q(c, W, c1, c1, Q, NULL);
      orc_conserve(c, Q);
      for (p = 0; p < n3; p++) { cc[p] = 0.5 * cc[p] + 0.5 * c1[p]; cc[p] += 0.5 * dt * Q[p] / Kn; }
    }
  }
  if (order == 2) orc_advect_two(c, nX, x, dx, dt, ic, fc, f, ft);
}

/* ------------------------------------------------------------------ initial data, mesh, output */

/* src/initializer.c:91-198 */
void orc_init_hom(const orc_ctx *c, int flag, double *f) {
  const int N = c->N;
  const double *v = c->v;
  int i, j, k;
  for (i = 0; i < N; i++)
    for (j = 0; j < N; j++)
      for (k = 0; k < N; k++) {
        const double r2 = v[i] * v[i] + v[j] * v[j] + v[k] * v[k];
        double val = 0.0;
        switch (flag) {
          case 0: {
            const double sigma = 0.3 * c->L_v, S = 10.0;
            val = exp(-1 * S * (sqrt(r2) - sigma) * (sqrt(r2) - sigma) / (sigma * sigma)) / (S * S);
            break;
          }
          case 1:
            val = exp(-r2) / (M_PI * sqrt(M_PI));
            if (i >= N / 2) val = exp(-r2) / (M_PI * sqrt(M_PI)) / 2;
            break;
          case 2: {
            const double K = 1 - exp(-5.5 / 6.0), Temp = 1.0;
            val = (exp(-r2 / (2 * K * Temp * Temp))) / (2.0 * pow(2 * M_PI * K * Temp * Temp, 1.5)) *
                  ((5 * K - 3) / K + (1 - K) * r2 / (K * K * Temp * Temp));
            break;
          }
          case 3: {
            const double sigma = M_PI * c->L_v / 10.0;
            const double pre = 0.5 / pow(2.0 * M_PI * sigma * sigma, 1.5);
            val = pre * (exp(-((v[i] - 2.0 * sigma) * (v[i] - 2.0 * sigma) + v[j] * v[j] + v[k] * v[k]) / (2.0 * sigma * sigma)) +
                         exp(-((v[i] + 2 * sigma) * (v[i] + 2 * sigma) + v[j] * v[j] + v[k] * v[k]) / (2.0 * sigma * sigma)));
            break;
          }
          case 4:
            val = (1.0 / 1.0) * pow(0.5 * 1.0 / (M_PI * 1.0 * 1.0), 1.5) * exp(-(0.5 * 1.0 / (1.0 * 1.0)) * r2);
            break;
          case 5:
            val = (1 + 0.1 * sin(r2)) * exp(-r2) / (M_PI * sqrt(M_PI));
            break;
          default:
            fprintf(stderr, "orc_init_hom: unknown Init_field %d\n", flag); exit(1);
        }
        f[k + N * (j + N * i)] = val;
      }
}

/* src/initializer.c:297-421; ghost cells are left untouched (zero-filled by the caller) */
void orc_init_inhom(const orc_ctx *c, int flag, int nX, int order, double *fs) {
  const int N = c->N;
  const double *v = c->v;
  const double Ma = 1;
  double rho_l = 1.0, ux_l = 0.0, T_l = 1.0, rho_r = 1.0, ux_r = 0.0, T_r = 1.0, maxTemp;
  int i, j, k, l;
  switch (flag) {
    case 0:
      rho_l = 4.0 * Ma * Ma / (Ma * Ma + 3.0);
      T_l = (5.0 * Ma * Ma - 1.0) * (Ma * Ma + 3.0) / (16.0 * Ma * Ma);
      ux_l = -sqrt(5.0 / 3.0) * (Ma * Ma + 3.0) / (4.0 * Ma);
      rho_r = 1.0; T_r = 1.0; ux_r = -Ma * sqrt(5.0 / 3.0);
      break;
    case 1: rho_r = 1.0; T_r = 1.0; break;
    case 2: rho_l = 1.0; T_r = 2.0; T_l = 1.0; ux_l = -1.0; ux_r = -1.0; break;
    case 3: rho_r = 1.0; T_r = 1.5; break;
    case 5: rho_l = 1.0; T_l = 1.0; break;
    case 6: rho_l = 1.0; ux_l = 1.2972; T_l = 1.0; rho_r = 1.297; ux_r = 1.0; T_r = 1.195; break;
    default: fprintf(stderr, "orc_init_inhom: unknown Init_field %d\n", flag); exit(1);
  }
  (void)ux_r;
  maxTemp = T_r;
  for (l = order; l < nX + order; l++)
    for (i = 0; i < N; i++)
      for (j = 0; j < N; j++)
        for (k = 0; k < N; k++) {
          const double r2 = v[i] * v[i] + v[j] * v[j] + v[k] * v[k];
          double val = 0.0;
          switch (flag) {
            case 0:
              if (l < nX / 2) val = rho_l * exp(-r2 / T_l) / ((T_l * M_PI) * sqrt(T_l * M_PI));
              else val = rho_r * exp(-r2 / T_r) / ((T_r * M_PI) * sqrt(T_r * M_PI));
              break;
            case 1:
              val = (rho_r / 1.0) * pow(0.5 * 1.0 / (M_PI * 1.0 * maxTemp), 1.5) * exp(-(0.5 * 1.0 / (1.0 * maxTemp)) * r2);
              break;
            case 2:
              val = rho_l * exp(-((v[i] - ux_l) * (v[i] - ux_l) + v[j] * v[j] + v[k] * v[k]) / T_l) / ((T_l * M_PI) * sqrt(T_l * M_PI));
              break;
            case 3:
              val = rho_r * exp(-(v[i] * v[i] + v[j] * v[j] + v[k] * v[k]) / T_r) / ((T_r * M_PI) * sqrt(T_r * M_PI));
              break;
            case 5:
              val = rho_l * exp(-(v[i] * v[i] + v[j] * v[j] + v[k] * v[k]) / T_l) / ((T_l * M_PI) * sqrt(T_l * M_PI));
              break;
            case 6:
              if (l < nX / 2)
                val = rho_l * exp(-((v[i] - ux_l) * (v[i] - ux_l) + v[j] * v[j] + v[k] * v[k]) / T_l) / ((T_l * M_PI) * sqrt(T_l * M_PI));
              else
                val = rho_r * exp(-((v[i] - ux_r) * (v[i] - ux_r) + v[j] * v[j] + v[k] * v[k]) / T_r) / ((T_r * M_PI) * sqrt(T_r * M_PI));
              break;
          }
          fs[(long)l * c->n3 + k + N * (j + N * i)] = val;
        }
}

/* src/mesh_setup.c:65-71,121-146 for rank 0; right ghosts as the last-rank branch :165-175 */
void orc_make_mesh(int nzones, const int *zn, const double *zl, int order, double *x, double *dx) {
  double edge = 0.0;
  int z, j, cnt = order;
  for (z = 0; z < nzones; z++) {
    const double d = zl[z] / (double)zn[z];
    for (j = 0; j < zn[z]; j++) { dx[cnt] = d; x[cnt] = edge + 0.5 * d; edge += d; cnt++; }
  }
  if (order == 1) { dx[0] = dx[1]; x[0] = x[1] - dx[1]; }
  else { dx[1] = dx[2]; x[1] = x[2] - dx[2]; dx[0] = dx[1]; x[0] = x[1] - dx[1]; }
  dx[cnt] = dx[cnt - 1]; x[cnt] = x[cnt - 1] + dx[cnt - 1];
  if (order == 2) { cnt++; dx[cnt] = dx[cnt - 1]; x[cnt] = x[cnt - 1] + dx[cnt - 1]; }
}

/* src/output.c:159-209 */
void orc_row_0d(const orc_ctx *c, const double *f, double *row) {
  const int N = c->N;
  double u[3], e[2];
  const double rho = orc_density(c, f);
  double T;
  int l;
  orc_bulk_velocity(c, f, rho, u);
  T = orc_temperature(c, f, u, rho);
  orc_energy(c, f, e);
  row[0] = rho; row[1] = u[0]; row[2] = T; row[3] = rho * T; row[4] = e[1] / e[0];
  for (l = 0; l < N; l++) row[5 + l] = f[N / 2 + N * (N / 2 + N * l)];
}

/* src/output.c:301-343 */
void orc_row_1d(const orc_ctx *c, const double *f, double *row) {
  double u[3];
  const double rho = orc_density(c, f);
  double T;
  orc_bulk_velocity(c, f, rho, u);
  T = orc_temperature(c, f, u, rho);
  row[0] = rho; row[1] = u[0]; row[2] = T; row[3] = rho * T;
}

/* ------------------------------------------------------------------ isotropic weights */

typedef struct { double a0, a1, a2, lam; } gh_args;

static double sinc_(double x) { return (x != 0.0) ? sin(x) / x : 1.0; } /* src/weights.c:136-143 */

/* src/weights.c:156-160 */
static double ghat_r(double r, void *p) {
  const gh_args *a = (const gh_args *)p;
  return pow(r, a->lam + 2) * (sinc_(r * a->a0) * sinc_(r * a->a2) - sinc_(r * a->a1));
}

/* src/weights.c:181-206,261 (mu = 1/2) and the scaling at :277 with d_ref = 2 (src/species.c:28-43) */
double orc_weight_one(const orc_ctx *c, double lambda, int zf, int xf) {
  const int N = c->N;
  const int i = zf / (N * N), j = (zf / N) % N, k = zf % N;
  const int l = xf / (N * N), m = (xf / N) % N, n = xf % N;
  const double *e = c->eta;
  const double mu = 1.0 / (1.0 + 1.0);
  const double prefactor = 16.0 * M_PI * M_PI * c->deta * c->deta * c->deta / pow(2.0 * M_PI, 1.5) / (4.0 * M_PI);
  gh_args a;
  double res = 0.0, err;
  a.lam = lambda;
  a.a0 = mu * sqrt(e[i] * e[i] + e[j] * e[j] + e[k] * e[k]);
  a.a1 = sqrt(e[l] * e[l] + e[m] * e[m] + e[n] * e[n]);
  a.a2 = sqrt((e[l] - mu * e[i]) * (e[l] - mu * e[i]) + (e[m] - mu * e[j]) * (e[m] - mu * e[j]) +
              (e[n] - mu * e[k]) * (e[n] - mu * e[k]));
  orc_qag21(ghat_r, &a, 0.0, c->L_v, 1e-8, 1e-8, 10000, &res, &err);
  return c->wt[l] * c->wt[m] * c->wt[n] * 0.25 * pow(0.5 * (2.0 + 2.0), 2) * (prefactor * res);
}

/* src/weights.c:265-281 */
void orc_weights_iso(const orc_ctx *c, double lambda, double *W) {
  const long n3 = c->n3;
  long z;
#pragma omp parallel for schedule(dynamic, 8)
  for (z = 0; z < n3; z++) {
    long x;
    for (x = 0; x < n3; x++) W[z * n3 + x] = orc_weight_one(c, lambda, (int)z, (int)x);
  }
}
