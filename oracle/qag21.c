/*
 * oracle/qag21.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Restatement of the adaptive Gauss-Kronrod 21-point quadrature that the reference reaches
 * through GSL: gsl_integration_qag(&F, 0.0, L_v, 1e-8, 1e-8, 10000, key=2, ...) at
 * /root/reference/src/weights.c:203.  GSL is an un-vendored, un-versioned dependency of the
 * reference (CMakeLists.txt:72-75, find_package(GSL)); it is absent from this image, so the
 * published QUADPACK QAG algorithm (Piessens et al., routine DQAGE + DQK21 + DQPSRT, which GSL's
 * qag.c/qk.c/qpsrt.c transcribe) is restated here:
 *   - 21-point Kronrod rule with embedded 10-point Gauss rule, QUADPACK error heuristic
 *     (200*err/resasc)^1.5 and the 50*eps*resabs floor;
 *   - bisection of the interval with the largest error estimate, bookkeeping in an
 *     error-descending order list, the QUADPACK round-off detectors;
 *   - final result = plain sum of the stored interval results in storage order.
 * Pinned by tests/test_oracle_weights.py against the reference's byte-exact golden files
 * tests/BKW8/target/N8_isotropic_L_v5_lambda0.wts and
 * tests/heat_transport/target/N8_isotropic_L_v9_lambda1.wts.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include "qag21.h"

/* Kronrod abscissae x_0 > x_1 > ... > x_10 = 0 ; odd entries are the 10-point Gauss abscissae */
static const double XGK[11] = {
    0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
    0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
    0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
    0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
    0.294392862701460198131126603103866, 0.148874338981631210884826001129720,
    0.000000000000000000000000000000000};
static const double WG[5] = {
    0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
    0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
    0.295524224714752870173815619188769};
static const double WGK[11] = {
    0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
    0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
    0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
    0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
    0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
    0.149445554002916905664936468389821};

static double rescale_err(double err, double resabs, double resasc) {
  err = fabs(err);
  if (resasc != 0.0 && err != 0.0) {
    double scale = pow((200.0 * err / resasc), 1.5);
    err = (scale < 1.0) ? resasc * scale : resasc;
  }
  if (resabs > DBL_MIN / (50.0 * DBL_EPSILON)) {
    double floor_err = 50.0 * DBL_EPSILON * resabs;
    if (floor_err > err) err = floor_err;
  }
  return err;
}

void orc_gk21(orc_integrand f, void *ctx, double a, double b, double *result, double *abserr,
              double *resabs, double *resasc) {
  const double center = 0.5 * (a + b);
  const double half = 0.5 * (b - a);
  const double ahalf = fabs(half);
  const double fc = f(center, ctx);
  double fv1[10], fv2[10];
  double rg = 0.0, rk = fc * WGK[10], rabs = fabs(rk);
  int j;
  for (j = 0; j < 5; j++) { /* Gauss nodes (odd Kronrod indices) */
    const int t = 2 * j + 1;
    const double dx = half * XGK[t];
    const double f1 = f(center - dx, ctx), f2 = f(center + dx, ctx);
    const double fs = f1 + f2;
    fv1[t] = f1; fv2[t] = f2;
    rg += WG[j] * fs;
    rk += WGK[t] * fs;
    rabs += WGK[t] * (fabs(f1) + fabs(f2));
  }
  for (j = 0; j < 5; j++) { /* Kronrod-only nodes (even indices) */
    const int t = 2 * j;
    const double dx = half * XGK[t];
    const double f1 = f(center - dx, ctx), f2 = f(center + dx, ctx);
    fv1[t] = f1; fv2[t] = f2;
    rk += WGK[t] * (f1 + f2);
    rabs += WGK[t] * (fabs(f1) + fabs(f2));
  }
  {
    const double mean = rk * 0.5;
    double rasc = WGK[10] * fabs(fc - mean);
    for (j = 0; j < 10; j++) rasc += WGK[j] * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));
    *result = rk * half;
    *resabs = rabs * ahalf;
    *resasc = rasc * ahalf;
    *abserr = rescale_err((rk - rg) * half, *resabs, *resasc);
  }
}

/* interval store: parallel arrays + order list descending in error estimate */
typedef struct {
  size_t limit, size, nrmax, imax;
  double *a, *b, *r, *e;
  size_t *order;
} qstore;

static void resort(qstore *w) { /* DQPSRT */
  const size_t last = w->size - 1;
  size_t i_nrmax = w->nrmax;
  size_t i_maxerr = w->order[i_nrmax];
  double errmax, errmin;
  long i, k, top;
  if (last < 2) {
    w->order[0] = 0; w->order[1] = 1;
    w->imax = i_maxerr;
    return;
  }
  errmax = w->e[i_maxerr];
  while (i_nrmax > 0 && errmax > w->e[w->order[i_nrmax - 1]]) {
    w->order[i_nrmax] = w->order[i_nrmax - 1];
    i_nrmax--;
  }
  top = (last < (w->limit / 2 + 2)) ? (long)last : (long)(w->limit - last + 1);
  i = (long)i_nrmax + 1;
  while (i < top && errmax < w->e[w->order[i]]) {
    w->order[i - 1] = w->order[i];
    i++;
  }
  w->order[i - 1] = i_maxerr;
  errmin = w->e[last];
  k = top - 1;
  while (k > i - 2 && errmin >= w->e[w->order[k]]) {
    w->order[k + 1] = w->order[k];
    k--;
  }
  w->order[k + 1] = last;
  w->imax = w->order[i_nrmax];
  w->nrmax = i_nrmax;
}

static int too_small(double a1, double a2, double b2) {
  const double tmp = (1.0 + 100.0 * DBL_EPSILON) * (fabs(a2) + 1000.0 * DBL_MIN);
  return (fabs(a1) <= tmp && fabs(b2) <= tmp);
}

int orc_qag21(orc_integrand f, void *ctx, double a, double b, double epsabs, double epsrel,
              size_t limit, double *result, double *abserr) {
  qstore w;
  double area, errsum, res0, err0, rabs0, rasc0, tol, roundoff;
  size_t iter = 0, k;
  int rt1 = 0, rt2 = 0, etype = 0;
  /* most integrals converge in a handful of bisections: grow storage lazily */
  size_t cap = 64;
  *result = 0.0; *abserr = 0.0;
  orc_gk21(f, ctx, a, b, &res0, &err0, &rabs0, &rasc0);
  tol = fmax(epsabs, epsrel * fabs(res0));
  roundoff = 50.0 * DBL_EPSILON * rabs0;
  if (err0 <= roundoff && err0 > tol) { *result = res0; *abserr = err0; return 2; }
  if ((err0 <= tol && err0 != rasc0) || err0 == 0.0) { *result = res0; *abserr = err0; return 0; }
  if (limit == 1) { *result = res0; *abserr = err0; return 1; }
  if (cap > limit) cap = limit;
  w.limit = limit; w.size = 1; w.nrmax = 0; w.imax = 0;
  w.a = malloc(cap * sizeof(double)); w.b = malloc(cap * sizeof(double));
  w.r = malloc(cap * sizeof(double)); w.e = malloc(cap * sizeof(double));
  w.order = malloc((cap + 1) * sizeof(size_t));
  w.a[0] = a; w.b[0] = b; w.r[0] = res0; w.e[0] = err0; w.order[0] = 0;
  area = res0; errsum = err0; iter = 1;
  do {
    const size_t im = w.imax;
    const double ai = w.a[im], bi = w.b[im], ri = w.r[im], ei = w.e[im];
    const double a1 = ai, b1 = 0.5 * (ai + bi), a2 = b1, b2 = bi;
    double ar1, ar2, e1, e2, ab1, ab2, as1, as2, ar12, e12;
    size_t inew;
    orc_gk21(f, ctx, a1, b1, &ar1, &e1, &ab1, &as1);
    orc_gk21(f, ctx, a2, b2, &ar2, &e2, &ab2, &as2);
    ar12 = ar1 + ar2; e12 = e1 + e2;
    errsum += (e12 - ei);
    area += ar12 - ri;
    if (as1 != e1 && as2 != e2) {
      const double delta = ri - ar12;
      if (fabs(delta) <= 1.0e-5 * fabs(ar12) && e12 >= 0.99 * ei) rt1++;
      if (iter >= 10 && e12 > ei) rt2++;
    }
    tol = fmax(epsabs, epsrel * fabs(area));
    if (errsum > tol) {
      if (rt1 >= 6 || rt2 >= 20) etype = 2;
      if (too_small(a1, a2, b2)) etype = 3;
    }
    if (w.size == cap) {
      cap = (cap * 2 > limit) ? limit : cap * 2;
      w.a = realloc(w.a, cap * sizeof(double)); w.b = realloc(w.b, cap * sizeof(double));
      w.r = realloc(w.r, cap * sizeof(double)); w.e = realloc(w.e, cap * sizeof(double));
      w.order = realloc(w.order, (cap + 1) * sizeof(size_t));
    }
    inew = w.size;
    if (e2 > e1) { /* larger-error half stays in slot im */
      w.a[im] = a2; w.r[im] = ar2; w.e[im] = e2;
      w.a[inew] = a1; w.b[inew] = b1; w.r[inew] = ar1; w.e[inew] = e1;
    } else {
      w.b[im] = b1; w.r[im] = ar1; w.e[im] = e1;
      w.a[inew] = a2; w.b[inew] = b2; w.r[inew] = ar2; w.e[inew] = e2;
    }
    w.size++;
    resort(&w);
    iter++;
  } while (iter < limit && !etype && errsum > tol);
  {
    double s = 0.0;
    for (k = 0; k < w.size; k++) s += w.r[k];
    *result = s;
    *abserr = errsum;
  }
  free(w.a); free(w.b); free(w.r); free(w.e); free(w.order);
  if (errsum <= tol) return 0;
  if (etype) return etype;
  return (iter == limit) ? 1 : 4;
}
