/* oracle/qag21.h -- TEST INFRASTRUCTURE ONLY. QUADPACK QAG / GK21 restatement (see qag21.c). */
#ifndef ORC_QAG21_H
#define ORC_QAG21_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef double (*orc_integrand)(double x, void *ctx);
void orc_gk21(orc_integrand f, void *ctx, double a, double b, double *result, double *abserr,
              double *resabs, double *resasc);
/* returns 0 on convergence, QUADPACK-style non-zero code otherwise (result is still set) */
int orc_qag21(orc_integrand f, void *ctx, double a, double b, double epsabs, double epsrel,
              size_t limit, double *result, double *abserr);
#ifdef __cplusplus
}
#endif
#endif
