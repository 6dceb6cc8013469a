/* oracle/shims/gsl/gsl_integration.h -- TEST INFRASTRUCTURE ONLY.
 * The slice of GSL the isotropic weight generator uses (/root/reference/src/weights.c:184-206):
 * gsl_function, a workspace handle and gsl_integration_qag; key 2 (GK21) is the only key the
 * reference passes.  Backed by oracle/qag21.c. */
#ifndef ORC_SHIM_GSL_INTEGRATION_H
#define ORC_SHIM_GSL_INTEGRATION_H
#include <stddef.h>
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
typedef struct { size_t limit; } gsl_integration_workspace;
gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *w, double *result,
                        double *abserr);
#endif
