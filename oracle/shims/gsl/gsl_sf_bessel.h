/* oracle/shims/gsl/gsl_sf_bessel.h -- TEST INFRASTRUCTURE ONLY. Included, never called, by
 * /root/reference/src/weights.c:7. */
#ifndef ORC_SHIM_GSL_SF_BESSEL_H
#define ORC_SHIM_GSL_SF_BESSEL_H
#endif
