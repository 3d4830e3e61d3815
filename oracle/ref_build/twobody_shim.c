/* twobody_shim.c -- supplies twobody's c_rv_from_elements to the compiled reference
 * Cython (see src/twobody.h).  TEST INFRASTRUCTURE ONLY. */
#include "src/twobody.h"
#include "../joker_oracle.h"

void c_rv_from_elements(double *t, double *rv, int N_t, double P, double K, double e,
                        double omega, double phi0, double t0, double tol, int maxiter) {
  orc_rv_from_elements(t, rv, N_t, P, K, e, omega, phi0, t0, tol, maxiter, 0);
}
