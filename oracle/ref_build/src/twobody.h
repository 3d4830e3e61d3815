/* twobody.h -- stand-in for the header of the third-party dependency `twobody`
 * (adrn/twobody, twobody/src/twobody.h), which is NOT under /root/reference and is not
 * installed in this image.  TEST INFRASTRUCTURE ONLY.
 *
 * thejoker/src/fast_likelihood.pyx:27-30 declares c_rv_from_elements from this header.
 * When the reference's Cython is compiled here (oracle/ref_build/build_ref.py) the symbol
 * is supplied by oracle/ref_build/twobody_shim.c, i.e. by the oracle's restatement of the
 * published twobody algorithm -- everything else in the resulting extension is the
 * reference's own code. */
#ifndef TJB_REF_TWOBODY_H
#define TJB_REF_TWOBODY_H
void c_rv_from_elements(double *t, double *rv, int N_t, double P, double K, double e,
                        double omega, double phi0, double t0, double tol, int maxiter);
#endif
