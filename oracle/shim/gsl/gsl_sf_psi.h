// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for GSL's gsl_sf_psi (digamma), called at
// /root/reference/src/vbnmf_update.cpp:59,63.  GSL is not installed in this container and
// the reference pins no GSL version (src/Makevars:2 links the system -lgsl).  gsl_sf_psi's
// documented contract is psi(x) to double precision for x != 0,-1,-2,...; this stand-in
// evaluates it in long double (upward recurrence to x >= 12, then the Stirling series with
// Bernoulli terms up to x^-16) and rounds once, so it is correct to <= 1 ulp of double for
// x > 0 (cross-checked against mpmath at 50 digits in tests/test_oracle.py).
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
static inline double gsl_sf_psi(double x_in) {
    long double x = (long double)x_in, acc = 0.0L;
    while (x < 12.0L) {
        acc -= 1.0L / x;
        x += 1.0L;
    }
    const long double xi = 1.0L / x, x2 = xi * xi;
    // sum_{k>=1} B_{2k} / (2k x^{2k})
    long double s = x2 * (1.0L / 12.0L -
                    x2 * (1.0L / 120.0L -
                    x2 * (1.0L / 252.0L -
                    x2 * (1.0L / 240.0L -
                    x2 * (1.0L / 132.0L -
                    x2 * (691.0L / 32760.0L -
                    x2 * (1.0L / 12.0L -
                    x2 * (3617.0L / 8160.0L))))))));
    return (double)(acc + __builtin_logl(x) - 0.5L * xi - s);
}
#ifdef __cplusplus
}
#endif
