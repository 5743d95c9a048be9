// Special functions used by the VB-NMF update, usable from host and device.
//
// The reference evaluates these through GSL and base R:
//   gsl_sf_psi      src/vbnmf_update.cpp:59,63          -> vb_digamma()
//   gsl_sf_lngamma  src/vbnmf_update.cpp:81,82,85,87,89 -> lgamma() (CUDA / libm double)
//   digamma, psigamma(.,1)  R/bayesian.R:19-24          -> vb_digamma(), vb_trigamma() (host)
// All double precision; arguments are always > 0 on this path (shape parameters).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define VB_HD __host__ __device__ __forceinline__
#else
#define VB_HD static inline
#endif

// psi(x), x > 0: upward recurrence to xs >= 10, Stirling series with Bernoulli terms to xs^-16
// (truncation < 1e-17), recurrence terms added smallest first so -1/x enters last.
VB_HD double vb_digamma(double x) {
    const int nstep = x < 10.0 ? (int)ceil(10.0 - x) : 0;
    const double xs = x + (double)nstep;
    const double xi = 1.0 / xs, x2 = xi * xi;
    const double s = x2 * (1.0 / 12.0 - x2 * (1.0 / 120.0 - x2 * (1.0 / 252.0 - x2 * (1.0 / 240.0 -
                     x2 * (1.0 / 132.0 - x2 * (691.0 / 32760.0 - x2 * (1.0 / 12.0 -
                     x2 * (3617.0 / 8160.0))))))));
    double acc = log(xs) - 0.5 * xi - s;
    for (int j = nstep - 1; j >= 0; j--) acc -= 1.0 / (x + (double)j);
    return acc;
}

// psi'(x), x > 0 (hyper-parameter Newton step only; host side)
VB_HD double vb_trigamma(double x) {
    const int nstep = x < 10.0 ? (int)ceil(10.0 - x) : 0;
    const double xs = x + (double)nstep;
    const double xi = 1.0 / xs, x2 = xi * xi;
    const double s = xi * x2 * (1.0 / 6.0 - x2 * (1.0 / 30.0 - x2 * (1.0 / 42.0 - x2 * (1.0 / 30.0 -
                     x2 * (5.0 / 66.0 - x2 * (691.0 / 2730.0 - x2 * (7.0 / 6.0)))))));
    double acc = xi + 0.5 * x2 + s;
    for (int j = nstep - 1; j >= 0; j--) acc += 1.0 / ((x + (double)j) * (x + (double)j));
    return acc;
}
