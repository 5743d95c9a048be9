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

// 1/x for the shift terms: hardware seed + two Newton steps on the device (~1 ulp, a quarter of the
// instructions of an IEEE division), plain division on the host
VB_HD double vb_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
#else
    return 1.0 / x;
#endif
}

// psi(x) and lgamma(x) together, x > 0 (the posterior update needs both of the same argument,
// src/vbnmf_update.cpp:59,63 and :85,89): ONE upward shift to xs = x + n >= 10 serves both --
//   psi(x)    = psi(xs) - sum_{j<n} 1/(x+j),   lgamma(x) = lgamma(xs) - log prod_{j<n} (x+j)
// and both Stirling series run in 1/xs^2 (truncation < 1e-17).  About a third of the instructions
// of vb_digamma() + the library lgamma(); same accuracy (checked against mpmath on the host twin,
// tests/test_special_cpu.py).
VB_HD void vb_psi_lgamma(double x, double *psi, double *lgam) {
    const int nstep = x < 10.0 ? (int)ceil(10.0 - x) : 0;
    const double xs = x + (double)nstep;
    const double xi = vb_rcp(xs), x2 = xi * xi, lx = log(xs);
    const double sp = x2 * (1.0 / 12.0 - x2 * (1.0 / 120.0 - x2 * (1.0 / 252.0 - x2 * (1.0 / 240.0 -
                      x2 * (1.0 / 132.0 - x2 * (691.0 / 32760.0 - x2 * (1.0 / 12.0 -
                      x2 * (3617.0 / 8160.0))))))));
    const double sl = xi * (1.0 / 12.0 - x2 * (1.0 / 360.0 - x2 * (1.0 / 1260.0 - x2 * (1.0 / 1680.0 -
                      x2 * (1.0 / 1188.0 - x2 * (691.0 / 360360.0 - x2 * (1.0 / 156.0 -
                      x2 * (3617.0 / 122400.0))))))));
    double acc = lx - 0.5 * xi - sp, prod = 1.0;
    for (int j = nstep - 1; j >= 0; j--) {
        const double t = x + (double)j;
        acc -= vb_rcp(t);
        prod *= t;
    }
    *psi = acc;
    // (xs - 1/2) log xs - xs + log(2 pi)/2 + series, minus the log of the shift product
    *lgam = (xs - 0.5) * lx - xs + 0.91893853320467274178 + sl - (nstep ? log(prod) : 0.0);
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

// hyper_update (R/bayesian.R:2-53): Newton iteration on the shapes aw, ah from the four means
// mn = {mean log lw, mean log lh, mean ew, mean eh}; bw <- mean(ew) if flags[1]; bh <- mean(eh) in
// BOTH branches of :50-51.  Returns 0, or 2 when `niter` steps are exhausted (:43).
VB_HD int vb_hyper_update(const int *flags, const double *mn, double *hyper, int niter, double tol) {
    if (flags[0] + flags[1] + flags[2] + flags[3] == 0) return 0;                 // :4
    const double lwm = mn[0], lhm = mn[1], ewm = mn[2], ehm = mn[3];
    double aw0 = hyper[0], ah0 = hyper[2];
    const double bw0 = hyper[1], bh0 = hyper[3];
    double aw1 = aw0, ah1 = ah0;
    if (flags[0] + flags[2] > 0) {                                                 // :15
        int i = 1;
        while (i < niter) {                                                        // :17
            double dw = 0.0, dh = 0.0;
            if (flags[0])
                dw = (log(aw0) - vb_digamma(aw0) - ewm / bw0 + 1.0 + lwm - log(bw0)) /
                     (1.0 / aw0 - vb_trigamma(aw0));                               // :19-20
            if (flags[2])
                dh = (log(ah0) - vb_digamma(ah0) - ehm / bh0 + 1.0 + lhm - log(bh0)) /
                     (1.0 / ah0 - vb_trigamma(ah0));                               // :23-24
            aw1 = aw0 - dw;
            ah1 = ah0 - dh;
            // :28-35 halve the step until the shape is positive; the cap only matters for an
            // infinite step, where the R loop would never end
            for (int g = 0; aw1 <= 0 && g < 1200; g++) { dw = dw / 2; aw1 = aw0 - dw; }
            for (int g = 0; ah1 <= 0 && g < 1200; g++) { dh = dh / 2; ah1 = ah0 - dh; }
            if (aw1 <= 0 || ah1 <= 0) return 2;
            const double df =
                (1 - aw1 / aw0) * (1 - aw1 / aw0) + (1 - ah1 / ah0) * (1 - ah1 / ah0);  // :37
            if (df < tol) break;                                                   // :38
            aw0 = aw1;
            ah0 = ah1;
            i++;
        }
        if (i == niter) return 2;                                                  // :43
    }
    hyper[0] = aw1;
    hyper[1] = flags[1] ? ewm : bw0;                                               // :48-49
    hyper[2] = ah1;
    hyper[3] = ehm;                                                                // :50-51
    return 0;
}
