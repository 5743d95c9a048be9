// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for GSL's gsl_sf_lngamma, called at
// /root/reference/src/vbnmf_update.cpp:81,82,85,87,89.  GSL is not installed here; the
// contract is log|Gamma(x)| to double precision.  glibc's lgammal (80-bit) rounded once to
// double meets that for x > 0.
#pragma once
#include <math.h>
#ifdef __cplusplus
extern "C" {
#endif
static inline double gsl_sf_lngamma(double x) { return (double)lgammal((long double)x); }
#ifdef __cplusplus
}
#endif
