// TEST INFRASTRUCTURE: the host twin of ccfindr_b200/csrc/special.cuh as a C library, so that the
// special functions the CUDA kernels use can be checked against mpmath without a GPU.
#include "../ccfindr_b200/csrc/special.cuh"
extern "C" {
double twin_digamma(double x) { return vb_digamma(x); }
double twin_trigamma(double x) { return vb_trigamma(x); }
void twin_psi_lgamma(double x, double *psi, double *lg) { vb_psi_lgamma(x, psi, lg); }
}
