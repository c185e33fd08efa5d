// TEST-ONLY host build of fitnoise_b200/csrc/fn_special.cuh (g++, no CUDA), so that the
// special functions the kernels use can be checked against scipy/mpmath on a box without a GPU.
// Never loaded by the product.
#include "../../fitnoise_b200/csrc/fn_special.cuh"

extern "C" {
void hc_digamma(int n, const double* x, double* out) { for (int i = 0; i < n; ++i) out[i] = fn::digamma(x[i]); }
void hc_lgamma2(int n, const double* x, double* out) { for (int i = 0; i < n; ++i) { double d; fn::lgamma_digamma(x[i], out[i], d); } }
void hc_digamma2(int n, const double* x, double* out) { for (int i = 0; i < n; ++i) { double l; fn::lgamma_digamma(x[i], l, out[i]); } }
void hc_f_sf(int n, const double* f, const double* dfn, const double* dfd, double* out) {
    for (int i = 0; i < n; ++i) out[i] = fn::f_sf(f[i], dfn[i], dfd[i]);
}
void hc_chi2_sf(int n, const double* q, const double* df, double* out) {
    for (int i = 0; i < n; ++i) out[i] = fn::chi2_sf(q[i], df[i]);
}
}
