// fp64 special functions used by the per-gene kernels: digamma, F and chi-square survival
// functions.  Replaces the scipy calls on the reference path:
//   scipy.stats.f.sf      distributions.py:155   (Mvt.p_value)
//   scipy.stats.chi2.sf   distributions.py:57    (Mvnormal.p_value)
//   scipy.special.gammaln env.py:173-175         (Mvt.log_density; lgamma() here)
// digamma is new: the reference got d/d(df) from Theano autodiff (fitnoise.py:93).
//
// Everything is `FN_HD` so that tests/hostcheck can compile the very same code with g++ and
// compare it with scipy/mpmath on a box without a GPU.  The product only ever runs it on device.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FN_HD __host__ __device__ __forceinline__
#define FN_HDN inline __host__ __device__ __noinline__
#else
#define FN_HD inline
#define FN_HDN inline
#endif

namespace fn {

// psi(x) for x > 0: upward recurrence to x >= 10, then the asymptotic series
// ln x - 1/(2x) - sum B_2k / (2k x^2k)  (k = 1..7; truncation < 1e-17 at x = 10).
FN_HDN double digamma(double x) {
    double acc = 0.0;
    while (x < 10.0) { acc -= 1.0 / x; x += 1.0; }
    const double r = 1.0 / x, r2 = r * r;
    double s = r2 * (1.0 / 12.0 - r2 * (1.0 / 120.0 - r2 * (1.0 / 252.0 - r2 * (1.0 / 240.0
             - r2 * (1.0 / 132.0 - r2 * (691.0 / 32760.0 - r2 * (1.0 / 12.0)))))));
    return acc + log(x) - 0.5 * r - s;
}

// ln Gamma(x) and psi(x) together for x > 0.  Both shift x up to X = x + n >= 10 and share the product
// P = x (x+1) ... (x+n-1):  ln Gamma(x) = ln Gamma(X) - ln P,  psi(x) = psi(X) - P'/P  (P' carried along
// by P'_{k+1} = P'_k (x+k) + P_k), then Stirling / the asymptotic series at X (next terms < 1e-16).
// One call costs two logarithms and one division; the per-launch density tables sit on the critical
// path of every evaluation, where the library lgamma plus ten dependent divisions took ~2 us.
FN_HD void lgamma_digamma(double x, double& lg, double& dg) {
    double P = 1.0, dP = 0.0;
    bool shifted = false;
    while (x < 10.0) { dP = dP * x + P; P *= x; x += 1.0; shifted = true; }
    const double r = 1.0 / x, r2 = r * r, lx = log(x);
    const double sl = r * (1.0 / 12.0 - r2 * (1.0 / 360.0 - r2 * (1.0 / 1260.0 - r2 * (1.0 / 1680.0
                    - r2 * (1.0 / 1188.0 - r2 * (691.0 / 360360.0 - r2 * (1.0 / 156.0)))))));
    const double sd = r2 * (1.0 / 12.0 - r2 * (1.0 / 120.0 - r2 * (1.0 / 252.0 - r2 * (1.0 / 240.0
                    - r2 * (1.0 / 132.0 - r2 * (691.0 / 32760.0 - r2 * (1.0 / 12.0)))))));
    lg = (x - 0.5) * lx - x + 0.91893853320467274178 + sl;
    dg = lx - 0.5 * r - sd;
    if (shifted) { lg -= log(P); dg -= dP / P; }
}

// 1/x for the continued fractions below.  Device: hardware seed and two Newton steps (one MUFU and four
// FMAs; within an ulp for normal x of either sign, +-inf for x = 0) instead of the ~15-instruction IEEE
// division -- the test and noise p-value kernels spend most of their time in the six divisions of a
// Lentz step.  Host (tests/hostcheck): the exact quotient.
FN_HD double cf_rcp(double x) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;
#endif
}

// Continued fraction for the incomplete beta function (modified Lentz; Numerical Recipes 6.4).
FN_HDN double betacf(double a, double b, double x) {
    const double tiny = 1e-300, eps = 1e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x * cf_rcp(qap);
    if (fabs(d) < tiny) d = tiny;
    d = cf_rcp(d);
    double h = d;
    for (int m = 1; m <= 2000; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x * cf_rcp((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa * cf_rcp(c); if (fabs(c) < tiny) c = tiny;
        d = cf_rcp(d);
        h *= d * c;
        aa = -(a + m) * (qab + m) * x * cf_rcp((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa * cf_rcp(c); if (fabs(c) < tiny) c = tiny;
        d = cf_rcp(d);
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) <= eps) break;
    }
    return h;
}

// Regularised incomplete beta I_x(a, b) given BOTH x and y = 1 - x (each formed without
// cancellation by the caller), a, b > 0.
FN_HDN double betainc_xy(double a, double b, double x, double y) {
    if (!(x > 0.0)) return (x == 0.0) ? 0.0 : NAN;
    if (!(y > 0.0)) return (y == 0.0) ? 1.0 : NAN;
    const double lx = (x > 0.5) ? log1p(-y) : log(x);
    const double ly = (y > 0.5) ? log1p(-x) : log(y);
    const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
    const double front = exp(a * lx + b * ly - lbeta);
    if (x < (a + 1.0) / (a + b + 2.0))
        return front * betacf(a, b, x) / a;
    return 1.0 - front * betacf(b, a, y) / b;
}

// Survival function of F(dfn, dfd) at f >= 0:  I_{dfd/(dfd+dfn f)}(dfd/2, dfn/2).
FN_HD double f_sf(double f, double dfn, double dfd) {
    if (f != f || dfn != dfn || dfd != dfd) return NAN;
    if (f <= 0.0) return 1.0;
    if (isinf(f)) return 0.0;
    const double t = dfn * f;
    const double den = dfd + t;
    return betainc_xy(0.5 * dfd, 0.5 * dfn, dfd / den, t / den);
}

// Regularised upper incomplete gamma Q(a, x), a > 0, x >= 0 (series / continued fraction, NR 6.2).
FN_HDN double gammaincc(double a, double x) {
    if (x != x || a != a) return NAN;
    if (x <= 0.0) return 1.0;
    if (isinf(x)) return 0.0;
    const double lead = exp(a * log(x) - x - lgamma(a));
    if (x < a + 1.0) {
        double ap = a, del = cf_rcp(a), sum = del;
        for (int n = 0; n < 5000; ++n) {
            ap += 1.0;
            del *= x * cf_rcp(ap);
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        return 1.0 - sum * lead;
    }
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = cf_rcp(b), h = d;
    for (int i = 1; i <= 5000; ++i) {
        const double an = -i * (i - a);
        b += 2.0;
        d = an * d + b; if (fabs(d) < tiny) d = tiny;
        c = b + an * cf_rcp(c); if (fabs(c) < tiny) c = tiny;
        d = cf_rcp(d);
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) <= 1e-16) break;
    }
    return lead * h;
}

// Survival function of chi-square(df) at q.
FN_HD double chi2_sf(double q, double df) { return gammaincc(0.5 * df, 0.5 * q); }

}  // namespace fn
