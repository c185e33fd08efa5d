// Host-side (CPU, O(m c^2)) helpers of the C-ABI layer: design whitening and theta unpacking.
// Pure C++ -- also compiled into tests/hostcheck.
#pragma once
#include <math.h>
#include <string.h>
#include <vector>
#include "fn_gene.cuh"

namespace fn {

// The template widths the kernels are instantiated for; a design with c columns runs in the
// next width up, its extra columns zero (Normal<C>::pad makes that exact).
inline int round_width(int c) {
    static const int widths[] = {1, 2, 3, 4, 6, 8, 12, 16};
    for (int w : widths) if (c <= w) return w;
    return -1;
}

// Thin QR of an m x c design by twice-iterated modified Gram-Schmidt.  The kernels work in the
// orthonormal basis Q (every likelihood term depends on the design only through the column space
// of its retained rows), which keeps the c x c normal equations as well conditioned as
// the weights allow.  coef = R^-1 coef_Q.
struct Design {
    int m = 0, c = 0, width = 0;
    std::vector<double> q;      // m x width, row-major, zero padded
    std::vector<double> r;      // c x c upper triangular (row-major)
    std::vector<double> rinv;   // c x c upper triangular (row-major)
    bool build(int m_, int c_, const double* x) {
        m = m_; c = c_; width = round_width(c_);
        if (width < 0 || c < 1 || m < 1) return false;
        std::vector<double> col((size_t)m * c);
        for (int j = 0; j < m; ++j) for (int i = 0; i < c; ++i) col[(size_t)i * m + j] = x[(size_t)j * c + i];
        r.assign((size_t)c * c, 0.0);
        for (int i = 0; i < c; ++i) {
            double* v = &col[(size_t)i * m];
            double norm0 = 0.0;
            for (int j = 0; j < m; ++j) norm0 += v[j] * v[j];
            norm0 = sqrt(norm0);
            for (int pass = 0; pass < 2; ++pass)
                for (int t = 0; t < i; ++t) {
                    const double* u = &col[(size_t)t * m];
                    double dot = 0.0;
                    for (int j = 0; j < m; ++j) dot += u[j] * v[j];
                    for (int j = 0; j < m; ++j) v[j] -= dot * u[j];
                    r[(size_t)t * c + i] += dot;
                }
            double norm = 0.0;
            for (int j = 0; j < m; ++j) norm += v[j] * v[j];
            norm = sqrt(norm);
            if (!(norm > 1e-10 * norm0) || !(norm0 > 0.0)) return false;     // rank deficient design
            for (int j = 0; j < m; ++j) v[j] /= norm;
            r[(size_t)i * c + i] = norm;
        }
        q.assign((size_t)m * width, 0.0);
        for (int j = 0; j < m; ++j) for (int i = 0; i < c; ++i) q[(size_t)j * width + i] = col[(size_t)i * m + j];
        rinv.assign((size_t)c * c, 0.0);
        for (int j = 0; j < c; ++j) {
            rinv[(size_t)j * c + j] = 1.0 / r[(size_t)j * c + j];
            for (int i = j - 1; i >= 0; --i) {
                double s = 0.0;
                for (int t = i + 1; t <= j; ++t) s -= r[(size_t)i * c + t] * rinv[(size_t)t * c + j];
                rinv[(size_t)i * c + j] = s / r[(size_t)i * c + i];
            }
        }
        return true;
    }
};

inline int theta_length(const Family& f) { return (f.dist == DIST_T ? 1 : 0) + f.na + f.nb; }

// theta (in the reference's _initial order: [df], theta_a block, theta_b block) -> the scalars and
// the two per-sample tables the kernels consume.  Returns false on a non-finite / non-positive theta.
inline bool unpack_theta(const Family& f, int m, const double* theta, double& v, int& is_t,
                         double& theta0, std::vector<double>& ta, std::vector<double>& tb) {
    int at = 0;
    is_t = f.dist != DIST_NORMAL;
    v = f.dist == DIST_T ? theta[at++] : f.fixed_df;
    if (is_t && !(v > 0.0 && v < INFINITY)) return false;
    const double* a = theta + at;
    const double* b = theta + at + f.na;
    for (int i = 0; i < f.na + f.nb; ++i) if (!(a[i] > 0.0 && a[i] < INFINITY)) return false;
    theta0 = 1.0;
    ta.assign(m, 0.0); tb.assign(m, 0.0);
    if (f.kind == KIND_SCALE1) {
        if (f.na == 1) theta0 = a[0];
    } else if (f.kind == KIND_SCALE_PS) {
        for (int j = 0; j < m; ++j) { ta[j] = 1.0 / a[j]; tb[j] = log(a[j]); }
    } else {
        for (int j = 0; j < m; ++j) { ta[j] = a[f.na == 1 ? 0 : j]; tb[j] = b[f.nb == 1 ? 0 : j]; }
    }
    return true;
}

// [design | factors | 0] table of width `width` for the factor models (fn_gene.cuh, eval_factors)
inline std::vector<double> augment_design(const Design& d, int width, int nf, const double* F) {
    std::vector<double> z((size_t)d.m * width, 0.0);
    for (int j = 0; j < d.m; ++j) {
        for (int i = 0; i < d.c; ++i) z[(size_t)j * width + i] = d.q[(size_t)j * d.width + i];
        for (int f = 0; f < nf; ++f) z[(size_t)j * width + d.c + f] = F[(size_t)j * nf + f];
    }
    return z;
}

#define FN_MAX_FACTORS 4

inline bool family_valid(const Family& f, int m, bool has_aux) {
    if (f.nf < 0 || f.nf > FN_MAX_FACTORS || (f.nf > 0 && f.kind != KIND_SCALE1)) return false;
    if (f.kind == KIND_SCALE1) return (f.na == 0 || f.na == 1) && f.nb == 0;
    if (f.kind == KIND_SCALE_PS) return f.na == m && f.nb == 0;
    if (f.kind == KIND_PATSEQ) return has_aux && (f.na == 1 || f.na == m) && (f.nb == 1 || f.nb == m);
    return false;
}

}  // namespace fn

#define FN_DISPATCH_WIDTH(W, ...)                                   \
    switch (W) {                                                    \
        case 1: { constexpr int C = 1; __VA_ARGS__; } break;        \
        case 2: { constexpr int C = 2; __VA_ARGS__; } break;        \
        case 3: { constexpr int C = 3; __VA_ARGS__; } break;        \
        case 4: { constexpr int C = 4; __VA_ARGS__; } break;        \
        case 6: { constexpr int C = 6; __VA_ARGS__; } break;        \
        case 8: { constexpr int C = 8; __VA_ARGS__; } break;        \
        case 12: { constexpr int C = 12; __VA_ARGS__; } break;      \
        case 16: { constexpr int C = 16; __VA_ARGS__; } break;      \
        default: break;                                             \
    }
