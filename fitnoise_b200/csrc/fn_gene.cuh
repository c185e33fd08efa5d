// Per-gene REML mathematics of the Fitnoise hot path (SURVEY.md section 8a "Math").
//
// What the reference does per gene with a complete QR, an explicit inverse and a determinant of
// the (k-c)x(k-c) transformed covariance (fitnoise.py:23-50, distributions.py:4-7,135-147,
// 167-215) is done here with the c x c weighted normal equations only:
//
//   w_j = 1/sigma2_j (0 for dropped samples), A = X'WX, b = X'Wy, beta = A^-1 b
//   d2        = z2' S^-1 z2        = sum_j w_j (y_j - x_j beta)^2 = y'Wy - b'beta
//   ln det S  = sum_ret ln sigma2_j + ln det A - ln det (X_ret' X_ret)
//   d(-ll)/d sigma2_j = 1/2 (w_j - w_j^2 x_j'A^-1 x_j) - kappa (w_j r_j)^2
//
// Thread mapping (device): a gene is handled by `lpg` adjacent lanes of one warp (lpg a power of
// two, 1..32); lane l owns samples l, l+lpg, ... of the gene's row, which lives in shared memory
// (tile staged by fn_tile.cuh).  Lane-partial sums are combined with xor-shuffles, after which
// every lane of the gene holds the full sums and redundantly does the small c x c algebra.
// On the host (tests/hostcheck only) lpg = 1 and the shuffles compile away.
#pragma once
#include <stdint.h>
#include <string.h>
#include "fn_special.cuh"

namespace fn {

enum { KIND_SCALE1 = 0, KIND_SCALE_PS = 1, KIND_PATSEQ = 2 };
// kernel-internal specialisation of KIND_PATSEQ: one shared theta_a and theta_b (eval_patseq_shared)
enum { KIND_PATSEQ_SHARED = 3 };
enum { DIST_NORMAL = 0, DIST_T = 1, DIST_T_FIXED = 2 };

// Variance family: sigma2_j = theta_a[j or 0] * U_j + theta_b[j or 0] * T_j   
//   KIND_SCALE1    U_j = 1/w_j (w = precision weight, or 1), one shared theta_a (na = 1), or
//                  theta_a == 1 fixed (na = 0: Model_independent)          fitnoise.py:530-535,573-578
//   KIND_SCALE_PS  U_j = 1/w_j, theta_a per sample (na = m)                 fitnoise.py:609-614
//   KIND_PATSEQ    U_j = rc_j (* gs if v3), T_j = y_j^2 (tail_own) or gs;   fitnoise.py:671-678,
//                  rc_j = 1/max(count_j,1), gs = avgtail^2; na, nb in {1, m} 717-724,769-776,825-832,875-882
struct Family {
    int kind;
    int na, nb;        // number of theta_a / theta_b entries: 0, 1 or m
    int tail_own;      // PATSEQ v1: T_j = y_j^2
    int v3;            // PATSEQ v3: sigma2_j = (theta_a rc_j + theta_b) * gs
    int dist;          // DIST_NORMAL, DIST_T (df = theta[0]), DIST_T_FIXED (df = fixed_df)
    int nf;            // unknown batch-effect factors (Model_factors_mixin, fitnoise.py:439-518); SCALE1 only
    double fixed_df;   // Model_independent: 1e-10  (fitnoise.py:577)
};

FN_HD bool finite_d(double x) { return fabs(x) < INFINITY; }   // false for NaN too

// ------------------------------------------------------------------------------------------
// lane cooperation
// ------------------------------------------------------------------------------------------
// (lpg is uniform over the CTA: the fall-through switch runs exactly log2(lpg) shuffle steps with one
// branch, where a loop over the distances spent more instructions on loop control than on shuffles --
// 10 % of the instructions of the headline kernel at 1M x 96.)
FN_HD double lane_sum(double v, int lpg) {
#ifdef __CUDA_ARCH__
    switch (lpg) {
        case 32: v += __shfl_xor_sync(0xffffffffu, v, 16);
        case 16: v += __shfl_xor_sync(0xffffffffu, v, 8);
        case 8: v += __shfl_xor_sync(0xffffffffu, v, 4);
        case 4: v += __shfl_xor_sync(0xffffffffu, v, 2);
        case 2: v += __shfl_xor_sync(0xffffffffu, v, 1);
        default: break;
    }
#else
    (void)lpg;
#endif
    return v;
}
// Sum N doubles over the lanes of a gene at once: one loop over the shuffle distances (the per-value
// loops of lane_sum cost more in loop control than in shuffles when called seven times per gene).
template <int N>
FN_HD void lane_sum_n(double* v, int lpg) {
#ifdef __CUDA_ARCH__
#define FN_LANE_STEP(off) _Pragma("unroll") for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
    switch (lpg) {
        case 32: FN_LANE_STEP(16)
        case 16: FN_LANE_STEP(8)
        case 8: FN_LANE_STEP(4)
        case 4: FN_LANE_STEP(2)
        case 2: FN_LANE_STEP(1)
        default: break;
    }
#undef FN_LANE_STEP
#else
    (void)v; (void)lpg;
#endif
}
FN_HD int lane_sum_i(int v, int lpg) {
#ifdef __CUDA_ARCH__
    for (int off = lpg >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
#else
    (void)lpg;
#endif
    return v;
}
FN_HD bool warp_any(bool p) {
#ifdef __CUDA_ARCH__
    return __any_sync(0xffffffffu, p);
#else
    return p;
#endif
}

// Product of many positive doubles kept as mantissa in [1,2) and an integer exponent, so that
// sum_j ln sigma2_j costs one log per gene instead of one per sample.
struct LogProd {
    double mant;
    int expo;
    FN_HD void init() { mant = 1.0; expo = 0; }
    FN_HD void renorm() {
#ifdef __CUDA_ARCH__
        long long bits = __double_as_longlong(mant);
#else
        long long bits; memcpy(&bits, &mant, 8);
#endif
        expo += (int)((bits >> 52) & 0x7ff) - 1023;
        bits = (bits & ~(0x7ffLL << 52)) | (1023LL << 52);
#ifdef __CUDA_ARCH__
        mant = __longlong_as_double(bits);
#else
        memcpy(&mant, &bits, 8);
#endif
    }
    // x must be a positive normal double; anything else poisons the product with NaN
    FN_HD void mul(double x) {
        mant *= x;
        if (!(mant >= 2.3e-308 && mant < 1.7e308)) { mant = NAN; return; }
        renorm();
    }
    FN_HD void lane_reduce(int lpg) {
#ifdef __CUDA_ARCH__
        for (int off = lpg >> 1; off > 0; off >>= 1) {
            const double om = __shfl_xor_sync(0xffffffffu, mant, off);
            const int oe = __shfl_xor_sync(0xffffffffu, expo, off);
            mant *= om; expo += oe;
            if (mant == mant) renorm();
        }
#else
        (void)lpg;
#endif
    }
    FN_HD double log_value() const { return log(mant) + 0.6931471805599453094 * (double)expo; }
};

// ------------------------------------------------------------------------------------------
// view of one gene's row for one lane
// ------------------------------------------------------------------------------------------
struct GeneView {
    double* y;          // row of y (m doubles); writable for the in-place per-sample gradient
    double* aux;        // row of the second array (precision weights or rc), or nullptr
    double gs;          // per-gene scalar (avgtail^2) or 0
    int m;
    int lane, lpg;      // this lane's index within the gene, lanes per gene
    int cnt;            // number of samples this lane owns
    int skew;           // per-gene rotation of the sample order (shared-memory bank spreading)
    int blocked;        // 1: rotate within blocks of `bsize` samples
    int bsize, steps, nblk;   // block size B (power of two dividing m, <= 16), B/lpg steps, m/B blocks
    // `gene_local` = index of the gene within the tile.  Rows of a tile are m doubles apart, so when
    // m is a multiple of 16 the lanes of a warp that work on different genes would all hit the same
    // shared-memory banks (8-byte bank pairs); for other even m part of them would.
    FN_HD void setup(int m_, int lane_, int lpg_, int gene_local) {
        m = m_; lane = lane_; lpg = lpg_;
        cnt = (m - lane + lpg - 1) / lpg;
        if (cnt < 0) cnt = 0;
        bsize = m & -m;                      // largest power of two dividing m
        if (bsize > 16) bsize = 16;
        blocked = (lpg <= bsize) ? 1 : 0;
        steps = blocked ? bsize / lpg : 0;
        nblk = blocked ? m / bsize : 0;
        // Rows are m doubles apart: with B < 16 the rows of 16/B consecutive genes start in different
        // bank pairs already, so the rotation only has to separate genes 16/B apart: gene_local * B / 16.
        // (Rotating by gene_local itself put genes g and g + 8/lpg of a half-warp into the same bank
        // pair whenever B = 8: two-way conflicts on every load at m = 24, 40, 56, ...)
        skew = blocked ? (gene_local * bsize) / 16 : 0;
    }
    // Visit this lane's samples, each exactly once; `on` = false visits nothing.  The order only
    // changes the association of the floating-point sums.  Every thread of a warp makes the same
    // number of visits (no divergence).
    //  blocked: step (r, t) -> sample B t + ((lane + lpg (skew + r)) mod B), r < B/lpg, t < m/B.
    //           With B = 16 the 16 threads of a half-warp read 16 distinct bank pairs at every step
    //           and a warp touches only 16 distinct rows of X; smaller B spreads as far as m allows.
    //  else:    lane, lane + lpg, ... (lpg does not divide a power-of-two factor of m).
    template <class F>
    FN_HD void for_each(bool on, F f) const {
        if (!on) return;
        if (blocked) {
            const int mask = bsize - 1;
            for (int r = 0; r < steps; ++r) {
                const int jr = (lane + lpg * (skew + r)) & mask;
#pragma unroll 2
                for (int t = 0; t < nblk; ++t) f(jr + bsize * t);
            }
            return;
        }
        int j = lane;
#pragma unroll 2
        for (int i = 0; i < cnt; ++i, j += lpg) f(j);
    }
    // The same visit for the one sweep that is hot at large m: the block loop is fully unrolled
    // for the common m/16, so that a rotation step issues all its shared-memory loads (constant
    // offsets from one base address) before the FMAs that consume them.
    template <int NB, class F>
    FN_HD void blocked_unrolled(F f) const {
        for (int r = 0; r < steps; ++r) {
            const int jr = (lane + lpg * (skew + r)) & 15;
#pragma unroll
            for (int t = 0; t < NB; ++t) f(jr + 16 * t);
        }
    }
    // The same visit with a hook after every rotation step, i.e. after at most 8 samples: the caller
    // can keep a running PRODUCT of per-sample values and renormalise it there instead of testing a
    // counter per sample.  The block loop is unrolled for the common block counts.
    template <int BS, int NB, class F, class G>
    FN_HD void steps_unrolled(F f, G step_end) const {
        for (int r = 0; r < steps; ++r) {
            const int jr = (lane + lpg * (skew + r)) & (BS - 1);
#pragma unroll
            for (int t = 0; t < NB; ++t) f(jr + BS * t);
            step_end();
        }
    }
    template <class F, class G>
    FN_HD void for_each_steps(bool on, F f, G step_end) const {
        if (!on) return;
        if (blocked) {
            if (bsize == 16) {
                switch (nblk) {
                    case 1: steps_unrolled<16, 1>(f, step_end); return;
                    case 2: steps_unrolled<16, 2>(f, step_end); return;
                    case 3: steps_unrolled<16, 3>(f, step_end); return;
                    case 4: steps_unrolled<16, 4>(f, step_end); return;
                    case 6: steps_unrolled<16, 6>(f, step_end); return;
                    default: break;
                }
            } else if (bsize == 8) {
                switch (nblk) {
                    case 1: steps_unrolled<8, 1>(f, step_end); return;
                    case 3: steps_unrolled<8, 3>(f, step_end); return;
                    case 5: steps_unrolled<8, 5>(f, step_end); return;
                    default: break;
                }
            }
            const int mask = bsize - 1;
            for (int r = 0; r < steps; ++r) {
                const int jr = (lane + lpg * (skew + r)) & mask;
                for (int t = 0; t < nblk; ++t) {
                    f(jr + bsize * t);
                    if ((t & 7) == 7) step_end();
                }
                step_end();
            }
            return;
        }
        int j = lane;
        for (int i = 0; i < cnt; ++i, j += lpg) { f(j); step_end(); }
    }
    template <class F>
    FN_HD void for_each_hot(bool on, F f) const {
        if (!on) return;
        if (blocked && bsize == 16) {
            switch (nblk) {
                case 1: blocked_unrolled<1>(f); return;
                case 2: blocked_unrolled<2>(f); return;
                case 3: blocked_unrolled<3>(f); return;
                case 4: blocked_unrolled<4>(f); return;
                case 6: blocked_unrolled<6>(f); return;
                case 8: blocked_unrolled<8>(f); return;
                default: break;
            }
        }
        for_each(true, f);
    }
};

// ------------------------------------------------------------------------------------------
// small dense algebra on packed lower-triangular storage: (i,j), j<=i at i(i+1)/2+j
// ------------------------------------------------------------------------------------------
template <int C> struct Tri { enum { N = C * (C + 1) / 2 }; };
#define FN_TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

// 1/sqrt(x) for a positive pivot.  FAST: the device's rsqrt (1 ulp) instead of a correctly rounded
// square root and division -- used where the factor only feeds sums over genes (the likelihood
// kernels); the coefficient kernel, whose per-gene output can amplify the last bit by the condition
// number of an almost collinear sub-design, keeps the two correctly rounded operations.
template <bool FAST>
FN_HD double rsqrt_pos(double x) {
#ifdef __CUDA_ARCH__
    if (FAST) return rsqrt(x);
#endif
    return 1.0 / sqrt(x);
}

// In-place Cholesky A = L L'.  On return a holds L below the diagonal and 1/L_ii on it;
// det = prod of pivots (so ln det A = ln det).  Returns false when a pivot is <= rel_tol * A_ii.
template <int C, bool FAST = false>
FN_HD bool chol(double* a, double& det, double rel_tol) {
    bool ok = true;
    det = 1.0;
#pragma unroll
    for (int i = 0; i < C; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double s = a[FN_TRI(i, j)];
#pragma unroll
            for (int t = 0; t < j; ++t) s -= a[FN_TRI(i, t)] * a[FN_TRI(j, t)];
            if (j < i) {
                a[FN_TRI(i, j)] = s * a[FN_TRI(j, j)];
            } else {
                const double orig = a[FN_TRI(i, i)];
                if (!(s > rel_tol * orig) || !(s > 0.0)) { ok = false; s = 1.0; }
                det *= s;
                a[FN_TRI(i, i)] = rsqrt_pos<FAST>(s);
            }
        }
    }
    return ok;
}

// Solve L L' x = b in place (L as produced by chol()).
template <int C>
FN_HD void chol_solve(const double* l, double* x) {
#pragma unroll
    for (int i = 0; i < C; ++i) {
        double s = x[i];
#pragma unroll
        for (int t = 0; t < i; ++t) s -= l[FN_TRI(i, t)] * x[t];
        x[i] = s * l[FN_TRI(i, i)];
    }
#pragma unroll
    for (int i = C - 1; i >= 0; --i) {
        double s = x[i];
#pragma unroll
        for (int t = i + 1; t < C; ++t) s -= l[FN_TRI(t, i)] * x[t];
        x[i] = s * l[FN_TRI(i, i)];
    }
}

// h = x' A^-1 x = |L^-1 x|^2
template <int C>
FN_HD double leverage(const double* l, const double* x) {
    double z[C];
    double h = 0.0;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        double s = x[i];
#pragma unroll
        for (int t = 0; t < i; ++t) s -= l[FN_TRI(i, t)] * z[t];
        z[i] = s * l[FN_TRI(i, i)];
        h += z[i] * z[i];
    }
    return h;
}

// A^-1 (packed lower) from the Cholesky factor.
template <int C>
FN_HD void chol_inverse(const double* l, double* ainv) {
    double li[Tri<C>::N];           // L^-1, lower triangular
#pragma unroll (C > 8 ? 1 : 64)
    for (int j = 0; j < C; ++j) {
        li[FN_TRI(j, j)] = l[FN_TRI(j, j)];
#pragma unroll (C > 8 ? 1 : 64)
        for (int i = j + 1; i < C; ++i) {
            double s = 0.0;
#pragma unroll (C > 8 ? 1 : 64)
            for (int t = j; t < i; ++t) s -= l[FN_TRI(i, t)] * li[FN_TRI(t, j)];
            li[FN_TRI(i, j)] = s * l[FN_TRI(i, i)];
        }
    }
#pragma unroll (C > 8 ? 1 : 64)
    for (int i = 0; i < C; ++i)
#pragma unroll (C > 8 ? 1 : 64)
        for (int j = 0; j <= i; ++j) {
            double s = 0.0;
#pragma unroll (C > 8 ? 1 : 64)
            for (int t = i; t < C; ++t) s += li[FN_TRI(t, i)] * li[FN_TRI(t, j)];
            ainv[FN_TRI(i, j)] = s;
        }
}

// Weighted normal-equation accumulators for one gene (one lane's share until reduce()).
template <int C>
struct Normal {
    double a[Tri<C>::N];
    double b[C];
    double yy;
    int k;
    FN_HD void init() {
#pragma unroll
        for (int i = 0; i < Tri<C>::N; ++i) a[i] = 0.0;
#pragma unroll
        for (int i = 0; i < C; ++i) b[i] = 0.0;
        yy = 0.0; k = 0;
    }
    FN_HD void add(double w, double y, const double* x) {
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const double wx = w * x[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) a[FN_TRI(i, j)] += wx * x[j];
            b[i] += wx * y;
        }
        yy += (w * y) * y;
    }
    FN_HD void reduce(int lpg) {
        if (lpg == 1) return;
#ifdef __CUDA_ARCH__
        for (int off = lpg >> 1; off > 0; off >>= 1) {
#pragma unroll
            for (int i = 0; i < Tri<C>::N; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], off);
#pragma unroll
            for (int i = 0; i < C; ++i) b[i] += __shfl_xor_sync(0xffffffffu, b[i], off);
            yy += __shfl_xor_sync(0xffffffffu, yy, off);
            k += __shfl_xor_sync(0xffffffffu, k, off);
        }
#endif
    }
    // columns c_real..C-1 of the design are zero padding: put 1 on their diagonal so that the
    // factorisation goes through with det and beta unchanged
    FN_HD void pad(int c_real) {
#pragma unroll
        for (int i = 0; i < C; ++i) if (i >= c_real) a[FN_TRI(i, i)] = 1.0;
    }
};

// Row stride of a design table in shared memory: the dense stride C.  Rows of an even width are then
// 16-byte aligned and a row is read with C/2 LDS.128; with the bank rotation of GeneView the quarter-
// warps of such a read hit distinct 16-byte units (3 j mod 8 is a permutation for C = 6, j mod 8 for
// C = 2).  (An odd stride removes the conflicts of LDS.64 reads but doubles the number of load
// instructions -- measured slower at c = 6 on B200: 67 vs 48 us, 200k x 48.)
template <int C> struct XStride {
    enum { V = C };
};

template <int C>
FN_HD void load_x(const double* X, int j, double* x) {
#pragma unroll
    for (int i = 0; i < C; ++i) x[i] = X[j * XStride<C>::V + i];
}

// ------------------------------------------------------------------------------------------
// per-sample variance model
// ------------------------------------------------------------------------------------------
// Per-evaluation tables in shared memory (device) / plain arrays (host), all of length m:
//   KIND_SCALE_PS: ta[j] = 1/theta_j, tb[j] = ln theta_j
//   KIND_PATSEQ:   ta[j] = theta_a(j), tb[j] = theta_b(j)
struct EvalConst {
    int kind, tail_own, v3;
    int is_t;             // multivariate t (else normal)
    double v;             // df
    double theta0;        // SCALE1 shared variance (1 when na == 0)
    double itheta0, ltheta0;   // 1/theta0, ln theta0
    double iv;            // 1/v
    FN_HD void finish_setup() { itheta0 = 1.0 / theta0; ltheta0 = log(theta0); iv = is_t ? 1.0 / v : 0.0; }
    const double* ta;
    const double* tb;
    const double* t0;     // t0[p]: -(log-density terms that depend on (v,p) only)
    const double* t1;     // t1[p]: their derivative with respect to v (t only)
};

// (v,p)-only terms of -log density (distributions.py:139-146 t / :44-46 normal), and d/dv.
FN_HD void density_tables(int is_t, double v, int p, double& t0, double& t1) {
    if (is_t) {
        double lg_hp, dg_hp, lg_v, dg_v;
        lgamma_digamma(0.5 * (v + p), lg_hp, dg_hp);
        lgamma_digamma(0.5 * v, lg_v, dg_v);
        t0 = -(lg_hp - lg_v - (0.5 * p) * log(3.14159265358979323846 * v));
        t1 = -0.5 * dg_hp + 0.5 * dg_v + 0.5 * p / v;
    } else {
        t0 = 0.5 * p * 1.8378770664093454836;      // p/2 ln(2 pi)
        t1 = 0.0;
    }
}

// Weight (1/sigma2) and validity of sample j.  U and T are the partials of sigma2 with respect to
// theta_a(j) and theta_b(j) (PATSEQ) -- for the SCALE kinds the gradient is formed without them.
struct SampleVar {
    double w;        // 1/sigma2, or 0 when dropped
    double s2;       // sigma2 (PATSEQ only)
    double U, T;     // PATSEQ only
    bool valid;
};

FN_HD SampleVar sample_var(const EvalConst& ec, const GeneView& gv, int j, double y) {
    SampleVar sv;
    sv.U = 0.0; sv.T = 0.0; sv.s2 = 1.0;
    const bool yok = finite_d(y);
    if (ec.kind == KIND_PATSEQ) {
        const double rc = gv.aux[j];
        sv.U = ec.v3 ? rc * gv.gs : rc;
        sv.T = ec.tail_own ? (yok ? y * y : 0.0) : gv.gs;
        sv.s2 = ec.ta[j] * sv.U + ec.tb[j] * sv.T;
        sv.valid = yok && finite_d(sv.s2);          // Mv*.good: distributions.py:31-36
        sv.w = sv.valid ? 1.0 / sv.s2 : 0.0;
    } else {
        // sigma2 = theta / w_prec : finite unless w_prec is 0 or NaN  (aux = 1/weights, fitnoise.py:540)
        const double wp = gv.aux ? gv.aux[j] : 1.0;
        sv.valid = yok && (wp != 0.0) && (wp == wp);
        const double it = (ec.kind == KIND_SCALE_PS) ? ec.ta[j] : 1.0;   // SCALE1 works at theta = 1
        sv.w = sv.valid ? wp * it : 0.0;
    }
    return sv;
}

// ------------------------------------------------------------------------------------------
// K1: theta-independent preparation of one gene (fitnoise.py:23-35 restated without the QR)
// ------------------------------------------------------------------------------------------
struct PrepOut {
    int k;              // retained samples
    bool eligible;      // k > c and the retained sub-design has full column rank
    double cst;         // theta-independent part of ln det S:  -ln det(X_ret'X_ret) [- sum_ret ln w_prec]
    double average;     // mean(y[retain]) when k > c, else NaN       (fitnoise.py:267)
    double ysum, yss;   // over finite y: sum and sum of squares about the row mean  (initial variance)
    int nfinite;
    int bad_counts;     // PATSEQ: samples with count == 0 but finite y   (fitnoise.py:647)
    double avg_tail;    // PATSEQ v2/v3: avgtail (not squared)            (fitnoise.py:747,853)
};

#define FN_RANK_TOL 1e-10

// `counts_raw`: for PATSEQ the aux row still holds RAW counts at prepare time.
template <int C>
FN_HD void prepare_gene(const Family& fam, const GeneView& gv, const double* X, int c_real, PrepOut& out) {
    LogProd lw; lw.init();
    double ysum = 0.0, csum = 0.0, ycsum = 0.0, ysel = 0.0;
    int nfin = 0, bad = 0, k = 0;
    const bool patseq = fam.kind == KIND_PATSEQ;
    // whether sample j is retained (theta-independent part of Mv*.good)
    auto retained = [&](double y, double aux) -> bool {
        if (patseq) return finite_d(y) && finite_d(aux);
        return finite_d(y) && aux != 0.0 && aux == aux;
    };
    // Unweighted, not PAT-Seq (the streaming case): ONE unrolled sweep gives the row sum and the squared
    // deviations about a shift taken from the row itself (its first value: a point of the sample, so
    // s2 - s1^2/m loses no more digits than the two-pass form has to spare); a non-finite value anywhere
    // makes s2 non-finite, and only such rows take the general two sweeps below.
    bool fast = false;
    double fast_sum = 0.0, fast_ss = 0.0;
    if (!gv.aux && !patseq) {
        const double y0 = gv.y[0];
        const double shift = finite_d(y0) ? y0 : 0.0;
        double s12[2] = {0.0, 0.0};
        gv.for_each_hot(true, [&](int j) {
            const double d = gv.y[j] - shift;
            s12[0] += d;
            s12[1] = fma(d, d, s12[1]);
        });
        lane_sum_n<2>(s12, gv.lpg);
        fast = finite_d(s12[1]);
        fast_sum = fma((double)gv.m, shift, s12[0]);
        fast_ss = s12[1] - s12[0] * s12[0] / (double)gv.m;
        if (fast_ss < 0.0) fast_ss = 0.0;
    }
    if (!warp_any(!fast)) {                      // every gene of this warp is complete
        out.k = gv.m;
        out.eligible = gv.m > c_real;
        out.cst = 0.0;
        out.average = (gv.m > c_real) ? fast_sum / gv.m : NAN;
        out.ysum = fast_sum; out.yss = fast_ss; out.nfinite = gv.m;
        out.bad_counts = 0;
        out.avg_tail = 0.0;
        return;
    }
    gv.for_each(true, [&](int j) {
        const double y = gv.y[j];
        const bool yok = finite_d(y);
        const double aux = gv.aux ? gv.aux[j] : 1.0;
        const bool valid = retained(y, aux);
        if (patseq) {
            if (aux == 0.0 && yok) ++bad;
            csum += aux;
            ycsum += (yok ? y : 0.0) * aux;
        } else if (valid && gv.aux) {
            lw.mul(fabs(aux));
        }
        if (yok) { ysum += y; ++nfin; }
        if (valid) { ysel += y; ++k; }
    });
    lw.lane_reduce(gv.lpg);
    ysum = lane_sum(ysum, gv.lpg);
    ysel = lane_sum(ysel, gv.lpg);
    nfin = lane_sum_i(nfin, gv.lpg);
    bad = lane_sum_i(bad, gv.lpg);
    k = lane_sum_i(k, gv.lpg);
    if (patseq) { csum = lane_sum(csum, gv.lpg); ycsum = lane_sum(ycsum, gv.lpg); }
    // The design columns are orthonormal over all m samples (Design::build), so a gene that keeps every
    // sample has X_ret'X_ret = I: determinant 1, full rank.  Only genes that lost a sample need the masked
    // Gram matrix and its factorisation (none at all in a data set without missing values).
    double det = 1.0;
    bool full_rank = true;
    const bool partial = k < gv.m;
    if (warp_any(partial)) {
        Normal<C> g;
        g.init();
        gv.for_each(partial, [&](int j) {
            if (retained(gv.y[j], gv.aux ? gv.aux[j] : 1.0)) {
                double x[C];
                load_x<C>(X, j, x);
                g.add(1.0, 0.0, x);
            }
        });
        g.reduce(gv.lpg);
        g.pad(c_real);
        double det2;
        const bool rank2 = chol<C>(g.a, det2, FN_RANK_TOL);
        if (partial) { det = det2; full_rank = rank2; }
    }
    // second sweep: squared deviations of the finite y about their mean (numpy.var, fitnoise.py:546)
    const double mean = nfin > 0 ? ysum / nfin : 0.0;
    double ss = 0.0;
    gv.for_each(true, [&](int j) {
        const double y = gv.y[j];
        if (finite_d(y)) { const double d = y - mean; ss += d * d; }
    });
    ss = lane_sum(ss, gv.lpg);

    out.k = k;
    out.eligible = (k > c_real) && full_rank;
    out.cst = (det == 1.0) ? 0.0 : -log(det);
    if (!patseq && gv.aux) out.cst -= lw.log_value();
    out.average = (k > c_real) ? ysel / k : NAN;
    out.ysum = ysum; out.yss = ss; out.nfinite = nfin;
    out.bad_counts = bad;
    out.avg_tail = 0.0;
    if (patseq) {
        double num = ycsum;
        if (!fam.v3) num = fmax(1.0, num);
        out.avg_tail = num / fmax(1.0, csum);
    }
}

// ------------------------------------------------------------------------------------------
// K2: -log likelihood and gradient of one gene
// ------------------------------------------------------------------------------------------
struct EvalOut {
    double value;       // -ll (without the theta-independent cst/2 unless cst was passed)
    double dv;          // d/d df
    double ga, gb;      // d/d theta_a, d/d theta_b when they are shared parameters
    double d2;          // Mahalanobis distance z2'S^-1 z2
    int p;              // residual df k - c
    int used;           // 1 when the gene contributed
};

FN_HD void eval_zero(EvalOut& o) { o.value = 0.0; o.dv = 0.0; o.ga = 0.0; o.gb = 0.0; o.d2 = NAN; o.p = 0; o.used = 0; }

// value, dv and kappa from (p, ln det S, d2)
// The terms that depend on (v, p) only come from the tables ec.t0 / ec.t1 when the caller wants
// complete per-gene values.  The sum over genes does not need them per gene: sum_g t0[p_g] =
// sum_p N_p t0[p] with N_p the theta-independent histogram of residual df that the preparation
// kernel counts once per fit, so the evaluation kernels run with ec.t0 == nullptr and the C-ABI
// layer adds the few table terms on the host (fn_api.cu, table_terms).
FN_HD double finish_density(const EvalConst& ec, int p, double logdet, double d2, double& dv, double& kappa) {
    const double t0 = ec.t0 ? ec.t0[p] : 0.0;
    if (ec.is_t) {
        const double v = ec.v;
        const double l1p = log1p(d2 * ec.iv);
        const double hp = 0.5 * (v + p);
        kappa = hp / (v + d2);
        dv = (ec.t1 ? ec.t1[p] : 0.0) + 0.5 * l1p - kappa * d2 * ec.iv;
        return t0 + 0.5 * logdet + hp * l1p;
    }
    kappa = 0.5;
    dv = 0.0;
    return t0 + 0.5 * (logdet + d2);
}

// KIND_SCALE1: sigma2_j = theta0 / w_prec_j.  All sums are taken at theta = 1 and rescaled;
// the gradient with respect to theta0 is closed-form (no second sweep):
//   d(-ll)/d theta0 = (p/2 - kappa d2) / theta0.
//
// X must have ORTHONORMAL columns over the m samples (Design::build whitens every design).  For a
// gene with no weights and no missing value the normal matrix is then the identity, so the sweep
// needs only b = X'y and y'y (C+1 FMAs per sample): beta = b, det A = 1, d2 = y'y - |b|^2.  A
// missing value makes y'y non-finite, which is how incomplete rows are found at no per-sample
// cost; they (and every weighted gene) take the masked sweep that also builds A.
struct Scale1Mid {
    double d2;          // Mahalanobis distance at theta = 1
    double det_a;       // det(X'WX) at theta = 1 (1 for a complete unweighted gene); its log is taken by scale1_finish
    int p;              // residual df
    int state;          // 0: gene not in the REML sum, 1: usable, 2: failed (non-finite)
};

// First half: everything that needs the gene's row (sums, solve, d2).  Cheap in the common case.
// LEAN: the caller guarantees that every eligible gene is complete and unweighted (the preparation
// kernel counted none that is not), so the masked sweep is not even compiled into the kernel -- its
// registers are what limits the streaming kernel to one CTA per SM at c >= 6.  A gene that turns out
// incomplete all the same is reported as failed (state 2), never silently mis-evaluated.
template <int C, bool LEAN = false>
FN_HD void scale1_sweep(const EvalConst& ec, const GeneView& gv, const double* X, int c_real,
                        bool eligible, Scale1Mid& mid) {
    const bool has_aux = gv.aux != nullptr;
    double b[C], beta[C];
    double yy = 0.0, det = 1.0;
    int k = gv.m;
    bool ok = true, complete = false;
#pragma unroll
    for (int i = 0; i < C; ++i) b[i] = 0.0;
    if (!has_aux) {
        gv.for_each_hot(eligible, [&](int j) {
            const double y = gv.y[j];
            double x[C];
            load_x<C>(X, j, x);
#pragma unroll
            for (int i = 0; i < C; ++i) b[i] += x[i] * y;
            yy += y * y;
        });
        if (gv.lpg > 1) {
            double sums[C + 1];
#pragma unroll
            for (int i = 0; i < C; ++i) sums[i] = b[i];
            sums[C] = yy;
            lane_sum_n<C + 1>(sums, gv.lpg);
#pragma unroll
            for (int i = 0; i < C; ++i) b[i] = sums[i];
            yy = sums[C];
        }
        complete = finite_d(yy);
#pragma unroll
        for (int i = 0; i < C; ++i) beta[i] = b[i];
    }
    if (LEAN) {
        ok = complete;
    } else if (has_aux || warp_any(eligible && !complete)) {
        Normal<C> nm;
        nm.init();
        // sigma2 = theta / w_prec at theta = 1: a sample counts when y is finite and its precision
        // weight is neither 0 nor NaN (aux = 1/weights, fitnoise.py:540; Mv*.good)
        auto visit = [&](int j) {
            const double y = gv.y[j];
            const double wp = has_aux ? gv.aux[j] : 1.0;
            const bool valid = finite_d(y) && wp != 0.0 && wp == wp;
            double x[C];
            load_x<C>(X, j, x);
            nm.add(valid ? wp : 0.0, valid ? y : 0.0, x);
            nm.k += valid ? 1 : 0;
        };
        if (C <= 4) gv.for_each_hot(eligible && !complete, visit);
        else gv.for_each(eligible && !complete, visit);
        nm.reduce(gv.lpg);
        nm.pad(c_real);
        double det2;
        const bool ok2 = chol<C, true>(nm.a, det2, 0.0);
        double beta2[C];
#pragma unroll
        for (int i = 0; i < C; ++i) beta2[i] = nm.b[i];
        chol_solve<C>(nm.a, beta2);
        if (!complete) {
#pragma unroll
            for (int i = 0; i < C; ++i) { b[i] = nm.b[i]; beta[i] = beta2[i]; }
            yy = nm.yy; k = nm.k; det = det2; ok = ok2;
        }
    }
    double d2 = yy;
#pragma unroll
    for (int i = 0; i < C; ++i) d2 -= b[i] * beta[i];
    // y'Wy - b'beta cancels badly when the fit explains almost everything: redo from residuals
    if (warp_any(eligible && d2 < 1e-4 * yy)) {
        double s = 0.0;
        gv.for_each(eligible && d2 < 1e-4 * yy, [&](int j) {
            const double y = gv.y[j];
            const SampleVar sv = sample_var(ec, gv, j, y);
            double x[C];
            load_x<C>(X, j, x);
            double r = sv.valid ? y : 0.0;
#pragma unroll
            for (int t = 0; t < C; ++t) r -= x[t] * beta[t];
            s += sv.w * r * r;
        });
        s = lane_sum(s, gv.lpg);
        if (d2 < 1e-4 * yy) d2 = s;
    }
    mid.p = k - c_real;
    mid.d2 = d2 < 0.0 ? 0.0 : d2;
    mid.det_a = complete ? 1.0 : det;
    mid.state = !eligible ? 0 : ((ok && mid.p > 0 && d2 == d2) ? 1 : 2);
}

// Second half: the density and its derivatives from (p, d2); needs no data.
FN_HD void scale1_finish(const EvalConst& ec, const Scale1Mid& mid, double cst, EvalOut& out) {
    eval_zero(out);
    if (mid.state == 0) return;
    out.used = 1;
    if (mid.state == 2) { out.value = NAN; return; }
    const int p = mid.p;
    const double d2 = mid.d2 * ec.itheta0;
    const double logdet = p * ec.ltheta0 + (mid.det_a == 1.0 ? 0.0 : log(mid.det_a)) + cst;
    double dv, kappa;
    out.value = finish_density(ec, p, logdet, d2, dv, kappa);
    out.dv = dv;
    out.ga = (0.5 * p - kappa * d2) * ec.itheta0;
    out.d2 = d2; out.p = p;
}

template <int C>
FN_HD void eval_scale1(const EvalConst& ec, const GeneView& gv, const double* X, int c_real,
                       bool eligible, double cst, EvalOut& out) {
    Scale1Mid mid;
    scale1_sweep<C>(ec, gv, X, c_real, eligible, mid);
    scale1_finish(ec, mid, cst, out);
}

// KIND_SCALE_PS / KIND_PATSEQ: general diagonal variance.  Two sweeps over the gene's row in shared
// memory:
//   1. normal equations, sum_ret ln sigma2_j, k.  PATSEQ pays its one division per sample here and
//      leaves w_j = 1/sigma2_j (0 for a dropped sample) in the sample's slot of the aux tile, so the
//      second sweep needs neither a division nor rc_j:  U_j w_j = (1 - theta_b(j) T_j w_j) / theta_a(j).
//   2. per-sample gradient terms  e~_j = 1/2 (1 - w_j h_j) - kappa w_j r_j^2  (kappa needs d2, which
//      comes from the first sweep as y'Wy - b'beta, re-derived from residuals if that cancelled).
// PATSEQ returns term a multiplied by theta_a(j) (saves a division per sample); whoever assembles
// the gradient divides block a by theta_a.
// When a theta block is per-sample (inplace_a / inplace_b) the per-sample gradient contribution is
// written over the sample's own slot in the y (term a) / aux (term b) tile for the caller to
// column-sum; otherwise it is accumulated into out.ga / out.gb.
// `write_slots` = false: this lane's gene is handled elsewhere (complete-row fast path): leave its slots alone.
template <int C, int KIND>
FN_HD void eval_general(const EvalConst& ec, GeneView& gv, const double* X, int c_real,
                        bool eligible, double cst, bool inplace_a, bool inplace_b, EvalOut& out,
                        bool write_slots = true) {
    Normal<C> nm;
    nm.init();
    LogProd lp; lp.init();
    double pend = 1.0;                 // PATSEQ: product of up to 4 variances between renormalisations
    int npend = 0;
    double slt = 0.0;                  // SCALE_PS: sum_ret ln theta_j
    gv.for_each(eligible, [&](int j) {
        const double y = gv.y[j];
        const SampleVar sv = sample_var(ec, gv, j, y);
        double x[C];
        load_x<C>(X, j, x);
        nm.add(sv.w, sv.valid ? y : 0.0, x);
        if (KIND == KIND_PATSEQ) gv.aux[j] = sv.w;
        if (sv.valid) {
            ++nm.k;
            if (KIND == KIND_PATSEQ) {
                pend *= sv.s2;
                if (++npend == 4) { lp.mul(pend); pend = 1.0; npend = 0; }
            } else {
                slt += ec.tb[j];
            }
        }
    });
    if (KIND == KIND_PATSEQ) { lp.mul(pend); lp.lane_reduce(gv.lpg); }
    else slt = lane_sum(slt, gv.lpg);
    nm.reduce(gv.lpg);
    nm.pad(c_real);
    double det;
    const bool ok = chol<C, true>(nm.a, det, 0.0);
    double beta[C];
#pragma unroll
    for (int i = 0; i < C; ++i) beta[i] = nm.b[i];
    chol_solve<C>(nm.a, beta);
    double d2 = nm.yy;
#pragma unroll
    for (int i = 0; i < C; ++i) d2 -= nm.b[i] * beta[i];

    // the weight of sample j as left by the first sweep
    auto weight = [&](int j, double y, bool& valid) -> double {
        if (KIND == KIND_PATSEQ) { const double w = gv.aux[j]; valid = w != 0.0; return w; }
        const SampleVar sv = sample_var(ec, gv, j, y);
        valid = sv.valid;
        return sv.w;
    };
    if (warp_any(eligible && d2 < 1e-4 * nm.yy)) {          // cancellation: d2 from residuals
        double s = 0.0;
        gv.for_each(eligible && d2 < 1e-4 * nm.yy, [&](int j) {
            const double y = gv.y[j];
            bool valid;
            const double w = weight(j, y, valid);
            double x[C];
            load_x<C>(X, j, x);
            double r = valid ? y : 0.0;
#pragma unroll
            for (int t = 0; t < C; ++t) r -= x[t] * beta[t];
            s += w * r * r;
        });
        s = lane_sum(s, gv.lpg);
        if (d2 < 1e-4 * nm.yy) d2 = s;
    }
    if (d2 < 0.0) d2 = 0.0;

    eval_zero(out);
    const int p = nm.k - c_real;
    const bool good = eligible && ok && p > 0 && d2 == d2;
    double kappa = 0.5;
    if (good) {
        const double lsig = (KIND == KIND_PATSEQ) ? lp.log_value() : slt;
        const double logdet = lsig + log(det) + cst;
        double dv;
        out.value = finish_density(ec, p, logdet, d2, dv, kappa);
        out.dv = dv; out.d2 = d2; out.p = p; out.used = 1;
        if (!(out.value == out.value)) { out.dv = 0.0; }
    } else if (eligible) {
        out.value = NAN; out.used = 1;
    }
    const bool grad_ok = good && out.value == out.value;

    // second sweep: gradient
    double ga = 0.0, gb = 0.0;
    gv.for_each(grad_ok || (write_slots && (inplace_a || (KIND == KIND_PATSEQ && inplace_b))), [&](int j) {
        const double y = gv.y[j];
        double ca = 0.0, cb = 0.0;
        if (grad_ok) {
            bool valid;
            const double w = weight(j, y, valid);
            if (valid) {
                double x[C];
                load_x<C>(X, j, x);
                double r = y;
#pragma unroll
                for (int t = 0; t < C; ++t) r -= x[t] * beta[t];
                const double h = leverage<C>(nm.a, x);
                const double et = 0.5 * (1.0 - w * h) - kappa * w * r * r;
                if (KIND == KIND_PATSEQ) {
                    const double T = ec.tail_own ? y * y : gv.gs;
                    const double tw = T * w;
                    cb = tw * et;                                         // T_j e_j
                    ca = (1.0 - ec.tb[j] * tw) * et;                      // theta_a(j) U_j e_j: the caller divides
                } else {
                    ca = et * ec.ta[j];          // d sigma2_j / d theta_j = sigma2_j / theta_j
                }
            }
        }
        if (inplace_a) gv.y[j] = ca; else ga += ca;
        if (KIND == KIND_PATSEQ) { if (inplace_b) gv.aux[j] = cb; else gb += cb; }
    });
    if (!inplace_a) out.ga = lane_sum(ga, gv.lpg);
    if (KIND == KIND_PATSEQ && !inplace_b) out.gb = lane_sum(gb, gv.lpg);
    // every lane of the gene now holds the same totals: only lane 0 may add them to the CTA sums
}

// ------------------------------------------------------------------------------------------
// KIND_SCALE_PS without observation weights: genes with no missing value share their normal matrix
// ------------------------------------------------------------------------------------------
// sigma2_j = theta_j for every gene, so for a COMPLETE gene  A = X'WX  (W = diag(1/theta_j)) does not
// depend on the gene at all.  Per evaluation and per design the CTA tabulates once
//   hx_j = w_j A^-1 x_j   (m x C),    cj_j = w_j/2 (1 - w_j x_j'A^-1 x_j)   (m),    ln det A,  sum_j ln theta_j
// and a complete gene then costs:  beta = sum_j hx_j y_j  (C FMAs per sample),  r_j = y_j - x_j'beta,
// d2 = sum_j w_j r_j^2  -- no normal equations, no Cholesky, no leverages.  Its per-sample gradient is
//   d(-ll)/d theta_j = cj_j - kappa_g (w_j r_j)^2:
// the gene leaves s_j = (w_j r_j)^2 in its y slot and kappa_g aside, the caller's column sum applies
// -kappa_g and adds (number of such genes) x cj_j at the end.  Genes with missing values take
// eval_general.  (200k x 48, c = 6: 0.19 ms -> see DESIGN.md.)
template <int C>
struct PsDesign {
    const double* hx;      // m x C
    const double* cj;      // m
    double logdet_a, slt;  // ln det A, sum_j ln theta_j
    bool ok;               // A positive definite
};

// one row of the tables: u = A^-1 x_j (ainv packed lower), w = 1/theta_j
template <int C>
FN_HD void ps_table_row(const double* ainv, const double* x, double w, double* hx_row, double& cj) {
    double h = 0.0;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        double u = 0.0;
#pragma unroll
        for (int q = 0; q < C; ++q) u += ainv[q > i ? FN_TRI(q, i) : FN_TRI(i, q)] * x[q];
        hx_row[i] = w * u;
        h += x[i] * u;
    }
    cj = 0.5 * w * (1.0 - w * h);
}

// First sweep: beta (whitened basis); `complete` comes back false when the row holds a non-finite y
// (0 x NaN and 0 x inf are NaN, so any such y poisons every beta_i).
template <int C>
FN_HD bool ps_complete_beta(const PsDesign<C>& pd, const GeneView& gv, bool on, double* beta) {
#pragma unroll
    for (int i = 0; i < C; ++i) beta[i] = 0.0;
    gv.for_each_hot(on, [&](int j) {
        const double y = gv.y[j];
        double hx[C];
        load_x<C>(pd.hx, j, hx);
#pragma unroll
        for (int i = 0; i < C; ++i) beta[i] += hx[i] * y;
    });
    lane_sum_n<C>(beta, gv.lpg);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += beta[i];
    return on && finite_d(s);
}

// Second sweep of a complete gene: residuals, d2, s_j = (w_j r_j)^2 into the y slots; then the density.
// `mine` = false: the gene is not handled here (not in the REML sum: its slots are zeroed when `zero`).
template <int C>
FN_HD void ps_complete_finish(const EvalConst& ec, const PsDesign<C>& pd, GeneView& gv, const double* X, int c_real,
                              bool mine, bool zero, const double* beta, double cst, EvalOut& out, double& kappa_out) {
    double d2 = 0.0;
    gv.for_each_hot(mine || zero, [&](int j) {
        double s = 0.0;
        if (mine) {
            const double y = gv.y[j], w = ec.ta[j];
            double x[C];
            load_x<C>(X, j, x);
            double r = y;
#pragma unroll
            for (int t = 0; t < C; ++t) r -= x[t] * beta[t];
            const double wr = w * r;
            d2 += wr * r;
            s = wr * wr;
        }
        gv.y[j] = s;
    });
    kappa_out = 0.0;
    d2 = lane_sum(d2, gv.lpg);          // every lane of the warp takes part (full-mask shuffles), whatever its gene does
    if (!mine) return;
    eval_zero(out);
    out.used = 1;
    const int p = gv.m - c_real;
    if (!(pd.ok && p > 0 && d2 == d2)) { out.value = NAN; return; }
    double dv, kappa;
    out.value = finish_density(ec, p, pd.slt + pd.logdet_a + cst, d2, dv, kappa);
    out.dv = dv; out.d2 = d2; out.p = p;
    if (out.value == out.value) kappa_out = kappa;
}

// ------------------------------------------------------------------------------------------
// The same complete-row fast path for the THREAD-PER-GENE layout with padded rows (fn_tile.cuh,
// TileSrc::row_stride): every thread of a warp walks its own gene's row through the SAME samples in
// the same order, two samples per step.  The pair (y_t, y_t+1) is one 16-byte load per thread (rows are
// 16-byte aligned and start in distinct bank groups), and the table rows hx_t, x_t, w_t are the same
// address for all 32 threads: broadcast reads.  The lanes-per-gene mapping of ps_complete_* reads a
// different table row in every thread instead (14 shared-memory wavefronts per warp and sample at
// c = 6); here a sample costs about 5 + 7.  No shuffles: each thread holds its gene's sums alone.
// ------------------------------------------------------------------------------------------
FN_HD void load_pair(const double* p, double& a, double& b) {
#ifdef __CUDA_ARCH__
    const double2 v = *reinterpret_cast<const double2*>(p);
    a = v.x; b = v.y;
#else
    a = p[0]; b = p[1];
#endif
}
FN_HD void store_pair(double* p, double a, double b) {
#ifdef __CUDA_ARCH__
    *reinterpret_cast<double2*>(p) = make_double2(a, b);
#else
    p[0] = a; p[1] = b;
#endif
}

// Row j of a design-shaped table whose rows are 2C doubles apart: the rows layout keeps the noise design's
// and the control design's tables INTERLEAVED (row j = [noise_j | control_j]), so that a warp whose genes
// are a mix of both reads two 16-byte-aligned pieces of one 16C-byte span (distinct banks) instead of two
// addresses a multiple of 128 bytes apart (the same banks: a two-way conflict on every table read).
template <int C>
FN_HD void load_x2(const double* X, int j, double* x) {
    const double* row = X + j * (2 * C);
#ifdef __CUDA_ARCH__
    if (C % 2 == 0) row = (const double*)__builtin_assume_aligned(row, 16);    // table base and both halves of a row are 16-byte aligned
#endif
#pragma unroll
    for (int i = 0; i < C; ++i) x[i] = row[i];
}

// beta = sum_j hx_j y_j; false when the row holds a non-finite y (m even)
template <int C>
FN_HD bool ps_rows_beta(const PsDesign<C>& pd, const double* yrow, int m, bool on, double* beta) {
#pragma unroll
    for (int i = 0; i < C; ++i) beta[i] = 0.0;
    if (!on) return false;
#pragma unroll 2
    for (int t = 0; t < m; t += 2) {
        double y0, y1;
        load_pair(yrow + t, y0, y1);
        double h0[C], h1[C];
        load_x2<C>(pd.hx, t, h0);
        load_x2<C>(pd.hx, t + 1, h1);
#pragma unroll
        for (int i = 0; i < C; ++i) beta[i] = fma(h0[i], y0, beta[i]);
#pragma unroll
        for (int i = 0; i < C; ++i) beta[i] = fma(h1[i], y1, beta[i]);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += beta[i];
    return finite_d(s);
}

// residuals, d2, s_j = (w_j r_j)^2 into the row's slots (zeros when `zero`), then the density
template <int C>
FN_HD void ps_rows_finish(const EvalConst& ec, const PsDesign<C>& pd, double* yrow, int m, const double* X, int c_real,
                          bool mine, bool zero, const double* beta, double cst, EvalOut& out, double& kappa_out) {
    kappa_out = 0.0;
    if (!mine) {
        if (zero)
            for (int t = 0; t < m; t += 2) store_pair(yrow + t, 0.0, 0.0);
        return;
    }
    double d2 = 0.0;
#pragma unroll 2
    for (int t = 0; t < m; t += 2) {
        double y0, y1, w0, w1;
        load_pair(yrow + t, y0, y1);
        load_pair(ec.ta + t, w0, w1);
        double x0[C], x1[C];
        load_x2<C>(X, t, x0);
        load_x2<C>(X, t + 1, x1);
        double r0 = y0, r1 = y1;
#pragma unroll
        for (int i = 0; i < C; ++i) { r0 = fma(-x0[i], beta[i], r0); r1 = fma(-x1[i], beta[i], r1); }
        const double wr0 = w0 * r0, wr1 = w1 * r1;
        d2 = fma(wr0, r0, d2);
        d2 = fma(wr1, r1, d2);
        store_pair(yrow + t, wr0 * wr0, wr1 * wr1);
    }
    eval_zero(out);
    out.used = 1;
    const int p = m - c_real;
    if (!(pd.ok && p > 0 && d2 == d2)) { out.value = NAN; return; }
    double dv, kappa;
    out.value = finish_density(ec, p, pd.slt + pd.logdet_a + cst, d2, dv, kappa);
    out.dv = dv; out.d2 = d2; out.p = p;
    if (out.value == out.value) kappa_out = kappa;
}

// 1/x for a positive normal double.  Device: hardware seed (>= 20 bits) and two Newton steps, i.e. one
// MUFU and four FMAs instead of the ~12-instruction correctly rounded division; the result is within
// an ulp, and it only ever feeds sums over samples and genes.
FN_HD double rcp_pos(double x) {
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

// KIND_PATSEQ with ONE shared theta_a and ONE shared theta_b (Model_*_patseq_v1 / _v2 / _v3 without
// per-sample parameters; fitnoise.py:659-678, 757-776, 863-882): value AND gradient from a single sweep.
//
// The second sweep of eval_general exists to form  G_b = sum_j T_j e_j,  e_j = d(-ll)/d sigma2_j =
// 1/2 (w_j - w_j^2 h_j) - kappa (w_j r_j)^2, which needs beta, A^-1 and kappa.  Expanded in moments that
// the first sweep can accumulate next to the normal equations (t_j = T_j w_j^2):
//   sum_j T_j w_j^2 h_j   = tr(A^-1 M_T),                M_T = sum_j t_j x_j x_j'
//   sum_j T_j w_j^2 r_j^2 = S_Tyy - 2 beta'b_T + beta'M_T beta,   b_T = sum_j t_j x_j y_j,  S_Tyy = sum_j t_j y_j^2
// so  G_b = 1/2 sum_j T_j w_j - 1/2 tr(A^-1 M_T) - kappa (S_Tyy - 2 beta'b_T + beta'M_T beta),  and because
// theta_a U_j + theta_b T_j = sigma2_j:
//   theta_a G_a + theta_b G_b = sum_j sigma2_j e_j = 1/2 (k - c) - kappa d2,
// the theta_a gradient needs no moments of its own.  One sweep, no per-sample leverage, and the staged
// tile is never written.  The quadratic form cancels like d2 = y'Wy - b'beta does; when that loses more
// than four digits both are recomputed from residuals (a second sweep for those genes only).
// Returns out.ga multiplied by theta_a, as eval_general does for this kind.
// The sweep: normal equations in nm, the T-weighted moments in mt / bt / sums, prod sigma2 in lp.
// TAIL_OWN (v1): T_j = y_j^2.  Otherwise T = gs for every sample of the gene, so it is factored out of
// the moments (they are accumulated with w_j^2 and multiplied by T once per gene) and
// sigma2_j = ua rc_j + tb0 with two per-gene scalars (v3 folds its avgtail^2 into ua).
// Written without a branch per sample: a dropped sample gets w = 0, y = 0, sigma2 = 1.
template <int C, bool TAIL_OWN>
FN_HD void patseq_shared_sweep(const EvalConst& ec, const GeneView& gv, const double* X, bool eligible,
                               double tha, double thb, Normal<C>& nm, double* mt, double* bt, double* sums, LogProd& lp) {
    const double ua = ec.v3 ? tha * gv.gs : tha;           // sigma2 = ua rc + thb T
    const double tb0 = TAIL_OWN ? 0.0 : thb * gv.gs;
    double pend = 1.0, sw = 0.0;
    gv.for_each_steps(eligible, [&](int j) {
        const double y = gv.y[j];
        const double rc = gv.aux[j];
        // rc is in (0, 1] or NaN and theta is finite: sigma2 is finite exactly when rc is, unless y^2 overflows
        const double T = TAIL_OWN ? y * y : 0.0;
        const double s2 = TAIL_OWN ? fma(ua, rc, thb * T) : fma(ua, rc, tb0);
        const bool valid = finite_d(TAIL_OWN ? s2 : y * rc);       // Mv*.good: distributions.py:31-36
        const double s2v = valid ? s2 : 1.0;
        const double w = valid ? rcp_pos(s2v) : 0.0;
        const double yv = valid ? y : 0.0;
        pend *= s2v;
        nm.k += valid ? 1 : 0;
        double x[C];
        load_x<C>(X, j, x);
        nm.add(w, yv, x);
        const double t2 = TAIL_OWN ? (valid ? T : 0.0) * (w * w) : w * w;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const double tx = t2 * x[i];
#pragma unroll
            for (int q = 0; q <= i; ++q) mt[FN_TRI(i, q)] += tx * x[q];
            bt[i] += tx * yv;
        }
        sums[0] += (t2 * yv) * yv;
        sw += TAIL_OWN ? t2 * s2v : w;                    // T_j w_j (T factored out below)
    }, [&]() { lp.mul(pend); pend = 1.0; });
    if (TAIL_OWN) {
        sums[1] = sw;
    } else {                                               // T = gs: put it back
        const double T = gv.gs;
#pragma unroll
        for (int i = 0; i < Tri<C>::N; ++i) mt[i] *= T;
#pragma unroll
        for (int i = 0; i < C; ++i) bt[i] *= T;
        sums[0] *= T;
        sums[1] = T * sw;
    }
}

template <int C>
FN_HD void eval_patseq_shared(const EvalConst& ec, const GeneView& gv, const double* X, int c_real,
                              bool eligible, double cst, EvalOut& out) {
    const double tha = ec.ta[0], thb = ec.tb[0];
    Normal<C> nm;
    nm.init();
    double mt[Tri<C>::N], bt[C];
#pragma unroll
    for (int i = 0; i < Tri<C>::N; ++i) mt[i] = 0.0;
#pragma unroll
    for (int i = 0; i < C; ++i) bt[i] = 0.0;
    double sums[2] = {0.0, 0.0};       // S_Tyy, sum_j T_j w_j
    LogProd lp; lp.init();
    if (ec.tail_own) patseq_shared_sweep<C, true>(ec, gv, X, eligible, tha, thb, nm, mt, bt, sums, lp);
    else patseq_shared_sweep<C, false>(ec, gv, X, eligible, tha, thb, nm, mt, bt, sums, lp);
    // sigma2, validity and the weight of sample j (the cancellation fallback below)
    auto variance = [&](int j, double y, double& T, bool& valid) -> double {
        const double rc = gv.aux[j];
        const bool yok = finite_d(y);
        const double U = ec.v3 ? rc * gv.gs : rc;
        T = ec.tail_own ? (yok ? y * y : 0.0) : gv.gs;
        const double s2 = tha * U + thb * T;
        valid = yok && finite_d(s2);
        return s2;
    };
    if (gv.lpg > 1) {
        lp.lane_reduce(gv.lpg);
        nm.reduce(gv.lpg);
        lane_sum_n<Tri<C>::N>(mt, gv.lpg);
        lane_sum_n<C>(bt, gv.lpg);
        lane_sum_n<2>(sums, gv.lpg);
    }
    nm.pad(c_real);
    double det;
    const bool ok = chol<C, true>(nm.a, det, 0.0);
    double beta[C];
#pragma unroll
    for (int i = 0; i < C; ++i) beta[i] = nm.b[i];
    chol_solve<C>(nm.a, beta);
    double d2 = nm.yy;
#pragma unroll
    for (int i = 0; i < C; ++i) d2 -= nm.b[i] * beta[i];
    // sum_j T_j w_j^2 r_j^2
    double quad = sums[0];
#pragma unroll
    for (int i = 0; i < C; ++i) {
        double mb = 0.0;
#pragma unroll
        for (int q = 0; q < C; ++q) mb += mt[q > i ? FN_TRI(q, i) : FN_TRI(i, q)] * beta[q];
        quad += beta[i] * (mb - 2.0 * bt[i]);
    }
    if (warp_any(eligible && d2 < 1e-4 * nm.yy)) {          // cancellation: both from residuals
        double r2[2] = {0.0, 0.0};
        gv.for_each(eligible && d2 < 1e-4 * nm.yy, [&](int j) {
            const double y = gv.y[j];
            double T;
            bool valid;
            const double s2 = variance(j, y, T, valid);
            if (valid) {
                const double w = 1.0 / s2;
                double x[C];
                load_x<C>(X, j, x);
                double r = y;
#pragma unroll
                for (int t = 0; t < C; ++t) r -= x[t] * beta[t];
                r2[0] += w * r * r;
                r2[1] += T * w * w * r * r;
            }
        });
        lane_sum_n<2>(r2, gv.lpg);
        if (d2 < 1e-4 * nm.yy) { d2 = r2[0]; quad = r2[1]; }
    }
    if (d2 < 0.0) d2 = 0.0;

    eval_zero(out);
    const int p = nm.k - c_real;
    const bool good = eligible && ok && p > 0 && d2 == d2;
    if (!good) {
        if (eligible) { out.value = NAN; out.used = 1; }
        return;
    }
    const double logdet = lp.log_value() + log(det) + cst;
    double dv, kappa;
    out.value = finish_density(ec, p, logdet, d2, dv, kappa);
    out.d2 = d2; out.p = p; out.used = 1;
    if (!(out.value == out.value)) return;
    out.dv = dv;
    // tr(A^-1 M_T) over the real columns (padded columns have M = 0)
    double ainv[Tri<C>::N];
    chol_inverse<C>(nm.a, ainv);
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        tr += ainv[FN_TRI(i, i)] * mt[FN_TRI(i, i)];
#pragma unroll
        for (int q = 0; q < i; ++q) tr += 2.0 * ainv[FN_TRI(i, q)] * mt[FN_TRI(i, q)];
    }
    out.gb = 0.5 * (sums[1] - tr) - kappa * quad;
    out.ga = 0.5 * p - kappa * d2 - thb * out.gb;          // theta_a G_a: the caller divides
}

// ------------------------------------------------------------------------------------------
// Factor models (Model_factors_mixin, fitnoise.py:439-518): covariance diag(sigma2) + F F'
// ------------------------------------------------------------------------------------------
// With Sigma = D + F F' (F: m x nf, shared by all genes, changes with theta) the REML quantities are
// those of the DIAGONAL model fitted with the augmented design Z = [X F] and a unit ridge on the
// factor block of the normal matrix (Henderson's mixed-model equations):
//   H = Z'WZ + diag(0_c, I_nf),  [beta; u] = H^-1 Z'Wy
//   d2 = y'Wy - (Z'Wy)'[beta; u] = sum_j w_j r~_j^2 + u'u            (r~ = y - X beta - F u)
//   ln det S = sum_ret ln sigma2_j + ln det H - ln det(X_ret'X_ret)
//   d(-ll)/d sigma2_j = 1/2 (w_j - w_j^2 z_j'H^-1 z_j) - kappa (w_j r~_j)^2
//   d(-ll)/d F[j,i]   = w_j z_j'(H^-1)_{:,c+i} - 2 kappa (w_j r~_j) u_i
// so the (k-c)x(k-c) dense covariance the reference inverts (plus_covar, distributions.py:74-76)
// never appears.  Z is m x C row-major with the design in columns [0,c), the factors in
// [c, c+nf) and zeros beyond.  Only the SCALE1 variance (Model_normal/_t/_independent) combines
// with factors in the reference (fitnoise.py:558-565, 592-594).
template <int C>
FN_HD void ridge_and_pad(double* a, int c_real, int nf) {
#pragma unroll
    for (int i = 0; i < C; ++i) {
        if (i >= c_real + nf) a[FN_TRI(i, i)] = 1.0;
        else if (i >= c_real) a[FN_TRI(i, i)] += 1.0;
    }
}

// `gscratch`: this gene's m x nf row block of the CTA's gradient scratch (written for every sample
// this lane owns: the dF contribution of the gene, zeros when the gene does not contribute).
template <int C>
FN_HD void eval_factors(const EvalConst& ec, const GeneView& gv, const double* Z, int c_real, int nf,
                        bool eligible, double cst, double* gscratch, EvalOut& out) {
    Normal<C> nm;
    nm.init();
    gv.for_each(eligible, [&](int j) {
        const double y = gv.y[j];
        const SampleVar sv = sample_var(ec, gv, j, y);
        double z[C];
        load_x<C>(Z, j, z);
        nm.add(sv.w * ec.itheta0, sv.valid ? y : 0.0, z);
        nm.k += sv.valid ? 1 : 0;
    });
    nm.reduce(gv.lpg);
    ridge_and_pad<C>(nm.a, c_real, nf);
    double det;
    const bool ok = chol<C, true>(nm.a, det, 0.0);
    double sol[C];
#pragma unroll
    for (int i = 0; i < C; ++i) sol[i] = nm.b[i];
    chol_solve<C>(nm.a, sol);
    double d2 = nm.yy;
#pragma unroll
    for (int i = 0; i < C; ++i) d2 -= nm.b[i] * sol[i];
    if (warp_any(eligible && d2 < 1e-4 * nm.yy)) {          // cancellation: d2 = sum w r~^2 + u'u
        double s = 0.0;
        gv.for_each(eligible && d2 < 1e-4 * nm.yy, [&](int j) {
            const double y = gv.y[j];
            const SampleVar sv = sample_var(ec, gv, j, y);
            double z[C];
            load_x<C>(Z, j, z);
            double r = sv.valid ? y : 0.0;
#pragma unroll
            for (int t = 0; t < C; ++t) r -= z[t] * sol[t];
            s += sv.w * ec.itheta0 * r * r;
        });
        s = lane_sum(s, gv.lpg);
#pragma unroll
        for (int i = 0; i < C; ++i) if (i >= c_real && i < c_real + nf) s += sol[i] * sol[i];
        if (d2 < 1e-4 * nm.yy) d2 = s;
    }
    if (d2 < 0.0) d2 = 0.0;

    eval_zero(out);
    const int p = nm.k - c_real;
    const bool good = eligible && ok && p > 0 && d2 == d2;
    double kappa = 0.5;
    if (good) {
        const double logdet = nm.k * ec.ltheta0 + log(det) + cst;
        double dv;
        out.value = finish_density(ec, p, logdet, d2, dv, kappa);
        out.dv = dv; out.d2 = d2; out.p = p; out.used = 1;
    } else if (eligible) {
        out.value = NAN; out.used = 1;
    }
    const bool grad_ok = good && out.value == out.value;

    // columns c..c+nf-1 of H^-1 (at most 4 factors)
    double hcol[4][C];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
#pragma unroll
        for (int i = 0; i < C; ++i) hcol[f][i] = (f < nf && i == c_real + f) ? 1.0 : 0.0;
        if (f < nf) chol_solve<C>(nm.a, hcol[f]);
    }
    double gsum = 0.0;
    gv.for_each(true, [&](int j) {
        double gf[4] = {0.0, 0.0, 0.0, 0.0};
        if (grad_ok) {
            const double y = gv.y[j];
            const SampleVar sv = sample_var(ec, gv, j, y);
            if (sv.valid) {
                const double w = sv.w * ec.itheta0;
                double z[C];
                load_x<C>(Z, j, z);
                double r = y;
#pragma unroll
                for (int t = 0; t < C; ++t) r -= z[t] * sol[t];
                const double h = leverage<C>(nm.a, z);
                gsum += 0.5 * (1.0 - w * h) - kappa * w * r * r;
                const double wr2k = 2.0 * kappa * w * r;
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    if (f < nf) {
                        double zh = 0.0, uf = 0.0;
#pragma unroll
                        for (int t = 0; t < C; ++t) { zh += z[t] * hcol[f][t]; if (t == c_real + f) uf = sol[t]; }
                        gf[f] = w * zh - wr2k * uf;
                    }
                }
            }
        }
        for (int f = 0; f < nf; ++f) gscratch[j * nf + f] = gf[f];
    });
    out.ga = lane_sum(gsum, gv.lpg) * ec.itheta0;      // d sigma2_j/d theta = sigma2_j/theta
}

// ------------------------------------------------------------------------------------------
// K3: GLS coefficients and their posterior (fit_coef, fitnoise.py:294-347)
// ------------------------------------------------------------------------------------------
struct CoefOut {
    bool ok;            // k >= c and full rank
    double d2;
    int p2;             // k - c
};

// beta (whitened basis) and A^-1 (packed lower, whitened basis).  The caller maps them back through
// R^-1 of the design's thin QR.  The variance kinds are evaluated at the FITTED theta
// (fitnoise.py:311-313).
// With factors (nf > 0) X is the augmented [design | factors] table and beta / ainv are the
// solution and the inverse of the ridge-augmented system; their first c_real entries / leading
// c_real x c_real block are the coefficients and (X'Sigma^-1 X)^-1.
template <int C>
FN_HD void coef_gene(const EvalConst& ec, const GeneView& gv, const double* X, int c_real, int nf,
                     double* beta, double* ainv, CoefOut& out) {
    const double wscale = nf > 0 ? ec.itheta0 : 1.0;       // the ridge does not scale with theta
    // SCALE1 without weights and factors works at theta = 1 with w_j = 1: for a gene that keeps every
    // sample the normal matrix is X'X = I (orthonormal columns), so beta = X'y and A^-1 = I -- no
    // accumulation of A, no factorisation, no inverse.  A missing value makes y'y non-finite, which is
    // how the other genes are found (as in scale1_sweep); they take the general route below.
    const bool unit = ec.kind == KIND_SCALE1 && gv.aux == nullptr && nf == 0;
    bool done = false, unit_d2_ok = false;
    double unit_d2 = 0.0;
    int k = 0;
    if (unit) {
        double b[C], yy = 0.0;
#pragma unroll
        for (int i = 0; i < C; ++i) b[i] = 0.0;
        gv.for_each_hot(true, [&](int j) {
            const double y = gv.y[j];
            double x[C];
            load_x<C>(X, j, x);
#pragma unroll
            for (int i = 0; i < C; ++i) b[i] += x[i] * y;
            yy += y * y;
        });
        lane_sum_n<C>(b, gv.lpg);
        yy = lane_sum(yy, gv.lpg);
        done = finite_d(yy);
        if (done) {
            k = gv.m;
            // d2 = y'y - |b|^2 as in scale1_sweep; redone from residuals below when it cancels
            unit_d2 = yy;
#pragma unroll
            for (int i = 0; i < C; ++i) if (i < c_real) unit_d2 -= b[i] * b[i];
            unit_d2_ok = unit_d2 >= 1e-4 * yy;
#pragma unroll
            for (int i = 0; i < C; ++i) beta[i] = (i < c_real) ? b[i] : 0.0;
#pragma unroll
            for (int i = 0; i < C; ++i)
#pragma unroll
                for (int q = 0; q <= i; ++q) ainv[FN_TRI(i, q)] = (i == q) ? 1.0 : 0.0;
        }
    }
    bool full_rank = true, okA = true;
    if (!unit || warp_any(!done)) {
        Normal<C> nm;
        nm.init();
        gv.for_each(!done, [&](int j) {
            const double y = gv.y[j];
            const SampleVar sv = sample_var(ec, gv, j, y);
            double x[C];
            load_x<C>(X, j, x);
            nm.add(sv.w * wscale, sv.valid ? y : 0.0, x);
            nm.k += sv.valid ? 1 : 0;
        });
        nm.reduce(gv.lpg);
        // rank of the retained sub-design: the unweighted masked Gram matrix -- the identity when nothing is
        // dropped, so only genes that lost a sample factorise it
        const bool partial = !done && nm.k < gv.m;
        if (warp_any(partial)) {
            Normal<C> gm;
            gm.init();
            gv.for_each(partial, [&](int j) {
                const SampleVar sv = sample_var(ec, gv, j, gv.y[j]);
                if (sv.valid) {
                    double x[C];
                    load_x<C>(X, j, x);
#pragma unroll
                    for (int i = 0; i < C; ++i) if (i >= c_real) x[i] = 0.0;     // rank of the design only
                    gm.add(1.0, 0.0, x);
                }
            });
            gm.reduce(gv.lpg);
            gm.pad(c_real);
            double detg;
            const bool rank2 = chol<C>(gm.a, detg, FN_RANK_TOL);
            if (partial) full_rank = rank2;
        }
        ridge_and_pad<C>(nm.a, c_real, nf);
        double det;
        const bool ok2 = chol<C>(nm.a, det, 0.0);
        if (!done) {
            okA = ok2;
            k = nm.k;
#pragma unroll
            for (int i = 0; i < C; ++i) beta[i] = nm.b[i];
            chol_solve<C>(nm.a, beta);
            chol_inverse<C>(nm.a, ainv);
        }
    }
    // residual sum of squares: a sweep of its own, except for complete unweighted genes whose d2 came
    // out of the first sweep without cancellation (every gene of a data set without missing values)
    double d2 = 0.0;
    const bool from_sums = done && unit_d2_ok;
    if (warp_any(!from_sums)) {
        gv.for_each(!from_sums, [&](int j) {
            const double y = gv.y[j];
            const SampleVar sv = sample_var(ec, gv, j, y);
            double x[C];
            load_x<C>(X, j, x);
            double r = sv.valid ? y : 0.0;
#pragma unroll
            for (int t = 0; t < C; ++t) r -= x[t] * beta[t];
            d2 += sv.w * wscale * r * r;
        });
        d2 = lane_sum(d2, gv.lpg);
    }
    if (from_sums) d2 = unit_d2;
#pragma unroll
    for (int i = 0; i < C; ++i) if (i >= c_real && i < c_real + nf) d2 += beta[i] * beta[i];   // + u'u
    if (ec.kind == KIND_SCALE1 && nf == 0) {  // sums were taken at theta = 1
        d2 /= ec.theta0;
#pragma unroll
        for (int i = 0; i < Tri<C>::N; ++i) ainv[i] *= ec.theta0;
    }
    out.p2 = k - c_real;
    out.ok = (k >= c_real) && full_rank && okA;
    out.d2 = (out.p2 > 0) ? d2 : 0.0;        // k == c: the fit is exact (conditional on nothing)
}

// ------------------------------------------------------------------------------------------
// K3b: contrasts and their test (Model.test, fitnoise.py:368-396; Mv*.p_value distributions.py:52-57,150-155)
// ------------------------------------------------------------------------------------------
#define FN_MAXC 16

// coef (c), V (c x c, row-major), df.  contrasts is c x kc row-major.  Writes contrasted (kc),
// optionally the kc x kc covariance C'VC, and the p-value of contrasted == 0.
FN_HDN void test_gene(int c, int kc, const double* coef, const double* V, double df, int is_t,
                      const double* contrasts, double* contrasted, double* ccov, double& pval) {
    double mean[FN_MAXC], vc[FN_MAXC * FN_MAXC], s[FN_MAXC * FN_MAXC], z[FN_MAXC];
    for (int a = 0; a < kc; ++a) {
        double acc = 0.0;
        for (int i = 0; i < c; ++i) acc += contrasts[i * kc + a] * coef[i];
        mean[a] = acc;
        contrasted[a] = acc;
    }
    for (int i = 0; i < c; ++i)                 // vc = V C   (c x kc)
        for (int a = 0; a < kc; ++a) {
            double acc = 0.0;
            for (int t = 0; t < c; ++t) acc += V[i * c + t] * contrasts[t * kc + a];
            vc[i * kc + a] = acc;
        }
    for (int a = 0; a < kc; ++a)                // s = C' V C  (kc x kc)
        for (int b = 0; b < kc; ++b) {
            double acc = 0.0;
            for (int i = 0; i < c; ++i) acc += contrasts[i * kc + a] * vc[i * kc + b];
            s[a * kc + b] = acc;
            if (ccov) ccov[a * kc + b] = acc;
        }
    // q = mean' s^-1 mean through a Cholesky factor of s
    bool ok = true;
    for (int a = 0; a < kc; ++a) {
        for (int b = 0; b <= a; ++b) {
            double acc = s[a * kc + b];
            for (int t = 0; t < b; ++t) acc -= s[a * kc + t] * s[b * kc + t];
            if (b < a) s[a * kc + b] = acc / s[b * kc + b];
            else { if (!(acc > 0.0)) { ok = false; acc = 1.0; } s[a * kc + a] = sqrt(acc); }
        }
    }
    double q = 0.0;
    for (int a = 0; a < kc; ++a) {
        double acc = mean[a];
        for (int t = 0; t < a; ++t) acc -= s[a * kc + t] * z[t];
        z[a] = acc / s[a * kc + a];
        q += z[a] * z[a];
    }
    if (!ok) { pval = NAN; return; }
    pval = is_t ? f_sf(q / kc, (double)kc, df) : chi2_sf(q, (double)kc);
}

// The same for a compile-time number of contrasts KC (1..4): C'VC is accumulated straight from V, so
// nothing is indexed at run time and everything stays in registers (test_gene keeps FN_MAXC^2 local
// arrays: a 4 KB stack frame per thread for what is usually c = 2, kc = 1).
template <int KC>
FN_HD void test_gene_k(int c, const double* coef, const double* V, double df, int is_t,
                       const double* contrasts, double* contrasted, double* ccov, double& pval) {
    double mean[KC], s[KC * KC];
#pragma unroll
    for (int a = 0; a < KC; ++a) mean[a] = 0.0;
#pragma unroll
    for (int a = 0; a < KC * KC; ++a) s[a] = 0.0;
    for (int i = 0; i < c; ++i) {
        const double ci = coef[i];
        double vc[KC];                       // row i of V C, summed in the order test_gene uses
#pragma unroll
        for (int b = 0; b < KC; ++b) vc[b] = 0.0;
        for (int t = 0; t < c; ++t) {
            const double v = V[i * c + t];
#pragma unroll
            for (int b = 0; b < KC; ++b) vc[b] += v * contrasts[t * KC + b];
        }
#pragma unroll
        for (int a = 0; a < KC; ++a) {
            const double cia = contrasts[i * KC + a];
            mean[a] += cia * ci;
#pragma unroll
            for (int b = 0; b < KC; ++b) s[a * KC + b] += cia * vc[b];
        }
    }
#pragma unroll
    for (int a = 0; a < KC; ++a) contrasted[a] = mean[a];
    if (ccov) {
#pragma unroll
        for (int a = 0; a < KC * KC; ++a) ccov[a] = s[a];
    }
    bool ok = true;
#pragma unroll
    for (int a = 0; a < KC; ++a) {
#pragma unroll
        for (int b = 0; b <= a; ++b) {
            double acc = s[a * KC + b];
#pragma unroll
            for (int t = 0; t < b; ++t) acc -= s[a * KC + t] * s[b * KC + t];
            if (b < a) s[a * KC + b] = acc / s[b * KC + b];
            else { if (!(acc > 0.0)) { ok = false; acc = 1.0; } s[a * KC + a] = sqrt(acc); }
        }
    }
    double q = 0.0, z[KC];
#pragma unroll
    for (int a = 0; a < KC; ++a) {
        double acc = mean[a];
#pragma unroll
        for (int t = 0; t < a; ++t) acc -= s[a * KC + t] * z[t];
        z[a] = acc / s[a * KC + a];
        q += z[a] * z[a];
    }
    if (!ok) { pval = NAN; return; }
    pval = is_t ? f_sf(q / KC, (double)KC, df) : chi2_sf(q, (double)KC);
}

}  // namespace fn
