// TEST-ONLY host build (g++, no CUDA) of the per-gene mathematics in
// fitnoise_b200/csrc/fn_gene.cuh, driven gene by gene with lpg = 1.  It lets the CPU test-suite
// compare the exact arithmetic the CUDA kernels run against the oracle on a box without a GPU.
// It is NOT a CPU fallback: the product never loads this library (fitnoise_b200/_native.py only
// loads libfitnoise_b200.so and raises when it or a CUDA device is missing).
#include <stdint.h>
#include <stdlib.h>
#include <vector>
#include "../../fitnoise_b200/csrc/fn_host.hpp"

using namespace fn;

namespace {

struct Prepared {
    std::vector<double> aux;     // transformed second array (precision weights, or rc)
    std::vector<double> gs, cst, average;
    std::vector<uint8_t> eligible;
    double rowvar_mean = 0.0;
    long bad_counts = 0;
};

template <int C>
void prepare_all(const Family& fam, long n, int m, const double* y, const double* aux_raw,
                 const uint8_t* controls, const Design& dn, const Design& dc, Prepared& pr) {
    pr.aux.assign(aux_raw ? (size_t)n * m : 0, 0.0);
    pr.gs.assign(n, 0.0); pr.cst.assign(n, 0.0); pr.average.assign(n, NAN); pr.eligible.assign(n, 0);
    double vsum = 0.0; long vcount = 0;
    std::vector<double> row(m), arow(m);
    for (long g = 0; g < n; ++g) {
        for (int j = 0; j < m; ++j) { row[j] = y[g * m + j]; if (aux_raw) arow[j] = aux_raw[g * m + j]; }
        GeneView gv; gv.y = row.data(); gv.aux = aux_raw ? arow.data() : nullptr; gv.gs = 0.0;
        gv.setup(m, 0, 1, 0);
        const Design& d = (controls && controls[g]) ? dc : dn;
        PrepOut po;
        prepare_gene<C>(fam, gv, d.q.data(), d.c, po);
        pr.eligible[g] = po.eligible; pr.cst[g] = po.cst; pr.average[g] = po.average;
        pr.bad_counts += po.bad_counts;
        if (po.nfinite > 0) { vsum += po.yss / po.nfinite; ++vcount; }
        if (fam.kind == KIND_PATSEQ) {
            pr.gs[g] = po.avg_tail * po.avg_tail;
            for (int j = 0; j < m; ++j) { const double cnt = aux_raw[g * m + j]; pr.aux[g * m + j] = (cnt == cnt) ? 1.0 / fmax(cnt, 1.0) : NAN; }
        } else if (aux_raw) {
            for (int j = 0; j < m; ++j) pr.aux[g * m + j] = aux_raw[g * m + j];
        }
    }
    pr.rowvar_mean = vcount ? vsum / vcount : NAN;
}

// host build of the complete-row tables of KIND_SCALE_PS (the CUDA kernel builds them per CTA:
// fn_kernels.cuh build_ps_design)
template <int C>
struct PsHost {
    std::vector<double> hx, cj;
    PsDesign<C> pd;
    void build(const Design& d, int m, const double* w, const double* lt) {
        constexpr int N = Tri<C>::N;
        double a[N], ainv[N];
        for (int e = 0; e < N; ++e) a[e] = 0.0;
        for (int j = 0; j < m; ++j)
            for (int i = 0; i < C; ++i)
                for (int k = 0; k <= i; ++k) a[FN_TRI(i, k)] += w[j] * d.q[(size_t)j * C + i] * d.q[(size_t)j * C + k];
        for (int i = d.c; i < C; ++i) a[FN_TRI(i, i)] = 1.0;
        double det;
        pd.ok = chol<C>(a, det, 0.0);
        chol_inverse<C>(a, ainv);
        hx.assign((size_t)m * C, 0.0); cj.assign(m, 0.0);
        double slt = 0.0;
        for (int j = 0; j < m; ++j) {
            ps_table_row<C>(ainv, &d.q[(size_t)j * C], w[j], &hx[(size_t)j * C], cj[j]);
            slt += lt[j];
        }
        pd.hx = hx.data(); pd.cj = cj.data(); pd.logdet_a = log(det); pd.slt = slt;
    }
};

// one gene through the complete-row fast path; returns false when the row is not complete
template <int C>
bool ps_gene(const EvalConst& ec, const PsHost<C>& ps, GeneView& gv, const Design& d, bool eligible, double cst,
             EvalOut& eo, double& kappa) {
    double beta[C];
    if (!ps_complete_beta<C>(ps.pd, gv, eligible, beta)) return false;
    ps_complete_finish<C>(ec, ps.pd, gv, d.q.data(), d.c, true, false, beta, cst, eo, kappa);
    return true;
}

}  // namespace

extern "C" {

// One likelihood + gradient evaluation over all genes.  grad has theta_length entries.
// Widths of the two designs are rounded up to the same template width (max of the two).
// factors: F is m x fam->nf (may be NULL when nf == 0); grad_f receives d/dF (m x nf).
int hc_eval(const Family* fam, long n, int m, const double* y, const double* aux_raw,
            const uint8_t* controls, int c_noise, const double* xn, int c_ctrl, const double* xc,
            const double* theta, double* score, double* d2, int* pdf, double* averages,
            uint8_t* eligible, double* value_total, double* grad, double* rowvar_mean, long* bad_counts,
            const double* F, double* grad_f) {
    Design dn, dc;
    if (!dn.build(m, c_noise, xn)) return -1;
    if (!dc.build(m, c_ctrl, xc)) return -2;
    const int nf = fam->nf;
    const int width = round_width((dn.c > dc.c ? dn.c : dc.c) + nf);
    if (width < 0) return -5;
    // re-pad both to the common width
    auto repad = [&](Design& d) {
        std::vector<double> q((size_t)m * width, 0.0);
        for (int j = 0; j < m; ++j) for (int i = 0; i < d.c; ++i) q[(size_t)j * width + i] = d.q[(size_t)j * d.width + i];
        d.q.swap(q); d.width = width;
    };
    repad(dn); repad(dc);
    if (!family_valid(*fam, m, aux_raw != nullptr)) return -3;
    double v, theta0; int is_t;
    std::vector<double> ta, tb;
    if (!unpack_theta(*fam, m, theta, v, is_t, theta0, ta, tb)) return -4;
    std::vector<double> t0(m + 1), t1(m + 1);
    for (int p = 0; p <= m; ++p) density_tables(is_t, v, p, t0[p], t1[p]);
    EvalConst ec;
    ec.kind = fam->kind; ec.tail_own = fam->tail_own; ec.v3 = fam->v3; ec.is_t = is_t; ec.v = v;
    ec.theta0 = theta0; ec.ta = ta.data(); ec.tb = tb.data(); ec.t0 = t0.data(); ec.t1 = t1.data(); ec.finish_setup();

    Prepared pr;
    FN_DISPATCH_WIDTH(width, prepare_all<C>(*fam, n, m, y, aux_raw, controls, dn, dc, pr));
    *rowvar_mean = pr.rowvar_mean; *bad_counts = pr.bad_counts;

    std::vector<double> zn, zc, gscratch((size_t)m * (nf ? nf : 1), 0.0);
    if (nf) {
        zn = augment_design(dn, width, nf, F);
        zc = augment_design(dc, width, nf, F);
        for (int i = 0; i < m * nf; ++i) grad_f[i] = 0.0;
    }
    const int P = theta_length(*fam);
    for (int i = 0; i < P; ++i) grad[i] = 0.0;
    const bool general_only = getenv("FN_NO_PATSEQ_SHARED") != nullptr;      // as the CUDA library: two-sweep path
    const int off_a = fam->dist == DIST_T ? 1 : 0, off_b = off_a + fam->na;
    double total = 0.0;
    std::vector<double> row(m), arow(m);
    for (long g = 0; g < n; ++g) {
        for (int j = 0; j < m; ++j) { row[j] = y[g * m + j]; if (aux_raw) arow[j] = pr.aux[g * m + j]; }
        GeneView gv; gv.y = row.data(); gv.aux = aux_raw ? arow.data() : nullptr; gv.gs = pr.gs[g];
        gv.setup(m, 0, 1, 0);
        const Design& d = (controls && controls[g]) ? dc : dn;
        EvalOut eo;
        const bool inpl_a = fam->na == m && fam->kind != KIND_SCALE1, inpl_b = fam->nb == m;
        if (fam->kind == KIND_SCALE_PS && !aux_raw && !nf && !getenv("FN_NO_PS_FAST")) {
            // complete genes: the shared-normal-matrix fast path, as the CUDA kernel does
            bool done = false;
            double kappa = 0.0;
            const bool ctl = controls && controls[g];
            FN_DISPATCH_WIDTH(width, {
                static thread_local PsHost<C> psn, psc;
                static thread_local const double* built_for = nullptr;
                if (built_for != theta || g == 0) { psn.build(dn, m, ta.data(), tb.data()); psc.build(dc, m, ta.data(), tb.data()); built_for = theta; }
                done = ps_gene<C>(ec, ctl ? psc : psn, gv, d, pr.eligible[g], pr.cst[g], eo, kappa);
                if (done) {
                    score[g] = eo.value; d2[g] = eo.d2; pdf[g] = eo.p; averages[g] = pr.average[g]; eligible[g] = pr.eligible[g];
                    total += eo.value;
                    if (fam->dist == DIST_T) grad[0] += eo.dv;
                    const std::vector<double>& cj = ctl ? psc.cj : psn.cj;
                    for (int j = 0; j < m; ++j) grad[off_a + j] += (kappa > 0.0 ? cj[j] : 0.0) - kappa * row[j];
                }
            });
            if (done) continue;
            for (int j = 0; j < m; ++j) row[j] = y[g * m + j];      // (untouched: the first sweep only reads)
        }
        FN_DISPATCH_WIDTH(width,
            if (nf) eval_factors<C>(ec, gv, ((controls && controls[g]) ? zc : zn).data(), d.c, nf, pr.eligible[g], pr.cst[g], gscratch.data(), eo);
            else if (fam->kind == KIND_SCALE1) eval_scale1<C>(ec, gv, d.q.data(), d.c, pr.eligible[g], pr.cst[g], eo);
            else if (fam->kind == KIND_SCALE_PS) eval_general<C, KIND_SCALE_PS>(ec, gv, d.q.data(), d.c, pr.eligible[g], pr.cst[g], inpl_a, inpl_b, eo);
            else if (fam->na == 1 && fam->nb == 1 && width <= 4 && !general_only) eval_patseq_shared<C>(ec, gv, d.q.data(), d.c, pr.eligible[g], pr.cst[g], eo);
            else eval_general<C, KIND_PATSEQ>(ec, gv, d.q.data(), d.c, pr.eligible[g], pr.cst[g], inpl_a, inpl_b, eo));
        score[g] = eo.used ? eo.value : NAN;
        d2[g] = eo.d2; pdf[g] = eo.p;
        averages[g] = pr.average[g];
        eligible[g] = pr.eligible[g];
        if (eo.used) {
            total += eo.value;
            if (fam->dist == DIST_T) grad[0] += eo.dv;
            if (fam->na == 1) grad[off_a] += eo.ga;
            if (fam->nb == 1) grad[off_b] += eo.gb;
        }
        if (nf) for (int i = 0; i < m * nf; ++i) grad_f[i] += gscratch[i];
        if (inpl_a) for (int j = 0; j < m; ++j) grad[off_a + j] += row[j];
        if (inpl_b) for (int j = 0; j < m; ++j) grad[off_b + j] += arow[j];
    }
    if (fam->kind == KIND_PATSEQ)            // eval_general returns term a times theta_a(j)
        for (int j = 0; j < fam->na; ++j) grad[off_a + j] /= theta[off_a + j];
    *value_total = total;
    return 0;
}

int hc_coef(const Family* fam, long n, int m, const double* y, const double* aux_raw,
            int c, const double* x, const double* theta, double* coef, double* covar, double* dfpost,
            const double* F) {
    Design d;
    if (!d.build(m, c, x)) return -1;
    const int nf = fam->nf;
    std::vector<double> z;
    if (nf) {
        const int width = round_width(c + nf);
        if (width < 0) return -5;
        z = augment_design(d, width, nf, F);
        d.width = width;
        d.q = z;
    }
    double v, theta0; int is_t;
    std::vector<double> ta, tb;
    if (!unpack_theta(*fam, m, theta, v, is_t, theta0, ta, tb)) return -4;
    EvalConst ec;
    ec.kind = fam->kind; ec.tail_own = fam->tail_own; ec.v3 = fam->v3; ec.is_t = is_t; ec.v = v;
    ec.theta0 = theta0; ec.ta = ta.data(); ec.tb = tb.data(); ec.t0 = nullptr; ec.t1 = nullptr; ec.finish_setup();
    Prepared pr;
    Design dummy = d;
    FN_DISPATCH_WIDTH(d.width, prepare_all<C>(*fam, n, m, y, aux_raw, nullptr, d, dummy, pr));
    std::vector<double> row(m), arow(m);
    for (long g = 0; g < n; ++g) {
        for (int j = 0; j < m; ++j) { row[j] = y[g * m + j]; if (aux_raw) arow[j] = pr.aux[g * m + j]; }
        GeneView gv; gv.y = row.data(); gv.aux = aux_raw ? arow.data() : nullptr; gv.gs = pr.gs[g];
        gv.setup(m, 0, 1, 0);
        double beta[FN_MAXC], ainv[FN_MAXC * (FN_MAXC + 1) / 2];
        CoefOut co;
        FN_DISPATCH_WIDTH(d.width, coef_gene<C>(ec, gv, d.q.data(), d.c, nf, beta, ainv, co));
        double* cf = coef + g * c; double* cv = covar + g * c * c;
        if (!co.ok) {
            for (int i = 0; i < c; ++i) cf[i] = NAN;
            for (int i = 0; i < c * c; ++i) cv[i] = NAN;
            dfpost[g] = NAN;
            continue;
        }
        const double s2 = is_t ? (v + co.d2) / (v + co.p2) : 1.0;
        dfpost[g] = is_t ? v + co.p2 : NAN;
        // coef = Rinv beta ; V = Rinv Ainv Rinv' s2
        for (int i = 0; i < c; ++i) {
            double acc = 0.0;
            for (int t = i; t < c; ++t) acc += d.rinv[i * c + t] * beta[t];
            cf[i] = acc;
        }
        double tmp[FN_MAXC * FN_MAXC];
        for (int i = 0; i < c; ++i) for (int j = 0; j < c; ++j) {
            double acc = 0.0;
            for (int t = i; t < c; ++t) { const int hi = t > j ? t : j, lo = t > j ? j : t; acc += d.rinv[i * c + t] * ainv[FN_TRI(hi, lo)]; }
            tmp[i * c + j] = acc;
        }
        for (int i = 0; i < c; ++i) for (int j = 0; j < c; ++j) {
            double acc = 0.0;
            for (int t = j; t < c; ++t) acc += tmp[i * c + t] * d.rinv[j * c + t];
            cv[i * c + j] = acc * s2;
        }
    }
    return 0;
}

void hc_test(long n, int c, int kc, const double* coef, const double* covar, const double* df, int is_t,
             const double* contrasts, double* contrasted, double* ccov, double* p) {
    for (long g = 0; g < n; ++g) {
        if (covar[g * c * c] != covar[g * c * c]) {
            for (int a = 0; a < kc; ++a) contrasted[g * kc + a] = NAN;
            for (int a = 0; a < kc * kc; ++a) ccov[g * kc * kc + a] = NAN;
            p[g] = NAN;
            continue;
        }
        // the same dispatch as the CUDA test_kernel: compile-time kc up to 4
        const double* cf = coef + g * c; const double* V = covar + g * c * c;
        double* out = contrasted + g * kc; double* cc = ccov + g * kc * kc;
        switch (kc) {
            case 1: test_gene_k<1>(c, cf, V, df[g], is_t, contrasts, out, cc, p[g]); break;
            case 2: test_gene_k<2>(c, cf, V, df[g], is_t, contrasts, out, cc, p[g]); break;
            case 3: test_gene_k<3>(c, cf, V, df[g], is_t, contrasts, out, cc, p[g]); break;
            case 4: test_gene_k<4>(c, cf, V, df[g], is_t, contrasts, out, cc, p[g]); break;
            default: test_gene(c, kc, cf, V, df[g], is_t, contrasts, out, cc, p[g]); break;
        }
    }
}

}  // extern "C"
