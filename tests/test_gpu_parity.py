"""Parity of the CUDA path (through the C ABI, libfitnoise_b200.so) with the reference's own
outputs (tests/golden) and with the oracle on seeded inputs.  Needs a B200: ``-m gpu``.

Tolerances (north_star): per-gene log-likelihood, coefficients and statistics 1e-9 relative;
p-values 1e-8 relative; fitted hyper-parameters 1e-6 relative; BH bit-exact given p.
"""
import os
import zlib

import numpy as np
import pytest

import fitnoise_b200 as fitnoise
from conftest import ROOT, assert_close, golden_context, golden_fit_kwargs, golden_names, load_golden
from fitnoise_b200 import _native
from helpers import check_against_golden, fit_golden, rank_deficient_rows, synthetic
from oracle import fitnoise_oracle as oracle

pytestmark = pytest.mark.gpu

NAMES = [n for n in golden_names() if n != "t_nan_quirk"]


@pytest.fixture(scope="module")
def engine():
    e = _native.Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("name", NAMES)
def test_golden_forced_param(name):
    """fit(force_param) + test on the GPU against the reference's outputs."""
    g = load_golden(name)
    fitted, tested = fit_golden(g, force_param=g["theta"])
    check_against_golden(g, fitted, tested)
    score, d2, p = fitted._session.per_gene(fitted._theta)
    ref = g["ref_score_rows"].copy()
    ref[rank_deficient_rows(g)] = np.nan
    assert_close(score, ref, 1e-9, what="per-gene score_row")


@pytest.mark.parametrize("name", NAMES)
def test_golden_value_and_gradient(name):
    """Summed objective and its gradient against the oracle (matrix-calculus gradient on the
    reference's own S, cross-checked against central differences in test_oracle_golden.py)."""
    g = load_golden(name)
    theta = g["theta"]
    if len(theta) == 0:
        pytest.skip("no hyper-parameters")
    fitted, _ = fit_golden(g, force_param=theta)
    value, grad, used, bad = fitted._objective(theta)
    kw = golden_fit_kwargs(g)
    nd = g.get("noise_design", g["design"])
    cfg = oracle.Configured(str(g["model"]), g["y"], golden_context(g), n_factors=int(g.get("n_factors", 0)),
                            design=nd, control_design=kw.get("control_design", np.ones((g["y"].shape[1], 1))),
                            controls=kw.get("controls"))
    out = oracle.fit_noise(cfg, nd, kw.get("control_design"), kw.get("controls"), force_param=theta,
                           skip_rank_deficient=True)
    ov, og = oracle.total_score_grad(cfg, out["prepared"], theta)
    assert used == len(out["prepared"]) and bad == 0
    assert_close(value, ov, 1e-11, what="objective")
    assert_close(grad, og, 1e-8, atol=1e-9 * abs(ov), what="gradient")


@pytest.mark.parametrize("lpg,threads", [(1, 128), (2, 128), (4, 256), (8, 256), (16, 256)])
def test_lane_mappings_agree(lpg, threads):
    """Every lanes-per-gene mapping of the tile kernels gives the same per-gene numbers."""
    g = load_golden("t_patseq_v2_per_sample")
    fam_model = getattr(fitnoise, str(g["model"]))()
    e = _native.Engine(0)
    try:
        m = g["y"].shape[1]
        fam = fam_model._family(m)
        e.set_tuning(lpg=0, threads=0)
        e.upload(g["y"], g["counts"], fam, None, g["design"])
        base = e.per_gene(g["theta"])
        bv, bg, _, _ = e.nll_grad(g["theta"])
        e.set_tuning(lpg=lpg, threads=threads)
        e.upload(g["y"], g["counts"], fam, None, g["design"])
        got = e.per_gene(g["theta"])
        v, gr, _, _ = e.nll_grad(g["theta"])
    finally:
        e.close()
    assert_close(got[0], base[0], 1e-12, what="score")
    assert_close(got[1], base[1], 1e-11, what="d2")
    assert np.array_equal(got[2], base[2])
    assert_close(v, bv, 1e-12)
    assert_close(gr, bg, 1e-9, atol=1e-12 * abs(bv))


@pytest.mark.parametrize("name", ["t_twogroup", "t_patseq_v2", "t_controls_opt", "t_factors1"])
def test_fitted_hyperparameters(name):
    g = load_golden(name)
    fitted, tested = fit_golden(g)
    ref = oracle.fit_and_test(str(g["model"]), g["y"], g["design"], golden_context(g),
                              coef_idx=[int(i) for i in g["test_coef"]], nan_aware=True,
                              skip_rank_deficient=True, n_factors=int(g.get("n_factors", 0)),
                              **golden_fit_kwargs(g))
    x, rx = fitted.optimization_result.x, ref["x"]
    free = ~(np.isclose(rx, 100.0) | np.isclose(rx, 1e-3))
    assert_close(x[free], rx[free], 1e-6, what="theta")
    assert_close(fitted.optimization_result.fun, ref["fun"], 1e-10, what="fun")
    assert_close(tested.p_values, ref["p_values"], 1e-5, what="p at fitted theta")


@pytest.mark.parametrize("lpg,threads", [(1, 128), (2, 256), (4, 256), (8, 128), (16, 256)])
@pytest.mark.parametrize("model", ["Model_t", "Model_t_per_sample", "Model_normal_patseq"])
def test_blocked_order_lane_mappings(model, lpg, threads):
    """m = 32 (a multiple of 16: blocked, bank-rotated sample order) under every lane mapping,
    against the oracle."""
    rng = np.random.default_rng(11)
    n, m = 300, 32
    design = np.stack([np.ones(m), (np.arange(m) >= 16) * 1.0, rng.normal(size=m)], axis=1)
    y = rng.normal(size=(n, m)) * 1.3 + 20.0
    context, aux = {}, None
    if model == "Model_normal_patseq":
        aux = rng.poisson(20.0, size=(n, m)).astype(float)
        aux[rng.random((n, m)) < 0.1] = 0
        y[aux == 0] = np.nan
        context = {"counts": aux}
        theta = np.array([300.0, 0.004])
    elif model == "Model_t_per_sample":
        theta = np.array([7.0] + list(rng.uniform(0.5, 2.0, size=m)))
    else:
        y[3, 5] = np.nan
        theta = np.array([7.0, 1.5])
    fam = getattr(fitnoise, model)()._family(m)
    e = _native.Engine(0)
    try:
        e.set_tuning(lpg=lpg, threads=threads)
        e.upload(y, aux, fam, None, design)
        score, d2, p = e.per_gene(theta)
        value, grad, used, bad = e.nll_grad(theta)
    finally:
        e.close()
    rows = np.arange(0, n, 6)
    cfg = oracle.Configured(model, y[rows], {k: v[rows] for k, v in context.items()}, nan_aware=True)
    prep = oracle.prepare_rows(cfg, [design] * len(rows), skip_rank_deficient=True)
    want = np.full(len(rows), np.nan)
    for r, item in prep.items():
        want[r] = oracle.score_row(cfg, theta, r, *item)
    assert_close(score[rows], want, 1e-9, what="score")
    assert_close(np.nansum(score), value, 1e-12, what="sum of scores")
    assert bad == 0


@pytest.mark.parametrize("config,n", [(1, 400), (2, 300), (3, 300), (4, 150), (5, 60)])
def test_synthetic_configs_against_oracle(config, n):
    """The five BASELINE.json shapes at a size the oracle finishes in seconds."""
    s = synthetic(config, n=n)
    model = getattr(fitnoise, s["model"])()
    kw = {k: s[k] for k in ("controls", "control_design") if k in s}
    init = model._family(s["y"].shape[1])
    cfg = oracle.Configured(s["model"], s["y"], s["context"], nan_aware=True)
    theta = cfg.initial.copy()
    if len(theta) and config == 3:
        theta = np.array([12.0, 700.0, 0.003])
    fitted = model.fit(fitnoise.Dataset(s["y"], s["context"]), s["design"], force_param=theta, **kw)
    tested = fitted.test(**s["test"])
    t = s["test"]
    ref = oracle.fit_and_test(s["model"], s["y"], s["design"], s["context"], coef_idx=t.get("coef"),
                              contrasts=t.get("contrasts"), force_param=theta, nan_aware=True,
                              skip_rank_deficient=True, **kw)
    assert_close(fitted.noise_p_values, ref["noise_p_values"], 1e-8, what="noise p")
    assert_close(fitted.averages, ref["averages"], 1e-12, what="averages")
    assert_close(fitted.coef, ref["coef"], 1e-9, atol=1e-10, what="coef")
    assert_close(fitted._coef_covar, ref["coef_covar"], 1e-9, atol=1e-13, what="covar")
    assert_close(tested.p_values, ref["p_values"], 1e-8, what="p")
    assert np.array_equal(tested.q_values, oracle.fdr(tested.p_values)), "BH not bit-exact"
    if len(theta):
        value, grad, used, bad = fitted._objective(theta)
        ov, og = oracle.total_score_grad(ref["cfg"], ref["prepared"], theta)
        assert_close(value, ov, 1e-11, what="objective")
        assert_close(grad, og, 1e-8, atol=1e-9 * abs(ov), what="gradient")


def test_bh_bit_exact(engine):
    rng = np.random.default_rng(7)
    g = load_golden("fdr")
    for i in "1234":
        assert np.array_equal(engine.bh(g["p" + i]), g["ref_q" + i]), "golden fdr vector %s" % i
    for n in (1, 2, 255, 256, 257, 2048, 2049, 100003, 1000000):
        p = rng.random(n) ** rng.choice([1, 3, 8])
        if n > 10:
            p[rng.integers(0, n, n // 50 + 1)] = np.nan
            p[rng.integers(0, n, n // 20 + 1)] = p[0]              # ties
            p[1] = 0.0
            p[2] = 1.0
        q = engine.bh(p)
        assert np.array_equal(q, oracle.fdr(p)), "n = %d" % n
    assert np.array_equal(fitnoise.fdr([0.01, np.nan, 0.04, 0.03]), oracle.fdr([0.01, np.nan, 0.04, 0.03]))


def test_full_size_properties():
    """BASELINE.json's largest shape at FULL size, 1,000,000 x 96 (the oracle would take days):
    size-independent properties -- the oracle on a slice, invariances on the whole."""
    n, m = 1000000, 96
    rng = np.random.default_rng(5)
    design = np.array([[1.0, 0.0]] * 48 + [[1.0, 1.0]] * 48)
    y = rng.normal(size=(n, m)) * np.sqrt(10.0 / rng.chisquare(10.0, size=(n, 1)))
    fam = fitnoise.Model_t()._family(m)
    theta = np.array([8.0, 1.2])
    e = _native.Engine(0)
    try:
        e.upload(y, None, fam, None, design)
        v_all, g_all, used, bad = e.nll_grad(theta)
        score, d2, p = e.per_gene(theta)
        assert used == n and bad == 0
        # 1. the sum of the per-gene scores is the objective
        assert_close(score.sum(), v_all, 1e-12, what="sum of per-gene scores")
        # 2. the oracle on a slice of 40 genes
        rows = np.arange(0, n, n // 40)[:40]
        cfg = oracle.Configured("Model_t", y[rows])
        prep = oracle.prepare_rows(cfg, [design] * len(rows))
        ref = np.array([oracle.score_row(cfg, theta, r, *prep[r]) for r in range(len(rows))])
        assert_close(score[rows], ref, 1e-9, what="per-gene score vs oracle")
        # 3. additivity over gene shards (what multi-GPU sharding relies on)
        half = n // 2
        e.upload(y[:half], None, fam, None, design)
        v1, g1, _, _ = e.nll_grad(theta)
        e.upload(y[half:], None, fam, None, design)
        v2, g2, _, _ = e.nll_grad(theta)
        assert_close(v1 + v2, v_all, 1e-12, what="shard additivity (value)")
        assert_close(g1 + g2, g_all, 1e-9, atol=1e-12 * abs(v_all), what="shard additivity (gradient)")
        # 4. invariance to a change of basis of the design and to adding a fitted effect
        e.upload(y + design @ np.array([3.0, -2.0]), None, fam, None, design @ np.array([[2.0, 1.0], [0.0, -3.0]]))
        v3, g3, _, _ = e.nll_grad(theta)
        assert_close(v3, v_all, 1e-10, what="invariance to design basis / fitted shift")
        # 5. scaling y by s moves the variance parameter by s^2
        e.upload(2.0 * y, None, fam, None, design)
        v4, _, _, _ = e.nll_grad(np.array([8.0, 4.8]))
        assert_close(v4, v_all + n * (m - 2) * np.log(2.0), 1e-11, what="scale equivariance")
    finally:
        e.close()


def test_more_than_2_31_elements():
    """23,000,000 genes x 96 samples = 2.2e9 doubles (17.7 GB) resident on one GPU: element offsets
    pass 2^31, so any 32-bit index in the kernels or the tile arithmetic would show up here.  Properties:
    shard additivity, the per-gene results of the last genes against a small upload of the same rows, the
    weights of the last gene."""
    torch = pytest.importorskip("torch")
    n, m = 23000000, 96
    assert n * m > 2 ** 31
    if torch.cuda.mem_get_info(0)[0] < 60e9:
        pytest.skip("needs 60 GB of free device memory")
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    y = torch.randn((n, m), dtype=torch.float64, device=dev, generator=gen)
    y[n - 3, 5] = float("nan")
    torch.cuda.synchronize()
    design = np.array([[1.0, 0.0]] * 48 + [[1.0, 1.0]] * 48)
    fam = fitnoise.Model_t()._family(m)
    theta = np.array([7.0, 1.1])
    e = _native.Engine(0)
    try:
        e.upload_dev(n, m, y.data_ptr(), 0, fam, 0, design)
        v_all, g_all, used, bad = e.nll_grad(theta)
        assert used == n and bad == 0
        score, d2, p = e.per_gene(theta)
        assert_close(score.sum(), v_all, 1e-11, what="sum of per-gene scores")
        coef, _, _ = e.coef(theta, design, want_covar=False)
        con, pv, q = e.test_bh(np.array([[0.0], [1.0]]))
        tail = y[n - 1000:].cpu().numpy()
        half = n // 2 + 7
        e.upload_dev(half, m, y.data_ptr(), 0, fam, 0, design)
        v1, g1, u1, _ = e.nll_grad(theta)
        e.upload_dev(n - half, m, y.data_ptr() + half * m * 8, 0, fam, 0, design)
        v2, g2, u2, _ = e.nll_grad(theta)
        assert u1 + u2 == n
        assert_close(v1 + v2, v_all, 1e-11, what="shard additivity (value)")
        assert_close(g1 + g2, g_all, 1e-8, atol=1e-11 * abs(v_all), what="shard additivity (gradient)")
        e.upload(tail, None, fam, None, design)
        s_tail, _, p_tail = e.per_gene(theta)
        c_tail, _, _ = e.coef(theta, design, want_covar=False)
        con_tail, pv_tail, _ = e.test_bh(np.array([[0.0], [1.0]]))
        # (a gene's place in its tile sets the order of its sums: equal to rounding, not bitwise)
        assert_close(score[n - 1000:], s_tail, 1e-12, what="scores of the last genes")
        assert np.array_equal(p[n - 1000:], p_tail)
        assert_close(coef[n - 1000:], c_tail, 1e-10, atol=1e-13, what="coefficients of the last genes")
        assert_close(pv[n - 1000:], pv_tail, 1e-9, what="p-values of the last genes")
        assert p[n - 3] == m - 1 - 2 and p[n - 1] == m - 2
        assert np.array_equal(q, oracle.fdr(pv))
    finally:
        e.close()
        del y


def _random_case(rng, n, m, c, model):
    design = np.concatenate([np.ones((m, 1)), rng.normal(size=(m, c - 1))], axis=1) if c > 1 else np.ones((m, 1))
    y = rng.normal(size=(n, m)) * 1.5 + 10.0
    y[rng.random((n, m)) < 0.05] = np.nan
    aux = None
    fam = getattr(fitnoise, model)()._family(m)
    P = fam.theta_length()
    if fam.kind == _native.KIND_PATSEQ:
        aux = rng.poisson(15.0, size=(n, m)).astype(float)
        aux[np.isnan(y)] = 0
        y[aux == 0] = np.nan
        theta = np.concatenate([[9.0] if fam.dist == 1 else [], rng.uniform(100, 300, fam.na), rng.uniform(1e-3, 5e-3, fam.nb)])
    else:
        if rng.random() < 0.5:
            aux = np.exp(rng.normal(0, 0.4, size=(n, m)))
            aux[rng.random((n, m)) < 0.02] = 0.0
        theta = np.concatenate([[9.0] if fam.dist == 1 else [], rng.uniform(0.5, 2.5, fam.na)])
    assert len(theta) == P
    return y, aux, fam, design, theta


SHAPES = [(1, 3, 1), (17, 1, 1), (5, 2, 2), (300, 5, 3), (1025, 7, 4), (333, 16, 5), (257, 33, 6),
          (200, 48, 8), (64, 100, 2), (40, 250, 3), (2049, 12, 7), (150, 40, 10), (99, 24, 12), (130, 64, 16), (20, 200, 16)]
SHAPE_MODELS = ["Model_t", "Model_normal", "Model_independent", "Model_t_per_sample", "Model_t_patseq",
                "Model_t_v1_patseq", "Model_normal_patseq_v3", "Model_t_patseq_per_sample",
                "Model_t_patseq_v1_per_sample"]


@pytest.mark.parametrize("model", SHAPE_MODELS)
def test_shapes_against_host_build(model):
    """Every kernel over odd shapes (n = 1 ... 2049, m = 1 ... 250, c = 1 ... 16, ~5 % missing, zero
    weights): the CUDA result against the host build of the same arithmetic (tests/hostcheck),
    which test_host_math.py pins to the reference goldens."""
    import hostcheck
    rng = np.random.default_rng(zlib.crc32(model.encode()))      # stable across processes, unlike hash()
    for n, m, c in SHAPES:
        if c >= m and m > 1:
            c = m - 1
        if c > m:
            c = m
        y, aux, fam, design, theta = _random_case(rng, n, m, c, model)
        coef_design = design
        host = hostcheck.HostEvaluator()
        host.upload(y, aux, fam, None, design)
        e = _native.Engine(0)
        try:
            info = e.upload(y, aux, fam, None, design)
            hs, hd, hp = host.per_gene(theta)
            gs, gd, gp = e.per_gene(theta)
            what = "%s n=%d m=%d c=%d" % (model, n, m, c)
            assert_close(gs, hs, 1e-10, what=what + " score")
            assert np.array_equal(gp, hp), what
            hv, hg, hu, _ = host.nll_grad(theta)
            gv, gg, gu, gb = e.nll_grad(theta)
            assert gu == hu and gb == 0, what
            if hu:
                assert_close(gv, hv, 1e-11, what=what + " value")
                assert_close(gg, hg, 1e-8, atol=1e-10 * max(abs(hv), 1.0), what=what + " grad")
            hn, ha = host.noise_stats(theta)
            gn, ga = e.noise_stats(theta)
            assert_close(gn, hn, 1e-9, what=what + " noise p")
            assert_close(ga, ha, 1e-12, what=what + " averages")
            hc, hcov, hdf = host.coef(theta, coef_design)
            gc, gcov, gdf = e.coef(theta, coef_design)
            # genes whose retained sub-design is almost collinear (huge coefficients) carry the
            # condition number times the fma-vs-no-fma rounding difference between device and host
            tame = np.all(np.abs(np.nan_to_num(hc)) < 100.0, axis=1)
            assert np.array_equal(np.isnan(gc), np.isnan(hc)), what + " coef NaN pattern"
            assert_close(gc[tame], hc[tame], 1e-9, atol=1e-9, what=what + " coef")
            assert_close(gc[~tame], hc[~tame], 1e-5, atol=1e-9, what=what + " coef (ill-conditioned rows)")
            assert_close(gcov[tame], hcov[tame], 1e-9, atol=1e-12, what=what + " covar")
            assert_close(gcov[~tame], hcov[~tame], 1e-5, atol=1e-12, what=what + " covar (ill-conditioned rows)")
            assert_close(gdf, hdf, 1e-12, what=what + " df")
            contrasts = np.eye(c)[:, : min(c, 2)]
            hcon, _, hpv = host.test(contrasts)
            gcon, gpv, gq = e.test_bh(contrasts)
            assert_close(gcon[tame], hcon[tame], 1e-9, atol=1e-9, what=what + " contrasts")
            assert_close(gpv[tame], hpv[tame], 1e-8, what=what + " p")
            assert_close(gpv[~tame], hpv[~tame], 1e-4, what=what + " p (ill-conditioned rows)")
            assert np.array_equal(gq, oracle.fdr(gpv)), what + " q"
            assert_close(e.weights(theta), host.weights(theta), 1e-12, what=what + " weights")
        finally:
            e.close()


def test_random_sweep():
    """A short run of tools/fuzz_parity.py: random models, shapes, missing rates, weights, control
    genes and tile geometries against the host build (the long runs are kept in profiles/)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(ROOT, "tools", "fuzz_parity.py"))
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    done, skipped, failed = fuzz.run(60.0, 2024, max_cases=150)
    assert failed == 0 and done - skipped >= 100


def test_degenerate_inputs():
    """All-missing data, all-zero weights, a single gene: nothing contributes, nothing crashes."""
    fam = fitnoise.Model_t()._family(6)
    design = np.array([[1.0, 0.0]] * 3 + [[1.0, 1.0]] * 3)
    e = _native.Engine(0)
    try:
        y = np.full((10, 6), np.nan)
        info = e.upload(y, None, fam, None, design)
        assert info.n_eligible == 0
        v, g, used, bad = e.nll_grad([5.0, 1.0])
        assert used == 0 and bad == 0 and v == 0.0 and np.all(g == 0.0)
        score, d2, p = e.per_gene([5.0, 1.0])
        assert np.all(np.isnan(score))
        coef, covar, df = e.coef([5.0, 1.0], design)
        assert np.all(np.isnan(coef))
        con, pv, q = e.test_bh(np.array([[0.0], [1.0]]))
        assert np.all(np.isnan(pv)) and np.all(q == 1.0)
        y = np.random.default_rng(0).normal(size=(10, 6))
        info = e.upload(y, np.zeros((10, 6)), fam, None, design)
        assert info.n_eligible == 0
        with pytest.raises(_native.NativeError):
            e.nll_grad([5.0, -1.0])
        with pytest.raises(_native.NativeError):
            e.nll_grad([5.0])
        with pytest.raises(_native.NativeError):
            e.upload(y, None, fam, None, np.ones((6, 2)))          # rank-deficient design
    finally:
        e.close()
    with pytest.raises(AssertionError):
        fitnoise.Model_t().fit(np.zeros((3, 6)), np.ones((5, 1)))   # design rows != samples


def test_empty_shard():
    """n = 0 (a rank that owns no gene when there are fewer genes than GPUs) goes through every call."""
    fam = fitnoise.Model_t()._family(6)
    design = np.array([[1.0, 0.0]] * 3 + [[1.0, 1.0]] * 3)
    e = _native.Engine(0)
    try:
        info = e.upload(np.zeros((0, 6)), None, fam, None, design)
        assert info.n_eligible == 0
        v, g, used, bad = e.nll_grad([5.0, 1.0])
        assert (v, used, bad) == (0.0, 0, 0) and np.all(g == 0.0)
        assert all(len(a) == 0 for a in e.per_gene([5.0, 1.0]))
        assert all(len(a) == 0 for a in e.noise_stats([5.0, 1.0]))
        coef, covar, df = e.coef([5.0, 1.0], design)
        assert coef.shape == (0, 2) and covar.shape == (0, 2, 2)
        con, pv, q = e.test_bh(np.array([[0.0], [1.0]]))
        assert con.shape == (0, 1) and len(pv) == 0 and len(q) == 0
        assert e.weights([5.0, 1.0]).shape == (0, 6)
        pat = fitnoise.Model_t_patseq()._family(6)
        e.upload(np.zeros((0, 6)), np.zeros((0, 6)), pat, None, design)
        v, g, used, bad = e.nll_grad([5.0, 100.0, 0.01])
        assert used == 0 and v == 0.0
    finally:
        e.close()
    fitted = fitnoise.Model_t().fit(np.zeros((0, 6)), design, force_param=[5.0, 1.0])
    tested = fitted.test(coef=[1])
    assert fitted.coef.shape == (0, 2) and len(tested.q_values) == 0 and fitted.noise_combined_p_value == 0.0


def test_top_table_order_bit_exact(engine):
    """The top-table order is the stable ascending order of p with missing p last (R's order(),
    backyard/old_fitnoise.R:822-823), bit-exact with numpy's stable argsort."""
    rng = np.random.default_rng(3)
    for n in (1, 5, 300, 4097, 250000):
        p = rng.random(n) ** 4
        if n > 4:
            p[rng.integers(0, n, n // 10 + 1)] = np.nan
            p[rng.integers(0, n, n // 5 + 1)] = p[1]        # many ties
            p[2] = 1.0
            p[3] = 0.0
        q, order = engine.bh_ranked(p)
        assert np.array_equal(order.astype(np.int64), np.argsort(p, kind="stable")), n
        assert np.array_equal(q, oracle.fdr(p)), n
    g = load_golden("t_per_sample_controls")
    fitted, tested = fit_golden(g, force_param=g["theta"])
    table = tested.top_table()
    assert np.array_equal(np.asarray(table.index), np.argsort(tested.p_values, kind="stable"))
    assert list(table.columns) == ["coef2", "AveExpr", "P.Value", "adj.P.Val", "noise.P.Value", "control"]
    assert_close(table["adj.P.Val"].values, g["ref_q_values"][np.asarray(table.index)], 1e-8)


@pytest.mark.parametrize("lpg,threads", [(0, 0), (1, 128), (2, 256), (4, 256), (8, 256)])
def test_factor_models_lane_mappings(lpg, threads):
    """The factor-model kernel (augmented ridge system, d/dF column sums) under every lane mapping and
    at m = 32 (blocked order), n = 700, against the host build of the same arithmetic."""
    import hostcheck
    rng = np.random.default_rng(21)
    n, m, nf = 700, 32, 2
    design = np.stack([np.ones(m), (np.arange(m) >= 16) * 1.0, rng.normal(size=m)], axis=1)
    F = rng.normal(size=(m, nf)) * 0.8
    y = rng.normal(size=(n, m)) + (rng.normal(size=(n, nf)) @ F.T) + 5.0
    y[rng.random((n, m)) < 0.03] = np.nan
    w = np.exp(rng.normal(0, 0.3, size=(n, m)))
    controls = rng.random(n) < 0.3
    control_design = design[:, [0, 2]]
    fam = fitnoise.Model_t_factors(nf)._family(m)
    theta = np.array([7.0, 1.2])
    host = hostcheck.HostEvaluator()
    host.upload(y, w, fam, controls, design, control_design)
    hv, hg, hgf, hu, _ = host.nll_grad_factors(theta, F)
    hs, hd, hp = host.per_gene(theta)
    hc, hcov, hdf = host.coef(theta, design)
    e = _native.Engine(0)
    try:
        e.set_tuning(lpg=lpg, threads=threads)
        e.upload(y, w, fam, controls, design, control_design)
        gv, gg, ggf, gu, gb = e.nll_grad_factors(theta, F)
        assert gu == hu and gb == 0
        assert_close(gv, hv, 1e-11, what="value")
        assert_close(gg, hg, 1e-8, atol=1e-10 * abs(hv), what="grad core")
        assert_close(ggf, hgf, 1e-8, atol=1e-10 * abs(hv), what="grad factors")
        gs, gd, gp = e.per_gene(theta)
        assert_close(gs, hs, 1e-10, what="score")
        assert np.array_equal(gp, hp)
        gc, gcov, gdf = e.coef(theta, design)
        assert_close(gc, hc, 1e-9, atol=1e-9, what="coef")
        assert_close(gcov, hcov, 1e-9, atol=1e-12, what="covar")
        assert_close(e.weights(theta), host.weights(theta), 1e-12, what="weights")
        with pytest.raises(_native.NativeError):
            e.nll_grad(theta)                      # factor families go through nll_grad_factors
    finally:
        e.close()


@pytest.mark.parametrize("config", [1, 2, 3, 4, 5])
def test_full_size_configs_against_host_build(config):
    """The five BASELINE.json configs at their FULL sizes: objective, gradient, per-gene scores,
    coefficients and p-values of the CUDA path against the host build of the same arithmetic
    (tests/hostcheck, pinned to the reference goldens by test_host_math.py), and BH against numpy."""
    import hostcheck
    s = synthetic(config)
    model = getattr(fitnoise, "Model_t" if config == 5 else s["model"])()
    n, m = s["y"].shape
    fam = model._family(m)
    aux = s["context"].get("weights", s["context"].get("counts"))
    controls, cdesign = s.get("controls"), s.get("control_design")
    theta = {1: [9.0, 1.0], 2: [], 3: [12.0, 750.0, 0.0025], 4: [8.0, 0.5], 5: [10.0, 1.0]}[config]
    theta = np.asarray(theta, dtype=float)
    host = hostcheck.HostEvaluator()
    host.upload(s["y"], aux, fam, controls, s["design"], cdesign)
    e = _native.Engine(0)
    try:
        info = e.upload(s["y"], aux, fam, controls, s["design"], cdesign)
        hs, hd, hp = host.per_gene(theta)
        gs, gd, gp = e.per_gene(theta)
        assert np.array_equal(gp, hp)
        assert_close(gs, hs, 1e-10, what="per-gene score")
        if len(theta):
            hv, hg, hu, _ = host.nll_grad(theta)
            gv, gg, gu, gb = e.nll_grad(theta)
            assert gu == hu == info.n_eligible and gb == 0
            assert_close(gv, hv, 1e-11, what="objective")
            assert_close(gg, hg, 1e-7, atol=1e-11 * abs(hv), what="gradient")
        hc, hcov, hdf = host.coef(theta, s["design"])
        gc, gcov, gdf = e.coef(theta, s["design"])
        assert_close(gc, hc, 1e-9, atol=1e-9, what="coef")
        assert_close(gcov, hcov, 1e-9, atol=1e-13, what="covar")
        t = s["test"]
        contrasts = t.get("contrasts")
        if contrasts is None:
            contrasts = np.zeros((s["design"].shape[1], len(t["coef"])))
            for i, j in enumerate(t["coef"]):
                contrasts[j, i] = 1.0
        hcon, _, hpv = host.test(contrasts)
        gcon, gpv, gq = e.test_bh(contrasts)
        assert_close(gpv, hpv, 1e-8, what="p-values")
        assert np.array_equal(gq, oracle.fdr(gpv)), "BH not bit-exact"
        assert np.array_equal(e.last_order.astype(np.int64), np.argsort(gpv, kind="stable")), "top-table order"
    finally:
        e.close()


@pytest.mark.parametrize("n,m,order", [(1, 3, 2), (500, 6, 1), (3000, 6, 2), (777, 33, 3), (2000, 96, 2),
                                       (300, 130, 2), (120, 250, 1)])
def test_transform_kernel_against_oracle(n, m, order):
    """fitnoise.transform: objective and gradient from the CUDA sums kernel against the oracle's
    dense matrix-calculus restatement of fittransform.py:103-137, and the applied transform."""
    from fitnoise_b200 import fittransform
    rng = np.random.default_rng(n + m + order)
    x = rng.poisson(np.exp(rng.normal(3.0, 1.5, size=(n, 1))) * rng.uniform(0.5, 2.0, size=m)).astype(float)
    c = min(3, m)
    design = np.concatenate([np.ones((m, 1)), rng.normal(size=(m, c - 1))], axis=1)
    pack = np.concatenate([rng.uniform(0.0, 2.0, size=(order - 1) * m), rng.uniform(0.5, 3.0, size=m)])
    q, _ = np.linalg.qr(design)
    e = _native.Engine(0)
    try:
        e.transform_upload(x, design)
        sums = e.transform_sums(order, pack)
        cost, grad, mu = fittransform._combine(sums, n, m, order, q @ q.T)
        y = e.transform_apply(order, pack, mu)
    finally:
        e.close()
    want = oracle.transform_apply(pack, x, order)
    assert_close(y, want - want.mean(axis=0), 1e-10, atol=1e-12, what="applied transform")
    if n > 1:
        ocost, ograd = oracle.transform_cost_grad(pack, x, design, order)
        assert_close(cost, ocost, 1e-10, what="objective")
        assert_close(grad, ograd, 1e-7, atol=1e-11, what="gradient")


def test_transform_fit_on_gpu():
    x = np.random.default_rng(1).poisson(np.exp(np.random.default_rng(2).normal(3, 1.5, size=(5000, 1)))
                                         * np.linspace(0.5, 2.0, 8)).astype(float)
    design = np.stack([np.ones(8), (np.arange(8) >= 4) * 1.0], axis=1)
    fitted = fitnoise.transform(x, design=design, order=2)
    ofit, oy = oracle.transform_fit(x, design, 2)
    assert abs(fitted.optimization_result.fun - ofit.fun) <= 2e-3 * abs(ofit.fun)
    want = oracle.transform_apply(fitted.optimization_result.x, x, 2)
    assert_close(fitted.y, want - want.mean(axis=0), 1e-10, atol=1e-12)


def test_integration_stub_from_the_docs():
    """The ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would add) runs as
    written against the built library and returns what the package's own binding returns."""
    import re
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# fitnoise/_b200.py.*?)```", text, re.S).group(1)
    block = block.replace('ctypes.CDLL("libfitnoise_b200.so")', 'ctypes.CDLL(%r)' % _native.library_path())
    ns = {}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    rng = np.random.default_rng(8)
    n, m = 500, 6
    y = rng.normal(size=(n, m))
    y[3, 2] = np.nan
    design = np.array([[1.0, 0.0]] * 3 + [[1.0, 1.0]] * 3)
    fam = ns["Family"](kind=0, na=1, nb=0, tail_own=0, v3=0, dist=1, nf=0, fixed_df=0.0)      # Model_t
    h = ns["Handle"](0)
    info = h.upload(y, None, fam, None, design, np.ones((m, 1)))
    value, grad = h.nll_grad(np.array([6.0, 1.2]))
    e = _native.Engine(0)
    try:
        info2 = e.upload(y, None, fitnoise.Model_t()._family(m), None, design)
        v2, g2, used, bad = e.nll_grad([6.0, 1.2])
    finally:
        e.close()
    assert info.n_eligible == info2.n_eligible == n and used == n
    assert value == v2 and np.array_equal(grad, g2)
    ns["_check"](ns["_lib"].fn_destroy(h.h))


def test_p_values_from_1_down_to_the_far_tails_against_host_build():
    """The device evaluates the F / chi-square continued fractions with a Newton-refined reciprocal
    (fn_special.cuh cf_rcp) where the host build divides: p-values from ~1 down to 1e-30 (t) and 1e-300 (normal; effects of
    0 ... 150 standard deviations), t and normal models, one and two contrasts, must agree to 1e-10."""
    import hostcheck
    rng = np.random.default_rng(77)
    n, m = 4000, 12
    group = (np.arange(m) >= m // 2) * 1.0
    batch = (np.arange(m) % 2) * 1.0
    design = np.stack([np.ones(m), group, batch], axis=1)
    y = rng.normal(size=(n, m))
    y += np.linspace(0.0, 150.0, n)[:, None] * group[None, :]
    y += rng.normal(size=(n, 1)) * batch[None, :]
    for model, theta in (("Model_t", [6.5, 1.1]), ("Model_normal", [0.9]), ("Model_independent", [])):
        fam = getattr(fitnoise, model)()._family(m)
        theta = np.asarray(theta, dtype=float)
        host = hostcheck.HostEvaluator()
        host.upload(y, None, fam, None, design)
        e = _native.Engine(0)
        try:
            e.upload(y, None, fam, None, design)
            hn, _ = host.noise_stats(theta)
            gn, _ = e.noise_stats(theta)
            assert_close(gn, hn, 1e-10, what=model + " noise p")
            host.coef(theta, design)
            e.coef(theta, design)
            for contrasts in (np.array([[0.0], [1.0], [0.0]]), np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])):
                _, _, hp = host.test(contrasts)
                _, gp, _ = e.test_bh(contrasts)
                assert np.nanmin(hp) < 1e-15 and np.nanmax(hp) > 0.5, (np.nanmin(hp), np.nanmax(hp))   # t tails are heavy; the normal models reach 1e-300
                positive = hp > 0
                assert_close(gp[positive], hp[positive], 1e-10, what="%s p, %d contrast(s)" % (model, contrasts.shape[1]))
                assert np.all(gp[~positive] < 1e-300)
        finally:
            e.close()
