"""BASELINE.json configs 1-4 at FULL size against the oracle's own full-size run (fixtures made by
tests/golden/make_fullsize.py: the oracle's tight optimiser loop and per-gene route on the same
seeded inputs).  Needs a B200: ``-m gpu``.

Tolerances (north_star): fitted hyper-parameters 1e-6 relative; per-gene log-likelihood,
coefficients and statistics 1e-9; p-values 1e-8; BH bit-exact given p.
"""
import os

import numpy as np
import pytest

import fitnoise_b200 as fitnoise
from conftest import GOLDEN_DIR, assert_close
from helpers import synthetic
from oracle import fitnoise_oracle as oracle

pytestmark = pytest.mark.gpu

CASES = ["cfg1", "cfg2", "cfg2t", "cfg3", "cfg4"]


def load_case(name):
    path = os.path.join(GOLDEN_DIR, "full_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated" % os.path.basename(path))
    with np.load(path, allow_pickle=False) as z:
        g = {k: z[k] for k in z.files}
    s = synthetic(int(g["config"]))
    parts = [np.nansum(s["y"]), float(np.isnan(s["y"]).sum()), np.nansum(s["y"][::97] ** 2)]
    for key in sorted(s["context"]):
        parts.append(np.nansum(s["context"][key]))
    if not np.allclose(parts, g["input_checksum"], rtol=1e-13, atol=0):
        pytest.skip("this numpy regenerates other inputs than the fixture was made from")
    return g, s


@pytest.mark.parametrize("name", CASES)
def test_full_size_against_oracle_fixture(name):
    g, s = load_case(name)
    model = getattr(fitnoise, str(g["model"]))()
    data = fitnoise.Dataset(s["y"], s["context"])
    kw = {k: s[k] for k in ("controls", "control_design") if k in s}
    rows = np.arange(0, s["y"].shape[0], int(g["stride"]))
    theta_ref = g["theta"]

    # (1) the optimiser on the GPU objective lands on the oracle's hyper-parameters
    fitted = model.fit(data, s["design"], **kw)
    if len(theta_ref):
        x = fitted.optimization_result.x
        free = ~(np.isclose(theta_ref, 100.0) | np.isclose(theta_ref, 1e-3))      # at a bound: compare the free ones
        assert_close(x[free], theta_ref[free], 1e-6, what="fitted theta")
        assert np.all(np.isclose(x[~free], theta_ref[~free], rtol=1e-6))
        assert_close(fitted.optimization_result.fun, g["fun"], 1e-10, what="objective at the optimum")
        # (2) objective and gradient at the ORACLE's theta
        value, grad, used, bad = fitted._objective(theta_ref)
        assert used == int(g["n_prepared"]) and bad == 0
        assert_close(value, g["total_at_theta"], 1e-11, what="objective at the oracle's theta")
        assert_close(grad, g["grad_at_theta"], 1e-8, atol=1e-9 * abs(float(g["total_at_theta"])), what="gradient")
    assert fitted.df == int(g["df"])
    fitted._session.close()

    # (3) per-gene results at the oracle's theta (identical theta on both sides)
    forced = model.fit(data, s["design"], force_param=theta_ref, **kw)
    tested = forced.test(**s["test"])
    score, d2, p = forced._session.per_gene(forced._theta)
    assert_close(score[rows], g["score_rows"], 1e-9, what="per-gene score_row")
    assert_close(np.nansum(score), g["score_sum"], 1e-11, what="sum of score_row")
    assert_close(forced.noise_p_values[rows], g["noise_p"], 1e-8, what="noise p-values")
    assert_close(np.nansum(forced.noise_p_values), g["noise_p_sum"], 1e-9, what="sum of noise p")
    assert_close(forced.noise_combined_p_value, g["noise_combined_p_value"], 1e-8, what="combined noise p")
    assert_close(forced.averages[rows], g["averages"], 1e-12, what="averages")
    assert_close(forced.coef[rows], g["coef"], 1e-9, atol=1e-10, what="coef")
    assert_close(np.nansum(forced.coef, axis=0), g["coef_sum"], 1e-9, what="column sums of coef")
    assert_close(forced._coef_df[rows], g["coef_df"], 1e-12, what="posterior df")
    assert_close(tested.contrasts[rows], g["contrasts"], 1e-9, atol=1e-10, what="contrasts")
    assert_close(tested.p_values[rows], g["p_values"], 1e-8, what="p-values")
    assert_close(np.nansum(tested.p_values), g["p_sum"], 1e-9, what="sum of p")
    assert int(np.isnan(tested.p_values).sum()) == int(g["n_nan_p"])
    # (4) BH: bit-exact given the device's own p; and the same discoveries as the oracle's run
    assert np.array_equal(tested.q_values, oracle.fdr(tested.p_values))
    assert int((tested.q_values < 0.05).sum()) == int(g["q_lt_0_05"])
    assert_close(float(np.sum(tested.q_values)), g["q_sum"], 1e-8, what="sum of q")
    # the top-table order is the stable ascending order of p, missing last
    key = np.where(np.isnan(tested.p_values), np.inf, tested.p_values)
    assert np.array_equal(tested._order, np.argsort(key, kind="stable"))
    forced._session.close()
