"""CPU half of the full-size parity check: the kernels' per-gene arithmetic compiled for the host
(tests/hostcheck) against the oracle's full-size fixture of BASELINE config 1 (20,000 x 6), and the
product's starting values / bounds against the reference's (``ref_initial`` / ``ref_bounds`` of the
golden fixtures)."""
import os

import numpy as np
import pytest

import fitnoise_b200 as fitnoise
import hostcheck
from conftest import GOLDEN_DIR, assert_close, golden_names, load_golden
from helpers import fit_golden, synthetic


def test_cfg1_full_size_host_arithmetic_against_oracle_fixture():
    path = os.path.join(GOLDEN_DIR, "full_cfg1.npz")
    with np.load(path, allow_pickle=False) as z:
        g = {k: z[k] for k in z.files}
    s = synthetic(1)
    parts = [np.nansum(s["y"]), float(np.isnan(s["y"]).sum()), np.nansum(s["y"][::97] ** 2)]
    if not np.allclose(parts, g["input_checksum"], rtol=1e-13, atol=0):
        pytest.skip("this numpy regenerates other inputs than the fixture was made from")
    model = fitnoise.Model_t()._with(_engine_factory=hostcheck.HostEvaluator)
    forced = model.fit(s["y"], s["design"], force_param=g["theta"])
    tested = forced.test(coef=[1])
    rows = np.arange(0, s["y"].shape[0], int(g["stride"]))
    value, grad, used, bad = forced._objective(g["theta"])
    assert used == int(g["n_prepared"]) and bad == 0
    assert_close(value, g["total_at_theta"], 1e-11, what="objective")
    assert_close(grad, g["grad_at_theta"], 1e-8, atol=1e-9 * abs(float(g["total_at_theta"])), what="gradient")
    assert_close(forced.noise_p_values[rows], g["noise_p"], 1e-8, what="noise p")
    assert_close(forced.coef[rows], g["coef"], 1e-9, atol=1e-10, what="coef")
    assert_close(tested.p_values[rows], g["p_values"], 1e-8, what="p")
    assert int((tested.q_values < 0.05).sum()) == int(g["q_lt_0_05"])


@pytest.mark.parametrize("name", [n for n in golden_names() if n != "t_nan_quirk"])
def test_product_initial_and_bounds_match_the_reference(name):
    """Model_*._configured of the PRODUCT (fitnoise_b200/models.py _initial_bounds) against the
    reference's _initial / _bounds (fitnoise.py:538-552 ... 843-861, t mix-in :431-436)."""
    g = load_golden(name)
    fitted, _ = fit_golden(g, engine_factory=hostcheck.HostEvaluator, force_param=g["theta"])
    initial = np.asarray(fitted._initial, dtype=float)
    bounds = np.asarray([[np.nan if v is None else v for v in b] for b in fitted._bounds], dtype=float).reshape(-1, 2)
    ref_initial, ref_bounds = g["ref_initial"], g["ref_bounds"].reshape(-1, 2)
    assert initial.shape == ref_initial.shape
    # with NaN in y the reference's own start is NaN (documented divergence 2): compare where it is finite
    ok = np.isfinite(ref_initial)
    assert_close(initial[ok], ref_initial[ok], 1e-12, what="initial")
    okb = np.isfinite(ref_bounds).all(axis=1) | np.isnan(ref_bounds).all(axis=1)
    assert_close(bounds[okb], ref_bounds[okb], 1e-12, what="bounds")
