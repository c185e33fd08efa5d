"""Full-size oracle fixtures for BASELINE.json configs 1-3 (parity gap 1 and 2).

    python tests/golden/make_fullsize.py [cfg1 cfg2 cfg2t cfg3]        (minutes of CPU each)

For each config the inputs are regenerated from the seed (tests/helpers.synthetic: 20,000 x 6,
20,000 x 12 with weights, 50,000 x 24 PAT-Seq with 15 % missing) and the ORACLE
(oracle/fitnoise_oracle.py: the reference's per-gene route, itself pinned by the 26 reference
fixtures of make_golden.py) runs the complete fit + test at full size: hyper-parameters by the
reference's tight optimiser loop (fitnoise.py:143-158), per-gene scores, noise p-values,
coefficients, p-values.  The fixture keeps what a parity test needs and no more (<= 0.4 MB each):

  * theta, objective value, df, number of genes in the REML sum;
  * per-gene score_row / noise_p / coef / p on every STRIDE-th gene, plus the sums of the full arrays;
  * a checksum of the regenerated inputs (the test refuses to compare if numpy regenerates others).

q-values are not stored: BH is checked bit-exactly against the oracle's fdr() on the device's own p.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STRIDE = 8

CASES = {
    "cfg1": dict(config=1, model="Model_t"),
    "cfg2": dict(config=2, model="Model_independent"),
    "cfg2t": dict(config=2, model="Model_t"),
    "cfg3": dict(config=3, model="Model_t_patseq"),
    "cfg4": dict(config=4, model="Model_t"),          # 200,000 x 48, c = 6, half the genes are controls (~1 h)
}


def input_checksum(s):
    parts = [np.nansum(s["y"]), float(np.isnan(s["y"]).sum()), np.nansum(s["y"][::97] ** 2)]
    for key in sorted(s["context"]):
        parts.append(np.nansum(s["context"][key]))
    return np.array(parts)


def make(name):
    from helpers import synthetic
    from oracle import fitnoise_oracle as oracle
    case = CASES[name]
    s = synthetic(case["config"])
    t0 = time.time()
    out = oracle.fit_and_test(case["model"], s["y"], s["design"], s["context"],
                              coef_idx=s["test"].get("coef"), contrasts=s["test"].get("contrasts"),
                              control_design=s.get("control_design"), controls=s.get("controls"),
                              nan_aware=True, skip_rank_deficient=True)
    theta = out["x"]
    cfg, prepared = out["cfg"], out["prepared"]
    n = s["y"].shape[0]
    score = np.full(n, np.nan)
    for row, item in prepared.items():
        score[row] = oracle.score_row(cfg, theta, row, *item)
    total, grad = (oracle.total_score_grad(cfg, prepared, theta) if len(theta) else (0.0, np.zeros(0)))
    rows = np.arange(0, n, STRIDE)
    np.savez_compressed(
        os.path.join(HERE, "full_%s.npz" % name),
        config=case["config"], model=case["model"], stride=STRIDE, input_checksum=input_checksum(s),
        theta=theta, fun=out["fun"], df=out["df"], n_prepared=len(prepared),
        total_at_theta=total, grad_at_theta=grad,
        noise_combined_p_value=out["noise_combined_p_value"],
        score_rows=score[rows], score_sum=np.nansum(score),
        noise_p=out["noise_p_values"][rows], noise_p_sum=np.nansum(out["noise_p_values"]),
        averages=out["averages"][rows],
        coef=out["coef"][rows], coef_sum=np.nansum(out["coef"], axis=0),
        coef_df=out["coef_df"][rows],
        contrasts=out["contrasts"][rows],
        p_values=out["p_values"][rows], p_sum=np.nansum(out["p_values"]),
        n_nan_p=int(np.isnan(out["p_values"]).sum()),
        q_lt_0_05=int((out["q_values"] < 0.05).sum()), q_sum=float(np.sum(out["q_values"])),
        seconds=time.time() - t0)
    print(name, "theta", theta, "fun", out["fun"], "in %.0f s" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or [nm for nm in CASES if nm != "cfg4"]
    if len(names) == 1:
        make(names[0])
    else:
        import subprocess
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), nm]) for nm in names]
        sys.exit(max(p.wait() for p in procs))
