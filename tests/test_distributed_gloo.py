"""The N > 1 path on CPU: two ranks over gloo, genes sharded, sums all-reduced, per-gene results
all-gathered.  Each rank's local evaluator is the host build of the kernel mathematics
(tests/hostcheck) -- the CUDA evaluator takes its place on a GPU box (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from conftest import ROOT, assert_close, golden_context, golden_fit_kwargs, golden_test_kwargs, load_golden

CASES = ["t_twogroup", "t_patseq_v2", "t_per_sample_controls", "t_factors1_controls"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run_case(name, forced):
    import fitnoise_b200 as fitnoise
    import hostcheck
    g = load_golden(name)
    nf = int(g.get("n_factors", 0))
    cls = getattr(fitnoise, str(g["model"]))
    model = (cls(nf) if nf else cls())._with(_engine_factory=hostcheck.HostEvaluator)
    kw = golden_fit_kwargs(g)
    if forced:
        kw["force_param"] = g["theta"]
    fitted = model.fit(fitnoise.Dataset(g["y"], golden_context(g)), design=g["design"], **kw)
    tested = fitted.test(**golden_test_kwargs(g))
    value, grad, used, bad = fitted._objective(g["theta"])
    return dict(x=np.asarray(fitted.optimization_result.x), value=value, grad=grad, used=used,
                noise_p=fitted.noise_p_values, averages=fitted.averages, coef=fitted.coef,
                covar=fitted._coef_covar, p=tested.p_values, q=tested.q_values, weights=fitted.get_weights(),
                combined=fitted.noise_combined_p_value, description=fitted.description,
                shard=(fitted._session.lo, fitted._session.hi, fitted._session.world))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        results = {}
        for name in CASES:
            results[name + ":forced"] = _run_case(name, True)
        results["t_twogroup:fit"] = _run_case("t_twogroup", False)
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), results, allow_pickle=True)
    finally:
        dist.destroy_process_group()


def test_two_ranks_match_one(tmp_path):
    single = {name + ":forced": _run_case(name, True) for name in CASES}
    single["t_twogroup:fit"] = _run_case("t_twogroup", False)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ranks = [np.load(os.path.join(str(tmp_path), "rank%d.npy" % r), allow_pickle=True).item() for r in range(2)]
    for key, one in single.items():
        for r in range(2):
            two = ranks[r][key]
            n = len(one["noise_p"])
            assert two["shard"][2] == 2 and two["shard"][1] - two["shard"][0] in (n // 2, n - n // 2)
            assert two["used"] == one["used"]
            assert_close(two["value"], one["value"], 1e-13, what=key + " value")
            assert_close(two["grad"], one["grad"], 1e-10, atol=1e-13 * abs(one["value"]), what=key + " grad")
            # per-gene results are computed by the owning rank and gathered: identical
            for field in ("noise_p", "averages", "coef", "covar", "p", "q", "weights"):
                if key.endswith(":forced"):
                    assert np.array_equal(two[field], one[field], equal_nan=True), "%s %s" % (key, field)
                else:       # theta itself differs in the last digits between 1 and 2 ranks
                    assert_close(two[field], one[field], 1e-7, what="%s %s" % (key, field))
            assert_close(two["combined"], one["combined"], 1e-7)
            assert_close(two["x"], one["x"], 1e-9, what=key + " theta")
        assert ranks[0][key]["description"] == ranks[1][key]["description"]


def _replicated_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["FN_SHARD_MIN_MB"] = "64"               # every case is "small": fitted whole by every rank
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        results = {"t_twogroup:fit": _run_case("t_twogroup", False), "independent_weights:forced": _run_case("independent_weights", True)}
        np.save(os.path.join(out_dir, "rep%d.npy" % rank), results, allow_pickle=True)
    finally:
        dist.destroy_process_group()


def test_small_problems_are_replicated_not_sharded(tmp_path):
    """Below FN_SHARD_MIN_MB every rank fits all genes itself (session.py): same numbers as one
    process, bit for bit, and no collective is entered (a hang here would be one)."""
    single = {"t_twogroup:fit": _run_case("t_twogroup", False), "independent_weights:forced": _run_case("independent_weights", True)}
    port = _free_port()
    mp.spawn(_replicated_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rep%d.npy" % r), allow_pickle=True).item()
        for key, one in single.items():
            two = got[key]
            assert two["shard"] == (0, len(one["noise_p"]), 1)
            assert two["value"] == one["value"] and np.array_equal(two["x"], one["x"])
            for field in ("noise_p", "averages", "coef", "covar", "p", "q", "weights", "grad"):
                assert np.array_equal(two[field], one[field], equal_nan=True), "%s %s" % (key, field)


def _tiny_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fitnoise_b200 as fitnoise
        import hostcheck
        rng = np.random.default_rng(5)
        y = rng.normal(size=(1, 6))                       # ONE gene on two ranks: rank 1 owns nothing
        design = np.array([[1.0, 0.0]] * 3 + [[1.0, 1.0]] * 3)
        model = fitnoise.Model_t()._with(_engine_factory=hostcheck.HostEvaluator)
        fitted = model.fit(y, design, force_param=[5.0, 1.0])
        tested = fitted.test(coef=[1])
        value, grad, used, bad = fitted._objective([5.0, 1.0])
        np.save(os.path.join(out_dir, "tiny%d.npy" % rank),
                dict(shard=(fitted._session.lo, fitted._session.hi), coef=fitted.coef, p=tested.p_values,
                     q=tested.q_values, value=value, used=used), allow_pickle=True)
    finally:
        dist.destroy_process_group()


def test_fewer_genes_than_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_tiny_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(os.path.join(str(tmp_path), "tiny%d.npy" % k), allow_pickle=True).item() for k in range(2)]
    assert r[0]["shard"] == (0, 1) and r[1]["shard"] == (1, 1)
    for k in range(2):
        assert r[k]["coef"].shape == (1, 2) and r[k]["used"] == 1
        assert np.array_equal(r[k]["p"], r[0]["p"]) and np.array_equal(r[k]["q"], r[0]["q"])
    assert r[0]["value"] == r[1]["value"]


def _transform_case():
    import fitnoise_b200 as fitnoise
    import hostcheck
    rng = np.random.default_rng(9)
    n, m = 301, 6
    group = (np.arange(m) >= 3) * 1.0
    mu = np.exp(rng.normal(3.0, 1.2, size=(n, 1)))
    x = rng.poisson(mu * np.exp(0.4 * group)[None, :]).astype(float)
    design = np.stack([np.ones(m), group], axis=1)
    fitted = fitnoise.Transform_polynomial(2)._with(_engine_factory=hostcheck.HostEvaluator).fit(x, design)
    return dict(y=fitted.y, pack=np.asarray(fitted.optimization_result.x), fun=fitted.optimization_result.fun)


def _transform_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        np.save(os.path.join(out_dir, "tr%d.npy" % rank), _transform_case(), allow_pickle=True)
    finally:
        dist.destroy_process_group()


def test_transform_two_ranks(tmp_path):
    """fitnoise.transform with the count matrix sharded over two ranks: the sums are all-reduced, so
    both ranks follow the same optimiser path and gather the same transformed matrix."""
    one = _transform_case()
    port = _free_port()
    mp.spawn(_transform_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(os.path.join(str(tmp_path), "tr%d.npy" % k), allow_pickle=True).item() for k in range(2)]
    assert np.array_equal(r[0]["pack"], r[1]["pack"]) and np.array_equal(r[0]["y"], r[1]["y"])
    assert r[0]["y"].shape == one["y"].shape
    assert abs(r[0]["fun"] - one["fun"]) <= 1e-3 * abs(one["fun"])
    assert np.corrcoef(r[0]["y"].reshape(-1), one["y"].reshape(-1))[0, 1] > 0.9999


def test_shard_bounds():
    from fitnoise_b200 import parallel
    for n, w in [(10, 1), (10, 3), (7, 8), (1000001, 8), (0, 2)]:
        b = parallel.shard_bounds(n, w)
        assert b[0] == 0 and b[-1] == n and len(b) == w + 1
        sizes = np.diff(b)
        assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1
