#!/usr/bin/env python
"""Benchmark of the Fitnoise hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): gene-likelihood-evals/sec.  A *step* is ONE evaluation of the REML
objective and its gradient over every gene of the workload (what the optimiser calls per proposed
theta; reference fitnoise.py:125-136).  Workload at N = 1: BASELINE.json configs[4], 1,000,000
genes x 96 samples, two-group design, on one GPU; the likelihood-and-gradient evaluation is timed
with Model_t (P = 2) because the literal Model_independent has no hyper-parameters and never
evaluates a likelihood; the literal Model_independent fit + test wall time is reported
beside it under "fit_test".  With N > 1 (torchrun, one rank per GPU) every rank holds its own
1M x 96 shard (weak scaling) and each step ends with the sum of the P+3 partial results over the
ranks: fused into the kernel epilogue over NVLink peer memory (default), or kernel + ncclAllReduce
(FN_REDUCE=nccl, also the fallback when peer mapping is unavailable).

`value`   : genes of all ranks / device time per step, inputs resident in HBM (y is 768 MB per GPU,
            six times the L2, so every step streams from DRAM; no extra L2 flush needed).
`e2e`     : the same through the C-ABI with HOST buffers: every step copies the step's y from pinned
            host memory (fn_upload: H2D + preparation kernel), evaluates, and reads the result back.
`roofline`: algorithmic bytes of the nll+grad kernel (8 m + 1 per gene: the y row and the status
            byte) / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline`: the oracle port (the reference's per-gene numpy algorithm) on a bounded gene sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GENES = int(os.environ.get("FN_BENCH_GENES", 1000000))
M = 96
THETA = np.array([8.0, 1.1])
WORKLOAD = "cfg5: %d genes x %d samples per GPU, 2-group design, Model_t likelihood+gradient evaluation" % (N_GENES, M)


def design_matrix():
    return np.array([[1.0, 0.0]] * (M // 2) + [[1.0, 1.0]] * (M // 2))


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if t.get("genes") == N_GENES and t.get("samples") == M:
            return float(t["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.path = gpu, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    out["sm_max_mhz"] = float(parts[2])
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [v for v in sm if v >= 0.5 * max(sm)] or sm
            out["sm_mhz"] = float(np.median(busy))
        out["reasons"] = sorted(reasons)
        return out


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's per-gene algorithm
# --------------------------------------------------------------------------------------------
def _cpu_chunk(args):
    seed, genes = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import fitnoise_oracle as oracle
    rng = np.random.default_rng(seed)
    y = rng.normal(size=(genes, M)) * np.sqrt(10.0 / rng.chisquare(10.0, size=(genes, 1)))
    design = design_matrix()
    cfg = oracle.Configured("Model_t", y)
    prepared = oracle.prepare_rows(cfg, [design] * genes)       # fitnoise.py:23-35 (not timed: done once per fit)
    t0 = time.perf_counter()
    value, grad = oracle.total_score_grad(cfg, prepared, THETA)  # one objective + gradient evaluation
    return time.perf_counter() - t0, value


def cpu_evals_per_sec(cores, genes_per_core, pool=None):
    """Gene-likelihood evaluations per second of the oracle port on `cores` processes."""
    if cores == 1:
        dt, _ = _cpu_chunk((1, genes_per_core))
        return genes_per_core / dt
    res = pool.map(_cpu_chunk, [(100 + i, genes_per_core) for i in range(cores)])
    # per-process evaluation time excludes data generation and the one-off QR prep
    slowest = max(r[0] for r in res)
    return cores * genes_per_core / slowest


def run_reference(args):
    """The reference arm: the oracle port of the reference's CPU algorithm (the reference itself is
    Python-2 source and cannot travel to the GPU box) on all host cores, each
    step one objective+gradient evaluation over a bounded sample of the workload's genes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = max(1, min(cores, 64))
    genes_per_core = int(os.environ.get("FN_REF_GENES_PER_CORE", 1500))
    import multiprocessing as mp
    pool = mp.get_context("spawn").Pool(cores) if cores > 1 else None
    try:
        for _ in range(min(args.warmup, 1)):
            cpu_evals_per_sec(cores, max(50, genes_per_core // 10), pool)
        rates, t0 = [], time.perf_counter()
        for _ in range(max(args.steps, 1)):
            rates.append(cpu_evals_per_sec(cores, genes_per_core, pool))
            if time.perf_counter() - t0 > 120:
                break
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    value = float(np.median(rates))
    sample = "%d genes x %d samples per step (%d processes x %d genes), one objective+gradient evaluation each" % (
        cores * genes_per_core, M, cores, genes_per_core)
    line = {
        "impl": "reference", "metric": "gene_likelihood_evals_per_sec", "value": value, "unit": "gene-evals/s",
        "n_gpus": args.gpus, "steps": len(rates), "warmup": args.warmup,
        "ms_per_step": 1000.0 * cores * genes_per_core / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "gene-evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "gene-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from fitnoise_b200 import _native
    import fitnoise_b200 as fitnoise

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; fitnoise_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    n, m = N_GENES, M
    design = design_matrix()
    fam = fitnoise.Model_t()._family(m)
    P = fam.theta_length()

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        gen = torch.Generator(device=dev)
        gen.manual_seed(2019 + rank)
        y_dev = torch.randn((n, m), dtype=torch.float64, device=dev, generator=gen)
        chi = torch.distributions.Chi2(torch.tensor(10.0, dtype=torch.float64, device=dev)).sample((n, 1))
        y_dev = y_dev * torch.sqrt(10.0 / chi)
        y_dev[: n // 10, m // 2:] += 1.0
        y_host = torch.empty((n, m), dtype=torch.float64, pin_memory=True)
        y_host.copy_(y_dev)
        stream.synchronize()
        y_np = y_host.numpy()

        engine = _native.Engine(local_rank, stream.cuda_stream)
        engine.upload_dev(n, m, y_dev.data_ptr(), 0, fam, 0, design)
        del y_dev
        reduce_buf = torch.zeros(P + 3, dtype=torch.float64, device=dev)

        # N > 1: the all-reduce of the P+3 sums is fused into the kernel epilogue (peer stores over
        # NVLink into IPC-mapped mailboxes); FN_REDUCE=nccl times kernel + ncclAllReduce instead.
        reduce_mode = os.environ.get("FN_REDUCE", "fused") if world > 1 else "none"
        if reduce_mode == "fused":
            # every rank must end up in the same mode: agree on whether the IPC mapping worked
            ok, why, mine = 1, "", None
            try:
                mine = engine.comm_mailbox(world, 8 + 2 * m)
            except Exception as exc:            # e.g. CUDA IPC not permitted on this box
                ok, why = 0, str(exc)
            handles = [None] * world
            dist.all_gather_object(handles, mine)       # every rank takes part, whatever happened
            if ok and all(h is not None for h in handles):
                try:
                    engine.comm_connect(rank, world, b"".join(handles))
                except Exception as exc:        # peer access not available
                    ok, why = 0, str(exc)
            else:
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int64, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                try:
                    engine.comm_disconnect()
                except Exception:
                    pass
                reduce_mode = "nccl"
                if rank == 0:
                    sys.stderr.write("fused exchange unavailable (%s); using ncclAllReduce\n" % why)
            dist.barrier()

        def step():
            if reduce_mode == "nccl":
                engine.nll_grad_dev(THETA, reduce_buf.data_ptr())
                dist.all_reduce(reduce_buf)
                return reduce_buf.cpu().numpy()
            return engine.nll_grad(THETA)          # (value, grad, n_used, n_bad): totals over all ranks, on the host

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # clocks are sampled from the start of the warm-up to the end of the timed region; the
        # warm-up runs for at least ~0.7 s so that the 100 ms sampler sees the GPU under this load
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        warm = max(args.warmup, 3)
        for _ in range(warm):
            first = step()
        # keep the GPU under this load for ~0.7 s more (same step count on every rank) so that the
        # 100 ms clock sampler sees it; still warm-up, not timed
        t_one = time.perf_counter()
        step()
        t_one = max(time.perf_counter() - t_one, 1e-5)
        extra = int(min(0.7 / t_one, 20000))
        if world > 1:
            t = torch.tensor([extra], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            extra = int(t.item())
        for _ in range(extra):
            step()
        if reduce_mode == "nccl":
            first = (first[0], first[1:P + 1], first[P + 1], first[P + 2])
        assert int(first[2]) == n * world and int(first[3]) == 0, first
        launches0 = engine.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kernel_ms = []
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            out = step()
        ev1.record(stream)
        barrier()
        elapsed_ms = ev0.elapsed_time(ev1)
        launches = engine.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        if world > 1:
            t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed_ms = float(t.item())
        ms_per_step = elapsed_ms / args.steps
        value = n * world / (ms_per_step / 1000.0)

        # kernel-only duration for the roofline: a second pass over the same launches with the
        # library's CUDA events around each kernel (on the stream it is launched on).  Kept out of the
        # timed region above because the extra event records and synchronisation slow the step down.
        engine.set_timing(True)
        for _ in range(max(10, min(args.steps, 200))):
            step()
            kernel_ms.append(engine.last_kernel_ms())
        engine.set_timing(False)
        kms_queued = engine.time_nll_grad(THETA, 100)      # back to back: no launch-latency gap in the events
        kms = float(np.mean(kernel_ms))
        alg_bytes = n * (8 * m + 1)
        peak, peak_src = measured_peak()
        achieved = alg_bytes / (kms * 1e-3) / 1e9

        # e2e through the C-ABI with host buffers (pinned): H2D + prepare + evaluate + read back
        e2e_steps = max(2, min(args.steps, 5))
        engine.upload(y_np, None, fam, None, design)
        engine.nll_grad(THETA)
        barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(e2e_steps):
            engine.upload(y_np, None, fam, None, design)
            res = step()
        ev1.record(stream)
        barrier()
        e2e_ms = ev0.elapsed_time(ev1) / e2e_steps
        e2e_wall_ms = (time.perf_counter() - t0) * 1000.0 / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_ms, e2e_wall_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms, e2e_wall_ms = float(t[0].item()), float(t[1].item())
        e2e_ms = max(e2e_ms, e2e_wall_ms)      # the host-side staging of pageable->device is wall time
        e2e_value = n * world / (e2e_ms / 1000.0)

        # the literal configs[4]: Model_independent fit + coef test through the public API (rank 0 only, N = 1)
        fit_test = None
        if world == 1 and not os.environ.get("FN_BENCH_SKIP_FIT"):
            engine.close()
            t_ind = []
            for _ in range(2):          # first call allocates the pooled engine's buffers; second is warm
                t0 = time.perf_counter()
                fitted = fitnoise.Model_independent().fit(y_np, design)
                tested = fitted.test(coef=[1])
                t_ind.append(time.perf_counter() - t0)
                n_sig = int((tested.q_values < 0.01).sum())
                fitted._session.close()
            t0 = time.perf_counter()
            fitted_t = fitnoise.Model_t().fit(y_np, design)
            tested_t = fitted_t.test(coef=[1])
            t_t = time.perf_counter() - t0
            fit_test = {
                "Model_independent_fit_test_wall_s": t_ind[1], "Model_independent_first_call_wall_s": t_ind[0],
                "q_lt_0.01": n_sig,
                "Model_t_fit_test_wall_s": t_t, "Model_t_evaluations": int(fitted_t._evaluations),
                "Model_t_df": float(fitted_t.param.df), "Model_t_sd": float(np.sqrt(fitted_t.param.variance)),
                "includes": "H2D of y (768 MB) from pinned host memory, preparation, optimiser, noise "
                            "p-values, coefficients, test, BH, D2H of all per-gene results",
            }
            fitted_t._session.close()

    cpu_baseline = None
    if rank == 0 and world == 1 and not os.environ.get("FN_BENCH_SKIP_CPU"):
        genes = int(os.environ.get("FN_CPU_GENES", 6000))
        os.environ["OMP_NUM_THREADS"] = "1"
        rate = cpu_evals_per_sec(1, genes)
        cpu_baseline = {"value": rate, "unit": "gene-evals/s", "cores": 1, "kind": "port",
                        "sample": "%d genes x %d samples, one objective+gradient evaluation of the oracle port "
                                  "(per-gene numpy loop: inv/det of the 94x94 covariance)" % (genes, m)}

    if rank == 0:
        line = {
            "metric": "gene_likelihood_evals_per_sec", "value": value, "unit": "gene-evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "genes_total": n * world, "samples": m, "design_columns": 2,
                       "theta_len": P, "l2": "inputs (768 MB per GPU) larger than L2; no flush needed",
                       "parallelism": "genes sharded, %d rank(s), all-reduce of %d doubles per step (%s)" % (world, P + 3, reduce_mode)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "gene-evals/s", "h2d_bytes_per_step": int(n * m * 8),
                    "d2h_bytes_per_step": int((P + 3) * 8), "ms_per_step": e2e_ms},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(), "kernel": "fn::eval_kernel<2, SCALE1>", "kernel_ms": kms,
                         "kernel_ms_queued": kms_queued, "achieved_queued": alg_bytes / (kms_queued * 1e-3) / 1e9,
                         "note": "achieved/frac use kernel_ms: CUDA events around each launch on an idle stream "
                                 "(includes the launch latency); *_queued: 100 launches back to back between two events",
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
            "cpu_baseline": cpu_baseline,
            "fit_test": fit_test,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
