#!/usr/bin/env python
"""Benchmark of the PR acquisition hot path (BASELINE.json: "PR acqf value+grad evals/sec (cand x MC)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one MC-PR EI value+gradient evaluation of `restarts` restarts x `mc` MC samples against an exact GP with
n_train training points (d = 13: 3 continuous + 10 binary, kernel "mixed") -- BASELINE.json configs[4] at the size its
target names (n_train = 1000, 64 restarts, 1024 MC samples).  One eval = one (candidate x MC sample) base-acquisition
value+gradient.  Multi-GPU: restarts are sharded, GP state replicated, no data-path collective (weak scaling: 64
restarts PER GPU); NCCL is used once, after the timed region, for the best-candidate allgather/argmax.

`value`      device-resident inputs, straight through the C ABI (prb_mc_value_grad), CUDA events, max over ranks.
`e2e`        the same metric through the reference-facing API (MCProbabilisticReparameterization.forward + autograd),
             theta in pinned HOST memory, host<->device copies inside the timed region.
`roofline`   the dominant kernel (the FP64 DMMA triangular products), timed live with CUDA events per launch.
`cpu_baseline` / `--impl reference`: the CPU oracle (port of the reference path: reference L1 semantics over restated
             BoTorch/GPyTorch arithmetic; the reference is pure Python and BoTorch is not installed, so there is no
             oracle/_ref) on the box's host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "pr_acqf_value_grad_evals_per_sec"
UNIT = "evals/s"
FP64_PEAK_TFLOPS = 37.1  # measured on this pool's B200: DMMA microbenchmark, profiles/r01_fp64_peak_microbench.txt
NCU_TRAFFIC_BYTES_PER_LAUNCH = (17.44e6 + 484.57e6 + 572.04e6 + 15.42e6) / 2  # profiles/r01_trgemm_tma_ncu.txt
FP64_PEAK_SOURCE = ("measured FP64 DMMA peak (tools/microbench/fp64_peak.cu, profiles/r01_fp64_peak_microbench.txt; "
                    "cuBLAS DGEMM 35.5 TF/s, profiles/r01_cublas_dgemm_peak.json); MEASURED_PEAKS.json has no FP64 entry")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-train", type=int, default=1000)
    ap.add_argument("--restarts", type=int, default=64, help="restarts per GPU")
    ap.add_argument("--mc", type=int, default=1024)
    ap.add_argument("--cpu-restarts", type=int, default=8, help="restarts in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-bo-iter", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"synthetic PR-EI value+grad: n_train={a.n_train}, d=13 (3 continuous + 10 binary, kernel 'mixed'), "
            f"{a.restarts} restarts/GPU x {a.mc} MC samples, tau=0.1, reinforce_ma")


def algorithmic_flops_per_eval(n: int) -> float:
    """F_vg(n) = 2 n^2 (two triangular products) + ~100 n (kernel, dots, gradient)."""
    return 2.0 * n * n + 100.0 * n


# ---------------------------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_oracle_step_fn(problem, mc, restarts, seed=1):
    from bench_workload import make_theta, sobol_base_samples
    from oracle import pr_oracle as P
    from tests.golden_util import oracle_acq, oracle_space

    torch.set_num_threads(os.cpu_count() or 1)
    space, oacq = oracle_space(problem), oracle_acq(problem)
    theta = make_theta(restarts, seed=seed)
    base = sobol_base_samples(mc)
    state = {"c": 0.0, "h": 0.0}

    def step():
        # restarts in chunks of 2 to bound the oracle's [B, n, d] temporaries (the reference chunks too: batch_limit)
        vals = []
        for s in range(0, restarts, 2):
            r = P.mc_pr(space, theta[s:s + 2], oacq, base, None, grad_estimator="reinforce_ma", ma_counter=state["c"],
                        ma_hidden=state["h"])
            vals.append(r.value)
        state["c"], state["h"] = r.ma_counter, r.ma_hidden
        return torch.cat(vals)

    return step


def time_cpu(step, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bench_workload import make_problem

    problem = make_problem(a.n_train, seed=0)
    problem["case"]["mc"] = a.mc
    b = a.cpu_restarts
    step = cpu_oracle_step_fn(problem, a.mc, b)
    sec = time_cpu(step, a.steps, a.warmup)
    value = b * a.mc / sec
    cores = torch.get_num_threads()
    sample = f"{b} of the {a.restarts} restarts x {a.mc} MC samples per step (value+grad), torch CPU float64"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "n_train": a.n_train, "restarts_per_gpu": a.restarts, "mc": a.mc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# Second metric of BASELINE.json: seconds per BO iteration on ackley13 pr__ei -- the AF-optimisation segment the reference
# runs at run_one_replication.py:494-528 (1024 raw samples screened in chunks of 32, 20 restarts x 200 Adam steps in
# batches of 5, candidate sampling, re-scoring with the base EI, argmax), GP fit excluded, at the end-of-run size n = 119.
# ---------------------------------------------------------------------------------------------------------------------
BO_OPTS = {"batch_limit": 5, "init_batch_limit": 32, "maxiter": 200, "nonnegative": True}


def bo_iter_problem():
    from bench_workload import make_problem, sobol_base_samples
    g = make_problem(119, seed=5)
    g["case"].update(mc=128, grad_estimator="reinforce_ma")
    g["base_samples"] = sobol_base_samples(128)
    g["base_samples_categorical"] = None
    return g


def bo_iter_ours(dev, one_batch=False, repeats=3):
    import bo_pr_b200 as B
    from tests.product_util import product_acq, product_mc_pr
    g = bo_iter_problem()
    acq = product_acq(g, dev)
    bounds = torch.stack([torch.zeros(13, dtype=torch.float64), torch.ones(13, dtype=torch.float64)]).to(dev)
    times = []
    for rep in range(repeats + 1):
        pr = product_mc_pr(g, dev, acq=acq)  # a fresh AF object per BO iteration, as the reference builds one
        opts = dict(BO_OPTS, seed=rep)
        if one_batch:
            opts.pop("batch_limit")
            opts.pop("init_batch_limit")
        torch.manual_seed(rep)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cand, vals = B.optimize_acqf(pr, bounds, q=1, num_restarts=20, raw_samples=1024, options=opts, stochastic=True,
                                     return_best_only=False)
        final = pr.sample_candidates(cand)
        with torch.no_grad():
            best = acq(final).argmax()
        chosen = final[best].cpu()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    return sorted(times[1:])[len(times[1:]) // 2], chosen


def bo_iter_cpu():
    from oracle import pr_oracle as P
    from bo_pr_b200.initializers import draw_sobol_samples, initialize_q_batch_nonneg
    from tests.golden_util import oracle_acq, oracle_space
    torch.set_num_threads(os.cpu_count() or 1)
    g = bo_iter_problem()
    space, oacq = oracle_space(g), oracle_acq(g)
    base = g["base_samples"]
    st = {"c": 0.0, "h": 0.0}

    def call(X, with_grad):
        r = P.mc_pr(space, X, oacq, base, None, grad_estimator="reinforce_ma", ma_counter=st["c"], ma_hidden=st["h"],
                    with_grad=with_grad)
        st["c"], st["h"] = r.ma_counter, r.ma_hidden
        return r

    torch.manual_seed(0)
    t0 = time.perf_counter()
    bounds = torch.stack([torch.zeros(13, dtype=torch.float64), torch.ones(13, dtype=torch.float64)])
    X_rnd = draw_sobol_samples(bounds, 1024, 1, seed=0)
    Y = torch.cat([call(X_rnd[s:s + 32], False).value for s in range(0, 1024, 32)])
    ics = initialize_q_batch_nonneg(X_rnd, Y, 20)
    outs = []
    for s in range(0, 20, 5):
        theta, _ = P.adam_multistart(lambda X: (lambda r: (r.value, r.grad))(call(X, True)),
                                     lambda X: call(X, False).value, ics[s:s + 5], bounds[0], bounds[1], maxiter=200)
        outs.append(theta)
    cand = torch.cat(outs)
    final = cand.clone()
    final[..., 3:] = torch.bernoulli(P.rounding_prob(space, cand))  # sample_candidates (:198-233), binary dims
    with torch.no_grad():
        oacq(final).argmax()
    return time.perf_counter() - t0


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(", ") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(a):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: bo_pr_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from bench_workload import make_problem, make_theta
    from bo_pr_b200 import _capi as capi
    from bo_pr_b200.acquisition import acq_desc_of
    from tests.product_util import product_acq, product_mc_pr  # thin constructors over the public API

    lib = capi.load_library()
    n, b, mc = a.n_train, a.restarts, a.mc
    problem = make_problem(n, seed=0, device=dev)
    problem["case"]["mc"] = mc
    problem["base_samples"] = None
    problem["base_samples_categorical"] = None
    acq = product_acq(problem, dev)
    state = acq.state  # GP state replicated on every rank (built from the same seeded data)
    theta = make_theta(b, seed=1 + rank)  # this rank's restarts
    theta_dev = theta.to(dev).reshape(b, -1).contiguous()
    d = theta_dev.shape[-1]

    # ---- cpu_baseline leg, part 1 (rank 0, N = 1): before the CPU port is timed as the baseline, a slice of this very
    # workload is checked against it (the oracle is the checker here, never the thing measured as `value`)
    pr = product_mc_pr(dict(problem, case=dict(problem["case"], grad_estimator="reinforce_ma")), dev, acq=acq)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import pr_oracle as P
        from tests.golden_util import oracle_acq, oracle_space
        from tests.product_util import close
        rnd = pr.input_transform["round"]
        rnd.ensure_base_samples(1, dev)
        Xs = theta[:2].clone()
        ref = P.mc_pr(oracle_space(problem), Xs, oracle_acq(problem), rnd.base_samples.cpu(), None,
                      grad_estimator="reinforce_ma")
        Xd = Xs.to(dev).requires_grad_(True)
        v = pr(Xd)
        gr = torch.autograd.grad(v.sum(), Xd)[0]
        ok1, w1 = close(v, ref.value, 1e-9, 1e-12)
        ok2, w2 = close(gr, ref.grad, 1e-9, 1e-11)
        if not (ok1 and ok2):
            raise AssertionError(f"parity gate failed: value {w1:.3g}x, grad {w2:.3g}x tolerance")
        pr._ma_state.zero_()

    # ---- device-resident throughput through the C ABI (Philox drawn on the fly: u_int = u_cat = NULL) -----------------
    sp, aq = pr._space(), acq_desc_of(acq)
    ones = torch.ones(b, dtype=torch.float64, device=dev)
    value = torch.empty(b, dtype=torch.float64, device=dev)
    grad = torch.empty(b, d, dtype=torch.float64, device=dev)
    ma = torch.zeros(2, dtype=torch.float64, device=dev)
    ws = state.workspace(b * mc, d, b)
    stream = torch.cuda.current_stream()

    def step():
        capi.check(lib.prb_mc_value_grad(state.handle, C.byref(sp), C.byref(aq), theta_dev.data_ptr(), b, mc, None, None,
                                         0, 0, capi.PRB_EST_REINFORCE_MA, ma.data_ptr(), 1, ones.data_ptr(),
                                         value.data_ptr(), grad.data_ptr(), None, ws.data_ptr(), ws.numel(), 0,
                                         stream.cuda_stream), "prb_mc_value_grad")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = lib.prb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    barrier()
    launches = int(lib.prb_launch_count() - l0)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / a.steps
    evals_per_step = world * b * mc
    value_evals = evals_per_step / (ms_per_step * 1e-3)

    # ---- value-only evaluation (raw-sample screening, gen_batch_initial_conditions): lower product only -----------------
    value_only = None
    if rank == 0:
        def vstep():
            capi.check(lib.prb_mc_value_grad(state.handle, C.byref(sp), C.byref(aq), theta_dev.data_ptr(), b, mc, None,
                                             None, 0, 0, capi.PRB_EST_REINFORCE_MA, ma.data_ptr(), 1, None,
                                             value.data_ptr(), None, None, ws.data_ptr(), ws.numel(), 0,
                                             stream.cuda_stream), "prb_mc_value_grad")
        for _ in range(a.warmup):
            vstep()
        torch.cuda.synchronize()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record(stream)
        for _ in range(a.steps):
            vstep()
        v1.record(stream)
        torch.cuda.synchronize()
        vms = v0.elapsed_time(v1) / a.steps
        # F_v(n) = n^2 + ~70 n flop per evaluation (one triangular product + kernel + dots)
        value_only = {"value": b * mc / (vms * 1e-3), "unit": UNIT, "ms_per_step": vms, "n_gpus": 1,
                      "roofline_frac": (b * mc * (n * n + 70.0 * n) / (vms * 1e-3)) / (FP64_PEAK_TFLOPS * 1e12),
                      "note": "value-only calls (no gradient): the screening of raw samples before the multi-start "
                              "optimisation; fraction of the FP64 DMMA peak on F_v(n) = n^2 + 70 n flop per evaluation"}

    # ---- the same step with MC de-duplication (PRB_OPT_DEDUP): reported separately, never as `value` ---------------------
    dedup = None
    if rank == 0:
        z = torch.empty(b, mc, pr._space().n_int, dtype=torch.uint8, device=dev)
        capi.check(lib.prb_pr_sample(C.byref(sp), theta_dev.data_ptr(), b, mc, None, None, 0, 0, None, z.data_ptr(), None,
                                     stream.cuda_stream), "prb_pr_sample")
        codes = (z.to(torch.int64) << torch.arange(z.shape[-1], device=dev)).sum(-1)
        n_unique = sum(int(torch.unique(codes[r]).numel()) for r in range(b))
        state.set_option(capi.PRB_OPT_DEDUP, 1)
        for _ in range(a.warmup):
            step()
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record(stream)
        for _ in range(a.steps):
            step()
        d1.record(stream)
        torch.cuda.synchronize()
        state.set_option(capi.PRB_OPT_DEDUP, 0)
        dms = d0.elapsed_time(d1) / a.steps
        dedup = {"value": b * mc / (dms * 1e-3), "unit": UNIT, "ms_per_step": dms, "n_gpus": 1,
                 "unique_columns": n_unique, "unique_fraction": n_unique / (b * mc),
                 "note": "PRB_OPT_DEDUP: one base-AF evaluation per unique (restart, discrete configuration), weighted by "
                         "multiplicity; identical results up to summation order; theta ~ U[0,1] (as PR converges the "
                         "unique fraction falls further). evals/s still counted per candidate x MC sample."}

    # ---- end to end through the reference-facing API, host buffers ----------------------------------------------------
    e2e = None
    if not a.no_e2e:
        theta_host = theta.clone().pin_memory()
        val_host = torch.empty(b, dtype=torch.float64).pin_memory()
        grad_host = torch.empty(b, 1, d, dtype=torch.float64).pin_memory()

        def e2e_step():
            X = theta_host.to(dev, non_blocking=True).requires_grad_(True)
            v = pr(X)
            g = torch.autograd.grad(v.sum(), X)[0]
            val_host.copy_(v.detach(), non_blocking=True)
            grad_host.copy_(g, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(a.warmup):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": evals_per_step / (float(t) / a.steps), "unit": UNIT,
               "h2d_bytes_per_step": world * theta_host.numel() * 8,
               "d2h_bytes_per_step": world * (val_host.numel() + grad_host.numel()) * 8,
               "ms_per_step": float(t) / a.steps * 1e3}
    clocks = sampler.stop() if sampler is not None else None

    # ---- roofline of the dominant kernel: per-launch CUDA events inside the library (separate pass) --------------------
    roofline, breakdown = None, None
    if rank == 0:
        torch.cuda.synchronize()
        lib.prb_profile_enable(1)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        msk = (C.c_double * capi.PRB_PROFILE_SLOTS)()
        cnt = (C.c_uint64 * capi.PRB_PROFILE_SLOTS)()
        capi.check(lib.prb_profile_read(msk, cnt, capi.PRB_PROFILE_SLOTS), "prb_profile_read")
        lib.prb_profile_enable(0)
        breakdown = {nm: {"ms_per_step": msk[i] / 3, "launches_per_step": int(cnt[i]) // 3}
                     for i, nm in enumerate(capi.PROFILE_SLOT_NAMES) if cnt[i]}
        gemm_ms = (msk[2] + msk[4]) / 3
        gemm_launches = int(cnt[2] + cnt[4]) // 3
        # algorithmic flops of the two triangular products: n(n+1) flop per column each (+ 2n for the alpha row)
        flops_per_step = b * mc * (2.0 * n * (n + 1) + 2.0 * n)
        achieved = flops_per_step / (gemm_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "trgemm_tma_kernel<LOWER_GEN | UPPER_GRAD> (TMA + mbarrier "
                    "pipeline, FP64 DMMA m8n8k4: Linv k* with k* generated in-kernel; Linv^T v with the gradient "
                    "contraction fused into the epilogue)",
                    "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
                    "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH if (n, b, mc) == (1000, 64, 1024) else None,
                    "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum averaged over the two "
                    "launches of a step (profiles/r01_trgemm_tma_ncu.txt: lower 0.017 + 0.485 GB, upper 0.572 + 0.015 GB);"
                    " algorithmic HBM bytes per launch = the v panel once, 0.537 GB (written by the lower product, read by "
                    "the upper one): 1.01x -- column tiles are scheduled in groups so that the row tiles of a column tile "
                    "re-read the panel from L2 (2.34 GB of DRAM reads per upper launch before the grouped order)",
                    "peak_source": FP64_PEAK_SOURCE,
                    "avg_launch_ms": gemm_ms / max(gemm_launches, 1), "launches_per_step": gemm_launches,
                    "share_of_step": gemm_ms / sum(v["ms_per_step"] for v in breakdown.values()),
                    "algorithmic_flop_per_launch": flops_per_step / max(gemm_launches, 1),
                    "whole_step_frac": evals_per_step / world * algorithmic_flops_per_eval(n) / (ms_per_step * 1e-3)
                    / 1e12 / FP64_PEAK_TFLOPS}

    # ---- the one collective of the path: best-candidate allgather + argmax (not timed) ---------------------------------
    best = None
    if dist is not None:
        from bo_pr_b200.distributed import gather_best
        best_val, best_theta, best_rank = gather_best(value, theta_dev)
        best = {"value": float(best_val), "rank": int(best_rank)}

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cb = a.cpu_restarts
        cpu_problem = dict(problem)
        cstep = cpu_oracle_step_fn(cpu_problem, mc, cb)
        sec = time_cpu(cstep, steps=24, warmup=2)
        cpu = {"value": cb * mc / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{cb} of the {b} restarts x {mc} MC samples per step (value+grad), 24 steps (~10 s), torch CPU float64"}

    # ---- seconds per BO iteration, ackley13 pr__ei (rank 0, N = 1 only) ---------------------------------------------------
    bo = None
    if rank == 0 and world == 1 and not a.no_bo_iter:
        t_ref_sem, _ = bo_iter_ours(dev, one_batch=False)
        t_one, _ = bo_iter_ours(dev, one_batch=True)
        bo = {"workload": "ackley13-shape pr__ei AF-optimisation segment at n_train=119: 1024 raw samples, 20 restarts x "
                          "200 Adam steps, mc=128, candidate sampling + base-EI re-scoring; GP fit excluded",
              "sec_reference_semantics": t_ref_sem, "note_reference_semantics": "restart batches of 5, screening chunks "
              "of 32 (EMA-baseline trajectory as the reference, SURVEY.md App. C.5); one prb_mc_adam call per batch",
              "sec_all_restarts_one_batch": t_one}
        if not a.no_cpu_baseline:
            bo["sec_cpu_port"] = bo_iter_cpu()
            bo["cpu_cores"] = torch.get_num_threads()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value_evals, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "n_train": n, "restarts_per_gpu": b, "mc": mc,
                       "evals_per_step": evals_per_step, "rng": "philox4x32-10 (seed 0, offset 0), drawn in the sampler",
                       "l2": "the per-step v panel (8 x n x b x mc bytes = "
                             f"{8 * n * b * mc / 1e9:.2f} GB written then read) exceeds the 126 MB L2; the 16 MB GP state "
                             "is L2-resident by design; no explicit flush",
                       "parallelism": f"restart-sharded x{world}, GP state replicated, no data-path collective"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "kernel_breakdown": breakdown, "bo_iter": bo, "dedup": dedup, "value_only": value_only,
        }
        if best is not None:
            line["best_candidate"] = best
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
