#!/usr/bin/env python
"""Benchmark of the PR acquisition hot path (BASELINE.json: "PR acqf value+grad evals/sec (cand x MC)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ackley13|rosenbrock|chemistry]

One "step" = one MC-PR value+gradient evaluation of `restarts` restarts x `mc` MC samples against an exact GP with n_train
training points -- BASELINE.json configs[4] at the size its target names (n_train = 1000, d = 13: 3 continuous + 10 binary,
kernel "mixed", EI, 64 restarts, 1024 MC samples).  One eval = one (candidate x MC sample) base-acquisition
value+gradient.

Multi-GPU is the north-star split: the 64 restarts are the TOTAL, restart r -> rank r mod G, the GP
state is replicated, no collective inside a step ("scaling": "strong").  The weak-scaling figure (64 restarts per rank)
is an extra key.  NCCL appears where the path has its exchange: the all-gathers of the sharded raw-sample screening and of
the per-restart results inside the `bo_iter_sharded` leg (optimize_acqf(shard=True)), timed there.

`value`      device-resident inputs, straight through the C ABI (prb_mc_value_grad), CUDA events, max over ranks.
`e2e`        the same metric with HOST buffers: MCProbabilisticReparameterization.value_and_grad_host -> the C ABI's host-buffer
             entry prb_mc_value_grad_host (pinned theta in, pinned value + gradient out; host->device copy, fused step,
             device->host copies and the wait for them inside the timed region, every step).  `e2e.via_forward_autograd`
             is the same step through nn.Module.forward + torch.autograd on device tensors made from the pinned buffers.
`roofline`   the dominant kernel (the FP64 DMMA triangular products), timed live with CUDA events per launch; the
             denominator is measured in the same run (`peak_live`: mma.sync.m8n8k4.f64 from registers) next to the
             library baseline for the same two products (cuBLAS dtrmm / dgemm).
`cpu_baseline` / `--impl reference`: the CPU oracle (port of the reference path: reference L1 semantics over restated
             BoTorch/GPyTorch arithmetic; the reference is pure Python and BoTorch is not installed, so there is no
             oracle/_ref) on the box's host cores.  `--impl reference` runs the SAME step (all restarts).
"""
from __future__ import annotations

import argparse
import ctypes as C
import glob
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "pr_acqf_value_grad_evals_per_sec"
UNIT = "evals/s"
FP64_PEAK_FALLBACK = 37.1  # DMMA microbenchmark of round 1 (profiles/r01_fp64_peak_microbench.txt); used if the live run fails
TRAFFIC_TABLE = os.path.join(ROOT, "profiles", "r02_traffic.json")  # ncu dram bytes per launch, keyed "workload,n,b,mc"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ackley13", choices=["ackley13", "rosenbrock", "chemistry"])
    ap.add_argument("--n-train", type=int, default=1000)
    ap.add_argument("--restarts", type=int, default=64, help="restarts in TOTAL (sharded r mod G over the ranks)")
    ap.add_argument("--mc", type=int, default=1024)
    ap.add_argument("--cpu-restarts", type=int, default=8, help="restarts in the bounded cpu_baseline sample (ours arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-bo-iter", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip value-only / dedup / generic / weak / library legs")
    return ap.parse_args()


WORKLOADS = {
    "ackley13": "d=13 (3 continuous + 10 binary, kernel 'mixed': Matern ARD x isotropic Matern over the binary dims), EI",
    "rosenbrock": "d=10 (4 continuous + 6 ordinals {0..3} inside ONE ARD Matern: the kernel does not separate), UCB",
    "chemistry": "d=22 one-hot / 5 numeric (2 continuous + categoricals 4, 12, 4; kernel 'mixed_categorical'), EI",
}


def workload_name(a):
    return (f"synthetic PR value+grad [{a.workload}]: n_train={a.n_train}, {WORKLOADS[a.workload]}, "
            f"{a.restarts} restarts in total x {a.mc} MC samples, tau=0.1, reinforce_ma")


def algorithmic_flops_per_eval(n: int) -> float:
    """F_vg(n) = 2 n^2 (two triangular products) + ~100 n (kernel, dots, gradient)."""
    return 2.0 * n * n + 100.0 * n


def make_workload(a, device="cpu"):
    """(problem dict, theta_all [restarts, 1, d]) for the named synthetic shape; everything seeded."""
    if a.workload == "ackley13":
        from bench_workload import make_problem, make_theta
        g = make_problem(a.n_train, seed=0, device=device)
        g["case"]["mc"] = a.mc
        g["base_samples"] = g["base_samples_categorical"] = None
        return g, make_theta(a.restarts, seed=1)
    from bench_workload import make_generic_problem
    return make_generic_problem(a.workload, a.n_train, a.restarts, a.mc)


# ---------------------------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_oracle_step_fn(problem, theta, mc, chunk=2, mc_chunk=None):
    """One value+grad step of the CPU oracle over `theta` (all of it), restarts in chunks of `chunk` to bound the
    [B, n, d] temporaries (the reference chunks too: batch_limit); `mc_chunk` = the reference's own pr_batch_limit."""
    from bench_workload import sobol_base_samples
    from oracle import pr_oracle as P
    from tests.golden_util import oracle_acq, oracle_space

    torch.set_num_threads(os.cpu_count() or 1)
    space, oacq = oracle_space(problem), oracle_acq(problem)
    c = problem["case"]
    base_i = problem.get("base_samples")
    base_c = problem.get("base_samples_categorical")
    if base_i is None and len(c["integer_indices"]) > 0:
        base_i = sobol_base_samples(mc, n_cols=len(c["integer_indices"]))
    if base_c is None and c["categorical_features"]:
        base_c = sobol_base_samples(mc, n_cols=len(c["categorical_features"]), seed=4321)
    acq = oacq
    if mc_chunk is not None:
        acq = lambda X: torch.cat([oacq(X[:, s:s + mc_chunk]) for s in range(0, X.shape[1], mc_chunk)], dim=1)  # noqa: E731
    state = {"c": 0.0, "h": 0.0}
    b = theta.shape[0]

    def step():
        vals = []
        for s in range(0, b, chunk):
            r = P.mc_pr(space, theta[s:s + chunk], acq, base_i, base_c, grad_estimator="reinforce_ma",
                        ma_counter=state["c"], ma_hidden=state["h"], apply_numeric=c["apply_numeric"])
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


def config_of(a, world, extra=None):
    cfg = {"workload": workload_name(a), "n_train": a.n_train, "restarts_total": a.restarts, "mc": a.mc,
           "evals_per_step": a.restarts * a.mc}
    cfg.update(extra or {})
    return cfg


def run_reference(a):
    """The reference arm: the CPU port of the path on all host threads, the SAME step as the ours-arm (all restarts)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    problem, theta = make_workload(a)
    step = cpu_oracle_step_fn(problem, theta, a.mc, chunk=2)
    sec = time_cpu(step, a.steps, a.warmup)
    value = a.restarts * a.mc / sec
    cores = torch.get_num_threads()
    # "as shipped" chunking: restart batches of 5, MC chunks of pr_batch_limit = 32 -- bounded sample
    nb = min(10, a.restarts)
    shipped = cpu_oracle_step_fn(problem, theta[:nb], a.mc, chunk=5, mc_chunk=32)
    sec_s = time_cpu(shipped, 2, 1)
    sample = f"all {a.restarts} restarts x {a.mc} MC samples per step (value+grad), restart chunks of 2, torch CPU float64"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_of(a, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "as_shipped": {"value": nb * a.mc / sec_s, "unit": UNIT, "sample": f"{nb} restarts in batches of 5, MC chunks of 32 "
                       "(the reference's batch_limit / pr_batch_limit), 2 steps"},
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
        chosen = B.finalize_candidates(pr, cand)[0].cpu()
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


def bo_iter_sharded(a, problem, acq, dev, dist, world, repeats=2):
    """One acquisition-optimisation segment at the benchmark's n_train with the work split over the ranks as `north_star`
    describes: GP state replicated (each rank builds it from the same data), raw samples AND restarts sharded r mod G, one
    all-gather for the screening values, one for the per-restart results, argmax + candidate sampling + base-AF re-scoring
    replicated.  64 restarts x 1024 raw samples x 128 MC samples, 200 Adam steps.  Wall clock, max over ranks."""
    import bo_pr_b200 as B
    from tests.product_util import product_acq, product_mc_pr
    d = problem["case"]["dim"]
    bounds = torch.stack([torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)]).to(dev)
    g128 = dict(problem, case=dict(problem["case"], mc=128, grad_estimator="reinforce_ma"))

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    best = None
    for rep in range(repeats + 1):
        torch.manual_seed(rep)
        sync()
        t0 = time.perf_counter()
        acq_r = product_acq(problem, dev)  # GP-state replication: K, Cholesky, inverse factor, alpha on this rank
        acq_r.state
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        pr = product_mc_pr(g128, dev, acq=acq_r)
        opts = {"maxiter": 200, "nonnegative": problem["case"]["acq"] == "ei", "seed": rep, "init_batch_limit": 128}
        ics = B.gen_batch_initial_conditions(pr, bounds, q=1, num_restarts=a.restarts, raw_samples=1024, options=opts,
                                             shard=world > 1)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        cand, vals = B.optimize_acqf(pr, bounds, q=1, num_restarts=a.restarts, options=opts, batch_initial_conditions=ics,
                                     stochastic=True, return_best_only=False, shard=world > 1)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        chosen, chosen_val = B.finalize_candidates(pr, cand, seed=rep)
        chosen.cpu()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        ph = torch.tensor([t4 - t0, t1 - t0, t2 - t1, t3 - t2, t4 - t3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ph, op=dist.ReduceOp.MAX)
        if rep > 0 and (best is None or float(ph[0]) < best[0]):
            best = ph.tolist()
    return {"workload": f"{a.workload} shape at n_train={a.n_train}: 1024 raw samples, {a.restarts} restarts x 200 Adam "
                        "steps, mc=128, candidate sampling + base-AF re-scoring on the device",
            "sec": best[0], "phases_sec": {"gp_state_build_replicated": best[1], "raw_sample_screening_sharded": best[2],
                                           "adam_sharded_plus_allgather": best[3], "finalise": best[4]},
            "n_gpus": world, "timing": "wall clock between device synchronisations, max over ranks, best of 2"}


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled every 20 ms WHILE the timed region runs, through NVML in a background thread
    (`nvidia-smi -lms` re-queries a dozen fields per sample, power draw included, and each query stalled the GPU for about a
    millisecond: 2-3 samples inside an 88 ms timed region cost 3 %).  Falls back to nvidia-smi if NVML cannot be loaded."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        import threading
        self.samples, self.mask, self.smax = [], 0, None
        self.stop_flag = False
        self.proc = self.tmp = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None
            self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms",
                                              "200", "-i", str(index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
            except Exception:
                self.proc = None

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": ["no samples"]}
            sm = sorted(self.samples)
            reasons = sorted(nm for bit, nm in self.REASONS.items() if self.mask & bit)
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.smax, "reasons": reasons, "samples": len(sm),
                    "source": "NVML, 20 ms period, during the timed regions"}
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
            for nm, v in zip(names, r[2:6]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 200"}


def cuda_time(fn, steps, warmup, stream):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def library_baseline(n, cols, dev):
    """cuBLAS on the same two products (outside the timed region): dtrmm for Linv K* and Linv^T V (what a library-only
    implementation would call once k* is materialised) and the dense dgemm GPyTorch itself runs.  ms per step (both
    products), or None where the call could not be made."""
    out = {"cublas_dtrmm_ms": None, "cublas_dgemm_ms": None, "cublas_dgemm_tflops_dense": None}
    stream = torch.cuda.current_stream()
    try:
        A = torch.randn(n, n, dtype=torch.float64, device=dev).tril_()
        Bm = torch.randn(cols, n, dtype=torch.float64, device=dev)  # row-major [cols, n] == column-major [n, cols]
        Cm = torch.empty_like(Bm)
        ms = cuda_time(lambda: (torch.matmul(Bm, A.t(), out=Cm), torch.matmul(Cm, A, out=Bm)), 3, 1, stream)
        out["cublas_dgemm_ms"] = ms
        out["cublas_dgemm_tflops_dense"] = 2 * 2.0 * n * n * cols / (ms * 1e-3) / 1e12
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcublas.so*"))
        for sp in sys.path:
            cands += glob.glob(os.path.join(sp, "nvidia", "cublas", "lib", "libcublas.so*"))
        lib = None
        for pth in cands:
            try:
                lib = C.CDLL(pth)
                break
            except OSError:
                continue
        if lib is not None:
            fn = lib.cublasDtrmm_v2
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                           C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            handle = torch.cuda.current_blas_handle()
            one = C.c_double(1.0)

            def trmm(trans, src, dst):  # side = left (0), uplo = lower (0), diag = non-unit (0); column-major [n, cols]
                rc = fn(handle, 0, 0, trans, 0, n, cols, C.addressof(one), A.data_ptr(), n, src.data_ptr(), n,
                        dst.data_ptr(), n)
                if rc != 0:
                    raise RuntimeError(f"cublasDtrmm_v2 status {rc}")

            out["cublas_dtrmm_ms"] = cuda_time(lambda: (trmm(0, Bm, Cm), trmm(1, Cm, Bm)), 3, 1, stream)
    except Exception as e:  # the baseline is a courtesy figure; never fail the bench over it
        out["error"] = repr(e)[:200]
    return out


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

    from bo_pr_b200 import _capi as capi
    from bo_pr_b200.acquisition import acq_desc_of
    from tests.product_util import product_acq, product_mc_pr  # thin constructors over the public API

    lib = capi.load_library()
    n, mc = a.n_train, a.mc
    problem, theta_all = make_workload(a, device=dev)
    acq = product_acq(problem, dev)
    state = acq.state  # GP state replicated on every rank (built from the same seeded data)
    theta = theta_all[rank::world].contiguous()  # restart r -> rank r mod G
    b = theta.shape[0]
    theta_dev = theta.to(dev).reshape(b, -1).contiguous()
    d = theta_dev.shape[-1]
    sep = a.workload == "ackley13"

    # ---- parity gate (rank 0, N = 1): before the CPU port is timed as the baseline, a slice of this very workload is
    # checked against it (the oracle is the checker here, never the thing measured as `value`)
    prob_ma = dict(problem, case=dict(problem["case"], grad_estimator="reinforce_ma"))
    pr = product_mc_pr(prob_ma, dev, acq=acq)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import pr_oracle as P
        from tests.golden_util import oracle_acq, oracle_space
        from tests.product_util import close
        rnd = pr.input_transform["round"]
        rnd.ensure_base_samples(1, dev)
        ui, uc = rnd.base_sample_ptrs()
        Xs = theta[:2].clone()
        ref = P.mc_pr(oracle_space(problem), Xs, oracle_acq(problem), None if ui is None else ui.cpu(),
                      None if uc is None else uc.cpu(), grad_estimator="reinforce_ma",
                      apply_numeric=problem["case"]["apply_numeric"])
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
    stream = torch.cuda.current_stream()

    def make_step(theta_d, with_grad=True):
        bb = theta_d.shape[0]
        ones = torch.ones(bb, dtype=torch.float64, device=dev)
        value = torch.empty(bb, dtype=torch.float64, device=dev)
        grad = torch.empty(bb, d, dtype=torch.float64, device=dev)
        ma = torch.zeros(2, dtype=torch.float64, device=dev)
        ws = state.workspace(bb * mc, d, bb)

        def step():
            capi.check(lib.prb_mc_value_grad(state.handle, C.byref(sp), C.byref(aq), theta_d.data_ptr(), bb, mc, None,
                                             None, 0, 0, capi.PRB_EST_REINFORCE_MA, ma.data_ptr(), 1,
                                             ones.data_ptr() if with_grad else None, value.data_ptr(),
                                             grad.data_ptr() if with_grad else None, None, ws.data_ptr(), ws.numel(), 0,
                                             stream.cuda_stream), "prb_mc_value_grad")
        step.keep = (ones, value, grad, ma, ws)
        return step

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
        barrier()
        l0 = lib.prb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step_fn()
        e1.record(stream)
        barrier()
        launches = torch.tensor([float(lib.prb_launch_count() - l0)], dtype=torch.float64, device=dev)
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(launches, op=dist.ReduceOp.SUM)
        return float(ms) / steps, int(launches)

    step = make_step(theta_dev)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_per_step, launches = timed(step, a.steps, a.warmup)
    evals_per_step = a.restarts * mc
    value_evals = evals_per_step / (ms_per_step * 1e-3)

    # ---- weak scaling (extra): 64 restarts on EVERY rank, as round 1 reported --------------------------------------------
    weak = None
    if world > 1 and not a.no_extras:
        from bench_workload import make_theta
        th_w = (make_theta(a.restarts, seed=1 + rank) if a.workload == "ackley13" else theta_all).to(dev)
        th_w = th_w.reshape(a.restarts, -1).contiguous()
        w_ms, _ = timed(make_step(th_w), max(5, a.steps // 2), a.warmup)
        weak = {"value": world * a.restarts * mc / (w_ms * 1e-3), "unit": UNIT, "ms_per_step": w_ms,
                "restarts_per_gpu": a.restarts, "scaling": "weak"}

    # ---- end to end through the reference-facing API, host buffers ----------------------------------------------------
    e2e = None
    if not a.no_e2e:
        theta_host = theta.clone().pin_memory()
        # one pinned block, [b, d] gradient then [b] values: the library returns both with a single device->host copy
        out_host = torch.empty(b * d + b, dtype=torch.float64).pin_memory()
        grad_host = out_host[: b * d].view(b, 1, d)
        val_host = out_host[b * d:]

        def e2e_step():
            # the host-buffer entry of the C ABI (prb_mc_value_grad_host) through its Python method: pinned theta in, pinned
            # value + gradient out, H2D copy -> fused step -> D2H copies in one call that returns when the results have landed
            pr.value_and_grad_host(theta_host, val_host, grad_host, synchronize=True)

        def e2e_autograd_step():
            # the same step through nn.Module.forward + torch.autograd with device tensors made from the pinned buffers
            X = theta_host.to(dev, non_blocking=True).requires_grad_(True)
            v = pr(X)
            g = torch.autograd.grad(v.sum(), X)[0]
            val_host.copy_(v.detach(), non_blocking=True)
            grad_host.copy_(g, non_blocking=True)
            torch.cuda.synchronize()

        def wall(fn):
            for _ in range(a.warmup):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                fn()
            barrier()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t) / a.steps

        sec = wall(e2e_step)
        sec_ag = wall(e2e_autograd_step)
        e2e = {"value": evals_per_step / sec, "unit": UNIT,
               "h2d_bytes_per_step": a.restarts * d * 8, "d2h_bytes_per_step": a.restarts * (1 + d) * 8,
               "ms_per_step": sec * 1e3,
               "api": "MCProbabilisticReparameterization.value_and_grad_host -> prb_mc_value_grad_host (C ABI, host buffers)",
               "via_forward_autograd": {"value": evals_per_step / sec_ag, "ms_per_step": sec_ag * 1e3,
                                        "api": "pinned theta .to(device) -> forward -> autograd.grad -> .copy_ to pinned"}}
    clocks = sampler.stop() if sampler is not None else None

    # ---- value-only evaluation (raw-sample screening) and MC de-duplication: rank 0, N = 1, reported separately ---------
    value_only = dedup = None
    if rank == 0 and world == 1 and not a.no_extras:
        vms = cuda_time(make_step(theta_dev, with_grad=False), a.steps, a.warmup, stream)
        value_only = {"value": b * mc / (vms * 1e-3), "unit": UNIT, "ms_per_step": vms, "n_gpus": 1,
                      "note": "value-only calls (no gradient): the screening of raw samples before the multi-start "
                              "optimisation; F_v(n) = n^2 + 70 n flop per evaluation (SURVEY.md 8d)",
                      "flop_per_eval": n * n + 70.0 * n}
        if sep:
            z = torch.empty(b, mc, sp.n_int, dtype=torch.uint8, device=dev)
            capi.check(lib.prb_pr_sample(C.byref(sp), theta_dev.data_ptr(), b, mc, None, None, 0, 0, None, z.data_ptr(),
                                         None, stream.cuda_stream), "prb_pr_sample")
            codes = (z.to(torch.int64) << torch.arange(z.shape[-1], device=dev)).sum(-1)
            n_unique = sum(int(torch.unique(codes[r]).numel()) for r in range(b))
            state.set_option(capi.PRB_OPT_DEDUP, 1)
            dms = cuda_time(step, a.steps, a.warmup, stream)
            state.set_option(capi.PRB_OPT_DEDUP, 0)
            dedup = {"value": b * mc / (dms * 1e-3), "unit": UNIT, "ms_per_step": dms, "n_gpus": 1,
                     "unique_columns": n_unique, "unique_fraction": n_unique / (b * mc),
                     "note": "PRB_OPT_DEDUP: one base-AF evaluation per unique (restart, discrete configuration), weighted "
                             "by multiplicity; identical results up to summation order; theta ~ U[0,1] (as PR converges "
                             "the unique fraction falls further). evals/s still counted per candidate x MC sample."}

    # ---- roofline of the dominant kernel: per-launch CUDA events inside the library (separate pass) --------------------
    roofline, breakdown = None, None
    if rank == 0:
        peaks = {}
        for mode, name in ((0, "dmma_m8n8k4"), (1, "dfma"), (2, "dmma_and_dfma_interleaved")):
            tf = C.c_double(0.0)
            if lib.prb_debug_fp64_peak(mode, C.byref(tf), stream.cuda_stream) == 0:
                peaks[name] = tf.value
        peak = peaks.get("dmma_m8n8k4") or FP64_PEAK_FALLBACK
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
        gemm_launches = max(int(cnt[2] + cnt[4]) // 3, 1)
        # algorithmic flops of the two triangular products: n(n+1) flop per column each (+ 2n for the alpha row)
        flops_per_step = b * mc * (2.0 * n * (n + 1) + 2.0 * n)
        achieved = flops_per_step / (gemm_ms * 1e-3) / 1e12
        traffic = None
        try:
            traffic = json.load(open(TRAFFIC_TABLE)).get(f"{a.workload},{n},{b},{mc}")
        except Exception:
            pass
        lib_base = library_baseline(n, b * mc, dev) if not a.no_extras else {}
        roofline = {"bound": "tensor",
                    "kernel": "trgemm_tma_kernel (TMA + mbarrier pipeline, FP64 DMMA m8n8k4): lower Linv k* "
                              + ("with k* generated in-kernel" if sep else "from the k* panel")
                              + "; upper Linv^T v with the gradient contraction fused into the epilogue",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": None if traffic is None else traffic.get("bytes_per_launch"),
                    "traffic_source": None if traffic is None else traffic.get("source"),
                    "algorithmic_hbm_bytes_per_step": 8.0 * (n * (n + 1) / 2 + n * (d + 1)) + 8.0 * b * d + 8.0 * b * (1 + d),
                    "peak_source": "peak_live.dmma_m8n8k4: mma.sync.m8n8k4.f64 from registers on every SM, measured in this "
                                   "run outside the timed region (prb_debug_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                    "peak_live": peaks, "library": lib_base,
                    "avg_launch_ms": gemm_ms / gemm_launches, "launches_per_step": gemm_launches,
                    "share_of_step": gemm_ms / sum(v["ms_per_step"] for v in breakdown.values()),
                    "algorithmic_flop_per_launch": flops_per_step / gemm_launches,
                    "whole_step_frac": b * mc * algorithmic_flops_per_eval(n) / (ms_per_step * 1e-3) / 1e12 / peak}
        if value_only is not None:
            value_only["roofline_frac"] = value_only["value"] * value_only["flop_per_eval"] / (peak * 1e12)

    # ---- one BO iteration with the north-star sharding (all ranks) ---------------------------------------------------------
    bo_sh = None
    if not a.no_bo_iter:
        bo_sh = bo_iter_sharded(a, problem, acq, dev, dist, world)

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cb = min(a.cpu_restarts, a.restarts)
        cpu_problem = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in problem.items()}
        cstep = cpu_oracle_step_fn(cpu_problem, theta_all[:cb], mc)
        sec = time_cpu(cstep, steps=24, warmup=2)
        cpu = {"value": cb * mc / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{cb} of the {a.restarts} restarts x {mc} MC samples per step (value+grad), 24 steps (~10 s), "
                         "torch CPU float64"}

    # ---- the generic (non-separable) kernel structure at the same size: rank 0, N = 1 ---------------------------------------
    generic = None
    if rank == 0 and world == 1 and not a.no_extras and a.workload == "ackley13":
        generic = {}
        for wl in ("rosenbrock", "chemistry"):
            a2 = argparse.Namespace(**dict(vars(a), workload=wl))
            p2, th2 = make_workload(a2, device=dev)
            acq2 = product_acq(p2, dev)
            pr2 = product_mc_pr(dict(p2, case=dict(p2["case"], grad_estimator="reinforce_ma")), dev, acq=acq2)
            sp2, aq2 = pr2._space(), acq_desc_of(acq2)
            th2d = th2.to(dev).reshape(a.restarts, -1).contiguous()
            d2 = th2d.shape[-1]
            st2 = acq2.state
            ones2 = torch.ones(a.restarts, dtype=torch.float64, device=dev)
            v2 = torch.empty(a.restarts, dtype=torch.float64, device=dev)
            g2 = torch.empty(a.restarts, d2, dtype=torch.float64, device=dev)
            ma2 = torch.zeros(2, dtype=torch.float64, device=dev)
            ws2 = st2.workspace(a.restarts * mc, d2, a.restarts)

            def step2():
                capi.check(lib.prb_mc_value_grad(st2.handle, C.byref(sp2), C.byref(aq2), th2d.data_ptr(), a.restarts, mc,
                                                 None, None, 0, 0, capi.PRB_EST_REINFORCE_MA, ma2.data_ptr(), 1,
                                                 ones2.data_ptr(), v2.data_ptr(), g2.data_ptr(), None, ws2.data_ptr(),
                                                 ws2.numel(), 0, stream.cuda_stream), "prb_mc_value_grad")
            gms = cuda_time(step2, max(5, a.steps // 2), a.warmup, stream)
            lib.prb_profile_enable(1)
            step2()
            torch.cuda.synchronize()
            msk = (C.c_double * capi.PRB_PROFILE_SLOTS)()
            cnt = (C.c_uint64 * capi.PRB_PROFILE_SLOTS)()
            capi.check(lib.prb_profile_read(msk, cnt, capi.PRB_PROFILE_SLOTS), "prb_profile_read")
            lib.prb_profile_enable(0)
            generic[wl] = {"workload": WORKLOADS[wl], "value": a.restarts * mc / (gms * 1e-3), "unit": UNIT,
                           "ms_per_step": gms,
                           "whole_step_frac": a.restarts * mc * algorithmic_flops_per_eval(n) / (gms * 1e-3) / 1e12
                           / (roofline["peak"] if roofline else FP64_PEAK_FALLBACK),
                           "kernel_breakdown_ms": {nm: msk[i] for i, nm in enumerate(capi.PROFILE_SLOT_NAMES) if cnt[i]}}
            del st2, ws2, acq2, pr2

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
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_of(a, world, {
                "restarts_per_gpu": b, "rng": "philox4x32-10 (seed 0, offset 0), drawn in the sampler",
                "l2": "the per-step v panel (8 x n x b x mc bytes = "
                      f"{8 * n * b * mc / 1e9:.2f} GB per GPU, written then read) "
                      + ("exceeds the 126 MB L2" if 8 * n * b * mc > 126e6 else "fits the 126 MB L2 (it is an intermediate "
                         "the step itself produces, not an input: inputs are theta and the 16 MB GP state)")
                      + "; no explicit flush",
                "parallelism": f"{a.restarts} restarts in total sharded r mod {world}, GP state replicated, no data-path "
                               "collective inside a step"}),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "kernel_breakdown": breakdown, "bo_iter": bo, "bo_iter_sharded": bo_sh, "weak": weak, "dedup": dedup,
            "value_only": value_only, "generic_path": generic,
        }
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
