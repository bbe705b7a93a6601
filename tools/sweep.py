"""Synthetic PR-EI sweep of BASELINE.json configs[4]: n_train 200-4000, 64 restarts IN TOTAL x {128, 1024} MC samples, at
1 / 2 / 4 / 8 GPUs (restart r -> rank r mod G, GP state replicated, no collective inside a step).

    python tools/sweep.py out.json                                                          # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py out.json   # 8 GPUs

One JSON object per shape: CUDA-event timing of prb_mc_value_grad (device-resident inputs), barrier + synchronise on both
sides, max over ranks.  Roofline fraction is against F_vg(n) = 2 n^2 + 100 n flop per evaluation and
the FP64 DMMA peak measured in the same process (prb_debug_fp64_peak)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench_workload import make_problem, make_theta  # noqa: E402
from bo_pr_b200 import _capi as capi  # noqa: E402
from bo_pr_b200.acquisition import acq_desc_of  # noqa: E402
from tests.product_util import product_acq, product_mc_pr  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
lib = capi.load_library()
tf = C.c_double(0.0)
PEAK = 37.1e12
if lib.prb_debug_fp64_peak(0, C.byref(tf), None) == 0:
    PEAK = tf.value * 1e12


def measure(n, total_restarts, mc, steps=5, warmup=3, grad=True):
    g = make_problem(n, seed=0, device=dev)
    g["case"]["mc"] = mc
    g["base_samples"] = g["base_samples_categorical"] = None
    acq = product_acq(g, dev)
    state = acq.state
    pr = product_mc_pr(dict(g, case=dict(g["case"], grad_estimator="reinforce_ma")), dev, acq=acq)
    theta = make_theta(total_restarts)[rank::world].to(dev)
    b = theta.shape[0]
    theta = theta.reshape(b, -1).contiguous()
    d = theta.shape[-1]
    sp, aq = pr._space(), acq_desc_of(acq)
    ones = torch.ones(b, dtype=torch.float64, device=dev)
    value = torch.empty(b, dtype=torch.float64, device=dev)
    gradt = torch.empty(b, d, dtype=torch.float64, device=dev)
    ma = torch.zeros(2, dtype=torch.float64, device=dev)
    ws = state.workspace(b * mc, d, b)
    st = torch.cuda.current_stream()

    def step():
        capi.check(lib.prb_mc_value_grad(state.handle, C.byref(sp), C.byref(aq), theta.data_ptr(), b, mc, None, None, 0, 0,
                                         capi.PRB_EST_REINFORCE_MA, ma.data_ptr(), 1, ones.data_ptr() if grad else None,
                                         value.data_ptr(), gradt.data_ptr() if grad else None, None, ws.data_ptr(),
                                         ws.numel(), 0, st.cuda_stream), "prb_mc_value_grad")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        step()
    e1.record(st)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    evals = total_restarts * mc / (ms * 1e-3)
    flop = (2.0 * n * n + 100.0 * n) if grad else (1.0 * n * n + 70.0 * n)
    return {"n_train": n, "restarts_total": total_restarts, "restarts_per_gpu": b, "mc": mc, "n_gpus": world,
            "mode": "value+grad" if grad else "value", "ms_per_step": ms, "evals_per_s": evals,
            "roofline_frac": evals * flop / (PEAK * world), "workspace_mb": ws.numel() / 2 ** 20}


if __name__ == "__main__":
    out = []
    for n in (200, 500, 1000, 2000, 4000):
        for mc in (128, 1024):
            r = measure(n, 64, mc)
            if rank == 0:
                print(json.dumps(r), flush=True)
            out.append(r)
    if world == 1:
        r = measure(1000, 64, 1024, grad=False)
        print(json.dumps(r), flush=True)
        out.append(r)
    if rank == 0 and len(sys.argv) > 1:
        json.dump({"peak_tflops_live": PEAK / 1e12, "rows": out}, open(sys.argv[1], "w"), indent=1)
    if dist is not None:
        dist.destroy_process_group()
