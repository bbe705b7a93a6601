import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from bo_pr_b200 import _capi as capi
from bo_pr_b200.acquisition import acq_desc_of
from tests.product_util import product_acq, product_mc_pr
from tests.random_cases import build_problem
from tests.config_timings import CONFIGS
dev = torch.device("cuda", 0)
lib = capi.load_library()
for name in ("rosenbrock10_scaled pr__ucb", "mixed_int_f1_TR pr__ei"):
    c = dict(CONFIGS[name], b=5, mc=128, seed=3)
    g = build_problem(c)
    acq = product_acq(g, dev); pr = product_mc_pr(g, dev, acq=acq); st = acq.state
    b, mc, d = 5, 128, c["dim"]
    theta = g["theta"].to(dev).reshape(b, d).contiguous()
    rnd = pr.input_transform["round"]; rnd.ensure_base_samples(1, dev); ui, uc = rnd.base_sample_ptrs()
    sp, aq = pr._space(), acq_desc_of(acq)
    lo = torch.zeros(d, dtype=torch.float64, device=dev); hi = torch.ones(d, dtype=torch.float64, device=dev)
    adam = torch.zeros(2 * b * d, dtype=torch.float64, device=dev); value = torch.empty(b, dtype=torch.float64, device=dev)
    ma = torch.zeros(2, dtype=torch.float64, device=dev)
    ws = st.workspace(b * mc, d, b)
    def run(stream, iters):
        th = theta.clone(); adam.zero_()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        capi.check(lib.prb_mc_adam(st.handle, C.byref(sp), C.byref(aq), th.data_ptr(), b, mc, capi.ptr(ui), capi.ptr(uc), 0, 0, 1, ma.data_ptr(), lo.data_ptr(), hi.data_ptr(), 0.025, iters, 0, adam.data_ptr(), value.data_ptr(), ws.data_ptr(), ws.numel(), 0, stream), "adam")
        torch.cuda.synchronize(); return time.perf_counter() - t0
    side = torch.cuda.Stream()
    run(side.cuda_stream, 200); print(name, "graph path 200 steps: %.2f ms" % (run(side.cuda_stream, 200) * 1e3))
    lib.prb_profile_enable(1); run(None, 50)
    msk = (C.c_double * 12)(); cnt = (C.c_uint64 * 12)(); lib.prb_profile_read(msk, cnt, 12); lib.prb_profile_enable(0)
    print({nm: (round(msk[i] / 51 * 1e3, 1), int(cnt[i]) // 51) for i, nm in enumerate(capi.PROFILE_SLOT_NAMES) if cnt[i]}, "us per step (plain launches, events)")
    ts = (C.c_longlong * 16)()
    run(side.cuda_stream, 200)
    lib.prb_debug_small_timestamps(ts)
    t = [ts[i] for i in range(14)]
    seq = [11, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13]
    names = ["state reads", "prep", "gen", "wait A1", "lower", "acq epilogue", "upper", "grad contraction", "partials",
             "fence+ticket", "finish chain", "grid barrier"]
    print("steady-state step of CTA 0 (cycles):", {names[i]: t[seq[i + 1]] - t[seq[i]] for i in range(12)}, "total", t[13] - t[11])
