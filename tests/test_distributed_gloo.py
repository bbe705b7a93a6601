"""world_size-2 gloo test of the restart sharding + best-candidate exchange (CPU tensors; the NCCL path is the same
code over CUDA tensors)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bo_pr_b200.distributed import gather_best, gather_restarts, shard_restarts


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_restarts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        X = torch.rand(n_restarts, 1, 13, generator=g, dtype=torch.float64)  # identical on every rank
        vals_all = (X.sum(dim=(-1, -2)) * 7.0).sin()  # stand-in objective, one value per restart
        mine = shard_restarts(X)
        my_vals = (mine.sum(dim=(-1, -2)) * 7.0).sin()
        back = gather_restarts(my_vals, n_restarts)
        backX = gather_restarts(mine, n_restarts)
        bv, bx, br = gather_best(my_vals, mine)
        i = int(torch.argmax(vals_all))
        ok = (torch.equal(back, vals_all) and torch.equal(backX, X) and float(bv) == float(vals_all[i])
              and torch.equal(bx, X[i]) and br == i % world)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


class _ToyAcq(torch.nn.Module):
    """Deterministic stand-in acquisition function on CPU tensors (the sharding logic is device-agnostic)."""

    def forward(self, X):
        return (X.sum(dim=(-1, -2)) * 3.0).sin() + 1.5


def _worker_init(rank, world, port, n_raw, q):
    """Sharded raw-sample screening: every rank must return the SAME initial conditions, equal to what a single process
    computes from the same seed and rank 0's generator state."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bo_pr_b200.initializers import gen_batch_initial_conditions
        bounds = torch.stack([torch.zeros(5, dtype=torch.float64), torch.ones(5, dtype=torch.float64)])
        torch.manual_seed(100 + rank)  # ranks deliberately disagree: only rank 0's generator may matter
        if rank == 0:
            torch.manual_seed(100)
        ics = gen_batch_initial_conditions(_ToyAcq(), bounds, q=1, num_restarts=6, raw_samples=n_raw,
                                           options={"seed": 11, "init_batch_limit": 8, "nonnegative": True}, shard=True)
        torch.manual_seed(100)
        ref = gen_batch_initial_conditions(_ToyAcq(), bounds, q=1, num_restarts=6, raw_samples=n_raw,
                                           options={"seed": 11, "init_batch_limit": 8, "nonnegative": True}, shard=False)
        bufs = [torch.empty_like(ics) for _ in range(world)]
        dist.all_gather(bufs, ics)
        same = all(torch.equal(b, bufs[0]) for b in bufs)
        q.put((rank, bool(same and torch.equal(ics, ref))))
    finally:
        dist.destroy_process_group()


def _worker_mc(rank, world, port, b, q):
    """MC-sharded Adam: the per-step all-reduce of the local MC means reproduces the unsharded trajectory (up to the order of
    the two-term sum)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bo_pr_b200.distributed import allreduce_mean_, mc_sharded_adam
        g = torch.Generator().manual_seed(3)
        mc, d = 16, 4
        u = torch.rand(mc, d, generator=g, dtype=torch.float64)  # "base samples", identical on every rank
        ics = torch.rand(b, 1, d, generator=g, dtype=torch.float64)
        lo, hi = torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)
        state = {"h": torch.zeros(1, dtype=torch.float64)}

        def make(us, st):
            def f(X):  # MC mean over the slice of a smooth per-sample objective
                return (-(X - us.unsqueeze(0)).pow(2).sum(-1) + 0.1 * st["h"]).mean(-1)

            def vg(X):
                Xg = X.detach().requires_grad_(True)
                v = f(Xg)
                gr = torch.autograd.grad(v.sum(), Xg)[0]
                st["h"] = st["h"] - 0.3 * (st["h"] - v.detach().mean())  # EMA-like state, linear in the call mean
                return v.detach(), gr

            return f, vg

        f_loc, vg_loc = make(u[rank::world], state)
        out, vals = mc_sharded_adam(vg_loc, lambda X: f_loc(X).detach(), lambda: allreduce_mean_(state["h"]), ics, lo, hi,
                                    lr=0.05, maxiter=25)
        st1 = {"h": torch.zeros(1, dtype=torch.float64)}
        f_all, vg_all = make(u, st1)
        theta = ics.clone().clamp(0, 1).requires_grad_(True)
        opt = torch.optim.Adam([theta], lr=0.05)
        for _ in range(25):
            X = theta.detach().clamp(0, 1)
            _, gr = vg_all(X)
            opt.zero_grad()
            theta.grad = -gr
            opt.step()
        ref = theta.detach().clamp(0, 1)
        ok = torch.allclose(out, ref, rtol=0, atol=1e-12) and torch.allclose(vals, f_all(ref).detach(), atol=1e-12) \
            and torch.allclose(state["h"], st1["h"], atol=1e-13)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _run(n_restarts, target=None):
    target = _worker if target is None else target
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, 2, port, n_restarts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_shard_gather_best_world2_even():
    _run(20)


def test_shard_gather_best_world2_ragged():
    _run(7)


def test_sharded_raw_sample_screening_world2():
    _run(64, _worker_init)
    _run(37, _worker_init)  # ragged shard


def test_mc_sharded_adam_world2():
    _run(1, _worker_mc)  # fewer restarts than ranks: the case MC sharding exists for


def test_single_process_passthrough():
    X = torch.arange(12, dtype=torch.float64).reshape(4, 1, 3)
    v = torch.tensor([0.1, 0.9, 0.3, 0.2], dtype=torch.float64)
    assert torch.equal(shard_restarts(X), X)
    bv, bx, br = gather_best(v, X)
    assert float(bv) == 0.9 and torch.equal(bx, X[1]) and br == 0
