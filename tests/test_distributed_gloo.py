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


def _run(n_restarts):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_restarts, q)) for r in range(2)]
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


def test_single_process_passthrough():
    X = torch.arange(12, dtype=torch.float64).reshape(4, 1, 3)
    v = torch.tensor([0.1, 0.9, 0.3, 0.2], dtype=torch.float64)
    assert torch.equal(shard_restarts(X), X)
    bv, bx, br = gather_best(v, X)
    assert float(bv) == 0.9 and torch.equal(bx, X[1]) and br == 0
