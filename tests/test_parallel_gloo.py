"""world_size-2 gloo test of the data-parallel host logic (CPU): the collective helpers and the algebra the
data-parallel learner relies on -- averaged per-rank mean-gradients == global-batch gradient, combined moments ==
global unbiased std, averaged KL == global KL -- checked against the single-process CPU oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from surreal_b200.parallel import LearnerDP, combine_moments
    from oracle import pd as PD
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    dp = LearnerDP()
    torch.manual_seed(0)
    B, D, A = 64, 6, 3
    W = torch.randn(A, D, dtype=torch.float64) * 0.3
    x = torch.randn(B, D, dtype=torch.float64)
    adv = torch.randn(B, dtype=torch.float64)
    lo, hi = rank * B // world, (rank + 1) * B // world
    # (1) parameters start from rank 0
    p = torch.full((5,), float(rank))
    dp.broadcast_(p)
    assert p.tolist() == [0.0] * 5
    # (2) advantage normalisation over the global batch from per-rank moments
    a = adv[lo:hi]
    mom = torch.tensor([a.sum(), (a * a).sum(), float(a.numel())], dtype=torch.float64)
    dp.sum_(mom)
    mean, std = combine_moments(mom)
    assert abs(mean - adv.mean().item()) < 1e-12 and abs(std - adv.std().item()) < 1e-12
    # (3) averaged per-rank mean-loss gradients == global-batch gradient
    Wl = W.clone().requires_grad_(True)
    loss = (torch.tanh(x[lo:hi] @ Wl.t()).pow(2).sum(1) * adv[lo:hi]).mean()
    loss.backward()
    g = Wl.grad.clone()
    dp.sum_(g)
    g /= world
    Wg = W.clone().requires_grad_(True)
    (torch.tanh(x @ Wg.t()).pow(2).sum(1) * adv).mean().backward()
    assert torch.allclose(g, Wg.grad, atol=1e-12)
    # (4) the KL scalar every rank branches on is the global mean
    p0 = torch.cat([torch.tanh(x @ W.t()), torch.full((B, A), 0.4, dtype=torch.float64)], 1)
    p1 = torch.cat([torch.tanh(x @ (W * 1.1).t()), torch.full((B, A), 0.5, dtype=torch.float64)], 1)
    kl = PD.kl(p0[lo:hi], p1[lo:hi], A).mean().reshape(1)
    dp.mean_(kl)
    assert abs(kl.item() - PD.kl(p0, p1, A).mean().item()) < 1e-12
    out_q.put((rank, float(kl.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_algebra_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get() for _ in range(2))
    assert res[0] == res[1]                      # identical branch decision on every rank
