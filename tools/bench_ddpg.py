#!/usr/bin/env python
"""DDPG half of the hot path at BASELINE configs[2] size (not a bench.py line: a measurement for DESIGN.md).

1 048 576-slot UniformReplay in HBM, batch 4096, nets 300-200 / 400-300, n_step 3: times `replay.sample(B)` (CPython-exact
index stream on the host + device gather) and `learner.learn(batch)` (one CUDA graph), and the sample -> learn loop.

    python tools/bench_ddpg.py [iters]
"""
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch  # noqa: E402

from helpers import ddpg_configs  # noqa: E402


def main():
    from surreal_b200.replay import UniformReplay
    from surreal_b200.learner import DDPGLearner
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    dev = 'cuda:0'
    D, A, B, CAP = 64, 8, 4096, 1 << 20
    lc, ec, sc = ddpg_configs(D=D, A=A, actor_h=(300, 200), critic_h=(400, 300), B=B, n_step=3, memory_size=CAP, start=3000)
    R = UniformReplay(lc, ec, sc)
    g = torch.Generator(device=dev).manual_seed(4)
    R.r_obs.copy_(torch.randn(CAP, D, device=dev, generator=g))
    R.r_obs_next.copy_(torch.randn(CAP, D, device=dev, generator=g))
    R.r_act.copy_(torch.rand(CAP, A, device=dev, generator=g) * 2 - 1)
    R.r_rew.copy_(torch.randn(CAP, device=dev, generator=g))
    R.r_done.copy_((torch.rand(CAP, device=dev, generator=g) < 0.005).float())
    R.state[0], R.state[1] = 0, CAP
    R.mark_device_inserts()
    L = DDPGLearner(lc, ec, sc)
    random.seed(5)
    for _ in range(5):                                            # warm-up: graph capture, allocator
        L.learn(R.sample(B))
    torch.cuda.synchronize()

    def timed(fn, n):
        torch.cuda.synchronize()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (time.time() - t0) * 1e3 / n

    batch = R.sample(B)
    learn_dev, learn_wall = timed(lambda: L.learn(batch), iters)
    samp_dev, samp_wall = timed(lambda: R.sample(B), iters)
    loop_dev, loop_wall = timed(lambda: L.learn(R.sample(B)), iters)
    rec = (2 * D + A + 2) * 4
    print(json.dumps({
        'workload': 'DDPG configs[2]: 1M-slot uniform ring, batch 4096, nets 300-200 / 400-300, n_step 3',
        'learn_ms_device': learn_dev, 'learn_ms_wall': learn_wall,
        'sample_ms_device': samp_dev, 'sample_ms_wall': samp_wall,
        'loop_ms_wall': loop_wall, 'learner_updates_per_s': 1e3 / loop_wall,
        'gather_bytes_per_sample_call': B * (2 * rec + 8),
        'gather_gbs_device': B * (2 * rec + 8) / (samp_dev / 1e3) / 1e9 if samp_dev > 0 else None}))


if __name__ == '__main__':
    main()
