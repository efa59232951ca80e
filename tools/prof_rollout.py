#!/usr/bin/env python
"""Drive the persistent rollout kernel stand-alone (for ncu captures and quick timing).

    python tools/prof_rollout.py [T] [reps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch  # noqa: E402

from helpers import ppo_configs  # noqa: E402


def main():
    from surreal_b200.agent import PPOAgent
    from surreal_b200.replay import FIFOReplay
    from surreal_b200.env import SyntheticEnv
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    N, D, A = 1024, 64, 8
    lc, ec, sc = ppo_configs(D=D, A=A, actor_h=(256, 256), critic_h=(256, 256), n_step=128, stride=128, B=N,
                             memory_size=4 * N)
    ec.num_envs = N
    R = FIFOReplay(lc, ec, sc)
    ag = PPOAgent(lc, ec, sc, 0, 'training')
    env = SyntheticEnv(N, D, A, limit_episode_length=200, seed=0)
    ag.env = w = ag.prepare_env_agent(env)
    w.reset()
    assert ag.rollout_chunk_supported()
    for _ in range(3):
        ag.rollout_chunk(T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ag.rollout_chunk(T)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('persistent rollout: T=%d  %.3f ms per chunk  %.2f us per env step  (%d windows queued)' %
          (T, ms, ms * 1e3 / T, len(R)))


if __name__ == '__main__':
    main()
