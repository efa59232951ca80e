"""The persistent rollout kernel (the kernel that carries the headline number) against the ORACLE at the headline
shape: BASELINE configs[1] -- 1024 actors x n_step 128, 64-dim obs, A = 8, 64-256-256-8 policy with a LIVE head,
z-filter on, per-actor exploration constants.  Every record the kernel staged into the HBM FIFO is checked:

  (a) pd rows   == oracle.nets.ppo_actor(OZFilter.forward(obs)) with stds scaled by exp(noise_i)   (ppo_agent.py:133-139,149)
  (b) actions   == clip(mean + std * eps, -1, 1) for the known Philox draws                         (ppo_agent.py:140-145)
  (c) env       == the synthetic env's dynamics on (obs, action) with the known Philox draws (SURVEY §8d cfg 2)
  (d) windows   == oracle.windowing.multistep_windows per actor (clear-on-done, tails discarded), queued in
                   (completion step, actor) arrival order through oracle.replay.FIFO               (exp_sender_wrapper.py:204-228,
                                                                                                    fifo_replay.py:27-39)
Tolerance: 1e-5 * max(1, rms) on pd / obs / rewards, 1e-5 absolute on actions (|a| <= 1)."""
import numpy as np
import pytest
import torch

from helpers import ppo_configs
from oracle import nets as onets
from oracle import philox
from oracle.filters import ZFilter as OZ
from oracle.replay import FIFO as OFIFO
from oracle.windowing import multistep_windows

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _close(got, exp, what, tol=1e-5):
    got, exp = np.asarray(got, dtype=np.float64), np.asarray(exp, dtype=np.float64)
    bar = tol * max(1.0, float(np.sqrt(np.mean(exp ** 2))))
    err = float(np.abs(got - exp).max())
    assert err <= bar, '%s: max |diff| %.3e > %.3e' % (what, err, bar)


@pytest.mark.parametrize('ep_len,chunks', [(200, 3), (128, 2)])
def test_persistent_rollout_matches_oracle_at_cfg2_shape(ep_len, chunks):
    from surreal_b200.agent import PPOAgent
    from surreal_b200.replay import FIFOReplay
    from surreal_b200.env import SyntheticEnv
    N, D, A, n, H = 1024, 64, 8, 128, (256, 256)
    lc, ec, sc = ppo_configs(D=D, A=A, actor_h=H, critic_h=H, n_step=n, stride=n, B=N, memory_size=4 * N)
    ec.num_envs = N
    R = FIFOReplay(lc, ec, sc)
    ag = PPOAgent(lc, ec, sc, 3, 'training')
    g = torch.Generator().manual_seed(11)
    dims = [D, H[0], H[1], A]
    layers = [((torch.rand(dims[i + 1], dims[i], generator=g) * 2 - 1) / np.sqrt(dims[i]),
               (torch.rand(dims[i + 1], generator=g) * 2 - 1) * 0.1) for i in range(3)]
    log_var = torch.linspace(-1.2, -0.6, A).view(1, A)
    ag.model.actor.load_layers(layers, extra=log_var)
    zf = OZ(D)
    zf.update(torch.randn(700, D, generator=g) * 0.8 + 0.1)
    ag.model.z_stats.copy_(torch.cat([zf.running_sum, zf.running_sumsq, zf.count]).to(DEV))
    rng = np.random.default_rng(5)
    noise = rng.uniform(-0.25, 0.25, N)
    ag.set_noise(noise)
    env = SyntheticEnv(N, D, A, limit_episode_length=ep_len, seed=3)
    ag.env = w = ag.prepare_env_agent(env)
    w.reset()
    s0 = env.state.cpu().numpy().copy()
    assert ag.rollout_chunk_supported()
    T = chunks * n
    for _ in range(chunks):
        assert ag.rollout_chunk(n)
    torch.cuda.synchronize()
    assert int(env.step_counter.item()) == T

    # (d) the oracle's window stream per actor, queued in (completion step, actor) order
    lens, left = [], T
    while left > 0:
        lens.append(min(ep_len, left))
        left -= lens[-1]
    wins = multistep_windows(lens, n, n)              # lockstep actors: the same for everyone
    arrivals = sorted((ids[-1], i, k) for k, (ids, _, _) in enumerate(wins) for i in range(N))
    q = OFIFO(4 * N, N)
    for a in arrivals:
        q.insert(a)
    st = R._read_state()
    assert st['count'] == len(q) == len(wins) * N and st['dropped'] == 0 and st['total_in'] == len(wins) * N
    Ws, Wa = env.Ws.cpu().numpy(), env.Wa.cpu().numpy()
    scale = np.exp(noise).astype(np.float32)
    seen_first = False
    while len(q) >= N:
        expect = q.sample(N)
        b = R.sample(N)
        actors = np.array([e[1] for e in expect])
        k = expect[0][2]
        assert all(e[2] == k for e in expect) and list(actors) == list(range(N))
        ids, nxt_id, dn = wins[k]
        obs_full = b['obs_full'].cpu().numpy()                   # [N, n+1, D]
        act, pd = b['actions'].cpu().numpy(), b['persistent_infos'][0].cpu().numpy()
        rew, done = b['rewards'].cpu().numpy(), b['dones'].cpu().numpy()
        # (d) done flags of the window, as the reference's wrapper would have sent them
        np.testing.assert_array_equal(done, np.tile(np.array(dn, dtype=np.float32), (N, 1)))
        # first observation of the window: the env's reset state
        if ids[0] == 0:
            np.testing.assert_array_equal(obs_full[:, 0], s0)
            seen_first = True
        else:
            _, _, rz = philox.env_noise(env.seed + 7, ids[0] - 1, actors, D)
            _close(obs_full[:, 0], rz, 'reset state of window %d' % k)
        # (a) behaviour policy rows
        with torch.no_grad():
            pdo = onets.ppo_actor(zf.forward(torch.tensor(obs_full[:, :n].reshape(-1, D))), layers, log_var).numpy()
        pdo = pdo.reshape(N, n, 2 * A)
        pdo[:, :, A:] *= scale[:, None, None]
        _close(pd, pdo, 'pd rows of window %d' % k)
        assert float(np.abs(pdo[:, :, :A]).max()) > 0.3             # a live, non-trivial policy head
        for j, t in enumerate(ids):
            # (b) sampled, clipped actions for the known draws (oracle.agent.ppo_act's arithmetic, batched)
            eps = philox.agent_eps(ag.seed, t, actors, A).astype(np.float64)
            a_or = np.clip(eps * pdo[:, j, A:] + pdo[:, j, :A], -1, 1)
            err = float(np.abs(act[:, j] - a_or).max())
            assert err <= 1e-5, 'actions at step %d: %.3e' % (t, err)
            # (c) env dynamics on what the kernel staged
            nxt, r_or = philox.synth_env_step(obs_full[:, j], act[:, j], Ws, Wa, env.seed + 7, t, actors)
            _close(obs_full[:, j + 1], nxt, 'successor obs at step %d' % t)
            _close(rew[:, j], r_or, 'reward at step %d' % t)
        assert float(np.mean(np.abs(act) >= 1.0)) > 0.0005              # the clip is exercised
    assert seen_first
