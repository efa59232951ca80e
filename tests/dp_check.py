"""2-rank data-parallel learner check (run under torchrun on 2 GPUs; see tests/test_dp_gpu.py).
Each rank feeds HALF of a global batch; the result must match the single-process CPU oracle on the FULL batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def run_check(mode='clip', big_lr=False):
    """One data-parallel learn() (every rank feeds 1/world of a global batch) against the single-process oracle on
    the FULL batch, plus replica drift.  Needs an initialised NCCL process group; returns (ok, [messages])."""
    from helpers import ppo_configs
    from oracle import nets as onets
    from oracle.filters import ZFilter as OZ
    from oracle.ppo import OraclePPOLearner
    from surreal_b200.learner import PPOLearner
    rank, world = dist.get_rank(), dist.get_world_size()
    B, n, D, A = 256, 16, 24, 4
    gen = torch.Generator().manual_seed(3)

    def layers(dims):
        out = []
        for i in range(len(dims) - 1):
            b = 1.0 / np.sqrt(dims[i])
            out.append(((torch.rand(dims[i + 1], dims[i], generator=gen) * 2 - 1) * b, (torch.rand(dims[i + 1], generator=gen) * 2 - 1) * b))
        return out
    al, cl = layers([D, 64, 48, A]), layers([D, 64, 48, 1])
    log_var = torch.zeros(1, A) - 1.0
    zf = OZ(D)
    zf.update(torch.randn(300, D, generator=gen) * 1.3 + 0.2)
    rng = np.random.default_rng(7)
    obs = (rng.standard_normal((B, n, D)) * 1.2).astype(np.float32)
    obs_next = (rng.standard_normal((B, 1, D)) * 1.2).astype(np.float32)
    with torch.no_grad():
        pd0 = onets.ppo_actor(zf.forward(torch.tensor(obs[:, 0])), al, log_var).numpy()
    pd = np.tile(pd0[:, None, :], (1, n, 1)).astype(np.float32)
    pd[:, :, A:] *= np.exp(rng.uniform(-0.25, 0.25, (B, 1, 1))).astype(np.float32)
    actions = np.clip(rng.standard_normal((B, n, A)) * pd[:, :, A:] + pd[:, :, :A], -1, 1)
    rewards = rng.standard_normal((B, n)) * 0.3
    dones = np.zeros((B, n), dtype=np.float32)
    dones[rng.random(B) < 0.3, n - 1] = 1
    lr = 3e-3 if big_lr else 1e-4
    O = OraclePPOLearner(al, log_var, cl, zf, A, n, B, ppo_mode=mode, lr_actor=lr, lr_critic=lr)
    st_o = O.learn(dict(obs=obs, obs_next=obs_next, actions=actions, rewards=rewards, dones=dones, pd=pd))
    Bl = B // world
    lc, ec, sc = ppo_configs(D=D, A=A, actor_h=(64, 48), critic_h=(64, 48), n_step=n, stride=n, B=Bl, mode=mode, lr=lr)
    L = PPOLearner(lc, ec, sc)
    if rank == 0:                                   # only rank 0 holds the reference weights: the broadcast must spread them
        L.model.actor.load_layers(al, extra=log_var)
        L.model.critic.load_layers(cl)
        L.model.z_stats.copy_(torch.cat([zf.running_sum, zf.running_sumsq, zf.count]).cuda())
        L.ref_target_model.update_target_params(L.model)
    L.enable_data_parallel(dist.group.WORLD)
    sl = slice(rank * Bl, (rank + 1) * Bl)
    st = L.learn({'obs': obs[sl], 'obs_next': obs_next[sl], 'actions': actions[sl], 'rewards': rewards[sl], 'dones': dones[sl],
                  'persistent_infos': [pd[sl]], 'onetime_infos': None})
    torch.cuda.synchronize()
    ok = True
    msgs = []

    def chk(name, a, b, tol):
        nonlocal ok
        if abs(a - b) > tol:
            ok = False
            msgs.append('%s: %g vs %g' % (name, a, b))
    chk('policy epochs', float(L.last_n_policy_epochs), float(O.n_policy_epochs[-1]), 0.0)
    adv = L._adv.cpu().view(-1)
    chk('adv', float((adv - O.last_adv.view(-1)[sl]).abs().max()), 0.0, 1e-5)
    for k in ['_surr_loss', '_clip_surr_loss', '_kl_loss_adapt', '_pol_kl', '_val_loss', '_entropy', '_avg_return_targ']:
        if k in st_o:
            chk(k, st[k], st_o[k], 1e-5 * max(1.0, abs(st_o[k])))
    for l in range(3):
        for got, exp in ((L.model.actor.get_layer(l), O.actor[l]), (L.model.critic.get_layer(l), O.critic[l])):
            chk('W%d' % l, float((got[0].cpu() - exp[0].detach()).abs().max()), 0.0, max(2e-6, 0.02 * lr))
    zs = L.model.z_stats.cpu()
    chk('zcount', float(zs[-1]), float(O.zf.count), 1e-3)
    chk('zsum', float((zs[:D] - O.zf.running_sum).abs().max()), 0.0, 1e-3)
    # every rank must end with identical parameters
    p = L.model.actor.params.clone()
    dist.broadcast(p, 0)
    chk('replica drift', float((p - L.model.actor.params).abs().max()), 0.0, 0.0)
    return ok, msgs


def main():
    rank = int(os.environ['RANK'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    os.environ.setdefault('NCCL_MAX_NCHANNELS', '4')
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    ok, msgs = run_check(os.environ.get('DP_MODE', 'clip'), bool(os.environ.get('DP_BIGLR')))
    print('rank %d %s %s' % (rank, 'DP_OK' if ok else 'DP_FAIL', '; '.join(msgs)), flush=True)
    dist.barrier()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)        # NCCL-in-graph teardown hangs: skip the destructors
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
