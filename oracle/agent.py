"""Actor-side policy evaluation (reference: surreal/agent/ppo_agent.py:106-154, ddpg_agent.py:155-184)."""
import numpy as np
import torch

from . import nets
from . import pd as PD


def ppo_act(obs_row, actor_layers, log_var, zfilter, log_noise, eps=None, deterministic=False):
    """One env step of one PPO actor.  ``eps`` are the N(0,1) draws np.random.randn would return
    (float64).  Returns (action float64 [A], pd float32 [2A]); pd stds are scaled by exp(noise)
    BEFORE sampling and before being recorded (ppo_agent.py:139,149)."""
    A = log_var.numel()
    with torch.no_grad():
        x = torch.tensor(obs_row, dtype=torch.float32).unsqueeze(0)
        if zfilter is not None:
            x = zfilter.forward(x)
        pdv = nets.ppo_actor(x, actor_layers, log_var).numpy()
    pdv[:, A:] *= np.exp(log_noise)
    if deterministic:
        act = PD.maxprob(pdv, A).copy()
    else:
        act = PD.sample(pdv, A, np.asarray(eps).reshape(1, A))
    act = np.clip(act, -1, 1)
    return act.reshape(-1), pdv.reshape(-1)


def ppo_act_rnn(obs_row, actor_layers, log_var, zfilter, lstm, cells, log_noise, eps=None, deterministic=False):
    """One env step of one PPO actor in RNN mode (ppo_agent.py:133-149, ppo_net.py:317-351): ``lstm`` is a
    torch.nn.LSTM(batch_first=True), ``cells`` = (h, c) of shape [layers, 1, hidden] carried between steps.
    Returns (action, pd, (h_before, c_before) as shipped in onetime_infos, new cells)."""
    A = log_var.numel()
    onetime = (cells[0].squeeze(1).numpy().copy(), cells[1].squeeze(1).numpy().copy())
    with torch.no_grad():
        x = torch.tensor(obs_row, dtype=torch.float32).unsqueeze(0).unsqueeze(0)     # [1, 1, D]
        if zfilter is not None:
            x = zfilter.forward(x)
        feat, new_cells = lstm(x, cells)
        feat = feat.contiguous()
        pdv = nets.ppo_actor(feat.view(-1, feat.shape[2]), actor_layers, log_var).numpy()
    pdv[:, A:] *= np.exp(log_noise)
    if deterministic:
        act = PD.maxprob(pdv, A).copy()
    else:
        act = PD.sample(pdv, A, np.asarray(eps).reshape(1, A))
    act = np.clip(act, -1, 1)
    return act.reshape(-1), pdv.reshape(-1), onetime, (new_cells[0].detach(), new_cells[1].detach())


def ddpg_act(obs_row, actor_layers, sigma, unit_noise=None, deterministic=False):
    """ddpg_agent.py:155-184 with NormalActionNoise(0, sigma) (action_noise.py:9-16):
    clip -> + noise -> clip.  ``unit_noise`` ~ N(0,1) so that noise = sigma * unit_noise."""
    with torch.no_grad():
        x = torch.tensor(obs_row, dtype=torch.float32).unsqueeze(0)
        a = nets.ddpg_actor(x, actor_layers).numpy()[0]
    a = a.clip(-1, 1)
    if not deterministic:
        a += sigma * np.asarray(unit_noise)        # in-place on the float32 array (ddpg_agent.py:181)
    return a.clip(-1, 1)


def ddpg_act_ou(obs_row, actor_layers, x_prev, sigma, theta, dt, unit_noise=None, deterministic=False):
    """ddpg_agent.py:155-184 with OrnsteinUhlenbeckActionNoise (action_noise.py:22-39): float64 state
    x <- x + theta*(mu - x)*dt + sigma*sqrt(dt)*N(0,1), mu = 0; ``unit_noise`` are the N(0,1) draws.
    Returns (action float32 [A], new state float64 [A])."""
    with torch.no_grad():
        x = torch.tensor(obs_row, dtype=torch.float32).unsqueeze(0)
        a = nets.ddpg_actor(x, actor_layers).numpy()[0]
    a = a.clip(-1, 1)
    x_new = np.asarray(x_prev, dtype=np.float64)
    if not deterministic:
        mu = np.zeros_like(x_new)
        x_new = x_new + theta * (mu - x_new) * dt + sigma * np.sqrt(dt) * np.asarray(unit_noise, dtype=np.float64)
        a += x_new                                  # in-place on the float32 array (ddpg_agent.py:181)
    return a.clip(-1, 1), x_new
