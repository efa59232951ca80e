"""DDPG learner restatement (reference: surreal/learner/ddpg.py:186-428; low-dim, single critic)."""
import torch
import torch.nn as nn

from . import nets


class OracleDDPGLearner:
    def __init__(self, actor, critic, actor_t, critic_t, gamma=0.99, n_step=3, lr_actor=1e-4, lr_critic=1e-3,
                 clip_actor=True, actor_clip=1.0, clip_critic=False, critic_clip=5.0,
                 target_type='hard', target_interval=500, tau=1e-3):
        rg = lambda ls: [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in ls]  # noqa: E731
        ng = lambda ls: [(w.clone(), b.clone()) for w, b in ls]  # noqa: E731
        self.actor, self.critic = rg(actor), rg(critic)
        self.actor_t, self.critic_t = ng(actor_t), ng(critic_t)
        self.gamma, self.n_step = gamma, n_step
        self.clip_actor, self.actor_clip = clip_actor, actor_clip
        self.clip_critic, self.critic_clip = clip_critic, critic_clip
        self.target_type, self.target_interval, self.tau = target_type, target_interval, tau
        self.target_counter = 0
        self.actor_params = [t for wb in self.actor for t in wb]
        self.critic_params = [t for wb in self.critic for t in wb]
        self.critic_optim = torch.optim.Adam(self.critic_params, lr=lr_critic)      # ddpg.py:145-156
        self.actor_optim = torch.optim.Adam(self.actor_params, lr=lr_actor)

    def optimize(self, obs, actions, rewards, obs_next, dones):
        """ddpg.py:244-352.  Inputs are numpy as the aggregator emits them (rewards/dones float64 [B,1])."""
        obs = torch.tensor(obs, dtype=torch.float32)                   # ddpg.py:203-222 (preprocess)
        obs_next = torch.tensor(obs_next, dtype=torch.float32)
        actions = torch.tensor(actions, dtype=torch.float32)
        rewards = torch.tensor(rewards, dtype=torch.float32)
        dones = torch.tensor(dones, dtype=torch.float32)
        assert actions.max().item() <= 1.0 and actions.min().item() >= -1.0
        with torch.no_grad():
            pol_t = nets.ddpg_actor(obs_next, self.actor_t)           # ddpg.py:266
            q_t = nets.ddpg_critic(obs_next, pol_t, self.critic_t)
            y = rewards + pow(self.gamma, self.n_step) * q_t * (1.0 - dones)   # ddpg.py:279
        y_policy = nets.ddpg_critic(obs, actions, self.critic)
        for p in self.critic_params:
            p.grad = None
        critic_loss = nn.MSELoss()(y_policy, y)
        critic_loss.backward()
        if self.clip_critic:
            nn.utils.clip_grad_value_(self.critic_params, self.critic_clip)
        self.critic_optim.step()
        for p in self.actor_params:
            p.grad = None
        actor_loss = -nets.ddpg_critic(obs, nets.ddpg_actor(obs, self.actor), self.critic).mean()
        actor_loss.backward()
        if self.clip_actor:
            nn.utils.clip_grad_value_(self.actor_params, self.actor_clip)
        self.actor_optim.step()
        stats = {'actor_loss': actor_loss.item(), 'critic_loss': critic_loss.item(),
                 'action_norm': actions.norm(2, 1).mean().item(), 'rewards': rewards.mean().item(),
                 'Q_target': y.mean().item(), 'Q_policy': y_policy.mean().item()}
        self._target_update()
        return stats

    def _target_update(self):                                          # ddpg.py:403-428
        pairs = list(zip(self.actor_t, self.actor)) + list(zip(self.critic_t, self.critic))
        if self.target_type == 'soft':
            for (wt, bt), (w, b) in pairs:
                wt.mul_(1.0 - self.tau).add_(w.detach(), alpha=self.tau)
                bt.mul_(1.0 - self.tau).add_(b.detach(), alpha=self.tau)
        else:
            self.target_counter += 1
            if self.target_counter % self.target_interval == 0:
                for (wt, bt), (w, b) in pairs:
                    wt.copy_(w.detach())
                    bt.copy_(b.detach())
