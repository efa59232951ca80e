"""DDPG learner restatement (reference: surreal/learner/ddpg.py:186-428; low-dim).  ``critic2`` / ``critic2_t`` switch on
the TD3 double critic (ddpg.py:279-283,298-321), ``policy_noise`` feeds the action regularisation (ddpg.py:267-278)."""
import numpy as np
import torch
import torch.nn as nn

from . import nets


class OracleDDPGLearner:
    def __init__(self, actor, critic, actor_t, critic_t, gamma=0.99, n_step=3, lr_actor=1e-4, lr_critic=1e-3,
                 clip_actor=True, actor_clip=1.0, clip_critic=False, critic_clip=5.0,
                 target_type='hard', target_interval=500, tau=1e-3, critic2=None, critic2_t=None):
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
        self.critic2 = self.critic2_t = None
        if critic2 is not None:                                            # model2 / model_target2 (ddpg.py:118-143)
            self.critic2, self.critic2_t = rg(critic2), ng(critic2_t)
            self.critic2_params = [t for wb in self.critic2 for t in wb]
            self.critic_optim2 = torch.optim.Adam(self.critic2_params, lr=lr_critic)

    def optimize(self, obs, actions, rewards, obs_next, dones, policy_noise=None):
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
            if policy_noise is not None:                               # ddpg.py:267-278: AFTER Q'_1 was computed
                noise = np.clip(np.asarray(policy_noise), -0.5, 0.5)
                pol_t = (pol_t + torch.tensor(noise, dtype=torch.float32)).clamp(-1, 1)
            y = rewards + pow(self.gamma, self.n_step) * q_t * (1.0 - dones)   # ddpg.py:279
            if self.critic2 is not None:                               # ddpg.py:280-283
                q_t2 = nets.ddpg_critic(obs_next, pol_t, self.critic2_t)
                y2 = rewards + pow(self.gamma, self.n_step) * q_t2 * (1.0 - dones)
                y = torch.min(y, y2)
        y_policy = nets.ddpg_critic(obs, actions, self.critic)
        for p in self.critic_params:
            p.grad = None
        critic_loss = nn.MSELoss()(y_policy, y)
        critic_loss.backward()
        if self.clip_critic:
            nn.utils.clip_grad_value_(self.critic_params, self.critic_clip)
        self.critic_optim.step()
        y_policy2 = None
        if self.critic2 is not None:                                   # ddpg.py:311-321 (critic_loss is overwritten)
            y_policy2 = nets.ddpg_critic(obs, actions, self.critic2)
            for p in self.critic2_params:
                p.grad = None
            critic_loss = nn.MSELoss()(y_policy2, y)
            critic_loss.backward()
            if self.clip_critic:
                nn.utils.clip_grad_value_(self.critic2_params, self.critic_clip)
            self.critic_optim2.step()
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
        if y_policy2 is not None:
            stats['Q_policy2'] = y_policy2.mean().item()
        self._target_update()
        return stats

    def _target_update(self):                                          # ddpg.py:403-428
        pairs = list(zip(self.actor_t, self.actor)) + list(zip(self.critic_t, self.critic))
        if self.critic2 is not None:
            pairs += list(zip(self.critic2_t, self.critic2))
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
