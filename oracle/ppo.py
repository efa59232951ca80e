"""PPO learner restatement (reference: surreal/learner/ppo.py:57-666; MLP / non-RNN branch).

State is explicit (lists of torch tensors); the arithmetic is the same stock-torch call sequence
the reference issues, so results match the reference bit-for-bit on CPU.
"""
import copy

import numpy as np
import torch
import torch.nn as nn

from . import pd as PD
from . import nets
from .filters import ZFilter, RewardFilter
from .gae import gae_from_values

DEFAULTS = dict(  # surreal/main/ppo_configs.py:15-94
    gamma=0.995, lam=0.97, norm_adv=True, reward_scale=1.0, use_z_filter=True, use_r_filter=False,
    ppo_mode='adapt', lr_actor=1e-4, lr_critic=1e-4, clip_actor_gradient=True, actor_gradient_norm_clip=5.0,
    clip_critic_gradient=True, critic_gradient_norm_clip=5.0, epoch_policy=10, epoch_baseline=10,
    kl_target=0.015, adjust_threshold=(0.5, 2.0), kl_cutoff_coeff=250, beta_init=1.0, beta_range=(1 / 35.0, 35.0),
    adapt_scale_constant=1.5, clip_epsilon_init=0.2, clip_range=(0.05, 0.3), clip_scale_constant=1.2,
    exp_interval=4096, init_log_sig=-1.0)


class OraclePPOLearner:
    def __init__(self, actor_layers, log_var, critic_layers, zfilter, action_dim, n_step, batch_size, **cfg):
        c = dict(DEFAULTS)
        c.update(cfg)
        self.c = c
        self.A = action_dim
        self.n_step = n_step
        self.batch_size = batch_size
        self.actor = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in actor_layers]
        self.log_var = log_var.clone().requires_grad_(True)
        self.critic = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in critic_layers]
        self.zf = zfilter.clone() if zfilter is not None else None
        self._sync_ref()                                             # ppo.py:151
        # parameter order == nn.Module.parameters(): own params (log_var) first, then the Sequential
        self.actor_params = [self.log_var] + [t for wb in self.actor for t in wb]
        self.critic_params = [t for wb in self.critic for t in wb]
        self.critic_optim = torch.optim.Adam(self.critic_params, lr=c['lr_critic'])   # ppo.py:159-168
        self.actor_optim = torch.optim.Adam(self.actor_params, lr=c['lr_actor'])
        if c['ppo_mode'] == 'adapt':
            self.beta = c['beta_init']
        else:
            self.clip_epsilon = c['clip_epsilon_init']
        self.exp_counter = 0
        self.kl_record = []
        self.current_iteration = 0
        self.rfilter = RewardFilter() if c['use_r_filter'] else None
        self.n_policy_epochs = []

    # -- model -------------------------------------------------------------------------------
    def _sync_ref(self):
        """PPOModel.update_target_params (ppo_net.py:226-242): actor, critic AND z-filter."""
        self.ref_actor = [(w.detach().clone(), b.detach().clone()) for w, b in self.actor]
        self.ref_log_var = self.log_var.detach().clone()
        self.ref_zf = self.zf.clone() if self.zf is not None else None

    def forward_actor(self, obs, ref=False):
        zf = self.ref_zf if ref else self.zf
        x = zf.forward(obs) if zf is not None else obs                # ppo_net.py:262-266
        if ref:
            return nets.ppo_actor(x, self.ref_actor, self.ref_log_var)
        return nets.ppo_actor(x, self.actor, self.log_var)

    def forward_critic(self, obs):
        x = self.zf.forward(obs) if self.zf is not None else obs
        return nets.ppo_critic(x, self.critic)

    # -- losses ------------------------------------------------------------------------------
    def _clip_loss(self, obs, actions, adv, behave_pol):              # ppo.py:194-225
        learn_pol = self.forward_actor(obs)
        learn_prob = PD.likelihood(actions, learn_pol, self.A)
        behave_prob = PD.likelihood(actions, behave_pol, self.A)
        ratio = learn_prob / behave_prob
        clipped = torch.clamp(ratio, 1 - self.clip_epsilon, 1 + self.clip_epsilon)
        surr = -ratio * adv.view(-1, 1)
        csurr = -clipped * adv.view(-1, 1)
        loss = torch.cat([surr, csurr], 1).max(1)[0].mean()
        stats = {'_surr_loss': surr.mean().item(), '_clip_surr_loss': loss.item(),
                 '_entropy': PD.entropy(learn_pol, self.A).mean().item(), '_clip_epsilon': self.clip_epsilon}
        return loss, stats

    def _adapt_loss(self, obs, actions, adv, behave_pol, ref_pol):    # ppo.py:250-285
        learn_pol = self.forward_actor(obs)
        prob_behave = PD.likelihood(actions, behave_pol, self.A)
        prob_learn = PD.likelihood(actions, learn_pol, self.A)
        kl = PD.kl(ref_pol, learn_pol, self.A).mean()
        surr = -(adv.view(-1, 1) * (prob_learn / torch.clamp(prob_behave, min=1e-2))).mean()
        loss = surr + self.beta * kl
        entropy = PD.entropy(learn_pol, self.A).mean()
        if kl.item() - 2.0 * self.c['kl_target'] > 0:
            loss = loss + self.c['kl_cutoff_coeff'] * (kl - 2.0 * self.c['kl_target']).pow(2)
        stats = {'_kl_loss_adapt': loss.item(), '_surr_loss': surr.item(), '_pol_kl': kl.item(),
                 '_entropy': entropy.item(), '_beta': self.beta}
        return loss, stats

    def _policy_update(self, obs, actions, adv, behave_pol, ref_pol):  # ppo.py:227-248 / 287-309
        if self.c['ppo_mode'] == 'clip':
            loss, stats = self._clip_loss(obs, actions, adv, behave_pol)
        else:
            loss, stats = self._adapt_loss(obs, actions, adv, behave_pol, ref_pol)
        for p in self.actor_params:
            p.grad = None
        loss.backward()
        if self.c['clip_actor_gradient']:
            stats['grad_norm_actor'] = float(nn.utils.clip_grad_norm_(self.actor_params,
                                                                      self.c['actor_gradient_norm_clip']))
        self.actor_optim.step()
        return stats

    def _value_update(self, obs, returns):                            # ppo.py:311-353
        values = self.forward_critic(obs)
        explained_var = 1 - torch.var(returns - values) / torch.var(returns)
        loss = (values - returns).pow(2).mean()
        stats = {'_val_loss': loss.item(), '_val_explained_var': explained_var.item()}
        for p in self.critic_params:
            p.grad = None
        loss.backward()
        if self.c['clip_critic_gradient']:
            stats['grad_norm_critic'] = float(nn.utils.clip_grad_norm_(self.critic_params,
                                                                       self.c['critic_gradient_norm_clip']))
        self.critic_optim.step()
        return stats

    # -- GAE ---------------------------------------------------------------------------------
    def gae_and_return(self, obs, obs_next, rewards, dones):          # ppo.py:355-418 (MLP branch)
        cat = torch.cat([obs, obs_next], dim=1)                       # [B, n+1, D]
        flat = cat.view(-1, cat.shape[-1])
        values = self.forward_critic(flat).view(self.batch_size, self.n_step + 1)
        self.last_values_raw = values.detach().clone()
        return gae_from_values(rewards, values, dones, self.c['gamma'], self.c['lam'],
                               norm_adv=self.c['norm_adv'])

    # -- learn -------------------------------------------------------------------------------
    def preprocess(self, batch):                                      # ppo.py:420-484
        obs = torch.tensor(batch['obs'], dtype=torch.float32)
        obs_next = torch.tensor(batch['obs_next'], dtype=torch.float32)
        actions = torch.tensor(batch['actions'], dtype=torch.float32)
        rewards = torch.tensor(batch['rewards'], dtype=torch.float32) * self.c['reward_scale']
        if self.rfilter is not None:
            normed = self.rfilter.forward(rewards)
            self.rfilter.update(rewards)
            rewards = normed
        dones = torch.tensor(batch['dones'], dtype=torch.float32)
        pds = torch.tensor(batch['pd'], dtype=torch.float32)
        return obs, obs_next, actions, rewards, dones, pds

    def learn(self, batch):                                           # ppo.py:588-613 + 487-586
        self.current_iteration += 1
        obs, obs_next, actions, rewards, dones, pds = self.preprocess(batch)
        adv, ret = self.gae_and_return(obs, obs_next, rewards, dones)
        adv, ret = adv.detach(), ret.detach()
        self.last_adv, self.last_ret = adv, ret
        behave_pol = pds[:, 0, :].contiguous()
        actions0 = actions[:, 0, :].contiguous()
        obs0 = obs[:, 0, :].contiguous()
        ref_pol = self.forward_actor(obs0, ref=True).detach()
        n_ep = 0
        for _ in range(self.c['epoch_policy']):
            stats = self._policy_update(obs0, actions0, adv, behave_pol, ref_pol)
            n_ep += 1
            curr_pol = self.forward_actor(obs0).detach()
            klv = PD.kl(ref_pol, curr_pol, self.A).mean()
            stats['_pol_kl'] = klv.item()
            if klv.item() > self.c['kl_target'] * 4:
                break
        self.n_policy_epochs.append(n_ep)
        self.kl_record.append(stats['_pol_kl'])
        for _ in range(self.c['epoch_baseline']):
            bstats = self._value_update(obs0, ret)
        stats.update(bstats)
        behave_lik = PD.likelihood(actions0, behave_pol, self.A)
        curr_lik = PD.likelihood(actions0, curr_pol, self.A)
        stats['_avg_return_targ'] = ret.mean().item()
        stats['_avg_log_sig'] = self.log_var.mean().item()
        stats['_avg_behave_likelihood'] = behave_lik.mean().item()
        stats['_avg_is_weight'] = (curr_lik / (behave_lik + 1e-4)).mean().item()
        stats['_ref_behave_diff'] = PD.kl(ref_pol, behave_pol, self.A).mean().item()
        stats['_lr'] = self.actor_optim.param_groups[0]['lr']
        if self.zf is not None:
            self.zf.update(obs0)                                      # ppo.py:578-579 (AFTER the updates)
            stats['obs_running_mean'] = float(np.mean(self.zf.running_mean()))
            stats['obs_running_square'] = float(np.mean(self.zf.running_square()))
            stats['obs_running_std'] = float(np.mean(self.zf.running_std()))
        if self.rfilter is not None:
            stats['reward_mean'] = self.rfilter.reward_mean()
        self.exp_counter += self.batch_size
        return stats

    def publish_parameter(self):                                      # ppo.py:623-666
        if self.exp_counter < self.c['exp_interval']:
            return False
        final_kl = np.mean(self.kl_record)
        c = self.c
        if c['ppo_mode'] == 'clip':
            if final_kl > c['kl_target'] * c['adjust_threshold'][1]:
                if c['clip_range'][0] < self.clip_epsilon:
                    self.clip_epsilon = self.clip_epsilon / c['clip_scale_constant']
            elif final_kl < c['kl_target'] * c['adjust_threshold'][0]:
                if c['clip_range'][1] > self.clip_epsilon:
                    self.clip_epsilon = self.clip_epsilon * c['clip_scale_constant']
        else:
            if final_kl > c['kl_target'] * c['adjust_threshold'][1]:
                if c['beta_range'][1] > self.beta:
                    self.beta = self.beta * c['adapt_scale_constant']
            elif final_kl < c['kl_target'] * c['adjust_threshold'][0]:
                if c['beta_range'][0] < self.beta:
                    self.beta = self.beta / c['adapt_scale_constant']
        self._sync_ref()
        self.kl_record = []
        self.exp_counter = 0
        return True
