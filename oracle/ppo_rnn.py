"""PPO learner restatement, RNN mode (reference: surreal/learner/ppo.py:389-406,507-525; model/ppo_net.py:143-152,
202-224,253-315).  The reference's DEFAULT PPO config: an LSTM stem shared by actor and critic, trained by BOTH
optimisers (each holds its own Adam state for the LSTM parameters), horizon-windowed GAE over
eff_len = n_step - horizon + 1 positions, initial cells handed over by the actors through ``onetime_infos``.

TEST INFRASTRUCTURE (like the rest of oracle/): stock torch on the CPU, the same call sequence as the reference.
"""
import copy

import torch
import torch.nn as nn

from . import pd as PD
from . import nets
from .gae import gae_from_values
from .ppo import OraclePPOLearner


class OraclePPOLearnerRNN(OraclePPOLearner):
    def __init__(self, actor_layers, log_var, critic_layers, zfilter, lstm_state, action_dim, n_step, batch_size,
                 horizon, rnn_hidden, rnn_layer=1, **cfg):
        super().__init__(actor_layers, log_var, critic_layers, zfilter, action_dim, n_step, batch_size, **cfg)
        self.horizon = horizon
        in_dim = lstm_state['weight_ih_l0'].shape[1]
        self.rnn = nn.LSTM(in_dim, rnn_hidden, rnn_layer, batch_first=True)         # ppo_net.py:143-149
        self.rnn.load_state_dict({k: torch.as_tensor(v) for k, v in lstm_state.items()})
        self.ref_rnn = copy.deepcopy(self.rnn)
        for p in self.ref_rnn.parameters():
            p.requires_grad_(False)
        # get_actor_params / get_critic_params (ppo_net.py:202-224): head parameters, then the LSTM's
        self.actor_params = self.actor_params + list(self.rnn.parameters())
        self.critic_params = self.critic_params + list(self.rnn.parameters())
        c = self.c
        self.critic_optim = torch.optim.Adam(self.critic_params, lr=c['lr_critic'])
        self.actor_optim = torch.optim.Adam(self.actor_params, lr=c['lr_actor'])
        self.cells = None

    def _sync_ref(self):
        super()._sync_ref()
        if hasattr(self, 'rnn'):                                                    # ppo_net.py:235-236
            self.ref_rnn.load_state_dict(self.rnn.state_dict())

    # obs: [B, L, D]
    def forward_actor(self, obs, ref=False):
        zf = self.ref_zf if ref else self.zf
        x = zf.forward(obs) if zf is not None else obs
        feat, _ = (self.ref_rnn if ref else self.rnn)(x, self.cells)                # ppo_net.py:277-279
        feat = feat.contiguous()
        shp = feat.size()
        flat = feat.view(-1, shp[2])                                                # builders.py:121-131
        out = nets.ppo_actor(flat, self.ref_actor if ref else self.actor, self.ref_log_var if ref else self.log_var)
        return out.view(shp[0], shp[1], -1)

    def forward_critic(self, obs):
        x = self.zf.forward(obs) if self.zf is not None else obs
        feat, _ = self.rnn(x, self.cells)
        feat = feat.contiguous()
        shp = feat.size()
        return nets.ppo_critic(feat.view(-1, shp[2]), self.critic).view(shp[0], shp[1], -1)

    def _value_update(self, obs, returns):                                          # ppo.py:311-353, 3-D values
        values = self.forward_critic(obs)
        if values.dim() == 3:
            values = values.squeeze(2)                                              # ppo.py:324-325
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

    def gae_and_return(self, obs, obs_next, rewards, dones):                        # ppo.py:376-406
        cat = torch.cat([obs, obs_next], dim=1)                                     # [B, n+1, D], NOT flattened
        values = self.forward_critic(cat).view(self.batch_size, self.n_step + 1)
        self.last_values_raw = values.detach().clone()
        return gae_from_values(rewards, values, dones, self.c['gamma'], self.c['lam'], horizon=self.horizon,
                               norm_adv=self.c['norm_adv'])

    def learn(self, batch):                                                         # ppo.py:588-613 + 487-586
        self.current_iteration += 1
        obs, obs_next, actions, rewards, dones, pds = self.preprocess(batch)
        h = torch.tensor(batch['h0'], dtype=torch.float32).transpose(0, 1).contiguous().detach()   # ppo.py:507-510
        c = torch.tensor(batch['c0'], dtype=torch.float32).transpose(0, 1).contiguous().detach()
        self.cells = (h, c)
        adv, ret = self.gae_and_return(obs, obs_next, rewards, dones)
        adv, ret = adv.detach(), ret.detach()
        self.last_adv, self.last_ret = adv, ret
        eff = self.n_step - self.horizon + 1
        behave_pol = pds[:, :eff, :].contiguous()
        actions_it = actions[:, :eff, :].contiguous()
        obs_it = obs[:, :eff, :].contiguous()
        ref_pol = self.forward_actor(obs_it, ref=True).detach()
        n_ep = 0
        for _ in range(self.c['epoch_policy']):
            stats = self._policy_update(obs_it, actions_it, adv, behave_pol, ref_pol)
            n_ep += 1
            curr_pol = self.forward_actor(obs_it).detach()
            klv = PD.kl(ref_pol, curr_pol, self.A).mean()
            stats['_pol_kl'] = klv.item()
            if klv.item() > self.c['kl_target'] * 4:
                break
        self.n_policy_epochs.append(n_ep)
        self.kl_record.append(stats['_pol_kl'])
        for _ in range(self.c['epoch_baseline']):
            bstats = self._value_update(obs_it, ret)
        stats.update(bstats)
        behave_lik = PD.likelihood(actions_it, behave_pol, self.A)
        curr_lik = PD.likelihood(actions_it, curr_pol, self.A)
        stats['_avg_return_targ'] = ret.mean().item()
        stats['_avg_log_sig'] = self.log_var.mean().item()
        stats['_avg_behave_likelihood'] = behave_lik.mean().item()
        stats['_avg_is_weight'] = (curr_lik / (behave_lik + 1e-4)).mean().item()
        stats['_ref_behave_diff'] = PD.kl(ref_pol, behave_pol, self.A).mean().item()
        stats['_lr'] = self.actor_optim.param_groups[0]['lr']
        if self.zf is not None:
            import numpy as np
            self.zf.update(obs_it)                                                  # ppo.py:578-579
            stats['obs_running_mean'] = float(np.mean(self.zf.running_mean()))
            stats['obs_running_square'] = float(np.mean(self.zf.running_square()))
            stats['obs_running_std'] = float(np.mean(self.zf.running_std()))
        self.exp_counter += self.batch_size
        return stats
