"""PPO learner restatement, pixel mode (reference: surreal/model/ppo_net.py:136-140,202-224,268-273,368-375;
model_builders/builders.py:8-33; BASELINE cfg 4).  uint8 frames are scaled by 1/255 and pass through a CNN stem
(Conv-ReLU x len(channels), Flatten, Linear-ReLU) that actor and critic SHARE and that BOTH optimisers train, each with
its own Adam state; the z-filter is off (it needs low-dim inputs, ppo_net.py:164-166).

TEST INFRASTRUCTURE (like the rest of oracle/): stock torch on the CPU, the same call sequence as the reference.
"""
import torch
import torch.nn.functional as F

from . import nets
from .ppo import OraclePPOLearner


class OraclePPOLearnerPixel(OraclePPOLearner):
    def __init__(self, actor_layers, log_var, critic_layers, conv_layers, strides, fc_layer, action_dim, n_step,
                 batch_size, **cfg):
        """conv_layers: [(weight [Co,Ci,k,k], bias [Co])], strides: [int], fc_layer: (weight [F, flat], bias [F])."""
        cfg = dict(cfg)
        cfg['use_z_filter'] = False
        super().__init__(actor_layers, log_var, critic_layers, None, action_dim, n_step, batch_size, **cfg)
        self.conv = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in conv_layers]
        self.strides = list(strides)
        self.fc = (fc_layer[0].clone().requires_grad_(True), fc_layer[1].clone().requires_grad_(True))
        self.ref_conv = [(w.detach().clone(), b.detach().clone()) for w, b in self.conv]
        self.ref_fc = (self.fc[0].detach().clone(), self.fc[1].detach().clone())
        stem = [t for wb in self.conv for t in wb] + list(self.fc)
        self.actor_params = self.actor_params + stem                       # get_actor_params (ppo_net.py:202-212)
        self.critic_params = self.critic_params + stem
        c = self.c
        self.critic_optim = torch.optim.Adam(self.critic_params, lr=c['lr_critic'])
        self.actor_optim = torch.optim.Adam(self.actor_params, lr=c['lr_actor'])

    def _sync_ref(self):
        super()._sync_ref()
        if hasattr(self, 'conv'):                                          # ppo_net.py:238-239
            self.ref_conv = [(w.detach().clone(), b.detach().clone()) for w, b in self.conv]
            self.ref_fc = (self.fc[0].detach().clone(), self.fc[1].detach().clone())

    def stem(self, frames, ref=False):
        """frames: [rows, C, H, W] float32 holding 0..255."""
        h = frames / 255.0                                                 # ppo_net.py:368-375
        for (w, b), s in zip(self.ref_conv if ref else self.conv, self.strides):
            h = torch.relu(F.conv2d(h, w, b, stride=s))
        h = h.flatten(1)
        w, b = self.ref_fc if ref else self.fc
        return torch.relu(F.linear(h, w, b))

    def forward_actor(self, obs, ref=False):
        feat = self.stem(obs, ref)
        if ref:
            return nets.ppo_actor(feat, self.ref_actor, self.ref_log_var)
        return nets.ppo_actor(feat, self.actor, self.log_var)

    def forward_critic(self, obs):
        return nets.ppo_critic(self.stem(obs), self.critic)

    def gae_and_return(self, obs, obs_next, rewards, dones):               # ppo.py:376-418 (MLP branch, 5-D obs)
        from .gae import gae_from_values
        cat = torch.cat([obs, obs_next], dim=1)
        flat = cat.view(-1, *cat.shape[2:])
        values = self.forward_critic(flat).view(self.batch_size, self.n_step + 1)
        self.last_values_raw = values.detach().clone()
        return gae_from_values(rewards, values, dones, self.c['gamma'], self.c['lam'], norm_adv=self.c['norm_adv'])
