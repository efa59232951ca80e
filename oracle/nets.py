"""Network forward passes, functional over explicit parameter lists.

reference: surreal/model/model_builders/builders.py:35-175 (ActorNetworkX, CriticNetworkX,
PPO_ActorNetwork, PPO_CriticNetwork) and surreal/model/ppo_net.py:253-315, ddpg_net.py:63-91.

A network is a list of ``(W [out,in], b [out])`` torch tensors (torch.nn.Linear convention).
"""
import torch
import torch.nn.functional as F


def params_from_state(sd, prefix, n_layers=3):
    """Pull ``[(W,b)]`` out of a reference-style state_dict whose Linear layers sit at
    ``<prefix>seq/{0,2,4}/{weight,bias}`` (golden fixtures use '/' as separator)."""
    out = []
    for i in range(n_layers):
        out.append((torch.tensor(sd['%sseq/%d/weight' % (prefix, 2 * i)]),
                    torch.tensor(sd['%sseq/%d/bias' % (prefix, 2 * i)])))
    return out


def mlp_trunk(x, layers):
    """Linear-ReLU-Linear-ReLU-Linear (no final activation)."""
    h = x
    for i, (w, b) in enumerate(layers):
        h = F.linear(h, w, b)
        if i < len(layers) - 1:
            h = torch.relu(h)
    return h


def ppo_actor(x, layers, log_var):
    """builders.py:114-132: mean = tanh(trunk); std = exp(log_var) broadcast; cat on dim 1."""
    mean = torch.tanh(mlp_trunk(x, layers))
    std = torch.exp(log_var) * torch.ones(mean.size())
    return torch.cat((mean, std), dim=1)


def ppo_critic(x, layers):
    """builders.py:160-175."""
    return mlp_trunk(x, layers)


def ddpg_actor(x, layers):
    """builders.py:35-56 without layernorm (ddpg_configs.py:21): tanh output."""
    return torch.tanh(mlp_trunk(x, layers))


def ddpg_critic(x, act, layers):
    """builders.py:58-84: h = relu(L0 x); cat(h, act); relu(L1 .); L2 -- action enters at layer 2."""
    (w0, b0), (w1, b1), (w2, b2) = layers
    h = torch.relu(F.linear(x, w0, b0))
    h = torch.relu(F.linear(torch.cat((h, act), 1), w1, b1))
    return F.linear(h, w2, b2)
