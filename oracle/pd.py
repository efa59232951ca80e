"""Diagonal-Gaussian policy distribution (reference: surreal/model/ppo_net.py:13-91).

``prob`` rows are ``[mean(A) | std(A)]``.  Formulas are reproduced as written in the reference,
including its quirks (entropy uses 0.5*sum(log std), ppo_net.py:72).
"""
import math

import numpy as np
import torch

_LOG_2PI = math.log(2.0 * math.pi)
_LOG_2PIE = math.log(2.0 * math.pi * math.e)


def _split(prob, d):
    prob = prob.reshape(-1, 2 * d)
    return prob[:, :d], prob[:, d:]


def loglikelihood(a, prob, d):
    """ppo_net.py:29-40 -> shape [rows, 1]."""
    a = a.reshape(-1, d)
    mean, std = _split(prob, d)
    quad = (((a - mean) / std) ** 2).sum(dim=1, keepdim=True)
    return -0.5 * quad - 0.5 * _LOG_2PI * d - std.log().sum(dim=1, keepdim=True)


def likelihood(a, prob, d):
    """ppo_net.py:42-46: exp(loglik) floored at 1e-5."""
    return torch.clamp(loglikelihood(a, prob, d).exp(), min=1e-5)


def kl(prob0, prob1, d):
    """ppo_net.py:48-62: KL(p0 || p1) -> shape [rows]."""
    m0, s0 = _split(prob0, d)
    m1, s1 = _split(prob1, d)
    return (s1 / s0).log().sum(dim=1) + ((s0 ** 2 + (m0 - m1) ** 2) / (2.0 * s1 ** 2)).sum(dim=1) - 0.5 * d


def entropy(prob, d):
    """ppo_net.py:64-72 (half the textbook log-std term, as in the reference)."""
    _, std = _split(prob, d)
    return 0.5 * std.log().sum(dim=1) + 0.5 * _LOG_2PIE * d


def sample(prob_np, d, eps):
    """ppo_net.py:74-83 with the N(0,1) draws ``eps`` injected (float64, like np.random.randn)."""
    prob_np = np.asarray(prob_np).reshape(-1, 2 * d)
    return eps * prob_np[:, d:] + prob_np[:, :d]


def maxprob(prob_np, d):
    """ppo_net.py:85-91 (2-D branch)."""
    return np.asarray(prob_np).reshape(-1, 2 * d)[:, :d]
