"""Batch aggregation (reference: surreal/learner/aggregator.py:33-103, 106-262), array-level."""
import numpy as np


def multistep_aggregate(windows):
    """``windows``: list of dicts with per-window arrays obs[n,D], obs_next[D], actions[n,A] (float64),
    rewards[n] (python floats), dones[n] (bool), pd[n,2A] (float32).  Output dtypes follow the
    reference: actions/rewards stay float64, dones become float32 (aggregator.py:176-183)."""
    return {
        'obs': np.stack([np.stack(w['obs']) for w in windows]),
        'obs_next': np.stack([np.stack([w['obs_next']]) for w in windows]),
        'actions': np.stack([np.stack(w['actions']) for w in windows]),
        'rewards': np.stack([np.array(w['rewards']) for w in windows]),
        'dones': np.stack([np.array(w['dones']) for w in windows]).astype('float32'),
        'pd': np.asarray([np.stack(w['pd']) for w in windows]),
    }


def ssar_aggregate(exps):
    """aggregator.py:52-103: actions float32; rewards / dones expand to [B,1] float64."""
    return {
        'obs': np.array([np.asarray(e['obs']) for e in exps]),
        'obs_next': np.array([np.asarray(e['obs_next']) for e in exps]),
        'actions': np.array([e['action'] for e in exps], dtype=np.float32),
        'rewards': np.expand_dims([e['reward'] for e in exps], axis=1),
        'dones': np.expand_dims([float(e['done']) for e in exps], axis=1),
    }
