"""Data-parallel learner plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink / NVSwitch on the
GPU box, gloo in the CPU tests) for the few exchanges the path needs:

  * ONE flat gradient all-reduce per optimiser step (the flat parameter layout makes it a single collective),
  * the moments of the advantage normalisation (global batch: ppo.py:413-416),
  * the mean-KL scalar that drives the early stop (so every rank takes the same branch),
  * the z-filter sums, and the reporting statistics.

The reference has no collective at all (its learner is single-process, README.md:24); actors / windows / replay
shards are independent, so nothing on the rollout side communicates."""
import torch
import torch.distributed as dist


class LearnerDP:
    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def broadcast_(self, t, src=0):
        dist.broadcast(t, src=src, group=self.group)
        return t

    def sum_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def mean_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        t.div_(self.world)
        return t


def combine_moments(moments):
    """[sum, sumsq, count] (already summed over ranks) -> (mean, unbiased std): what sb200_normalize_f32 applies."""
    s, q, n = float(moments[0]), float(moments[1]), float(moments[2])
    mean = s / n
    var = (q - s * s / n) / (n - 1.0)
    return mean, max(var, 0.0) ** 0.5
