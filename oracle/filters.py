"""Running-statistics filters (reference: surreal/model/z_filter.py, surreal/model/reward_filter.py)."""
import torch


class ZFilter:
    """z_filter.py:7-107.  Buffers start at sum=0, sumsq=eps, count=eps (z_filter.py:40-42)."""

    def __init__(self, dim, eps=1e-5):
        self.eps = eps
        self.running_sum = torch.zeros(dim)
        self.running_sumsq = eps * torch.ones(dim)
        self.count = torch.tensor([eps], dtype=torch.float32)

    def load(self, running_sum, running_sumsq, count):
        self.running_sum = torch.as_tensor(running_sum, dtype=torch.float32).clone()
        self.running_sumsq = torch.as_tensor(running_sumsq, dtype=torch.float32).clone()
        self.count = torch.as_tensor(count, dtype=torch.float32).reshape(1).clone()
        return self

    def clone(self):
        return ZFilter(self.running_sum.numel(), self.eps).load(self.running_sum, self.running_sumsq, self.count)

    def update(self, x):
        """z_filter.py:44-57."""
        x = x.reshape(-1, self.running_sum.numel())
        self.running_sum += x.sum(dim=0)
        self.running_sumsq += (x * x).sum(dim=0)
        self.count += float(len(x))

    def forward(self, x):
        """z_filter.py:59-79: whiten, floor std at eps, clamp to +-5."""
        shape = x.shape
        x = x.reshape(-1, shape[-1])
        mean = self.running_sum / self.count
        std = torch.clamp((self.running_sumsq / self.count - mean.pow(2)).pow(0.5), min=self.eps)
        return torch.clamp((x - mean) / std, -5.0, 5.0).reshape(shape)

    def running_mean(self):
        return (self.running_sum / self.count).numpy()

    def running_square(self):
        return (self.running_sumsq / self.count).numpy()

    def running_std(self):
        return ((self.running_sumsq / self.count) - (self.running_sum / self.count).pow(2)).pow(0.5).numpy()


class RewardFilter:
    """reward_filter.py:5-63.  NOTE the reference OVERWRITES running_sumsq (reward_filter.py:42)."""

    def __init__(self, eps=1e-5):
        self.eps = eps
        self.count = torch.tensor(eps, dtype=torch.float32)
        self.running_sum = torch.tensor(0.0, dtype=torch.float32)
        self.running_sumsq = torch.tensor(0.0, dtype=torch.float32)

    def update(self, x):
        self.count += float(x.numel())
        self.running_sum += x.sum()
        self.running_sumsq = (x * x).sum()

    def forward(self, x):
        mean = self.running_sum / self.count
        std = torch.clamp((self.running_sumsq / self.count - mean.pow(2)).pow(0.5), min=self.eps)
        return torch.clamp((x - mean) / std, -5.0, 5.0)

    def reward_mean(self):
        return (self.running_sum / self.count).item()
