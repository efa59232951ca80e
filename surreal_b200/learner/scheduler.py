"""Learning-rate schedule named by ``algo.network.anneal.lr_scheduler`` (ppo.py:121-125,171-178).

The reference resolves the name through ``torchx.nn.hyper_scheduler``, whose source is not vendored
(the exact LinearWithMinLR formula is UNPINNED).  This implementation follows the name:
linear decay from the initial rate to zero over ``num_updates`` scheduler steps, refreshed every
``update_freq`` steps and floored at ``min_lr``.  The rate lives in the optimiser's DEVICE lr slot, so
a step() is one tiny host->device write at publish time."""


class LinearWithMinLR:
    def __init__(self, optim, base_lr, num_updates, update_freq=1, min_lr=0.0):
        self.optim = optim
        self.base_lr = float(base_lr)
        self.num_updates = max(int(num_updates), 1)
        self.update_freq = max(int(update_freq), 1)
        self.min_lr = float(min_lr)
        self.n_step = 0
        self.lr = float(base_lr)

    def step(self):
        self.n_step += 1
        if self.n_step % self.update_freq == 0:
            frac = max(0.0, 1.0 - self.n_step / self.num_updates)
            self.lr = max(self.min_lr, self.base_lr * frac)
            self.optim.set_lr(self.lr)

    def get_lr(self):
        return [self.lr]

    def state_dict(self):
        return {'n_step': self.n_step, 'lr': self.lr}

    def load_state_dict(self, sd):
        """Own checkpoints carry {n_step, lr}.  A checkpoint written by the REFERENCE holds torchx's scheduler dict
        (its source is not vendored; torch-style schedulers use `last_epoch` / `_step_count` / `base_lrs`): the step
        count is taken from whichever of those keys exists and the rate is recomputed with this class's formula;
        a dict with none of them leaves the schedule at its initial state (with a warning) instead of raising."""
        n = None
        for k in ('n_step', 'last_epoch', '_step_count', 'step_count', 'num_steps'):
            if k in sd:
                n = int(sd[k]) - (1 if k == '_step_count' else 0)
                break
        if n is None:
            import warnings
            warnings.warn('LinearWithMinLR.load_state_dict: foreign scheduler state %s has no step count; the '
                          'schedule restarts' % sorted(sd.keys()))
            return
        self.n_step = max(n, 0)
        base = sd.get('base_lrs')
        if base:
            self.base_lr = float(base[0])
        if 'lr' in sd:
            self.lr = float(sd['lr'])
        else:                                   # recompute with this class's formula
            steps = self.n_step // self.update_freq * self.update_freq
            self.lr = self.base_lr if steps == 0 else max(self.min_lr, self.base_lr * max(0.0, 1.0 - steps / self.num_updates))
        self.optim.set_lr(self.lr)


def make_lr_scheduler(name, optim, base_lr, num_updates, update_freq, min_lr):
    if name != 'LinearWithMinLR':
        raise ValueError('unknown lr_scheduler "%s" (only LinearWithMinLR is referenced by the reference '
                         'configs, ppo_configs.py:43)' % name)
    return LinearWithMinLR(optim, base_lr, num_updates, update_freq=update_freq, min_lr=min_lr)
