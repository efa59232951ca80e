"""Periodic checkpointing in the reference's on-disk format (surreal/utils/checkpoint.py:18-347):

  <folder>/<name>.<global_steps>.ckpt   pickle of OrderedDict {attr: state_dict() or plain value}
  <folder>/metadata.<name>.yml          version, save_counter, history_ckpt_files (newest first), ckpt{...},
                                        tracked_attrs, keep_history, keep_best, best_ckpt_files, best_scores

so model / target-model / counter attributes written by either implementation restore into the other.
Attributes exposing ``state_dict()/load_state_dict()`` are stored through them (tensors moved to the CPU);
everything else is pickled as is.  One asymmetry is unavoidable: the reference pickles its LR schedulers as
OBJECTS of an un-vendored class (torchx); reading such a file works here (the object's step count is mapped onto
our schedule, unknown classes load as placeholders), while the reference would need torchx importable to read its
own file and cannot rebuild its scheduler objects from the state dicts written here.  Saving happens every ``period`` calls AND at least ``min_interval`` seconds apart."""
import datetime
import os
import pickle
import shutil
import time
from collections import OrderedDict

import torch
import yaml

CHECKPOINT_VERSION = '0.0.1'


class ForeignObject:
    """Placeholder for an instance of a class that cannot be imported here.  The reference pickles non-Module tracked
    attributes AS OBJECTS (utils/checkpoint.py:239-246) -- a PPO learner checkpoint holds two
    ``torchx.nn.hyper_scheduler.LinearWithMinLR`` instances -- and torchx is not installable; the rest of such a file
    must still load.  The instance keeps whatever state the pickle carried in ``__dict__``."""
    _foreign_name = '?'

    def __init__(self, *args, **kwargs):
        self._foreign_args = args

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self._foreign_state = state


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            return type(name, (ForeignObject,), {'_foreign_name': '%s.%s' % (module, name)})


def _as_state_dict(value):
    """What to hand to ``load_state_dict`` for a tracked attribute: dicts pass through; objects (the reference's pickled
    schedulers, or their ForeignObject stand-ins) contribute their plain scalar fields."""
    if isinstance(value, dict):
        return value
    if hasattr(value, 'state_dict') and not isinstance(value, ForeignObject):
        try:
            return value.state_dict()
        except Exception:
            pass
    fields = getattr(value, '__dict__', {})
    return {k: v for k, v in fields.items() if isinstance(v, (int, float, bool, list, tuple, str))}


def _to_cpu(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu()
    if isinstance(obj, dict):
        return type(obj)((k, _to_cpu(v)) for k, v in obj.items())
    return obj


class Checkpoint:
    def __init__(self, folder, name, *, tracked_obj, tracked_attrs=None, keep_history=1, keep_best=1, mkdir=True):
        self.folder = os.path.expanduser(folder)
        if mkdir:
            os.makedirs(self.folder, exist_ok=True)
        self.name = name
        self.tracked_obj = tracked_obj
        if os.path.exists(self.metadata_path()):
            self._load_metadata()
        else:
            if tracked_attrs is not None:
                assert isinstance(tracked_attrs, (list, tuple)) and all(isinstance(a, str) for a in tracked_attrs), \
                    'tracked_attrs must be a list of attribute name strings or None'
            assert keep_history >= 1 and keep_best >= 0
            self.metadata = dict(version=CHECKPOINT_VERSION, save_counter=0, history_ckpt_files=[], ckpt={},
                                 tracked_attrs=list(tracked_attrs) if tracked_attrs is not None else None,
                                 keep_history=keep_history, keep_best=keep_best, best_ckpt_files=[], best_scores=[])

    # -- paths ---------------------------------------------------------------------------------------
    def metadata_name(self):
        return 'metadata.{}.yml'.format(self.name)

    def metadata_path(self):
        return os.path.join(self.folder, self.metadata_name())

    def ckpt_name(self, suffix):
        return '{}.{}.ckpt'.format(self.name, suffix)

    def ckpt_path(self, suffix):
        return os.path.join(self.folder, self.ckpt_name(suffix))

    def _load_metadata(self):
        with open(self.metadata_path()) as fp:
            self.metadata = yaml.safe_load(fp)
        if self.metadata.get('version') != CHECKPOINT_VERSION:
            raise ValueError('checkpoint version incompatible, please examine {} and make sure it is {}'
                             .format(self.metadata_path(), CHECKPOINT_VERSION))

    def _save_metadata(self):
        with open(self.metadata_path(), 'w') as fp:
            yaml.safe_dump(self.metadata, fp, default_flow_style=False)

    # -- save ----------------------------------------------------------------------------------------
    def _dump(self, suffix):
        attrs = self.metadata['tracked_attrs']
        assert attrs is not None, 'tracked_attrs must not be None for save()'
        data = OrderedDict()
        for a in attrs:
            v = getattr(self.tracked_obj, a)
            data[a] = _to_cpu(v.state_dict()) if hasattr(v, 'state_dict') else v
        with open(self.ckpt_path(suffix), 'wb') as fp:
            pickle.dump(data, fp)

    def save(self, score=None, global_steps=None, reload_metadata=False, **ckpt_info):
        if reload_metadata:
            self._load_metadata()
        meta = self.metadata
        meta['save_counter'] += 1
        if global_steps is None:
            global_steps = meta['save_counter']
        self._dump(global_steps)
        meta['global_steps'] = global_steps
        files = [self.ckpt_name(global_steps)] + [f for f in meta['history_ckpt_files']
                                                  if f != self.ckpt_name(global_steps)]
        for old in files[meta['keep_history']:]:
            p = os.path.join(self.folder, old)
            if os.path.exists(p):
                os.remove(p)
        meta['history_ckpt_files'] = files[:meta['keep_history']]
        entry = dict(score=score, global_steps=global_steps, save_counter=meta['save_counter'], time=time.time(),
                     datetime=str(datetime.datetime.now()))
        entry.update(ckpt_info)
        meta['ckpt'][self.ckpt_name(global_steps)] = entry
        if meta['keep_best'] > 0:
            assert score is not None, 'score cannot be None if keep_best is enabled'
            best_name = self.ckpt_name('best-{}'.format(global_steps))
            ranked = sorted(list(zip(meta['best_scores'], meta['best_ckpt_files'])) + [(score, best_name)],
                            key=lambda t: -t[0])
            keep, drop = ranked[:meta['keep_best']], ranked[meta['keep_best']:]
            if (score, best_name) in keep:
                shutil.copy(self.ckpt_path(global_steps), os.path.join(self.folder, best_name))
                meta['ckpt'][best_name] = entry
            for _, f in drop:
                p = os.path.join(self.folder, f)
                if os.path.exists(p):
                    os.remove(p)
                meta['ckpt'].pop(f, None)
            meta['best_scores'] = [s for s, _ in keep]
            meta['best_ckpt_files'] = [f for _, f in keep]
        self._save_metadata()

    # -- restore -------------------------------------------------------------------------------------
    def _restore_file(self, ckpt_file, check_ckpt_exists):
        path = os.path.join(self.folder, ckpt_file)
        if not os.path.exists(path):
            if check_ckpt_exists:
                raise FileNotFoundError(path + ' missing.')
            return None
        with open(path, 'rb') as fp:
            data = _TolerantUnpickler(fp).load()
        for a in self.metadata['tracked_attrs']:
            cur = getattr(self.tracked_obj, a)
            if hasattr(cur, 'load_state_dict'):
                cur.load_state_dict(_as_state_dict(data[a]))
            elif isinstance(data[a], ForeignObject):
                raise TypeError('checkpoint attribute %r is an instance of %s, which cannot be imported here'
                                % (a, data[a]._foreign_name))
            else:
                setattr(self.tracked_obj, a, data[a])
        return path

    def restore(self, target, mode, reload_metadata=True, check_ckpt_exists=False, restore_folder=None):
        assert mode in ('best', 'history')
        old = self.folder
        if restore_folder:
            assert os.path.exists(restore_folder)
            self.folder = os.path.expanduser(restore_folder)
        try:
            if reload_metadata or restore_folder:
                if not os.path.exists(self.metadata_path()):
                    if check_ckpt_exists:
                        raise FileNotFoundError(self.metadata_path())
                    return None
                self._load_metadata()
            meta = self.metadata
            if isinstance(target, int):
                assert target >= 0
                files = meta['best_ckpt_files'] if mode == 'best' else meta['history_ckpt_files']
                if target >= len(files):
                    if check_ckpt_exists:
                        raise FileNotFoundError('{} [{}] ckpt file missing'.format(mode.capitalize(), target))
                    return None
                ckpt_file = files[target]
            else:
                assert '.ckpt' not in target
                ckpt_file = self.ckpt_name('best-{}'.format(target) if mode == 'best' else target)
            return self._restore_file(ckpt_file, check_ckpt_exists)
        finally:
            self.folder = old


class PeriodicCheckpoint(Checkpoint):
    def __init__(self, *args, period, min_interval=0, **kwargs):
        super().__init__(*args, **kwargs)
        assert period >= 1
        self.period = period
        self._period_counter = 0
        self.min_interval = min_interval
        self.last_update_time = time.time()

    def save(self, *args, **kwargs):
        self._period_counter += 1
        if self._period_counter % self.period == 0 and time.time() - self.last_update_time >= self.min_interval:
            super().save(*args, **kwargs)
            self.last_update_time = time.time()
            return True
        return False

    def reset_period(self):
        self._period_counter = 0
