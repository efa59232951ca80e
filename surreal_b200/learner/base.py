"""Learner base class: same constructor, hooks and main loop as surreal/learner/base.py:21-389, with the
three ZeroMQ wires collapsed to in-process objects:

  * batches     : ``attach_replay(replay)`` -- ``fetch_batch()`` pulls straight from the HBM replay
                  (replaces LearnerDataPrefetcher + REQ/REP, data_fetcher.py:9-58);
  * parameters  : ``ParameterPublisher`` is a device-side versioned snapshot (parameter_server.py:20-55);
  * metrics     : ``self.tensorplex.add_scalars`` lands in a ScalarSink.
"""
import os
import time
from pathlib import Path

from .. import utils as U
from ..distributed import ParameterPublisher, LocalHub
from ..session import Config


class Learner(metaclass=U.AutoInitializeMeta):
    def __init__(self, learner_config, env_config, session_config):
        self.learner_config = learner_config
        self.env_config = env_config
        self.session_config = session_config
        self.current_iter = 0
        self._replay = None
        self._setup_logging()

    # -- abstract -----------------------------------------------------------------------------------
    def learn(self, batch_exp):
        raise NotImplementedError

    def module_dict(self):
        raise NotImplementedError

    def checkpoint_attributes(self):
        return []

    # -- wiring -------------------------------------------------------------------------------------
    def _initialize(self):
        from ..checkpoint import PeriodicCheckpoint
        ck = self.session_config.checkpoint
        self._periodic_checkpoint = PeriodicCheckpoint(
            os.path.join(self.session_config.folder, 'checkpoint'), name='learner',
            period=ck.learner.periodic, min_interval=ck.learner.min_interval, tracked_obj=self,
            tracked_attrs=self.checkpoint_attributes(), keep_history=ck.learner.keep_history,
            keep_best=ck.learner.keep_best)
        if ck.restore:
            self.restore_checkpoint()
        self._ps_publish_tracker = U.TimedTracker(self.learner_config.parameter_publish.min_publish_interval)
        self._ps_publisher = ParameterPublisher(module_dict=self.module_dict())
        LocalHub.get(self.session_config).publisher = self._ps_publisher

    def attach_replay(self, replay):
        self._replay = replay

    @property
    def publisher(self):
        return self._ps_publisher

    def should_publish_parameter(self):
        return self._ps_publish_tracker.track_increment()

    def publish_parameter(self, iteration, message=''):
        self._ps_publisher.publish(iteration, message=message)

    def fetch_batch(self):
        if self._replay is None:
            raise RuntimeError('no replay attached: call learner.attach_replay(replay) (the ZeroMQ prefetcher '
                               'of the reference is collapsed to an in-process pull)')
        out = self.replay_out_buffers() if hasattr(self, 'replay_out_buffers') else None
        if out is not None:
            batch = self._replay.sample(self.learner_config.replay.batch_size, out=out)
        else:
            batch = self._replay.sample(self.learner_config.replay.batch_size)
        return self.preprocess(self._prefetcher_preprocess(batch))

    def fetch_iterator(self):
        while True:
            yield self.fetch_batch()

    def preprocess(self, batch):
        return batch

    def _prefetcher_preprocess(self, batch):
        return batch

    # -- logging / checkpoint -----------------------------------------------------------------------
    def _setup_logging(self):
        self.learn_timer = U.TimeRecorder()
        self.iter_timer = U.TimeRecorder()
        self.publish_timer = U.TimeRecorder()
        self.init_time = time.time()
        self.log = U.get_logger('learner')
        self.tensorplex = U.ScalarSink('learner/learner')

    def periodic_checkpoint(self, global_steps, score=None, **info):
        return self._periodic_checkpoint.save(score=score, global_steps=global_steps, **info)

    def restore_checkpoint(self):
        SC = self.session_config
        folder = SC.checkpoint.restore_folder
        if folder and os.path.basename(os.path.normpath(folder)) != 'checkpoint':
            folder = os.path.join(folder, 'checkpoint')
        restored = self._periodic_checkpoint.restore(target=SC.checkpoint.learner.restore_target,
                                                     mode=SC.checkpoint.learner.mode, restore_folder=folder)
        if restored:
            self.log.info('successfully restored from checkpoint %s', restored)

    # -- main loop (learner/base.py:348-376) ----------------------------------------------------------
    def main(self):
        self.main_setup()
        while True:
            self.main_loop()

    def main_setup(self):
        self.save_config()
        self.iter_timer.start()
        self.publish_parameter(0, message='batch ' + str(0))

    def main_loop(self):
        data = self.fetch_batch()
        with self.learn_timer.time():
            self.learn(data)
        if self.should_publish_parameter():
            with self.publish_timer.time():
                self.publish_parameter(self.current_iter, message='batch ' + str(self.current_iter))
        self.iter_timer.lap()
        self.current_iter += 1

    def save_config(self):
        folder = Path(self.session_config.folder)
        folder.mkdir(exist_ok=True, parents=True)
        Config(learner_config=self.learner_config, env_config=self.env_config,
               session_config=self.session_config).dump_file(str(folder / 'config.yml'))
