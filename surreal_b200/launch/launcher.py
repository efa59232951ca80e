"""Component launcher with the reference's surface (surreal/launch/launcher.py:20-440):
``<exe> <component> -- <config flags>``, ``SurrealDefaultLauncher(agent_class, learner_class, replay_class,
session_config, env_config, learner_config)``, ``setup_agent`` / ``setup_learner`` / ``run_replay_worker``.

In the reference every component is its own OS process wired over ZeroMQ by symphony.  Here agents, replay
and learner co-locate in ONE process per GPU, so the launcher's job collapses to constructing the three
plugin objects in the right order and running the engine loop; the pure-plumbing components (ps, tensorplex,
loggerplex, tensorboard, replay load balancer) are accepted and are no-ops."""
import sys
from argparse import ArgumentParser

import numpy as np

from .. import utils as U

_NOOP_COMPONENTS = ('ps', 'tensorplex', 'loggerplex', 'tensorboard', 'replay_loadbalancer')


class Launcher:
    def main(self, argv=None):
        argv = sys.argv[1:] if argv is None else list(argv)
        parser_args, config_args = argv, []
        if '--' in argv:
            i = argv.index('--')
            parser_args, config_args = argv[:i], argv[i + 1:]
        parser = ArgumentParser(description='launch a surreal component')
        parser.add_argument('component_name', type=str, help='which component to launch')
        args = parser.parse_args(parser_args)
        self.config_args = config_args
        self.setup(config_args)
        return self.launch(args.component_name)

    def launch(self, component_name):
        raise NotImplementedError

    def setup(self, args):
        pass


class PipelinedEngine:
    """Actors and learner as two concurrent CUDA streams of ONE process (the B200 form of Surreal's asynchronous
    actor / learner processes): while the learner works on batch k, the actors already roll out batch k+1 with the
    weights published after batch k-1 -- exactly the policy lag Surreal's actors run with (docs/ppo.md:16).

      stream A:  [wait pop_k, pub]  rollout_{k+1} .........................  -> roll_{k+1}
      stream L:  [wait roll_k] pop+gather_k -> pop_k | learn_k (CUDA graph) ... publish -> pub

    The rollout is a chain of small latency-bound kernels that occupies a handful of SMs; the learner's kernels fill
    the rest, so the step time tends to max(rollout, learn) instead of their sum.  Ordering between the streams is
    carried by events only (FIFO control block: push happens-before pop happens-before next push; parameter
    snapshot: publish happens-before fetch happens-before next publish)."""

    def __init__(self, agent, replay, learner, rollout_steps):
        import torch
        self.torch = torch
        self.agent, self.replay, self.learner, self.T = agent, replay, learner, rollout_steps
        self.sA, self.sL = torch.cuda.Stream(), torch.cuda.Stream()
        self.ev_roll = self.ev_pop = self.ev_pub = self.ev_fetch = None
        self.iterations = 0
        if hasattr(learner, 'before_epochs_hook'):
            # learners whose epochs are ONE all-SM persistent kernel take the GPU after the rollout, not beside it
            learner.before_epochs_hook = self._wait_rollout
        cur = torch.cuda.current_stream()
        self.sA.wait_stream(cur)
        self.sL.wait_stream(cur)

    def _rollout(self):
        torch = self.torch
        with torch.cuda.stream(self.sA):
            if self.ev_pop is not None:
                self.sA.wait_event(self.ev_pop)
            if self.ev_pub is not None:
                self.sA.wait_event(self.ev_pub)
            self.agent.main_loop(max_steps=self.T)          # parameter fetch (if due) + T batched env steps
            self.ev_fetch = torch.cuda.Event()
            self.ev_fetch.record(self.sA)
            self.ev_roll = torch.cuda.Event()
            self.ev_roll.record(self.sA)

    def _wait_rollout(self):
        """Called by the learner between its head graph and its epochs graph (learner stream current): the rollout that
        was enqueued before this learn() must have finished."""
        if self.ev_roll is not None:
            self.torch.cuda.current_stream().wait_event(self.ev_roll)

    def prime(self):
        """First rollout (nothing to overlap with yet)."""
        self._rollout()

    def step(self):
        """One learner iteration, overlapped with the actors' next rollout.  Returns the learner statistics."""
        torch = self.torch
        L = self.learner
        with torch.cuda.stream(self.sL):
            self.sL.wait_event(self.ev_roll)
            if not self.replay.start_sample_condition():
                raise RuntimeError('replay under-filled: rollout_steps too small for one learner batch')
            data = L.fetch_batch()                           # pop + gather into the learner's buffers
            self.ev_pop = torch.cuda.Event()
            self.ev_pop.record(self.sL)
        self._rollout()                                      # enqueue rollout k+1 BEFORE blocking on learn k
        with torch.cuda.stream(self.sL):
            with L.learn_timer.time():
                stats = L.learn(data)                        # graph replay + one statistics read-back
            if self.ev_fetch is not None:
                self.sL.wait_event(self.ev_fetch)            # the snapshot buffers may still be read by a fetch
            if L.should_publish_parameter():
                L.publish_parameter(L.current_iter, message='batch ' + str(L.current_iter))
            self.ev_pub = torch.cuda.Event()
            self.ev_pub.record(self.sL)
            L.iter_timer.lap()
            L.current_iter += 1
        self.iterations += 1
        return stats

    def drain(self):
        self.torch.cuda.synchronize()


class SurrealDefaultLauncher(Launcher):
    def __init__(self, agent_class, learner_class, replay_class, session_config, env_config, learner_config,
                 eval_mode='eval_stochastic', agent_batch_size=8, eval_batch_size=8, render=False):
        self.agent_class = agent_class
        self.learner_class = learner_class
        self.replay_class = replay_class
        self.session_config = session_config
        self.env_config = env_config
        self.learner_config = learner_config
        self.eval_mode = eval_mode
        self.render = render
        self.agent_batch_size = agent_batch_size
        self.eval_batch_size = eval_batch_size
        self.log = U.get_logger('launcher')
        self.replay = self.learner = self.agent = None

    # -- constructors the reference exposes (launcher.py:192-217, 317-332, 369-381) ----------------------
    def setup_replay(self, replay_id=0):
        self.replay = self.replay_class(self.learner_config, self.env_config, self.session_config, index=replay_id)
        return self.replay

    def setup_learner(self):
        self.learner = self.learner_class(learner_config=self.learner_config, env_config=self.env_config,
                                          session_config=self.session_config)
        if self.replay is not None:
            self.learner.attach_replay(self.replay)
        return self.learner

    def setup_agent(self, agent_id):
        np.random.seed(int(__import__('time').time() * 100000 % 100000))       # launcher.py:201
        self.agent = self.agent_class(learner_config=self.learner_config, env_config=self.env_config,
                                      session_config=self.session_config, agent_id=agent_id, agent_mode='training')
        return self.agent

    def setup_eval(self, eval_id):
        return self.agent_class(learner_config=self.learner_config, env_config=self.env_config,
                                session_config=self.session_config, agent_id=eval_id, agent_mode=self.eval_mode,
                                render=self.render)

    # -- engine -------------------------------------------------------------------------------------------
    def setup_engine(self, env=None):
        """replay -> learner -> agent, then the initial publish + fetch (learner/base.py:356-362,
        agent/base.py:234-242)."""
        if self.replay is None:
            self.setup_replay(0)
        if self.learner is None:
            self.setup_learner()
        if self.agent is None:
            self.setup_agent(0)
        self.learner.main_setup()
        self.agent.main_setup(env)
        return self.agent, self.replay, self.learner

    def run_engine(self, iterations=None, rollout_steps=None, pipelined=None):
        """Run actors + learner.  pipelined (default when one rollout chunk yields exactly one learner batch):
        PipelinedEngine -- actors and learner on two concurrent streams.  Otherwise alternate a rollout chunk with
        as many learner iterations as the replay can feed."""
        agent, replay, learner = self.setup_engine()
        T = rollout_steps or self.learner_config.algo.stride
        n, stride = self.learner_config.algo.n_step, self.learner_config.algo.stride
        one_batch_per_chunk = (n == stride == T and agent.num_envs == self.learner_config.replay.batch_size and
                               hasattr(learner, 'replay_out_buffers'))
        if pipelined is None:
            pipelined = one_batch_per_chunk
        if pipelined:
            eng = PipelinedEngine(agent, replay, learner, T)
            eng.prime()
            while iterations is None or eng.iterations < iterations:
                eng.step()
            eng.drain()
            return learner
        it = 0
        while iterations is None or it < iterations:
            agent.main_loop(max_steps=T)
            while replay.start_sample_condition():
                learner.main_loop()
                it += 1
                if iterations is not None and it >= iterations:
                    break
        return learner

    def launch(self, component_name_in):
        name = component_name_in.split('-')[0] if '-' in component_name_in else component_name_in
        if name in ('learner', 'engine', 'all'):
            return self.run_engine()
        if name in ('agent', 'agents', 'eval', 'evals', 'replay', 'replay_worker') or name in _NOOP_COMPONENTS:
            self.log.warning('component "%s" is collapsed into the learner process of surreal_b200 '
                             '(one process per GPU); launch `learner` instead', component_name_in)
            return None
        raise ValueError('Unexpected component {}'.format(component_name_in))

    def run_learner(self, iterations=None):
        return self.run_engine(iterations)

    def run_agent(self, agent_id):
        return self.launch('agent-%s' % agent_id)

    def run_replay_worker(self, replay_id):
        return self.setup_replay(replay_id)
