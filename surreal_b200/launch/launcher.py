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

    def run_engine(self, iterations=None, rollout_steps=None):
        """Alternate a rollout chunk of all co-located actors with as many learner iterations as the replay
        can feed (Surreal's actors and learner run concurrently in separate processes; on one GPU the two
        phases share the device and are time-sliced)."""
        agent, replay, learner = self.setup_engine()
        T = rollout_steps or self.learner_config.algo.stride
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
