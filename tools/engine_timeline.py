#!/usr/bin/env python
"""Where does a pipelined step go?  Replays PipelinedEngine.step() at the bench shape with CUDA events on both streams
(actors: stream A, learner: stream L) and prints, per step, when the rollout and the learn() graph start and end relative
to the start of the step.  Events cannot sit inside the learn() graph, so the learner shows as fetch (pop + gather) | learn."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    from surreal_b200.launch import SurrealDefaultLauncher, PipelinedEngine
    from surreal_b200.agent import PPOAgent
    from surreal_b200.learner import PPOLearner
    from surreal_b200.replay import FIFOReplay
    lc, ec, sc = bench.build_configs(bench.N_ACTORS, bench.HORIZON)
    la = SurrealDefaultLauncher(PPOAgent, PPOLearner, FIFOReplay, sc, ec, lc)
    agent, replay, learner = la.setup_engine()
    T = bench.HORIZON
    eng = PipelinedEngine(agent, replay, learner, T)
    eng.prime()
    for _ in range(5):
        eng.step()
    torch.cuda.synchronize()
    E = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    rows = []
    L = learner
    for _ in range(6):
        ev = {k: E() for k in ('l0', 'l1', 'l2', 'l3', 'a0', 'a1')}
        with torch.cuda.stream(eng.sL):
            eng.sL.wait_event(eng.ev_roll)
            ev['l0'].record(eng.sL)
            data = L.fetch_batch()
            ev['l1'].record(eng.sL)
            eng.ev_pop = torch.cuda.Event()
            eng.ev_pop.record(eng.sL)
        with torch.cuda.stream(eng.sA):
            eng.sA.wait_event(eng.ev_pop)
            if eng.ev_pub is not None:
                eng.sA.wait_event(eng.ev_pub)
            ev['a0'].record(eng.sA)
            agent.main_loop(max_steps=T)
            ev['a1'].record(eng.sA)
            eng.ev_fetch = torch.cuda.Event()
            eng.ev_fetch.record(eng.sA)
            eng.ev_roll = torch.cuda.Event()
            eng.ev_roll.record(eng.sA)
        with torch.cuda.stream(eng.sL):
            L.learn(data)
            ev['l2'].record(eng.sL)
            eng.sL.wait_event(eng.ev_fetch)
            if L.should_publish_parameter():
                L.publish_parameter(L.current_iter, message='batch ' + str(L.current_iter))
            ev['l3'].record(eng.sL)
            eng.ev_pub = torch.cuda.Event()
            eng.ev_pub.record(eng.sL)
            L.current_iter += 1
        rows.append(ev)
    torch.cuda.synchronize()
    print('epoch kernel generation:', L.epoch_kernel_gen if L.use_epoch_kernel else 0)
    print('per step, ms after the step began (l0):  gather done | rollout start .. end | learn done | publish done || next step begins')
    for i, ev in enumerate(rows):
        t = lambda k: ev['l0'].elapsed_time(ev[k])  # noqa: E731
        nxt = ev['l0'].elapsed_time(rows[i + 1]['l0']) if i + 1 < len(rows) else float('nan')
        print('  step %d: %.3f | %.3f .. %.3f | %.3f | %.3f || %.3f' % (i, t('l1'), t('a0'), t('a1'), t('l2'), t('l3'), nxt))


if __name__ == '__main__':
    main()
