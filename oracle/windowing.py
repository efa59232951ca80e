"""Experience windowing on the actor side (reference: surreal/env/exp_sender_wrapper.py).

Both models consume a flat stream of ``(step_id, reward, done)`` per episode and emit what the
reference's wrappers would ``send`` -- expressed over step ids so they can be compared with the
HBM staging kernels index-for-index."""
from collections import deque


def multistep_windows(ep_lens, n_step, stride, reward_fn=None):
    """exp_sender_wrapper.py:153-264.  Returns a list of (obs_step_ids[n], obs_next_id, done_flags[n]).
    The deque is cleared on reset (:204-207), so windows never span episodes and the tail of an
    episode that does not fill a window is discarded; after a send, ``stride`` items are popped (:222-226)."""
    out = []
    g = 0
    for L in ep_lens:
        last = deque()
        for t in range(L):
            done = (t + 1 >= L)
            last.append((g, done))
            g += 1
            if len(last) == n_step:
                out.append(([s for s, _ in last], g, [d for _, d in last]))
                for _ in range(stride):
                    if last:
                        last.popleft()
    return out


def ssar_nstep(ep_lens, n_step, gamma, reward_fn):
    """exp_sender_wrapper.py:72-112.  Emits (obs_id, obs_next_id, action_id, reward, done).
    Reward accumulation uses exponent ``n_step - i - 1`` for deque position i (:105) -- the true age
    only once the deque is full; reproduced as written."""
    out = []
    g = 0
    for L in ep_lens:
        last = deque()
        for t in range(L):
            obs_id = g
            g += 1
            reward = reward_fn(g)
            done = (t + 1 >= L)
            for i, e in enumerate(last):
                e[1] = g
                e[3] += pow(gamma, n_step - i - 1) * reward
                e[4] = done
            last.append([obs_id, g, obs_id, reward, done])
            if len(last) == n_step:
                e = last.popleft()
                out.append(tuple(e))
    return out
