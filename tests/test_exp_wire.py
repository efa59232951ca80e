"""experience frames from external CPU actors (exp_sender.py:10-98, exp_collector.py:22-65,
serializer.py:11-70).  CPU part: message structure, content-hash de-duplication, ZeroMQ PUSH -> PULL hand-off.  GPU
part: the frames end up as windows in the HBM FIFO."""
import time

import numpy as np
import pytest


def _window(rng, n, D, A, first_obs=None):
    obs = [{'low_dim': {'flat_inputs': (first_obs if (i == 0 and first_obs is not None) else rng.standard_normal(D)).astype(np.float32)}}
           for i in range(n)]
    nxt = {'low_dim': {'flat_inputs': rng.standard_normal(D).astype(np.float32)}}
    pinfo = [[rng.standard_normal(2 * A).astype(np.float32)] for _ in range(n)]
    return ({'obs': obs, 'obs_next': nxt},
            {'actions': [rng.standard_normal(A) for _ in range(n)], 'onetime_infos': [], 'persistent_infos': pinfo,
             'rewards': [float(r) for r in rng.standard_normal(n)], 'dones': [False] * (n - 1) + [True], 'infos': [{}] * n,
             'n_step': n})


def test_frame_structure_and_dedup():
    from surreal_b200 import exp_wire as W
    rng = np.random.default_rng(0)
    buf = W.ExpBuffer()
    h1, n1 = _window(rng, 4, 6, 2)
    h2, n2 = _window(rng, 4, 6, 2)
    h2['obs'][0] = h1['obs'][3]                     # overlapping windows share an observation (same content, same hash)
    buf.add(h1, n1)
    buf.add(h2, n2)
    assert len(buf.exp_list) == 2 and set(buf.exp_list[0]) >= {'obs_hash', 'obs_next_hash', 'actions', 'rewards'}
    assert len(buf.ob_storage) == 4 + 1 + 3 + 1     # the shared observation is stored once
    assert all(len(k) == 16 for k in buf.ob_storage)
    frame = buf.flush()
    assert buf.exp_list == [] and buf.ob_storage == {}
    exps = W.inflate(frame)
    assert len(exps) == 2 and 'obs' in exps[0] and 'obs_hash' not in exps[0]
    for got, (h, nn) in zip(exps, ((h1, n1), (h2, n2))):
        for a, b in zip(got['obs'], h['obs']):
            np.testing.assert_array_equal(a['low_dim']['flat_inputs'], b['low_dim']['flat_inputs'])
        np.testing.assert_array_equal(got['obs_next']['low_dim']['flat_inputs'], h['obs_next']['low_dim']['flat_inputs'])
        assert got['rewards'] == nn['rewards'] and got['dones'] == nn['dones'] and got['n_step'] == 4
    # 16-character base64(md5) content hash (serializer.py:55-67)
    import base64
    import hashlib
    assert W.binary_hash(b'abc') == base64.b64encode(hashlib.md5(b'abc').digest())[:16].decode()


def test_push_pull_collector_delivers_in_order():
    from surreal_b200 import exp_wire as W
    got = []
    port = 17000 + (int(time.time() * 1000) % 2000)
    srv = W.ExperienceCollectorServer('127.0.0.1', port, got.append)
    srv.start()
    time.sleep(0.2)
    snd = W.ExpSender(host='127.0.0.1', port=port, flush_iteration=3)
    rng = np.random.default_rng(1)
    sent = []
    for k in range(9):
        h, nn = _window(rng, 3, 5, 2)
        nn['tag'] = k
        sent.append(h)
        snd.send(h, nn)
    t0 = time.time()
    while len(got) < 9 and time.time() - t0 < 10:
        time.sleep(0.05)
    srv.stop()
    snd.close()
    assert srv.error is None and srv.frames == 3 and [e['tag'] for e in got] == list(range(9))
    np.testing.assert_array_equal(got[4]['obs'][1]['low_dim']['flat_inputs'], sent[4]['obs'][1]['low_dim']['flat_inputs'])


@pytest.mark.gpu
def test_external_actor_frames_land_in_the_hbm_fifo():
    import torch
    from helpers import ppo_configs
    from surreal_b200 import exp_wire as W
    from surreal_b200.replay import FIFOReplay
    n, D, A = 4, 6, 2
    lc, ec, sc = ppo_configs(D=D, A=A, n_step=n, stride=n, B=4, memory_size=32)
    R = FIFOReplay(lc, ec, sc)
    port = 19000 + (int(time.time() * 1000) % 2000)
    srv = W.ExperienceCollectorServer('127.0.0.1', port, R.insert)
    srv.start()
    time.sleep(0.2)
    snd = W.ExpSender(host='127.0.0.1', port=port, flush_iteration=2)
    rng = np.random.default_rng(2)
    wins = [_window(rng, n, D, A) for _ in range(4)]
    for h, nn in wins:
        snd.send(h, nn)
    t0 = time.time()
    while srv.experiences < 4 and time.time() - t0 < 20:
        time.sleep(0.05)
    srv.stop()
    snd.close()
    assert srv.error is None and len(R) == 4
    b = R.sample(4)
    torch.cuda.synchronize()
    for k, (h, nn) in enumerate(wins):
        exp_obs = np.stack([o['low_dim']['flat_inputs'] for o in h['obs']])
        np.testing.assert_array_equal(b['obs']['low_dim']['flat_inputs'][k].cpu().numpy(), exp_obs)
        np.testing.assert_array_equal(b['actions'][k].cpu().numpy(), np.stack(nn['actions']).astype(np.float32))
        np.testing.assert_array_equal(b['persistent_infos'][0][k].cpu().numpy(), np.stack([p[-1] for p in nn['persistent_infos']]))
        np.testing.assert_array_equal(b['dones'][k].cpu().numpy(), np.array(nn['dones'], dtype=np.float32))
