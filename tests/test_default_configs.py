"""The default config trees must equal the reference's (captured from the real reference in
tests/golden/configs.npz by make_golden.py:gen_configs)."""
import json


def _js(c):
    return json.loads(json.dumps(c.to_dict() if hasattr(c, 'to_dict') else c, default=str))


def test_default_trees_equal_reference(golden):
    from surreal_b200.session import (BASE_LEARNER_CONFIG, BASE_ENV_CONFIG, BASE_SESSION_CONFIG,
                                      LOCAL_SESSION_CONFIG)
    from surreal_b200.main.ppo_configs import (PPO_DEFAULT_LEARNER_CONFIG, PPO_DEFAULT_ENV_CONFIG,
                                               PPO_DEFAULT_SESSION_CONFIG)
    from surreal_b200.main.ddpg_configs import (DDPG_DEFAULT_LEARNER_CONFIG, DDPG_DEFAULT_ENV_CONFIG,
                                                DDPG_DEFAULT_SESSION_CONFIG)
    g = golden('configs')
    mine = dict(ppo_learner=PPO_DEFAULT_LEARNER_CONFIG, ppo_env=PPO_DEFAULT_ENV_CONFIG,
                ppo_session=PPO_DEFAULT_SESSION_CONFIG, ddpg_learner=DDPG_DEFAULT_LEARNER_CONFIG,
                ddpg_env=DDPG_DEFAULT_ENV_CONFIG, ddpg_session=DDPG_DEFAULT_SESSION_CONFIG,
                base_learner=BASE_LEARNER_CONFIG, base_env=BASE_ENV_CONFIG, base_session=BASE_SESSION_CONFIG,
                local_session=dict(LOCAL_SESSION_CONFIG))
    for k, v in mine.items():
        assert _js(v) == g.js(k), k
