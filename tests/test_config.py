"""Config-tree behaviour, modelled on the reference's own config tests (test-old/test_config.py) and
cross-checked key-by-key against the reference defaults where those were captured in goldens."""
import pytest

from surreal_b200.session.config import Config, ConfigError, extend_config


def base():
    return Config({
        'redis': {'replay': {'host': 'localhost', 'port': 6379},
                  'ps': {'host': '_dict_', 'port': '_list_', 'single': '_singleton_'}},
        'log': {'files': ['f1.txt', 'f2.txt'], 'outputs': [{'stdout1': 1}, {'stderr1': 10, 'stderr2': 20}]},
    })


def test_attribute_access_and_nested_lists():
    C = base()
    assert C.redis.replay.host == 'localhost' and C.log.files[1] == 'f2.txt' and C.log.outputs[1].stderr2 == 20
    with pytest.raises(ConfigError):
        C.redis.ps.badkey


def test_extend_fills_defaults_and_checks_required():
    filled = {'redis': {'ps': {'host': {'s': 2}, 'port': [1, 2], 'single': 'one-value'}}}
    C = extend_config(dict(filled), base())
    assert C.redis.replay.port == 6379 and C.redis.ps.host.s == 2
    C2 = Config(filled)
    C2.extend(base())
    assert C2 == C2.to_dict() == C.to_dict()


@pytest.mark.parametrize('bad', [
    {'redis': {'ps': {'host': 3, 'port': [1, 2], 'single': 1}}},                # not a dict
    {'redis': {'ps': {'host': {'s': 2}, 'port': {'t': 'x'}, 'single': 1}}},     # not a list
    {'redis': {'ps': {'host': {'s': 2}, 'port': [1, 2], 'single': {}}}},        # not a singleton
    {'redis': {'ps': {'host': {'s': 2}, 'port': [1, 2]}}},                      # required missing
    {'redis': {'replay': {'host': {'a': 1}}, 'ps': {'host': {}, 'port': [], 'single': 1}}},   # dict for singleton
    {'redis': {'replay': 3, 'ps': {'host': {}, 'port': [], 'single': 1}}},      # singleton for dict
    {},                                                                          # sub-dict holds a required key
])
def test_extend_errors(bad):
    with pytest.raises(ConfigError):
        extend_config(bad, base())


def test_numeric_placeholders_and_enum():
    D = {'a': '_int_', 'b': '_float_', 'c': '_num_', 'd': '_str_', 'e': '_bool_', 'f': '_object_',
         'g': '_enum[best,history]_'}
    ok = extend_config({'a': 1, 'b': 1.5, 'c': 2, 'd': 's', 'e': True, 'f': [1], 'g': 'best'}, D)
    assert ok.g == 'best'
    for k, v in [('a', 1.5), ('b', 1), ('c', 'x'), ('d', 3), ('e', 1), ('g', 'worst')]:
        good = {'a': 1, 'b': 1.5, 'c': 2, 'd': 's', 'e': True, 'f': None, 'g': 'best'}
        good[k] = v
        with pytest.raises(ConfigError):
            extend_config(good, D)
    # a placeholder may be inherited unchanged, but not changed
    assert extend_config({'a': '_int_', 'b': 1.0, 'c': 1, 'd': '', 'e': False, 'f': 0, 'g': 'best'}, D).a == '_int_'
    with pytest.raises(ConfigError):
        extend_config({'a': '_float_', 'b': 1.0, 'c': 1, 'd': '', 'e': False, 'f': 0, 'g': 'best'}, D)


def test_reserved_and_copy_and_dump(tmp_path):
    C = base()
    with pytest.raises(ConfigError):
        C.keys = 3
    D = C.copy()
    D.redis.replay.port = 1
    assert C.redis.replay.port == 6379
    p = str(tmp_path / 'c.yml')
    C.dump_file(p)
    assert Config.load_file(p).to_dict() == C.to_dict()
