"""Attribute-dict config trees with typed required placeholders.

Behavioural mirror of the reference's config surface (surreal/session/config.py:154-255): every
plugin constructor receives three such trees (learner_config, env_config, session_config), so a
drop-in must accept and produce the same objects:

  * ``Config(d)``: nested dict with attribute access; nested dicts (also inside lists/tuples) become
    Config; unknown attribute -> ConfigError.
  * ``cfg.extend(defaults)`` / ``extend_config(cfg, defaults)``: fill missing keys from ``defaults``;
    placeholders ``'_int_' '_float_' '_num_' '_str_' '_bool_' '_list_' '_dict_' '_singleton_'
    '_object_' '_enum[a,b]_'`` mark REQUIRED keys and are type-checked (config.py:24-53,100-147).

Written from the documented behaviour and the reference's own tests (test-old/test_config.py),
not from its source.
"""
import json
import re

import yaml


class ConfigError(Exception):
    pass


_ENUM = re.compile(r'_enum\[(.*)\]_')
_RESERVED = ('keys', 'items', 'values', 'get', 'copy', 'update', 'extend', 'load_file', 'dump_file', 'to_dict')

_SIMPLE = {
    '_object_': (lambda v: True, 'filled'),
    '_singleton_': (lambda v: not isinstance(v, (list, dict)), 'a singleton (non-list/dict)'),
    '_list_': (lambda v: isinstance(v, list), 'a list'),
    '_dict_': (lambda v: isinstance(v, dict), 'a dict'),
    '_int_': (lambda v: isinstance(v, int), 'an integer'),
    '_float_': (lambda v: isinstance(v, float), 'a float'),
    '_num_': (lambda v: isinstance(v, (int, float)), 'a numeric value'),
    '_str_': (lambda v: isinstance(v, str), 'a string'),
    '_bool_': (lambda v: isinstance(v, bool), 'a boolean'),
}


def _placeholder(value):
    """(checker, description) when ``value`` is a required-placeholder string, else None."""
    if not isinstance(value, str):
        return None
    v = value.lower()
    if v in _SIMPLE:
        return _SIMPLE[v]
    m = _ENUM.match(v)
    if m:
        if not m.group(1):
            raise ConfigError('_enum[...]_ cannot be empty')
        options = [o.strip() for o in m.group(1).split(',')]
        return (lambda x: x in options), 'an enum in [%s]' % m.group(1)
    return None


def _contains_placeholder(d):
    return any(_placeholder(v) is not None or (isinstance(v, dict) and _contains_placeholder(v)) for v in d.values())


def _path(trace, key):
    return 'key "%s" ' % '.'.join(trace + [key])


def _merge(config, default, trace):
    for key, dval in default.items():
        req = _placeholder(dval)
        if key not in config:
            if req is not None:
                raise ConfigError('Required entry missing: %smust be %s.' % (_path(trace, key), req[1]))
            if isinstance(dval, dict) and _contains_placeholder(dval):
                raise ConfigError('Sub-dict under %scontains a required config: %s.' % (_path(trace, key), dval))
            config[key] = dval
            continue
        val = config[key]
        if req is not None:
            if _placeholder(val) is not None:              # still a placeholder after extension
                if val != dval:
                    raise ConfigError('inherited %s: "%s" must match default "%s"' % (_path(trace, key), val, dval))
            elif not req[0](val):
                raise ConfigError('Wrong type: %smust be %s.' % (_path(trace, key), req[1]))
        elif isinstance(val, dict) and not isinstance(dval, dict):
            raise ConfigError(_path(trace, key) + 'must be a singleton instead of a sub-dict')
        elif isinstance(dval, dict):
            if not isinstance(val, dict):
                raise ConfigError(_path(trace, key) + 'must have a sub-dict instead of a singleton')
            config[key] = _merge(val, dval, trace + [key])
    return config


class Config(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = {} if d is None else d
        if kwargs:
            d = dict(d)
            d.update(kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if name in _RESERVED:
            raise ConfigError('"%s()" is a reserved method, cannot override' % name)
        if isinstance(value, (list, tuple)):
            value = [type(self)(x) if isinstance(x, dict) else x for x in value]
        elif isinstance(value, dict) and not isinstance(value, type(self)):
            value = type(self)(value)
        super().__setattr__(name, value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def __getattr__(self, key):
        try:
            return super().__getattribute__(key)
        except AttributeError:
            raise ConfigError('config key "%s" missing.' % key)

    def update(self, other):
        for k, v in other.items():
            setattr(self, k, v)

    def to_dict(self):
        out = {}
        for k, v in self.items():
            if isinstance(v, Config):
                out[k] = v.to_dict()
            elif isinstance(v, (list, tuple)):
                out[k] = type(v)(x.to_dict() if isinstance(x, Config) else x for x in v)
            else:
                out[k] = v
        return out

    def copy(self):
        return Config(self.to_dict())

    def extend(self, default_config):
        if not isinstance(default_config, dict):
            raise TypeError('default_config must be a dict')
        return _merge(self, default_config, [])

    @classmethod
    def load_file(cls, file_path):
        assert file_path.endswith(('.json', '.yaml', '.yml'))
        with open(file_path, 'r') as fp:
            return cls(json.load(fp) if file_path.endswith('.json') else yaml.safe_load(fp))

    def dump_file(self, file_path):
        assert file_path.endswith(('.json', '.yaml', '.yml'))
        with open(file_path, 'w') as fp:
            if file_path.endswith('.json'):
                json.dump(self, fp, indent=4)
            else:
                yaml.dump(self.to_dict(), stream=fp, indent=4, default_flow_style=False)


def extend_config(config, default_config):
    if not isinstance(config, dict) or not isinstance(default_config, dict):
        raise TypeError('extend_config expects two dicts')
    return Config(_merge(config, default_config, []))
