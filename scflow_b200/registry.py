"""mmcv-style Registry / build_from_cfg / Config, enough for the reference's configs to drive this package.

Mirrors the interface the reference uses from ``mmcv.utils`` (models/*/builder.py:1-6,
train.py:95 ``Config.fromfile``): classes are registered by NAME and built from dicts whose ``type`` key selects
the class and whose other keys are constructor kwargs.  mmcv itself is not installable offline.
"""
import copy
import importlib.util
import os
from typing import Any, Dict, Optional


class Registry:
    def __init__(self, name: str):
        self._name = name
        self._module_dict: Dict[str, type] = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return key in self._module_dict

    def __repr__(self):
        return f'Registry(name={self._name}, items={sorted(self._module_dict)})'

    def get(self, key: str) -> Optional[type]:
        return self._module_dict.get(key)

    def _register(self, cls, name=None, force=False):
        if not isinstance(cls, type):
            raise TypeError(f'module must be a class, but got {type(cls)}')
        names = [name or cls.__name__] if not isinstance(name, (list, tuple)) else list(name)
        for n in names:
            if not force and n in self._module_dict:
                raise KeyError(f'{n} is already registered in {self._name}')
            self._module_dict[n] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def _decorator(cls):
            self._register(cls, name, force)
            return cls
        return _decorator


def build_from_cfg(cfg: Dict[str, Any], registry: Registry, default_args: Optional[Dict[str, Any]] = None):
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
    if 'type' not in cfg and not (default_args and 'type' in default_args):
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}\n{default_args}')
    if not isinstance(registry, Registry):
        raise TypeError(f'registry must be a Registry object, but got {type(registry)}')
    args = copy.deepcopy(dict(cfg))
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f'{obj_type} is not in the {registry.name} registry')
    elif isinstance(obj_type, type):
        obj_cls = obj_type
    else:
        raise TypeError(f'type must be a str or valid type, but got {type(obj_type)}')
    try:
        return obj_cls(**args)
    except Exception as e:
        raise type(e)(f'{obj_cls.__name__}: {e}')


class ConfigDict(dict):
    """dict with attribute access (mmcv.utils.ConfigDict)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def _to_cfgdict(v):
    if isinstance(v, dict):
        return ConfigDict({k: _to_cfgdict(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return type(v)(_to_cfgdict(x) for x in v)
    return v


def _merge(base: dict, child: dict) -> dict:
    out = dict(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get('_delete_', False):
            out[k] = _merge(out[k], v)
        else:
            out[k] = {kk: vv for kk, vv in v.items() if kk != '_delete_'} if isinstance(v, dict) else v
    return out


class Config:
    """Python-file configs with ``_base_`` inheritance (configs/refine_models/scflow.py:1)."""

    def __init__(self, cfg_dict: Optional[dict] = None, filename: Optional[str] = None):
        object.__setattr__(self, '_cfg_dict', _to_cfgdict(cfg_dict or {}))
        object.__setattr__(self, 'filename', filename)

    @staticmethod
    def _file2dict(filename: str) -> dict:
        filename = os.path.abspath(filename)
        if not os.path.isfile(filename):
            raise FileNotFoundError(filename)
        spec = importlib.util.spec_from_file_location('_scflow_cfg', filename)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        cfg = {k: v for k, v in vars(mod).items() if not k.startswith('__') and not callable(v) and not isinstance(v, type(os))}
        base = cfg.pop('_base_', None)
        if base is not None:
            merged: dict = {}
            for b in ([base] if isinstance(base, str) else base):
                merged = _merge(merged, Config._file2dict(os.path.join(os.path.dirname(filename), b)))
            cfg = _merge(merged, cfg)
        return cfg

    @staticmethod
    def fromfile(filename: str) -> 'Config':
        return Config(Config._file2dict(filename), filename)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _to_cfgdict(value)

    def __contains__(self, name):
        return name in self._cfg_dict

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)
