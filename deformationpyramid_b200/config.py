"""Flat YAML config -> attribute dictionary, as eval_nolearned.py:17-40 builds it (easydict is not a
dependency: AttrDict covers the attribute access the driver uses), including the custom `!join`
tag (eval_nolearned.py:17-20)."""
from __future__ import annotations

import yaml


class AttrDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class _Loader(yaml.Loader):
    pass


def _join(loader, node):
    return "_".join(str(i) for i in loader.construct_sequence(node))


_Loader.add_constructor("!join", _join)

# config/NDP.yaml of the reference, verbatim values (the file itself is not copied)
NDP_DEFAULTS = dict(gpu_mode=True, deformation_model="NDP", use_ldmk=False, use_depth=False, iters=500, lr=0.01,
                    max_break_count=15, break_threshold_ratio=0.001, w_reg=0.0, samples=2000, m=9, k0=-8, depth=3,
                    width=128, act_fn="relu", motion_type="SE3", rotation_format="axis_angle")


def load_config(path: str) -> AttrDict:
    with open(path, "r") as f:
        return AttrDict(yaml.load(f, Loader=_Loader))


def ndp_config(**overrides) -> AttrDict:
    cfg = AttrDict(NDP_DEFAULTS)
    for k, v in overrides.items():
        cfg[k] = v
    return cfg
