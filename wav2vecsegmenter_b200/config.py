"""Hydra-free config composition for segment.py / inference.py (the GPU box has no hydra /
omegaconf). Supports what the reference's configs use: a `defaults` list of config groups,
`key=value` / `group=name` command-line overrides with dotted keys, and merging the training
run's saved `.hydra/config.yaml` underneath (reference segment.py:161-163). Hydra stays the
primary interface when it is importable."""
from __future__ import annotations

from pathlib import Path

import yaml


class Cfg(dict):
    """dict with attribute access (enough of DictConfig for generate())"""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(x):
    if isinstance(x, dict):
        return Cfg({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def merge(base: dict, over: dict) -> dict:
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = merge(out[k], v)
        else:
            out[k] = v
    return out


def _set_dotted(d: dict, key: str, value):
    parts = key.split(".")
    for p in parts[:-1]:
        d = d.setdefault(p, {})
    d[parts[-1]] = value


def compose(conf_dir, config_name: str, overrides: list[str]) -> Cfg:
    conf_dir = Path(conf_dir)
    root = yaml.safe_load((conf_dir / f"{config_name}.yaml").read_text()) or {}
    defaults = root.pop("defaults", [])
    root.pop("hydra", None)
    groups = {}
    for d in defaults:
        if isinstance(d, dict):
            groups.update(d)
    plain = []
    for ov in overrides:
        k, _, v = ov.partition("=")
        k = k.lstrip("+")
        if k in groups or (conf_dir / k).is_dir():
            groups[k] = v
        else:
            plain.append((k, yaml.safe_load(v) if v != "" else None))
    cfg = dict(root)
    for g, name in groups.items():
        f = conf_dir / g / f"{name}.yaml"
        if f.exists():
            cfg[g] = merge(cfg.get(g, {}) or {}, yaml.safe_load(f.read_text()) or {})
    for k, v in plain:
        _set_dotted(cfg, k, v)
    if cfg.get("config_path") not in (None, "???"):
        prev = yaml.safe_load(Path(cfg["config_path"]).read_text()) or {}
        prev.pop("hydra", None)
        cfg = merge(prev, cfg)
    missing = [k for k, v in cfg.items() if v == "???"]
    if missing:
        raise SystemExit(f"missing mandatory config values: {', '.join(missing)} (pass key=value)")
    return _wrap(cfg)


def instantiate(node: dict, **extra):
    """`_target_`-style construction (hydra.utils.instantiate for the one case generate() needs)"""
    import importlib

    kw = {k: v for k, v in node.items() if k != "_target_"}
    kw.update(extra)
    mod, _, name = node["_target_"].rpartition(".")
    return getattr(importlib.import_module(mod), name)(**kw)
