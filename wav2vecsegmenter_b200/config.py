"""Config composition for segment.py / inference.py when Hydra is not installed (the GPU box has
neither hydra nor omegaconf). It implements the subset of Hydra / OmegaConf behaviour the reference's
`conf/` tree and command lines rely on, so that `conf/` ships byte-identical to the reference's and
the reference's own command lines (README.md:73-79,105-160) compose unchanged:

* a `defaults` list with `_self_` and config groups (`- algorithm: dac`), group selection on the
  command line (`algorithm=strm`), dotted value overrides (`algorithm.threshold=0.3`), `+new.key=v`;
* lazy interpolation at access time: `${key.sub}`, relative `${.sibling}` / `${..parent_key}`,
  the resolvers `${hydra:runtime.cwd}` and `${oc.env:VAR[,default]}`, and
  `${hydra.job.override_dirname}` (sorted `key=value` items minus `exclude_keys`);
* mandatory values `???` that raise on ACCESS, not at load (OmegaConf's MissingMandatoryValue);
* `merge(prev, cfg)` = OmegaConf.merge (reference segment.py:161-163, inference.py:158-160);
* the job's run directory `hydra.run.dir` with `.hydra/{config,overrides}.yaml` written into it.

When Hydra IS importable, segment.py / inference.py use `@hydra.main` exactly like the reference and
this module is only used for `instantiate`.
"""
from __future__ import annotations

import copy
import os
import re
from pathlib import Path

import yaml


class MissingMandatoryValue(KeyError):
    """a `???` value was read (omegaconf.errors.MissingMandatoryValue)"""


class InterpolationError(KeyError):
    pass


_INTERP = re.compile(r"\$\{([^${}]+)\}")


class Cfg:
    """Read-mostly view of one dict node of the composed config (enough of DictConfig for the two
    CLIs): attribute / item access, `.get`, `in`, iteration, `dict(node)`, lazy interpolation."""

    __slots__ = ("_data", "_root", "_path")

    def __init__(self, data: dict, root: "Cfg | None" = None, path: tuple = ()):
        object.__setattr__(self, "_data", data)
        object.__setattr__(self, "_root", root if root is not None else self)
        object.__setattr__(self, "_path", path)

    # ---- resolution ------------------------------------------------------------------------
    def _lookup(self, dotted: str, at: tuple):
        """value of an interpolation key, absolute or relative to the node at path `at`"""
        if dotted.startswith("."):
            up = len(dotted) - len(dotted.lstrip("."))
            base = at[: len(at) - (up - 1)] if up > 1 else at
            parts = base + tuple(p for p in dotted.lstrip(".").split(".") if p)
        else:
            parts = tuple(dotted.split("."))
        node = self._root
        for i, p in enumerate(parts):
            if not isinstance(node, Cfg) or p not in node._data:
                raise InterpolationError(f"interpolation key '{dotted}' not found")
            node = node._child(p)
        return node

    def _resolver(self, name: str, arg: str):
        if name == "hydra":
            if arg == "runtime.cwd":
                return self._root._data.get("hydra", {}).get("runtime", {}).get("cwd", os.getcwd())
            return self._lookup("hydra." + arg, ())
        if name == "oc.env":
            var, _, default = arg.partition(",")
            if var.strip() in os.environ:
                return os.environ[var.strip()]
            if _:
                return default.strip()
            raise InterpolationError(f"environment variable '{var}' not set")
        raise InterpolationError(f"unsupported resolver '{name}'")

    def _resolve(self, value, at: tuple):
        if not isinstance(value, str):
            return value
        if value == "???":
            raise MissingMandatoryValue("Missing mandatory value: " + ".".join(at))
        if "${" not in value:
            return value

        def one(m):
            inner = m.group(1).strip()
            if ":" in inner:
                name, _, arg = inner.partition(":")
                return self._resolver(name, arg)
            return self._lookup(inner, at[:-1])

        whole = _INTERP.fullmatch(value)
        if whole:                      # the value IS one interpolation: keep the referenced type
            out = one(whole)
            return out
        prev = None
        while prev != value and "${" in value:   # nested / several interpolations in one string
            prev = value
            value = _INTERP.sub(lambda m: str(one(m)), value)
        return value

    def _child(self, key):
        v = self._data[key]
        path = self._path + (key,)
        if isinstance(v, dict):
            return Cfg(v, self._root, path)
        if isinstance(v, list):
            return [Cfg(x, self._root, path) if isinstance(x, dict) else self._resolve(x, path) for x in v]
        return self._resolve(v, path)

    # ---- mapping protocol ------------------------------------------------------------------
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        try:
            return self[k]
        except MissingMandatoryValue:
            raise
        except KeyError as e:
            raise AttributeError(k) from e

    def __getitem__(self, k):
        if k not in self._data:
            raise KeyError(k)
        return self._child(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __setitem__(self, k, v):
        self._data[k] = v._data if isinstance(v, Cfg) else v

    def __contains__(self, k):
        return k in self._data

    def __iter__(self):
        return iter(self._data)

    def __len__(self):
        return len(self._data)

    def keys(self):
        return self._data.keys()

    def items(self):
        return [(k, self[k]) for k in self._data]

    def get(self, k, default=None):
        """like DictConfig.get: default for absent keys and for None values; `???` still raises"""
        if k not in self._data:
            return default
        v = self[k]
        return default if v is None else v

    def pop(self, k, *default):
        v = self[k] if k in self._data else (default[0] if default else None)
        if k not in self._data and not default:
            raise KeyError(k)
        self._data.pop(k, None)
        return v

    def to_container(self, resolve: bool = True):
        """plain dict (OmegaConf.to_container / to_object)"""
        if not resolve:
            return copy.deepcopy(self._data)

        def conv(x):
            if isinstance(x, Cfg):
                return {k: conv(x[k]) for k in x}
            if isinstance(x, list):
                return [conv(v) for v in x]
            return x

        return conv(self)

    def __repr__(self):
        return f"Cfg({self._data!r})"


def to_yaml(cfg: Cfg) -> str:
    """unresolved dump, like OmegaConf.to_yaml(config)"""
    return yaml.safe_dump(cfg.to_container(resolve=False), sort_keys=False)


def merge_dicts(base: dict, over: dict) -> dict:
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = merge_dicts(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


def merge(prev, cfg: Cfg) -> Cfg:
    """OmegaConf.merge(prev_cfg, config): `cfg` wins, dict nodes merge recursively"""
    pd = prev._data if isinstance(prev, Cfg) else prev
    return Cfg(merge_dicts(pd, cfg._data))


def load(path) -> Cfg:
    """OmegaConf.load"""
    return Cfg(yaml.safe_load(Path(path).read_text()) or {})


def _set_dotted(d: dict, key: str, value):
    parts = key.split(".")
    for p in parts[:-1]:
        nxt = d.get(p)
        if not isinstance(nxt, dict):
            nxt = {}
            d[p] = nxt
        d = nxt
    d[parts[-1]] = value


def _parse_value(text: str):
    if text == "":
        return None
    try:
        return yaml.safe_load(text)
    except yaml.YAMLError:
        return text


def override_dirname(overrides: list[str], exclude_keys) -> str:
    """hydra.job.override_dirname: the override lines, minus excluded keys, sorted, comma-joined
    (hydra.core.override_parser.types.Override / hydra._internal.utils: get_overrides_dirname)"""
    keep = []
    for ov in overrides:
        key = ov.partition("=")[0].lstrip("+~")
        if key not in set(exclude_keys or []):
            keep.append(ov)
    return ",".join(sorted(keep))


def compose(conf_dir, config_name: str, overrides: list[str]) -> Cfg:
    """what `@hydra.main(config_path=conf_dir, config_name=config_name)` hands to the task function
    for the given command-line overrides (single run; sweeps `-m` are not supported)"""
    conf_dir = Path(conf_dir)
    if any(o in ("-m", "--multirun") for o in overrides):
        raise SystemExit("multirun (-m) needs Hydra; this fallback composer runs single jobs only")
    root = yaml.safe_load((conf_dir / f"{config_name}.yaml").read_text()) or {}
    defaults = root.pop("defaults", [])
    groups: dict = {}
    order = []          # merge order of the defaults list
    for d in defaults:
        if d == "_self_":
            order.append("_self_")
        elif isinstance(d, dict):
            for g, name in d.items():
                groups[g] = name
                order.append(g)
    if "_self_" not in order:
        order.append("_self_")     # Hydra >= 1.1 default: the primary config is merged last
    plain = []
    for ov in overrides:
        if "=" not in ov:
            raise SystemExit(f"cannot parse override '{ov}' (expected key=value)")
        k, _, v = ov.partition("=")
        k = k.lstrip("+")
        if k in groups or ("." not in k and (conf_dir / k).is_dir()):
            if k not in groups:
                order.append(k)
            groups[k] = v
        else:
            plain.append((k, _parse_value(v)))
    cfg: dict = {}
    for item in order:
        if item == "_self_":
            cfg = merge_dicts(cfg, root)
        elif groups.get(item) is not None:
            f = conf_dir / item / f"{groups[item]}.yaml"
            if not f.exists():
                options = sorted(p.stem for p in (conf_dir / item).glob("*.yaml"))
                raise SystemExit(f"Could not find '{item}/{groups[item]}'. Available options in '{item}': {options}")
            cfg = merge_dicts(cfg, {item: yaml.safe_load(f.read_text()) or {}})
    for k, v in plain:
        _set_dotted(cfg, k, v)
    hydra_node = cfg.setdefault("hydra", {})
    hydra_node.setdefault("runtime", {})["cwd"] = os.getcwd()
    job = hydra_node.setdefault("job", {})
    job["name"] = config_name
    excl = (((job.get("config") or {}).get("override_dirname") or {}).get("exclude_keys")) or []
    job["override_dirname"] = override_dirname(overrides, excl)
    hydra_node["overrides"] = {"task": list(overrides)}
    return Cfg(cfg)


def run_dir(cfg: Cfg) -> Path:
    """the job's working directory (`hydra.run.dir`); created, with `.hydra/{config,overrides}.yaml`
    like Hydra writes them. The reference's scripts write their outputs relative to it."""
    d = Path(str(cfg.hydra.run.dir))
    (d / ".hydra").mkdir(parents=True, exist_ok=True)
    user = {k: v for k, v in cfg.to_container(resolve=False).items() if k != "hydra"}
    (d / ".hydra" / "config.yaml").write_text(yaml.safe_dump(user, sort_keys=False))
    (d / ".hydra" / "overrides.yaml").write_text(yaml.safe_dump(list(cfg.hydra.overrides.task)))
    return d


def to_object(node):
    """OmegaConf.to_object for Cfg nodes, DictConfig nodes and plain dicts alike"""
    if isinstance(node, Cfg):
        return node.to_container(resolve=True)
    try:
        from omegaconf import OmegaConf  # type: ignore

        if OmegaConf.is_config(node):
            return OmegaConf.to_object(node)
    except ImportError:
        pass
    return dict(node)


def instantiate(node, **extra):
    """`_target_`-style construction (hydra.utils.instantiate for the one case generate() needs)"""
    import importlib

    node = to_object(node)
    kw = {k: v for k, v in node.items() if k != "_target_"}
    kw.update(extra)
    mod, _, name = node["_target_"].rpartition(".")
    return getattr(importlib.import_module(mod), name)(**kw)
