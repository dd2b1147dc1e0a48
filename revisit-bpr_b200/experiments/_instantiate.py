"""The subset of `hydra.utils.instantiate` the reference's configs rely on (hydra is preferred when
importable): dicts with `_target_` (dotted path to a callable), optional `_partial_: true`, nested
dicts / lists instantiated recursively, keyword overrides from the call site."""
from __future__ import annotations

import functools
import importlib
from typing import Any


def _locate(path: str) -> Any:
    parts = path.split(".")
    for cut in range(len(parts), 0, -1):
        try:
            obj = importlib.import_module(".".join(parts[:cut]))
        except ImportError:
            continue
        for attr in parts[cut:]:
            obj = getattr(obj, attr)
        return obj
    raise ImportError(f"cannot locate {path!r}")


def instantiate(config: Any, *args: Any, **overrides: Any) -> Any:
    if isinstance(config, list):
        return [instantiate(c) for c in config]
    if not isinstance(config, dict):
        return config
    if "_target_" not in config:
        return {k: instantiate(v) for k, v in config.items()}
    kwargs = {k: instantiate(v) for k, v in config.items() if k not in ("_target_", "_partial_", "_args_")}
    kwargs.update(overrides)
    target = _locate(config["_target_"])
    pos = [instantiate(a) for a in config.get("_args_", [])] + list(args)
    if config.get("_partial_", False):
        return functools.partial(target, *pos, **kwargs)
    return target(*pos, **kwargs)
