"""Stand-in for gymnasium.envs.registration (register / make) — test infrastructure only."""
import importlib

registry = {}


def register(id, entry_point=None, **kwargs):
    registry[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, reg_kwargs = registry[id]
    if isinstance(entry_point, str):
        mod, _, attr = entry_point.partition(":")
        entry_point = getattr(importlib.import_module(mod), attr)
    reg_kwargs = {k: v for k, v in reg_kwargs.items() if k not in ("max_episode_steps",)}
    return entry_point(**{**reg_kwargs.get("kwargs", {}), **kwargs})
