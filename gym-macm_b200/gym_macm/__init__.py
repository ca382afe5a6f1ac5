"""gym_macm -- B200-native batched replacement for the hot path of siyarvurucu/gym-macm.

    env = gym_macm.make("gym_macm:cm-flock-v0", n_agents=[4])          # dict API, one world
    envs = gym_macm.BatchedFlock(4096, n_agents=[64], reward_mode="linear")   # tensor API

The ids of gym_macm/__init__.py:3-16 are registered with gym / gymnasium when one of them is
installed; `make` works either way.
"""
from gym_macm import _lib, settings  # noqa: F401
from gym_macm.batched import BatchedFlock, BatchedTDM, BatchPool  # noqa: F401

_ENTRY = {"cm-flock-v0": "gym_macm.envs:Flock", "cm-tdm-v0": "gym_macm.envs:TDM"}


def make(env_id, **kwargs):
    """gym.make for this package: accepts "gym_macm:cm-flock-v0" or "cm-flock-v0"."""
    name = env_id.split(":")[-1]
    if name not in _ENTRY:
        raise KeyError("unknown gym_macm environment id %r (known: %s)" % (env_id, ", ".join(sorted(_ENTRY))))
    mod, cls = _ENTRY[name].split(":")
    import importlib
    return getattr(importlib.import_module(mod), cls)(**kwargs)


def _register():  # pragma: no cover - gym is absent in the build container
    for pkg in ("gym", "gymnasium"):
        try:
            reg = __import__(pkg + ".envs.registration", fromlist=["register"])
        except Exception:
            continue
        for name, entry in _ENTRY.items():
            for extra in (dict(order_enforce=False, disable_env_checker=True), dict()):
                try:
                    reg.register(id=name, entry_point=entry, **extra)
                    break
                except TypeError:
                    continue
                except Exception:
                    break


_register()
