"""gym.envs.registration [upstream, recalled]: register / make with the TimeLimit wrapper that
max_episode_steps implies (gym_blocks/__init__.py:6-53 registers seven ids with max_episode_steps=50)."""
import importlib

from .. import error


class EnvSpec(object):
    def __init__(self, id, entry_point=None, kwargs=None, max_episode_steps=None, **_):
        self.id = id
        self._entry_point = entry_point
        self._kwargs = {} if kwargs is None else kwargs
        self.max_episode_steps = max_episode_steps

    def make(self):
        if callable(self._entry_point):
            cls = self._entry_point
        else:
            mod_name, attr = self._entry_point.split(":")
            cls = getattr(importlib.import_module(mod_name), attr)
        env = cls(**self._kwargs)
        env.unwrapped.spec = self
        return env


class EnvRegistry(object):
    def __init__(self):
        self.env_specs = {}

    def register(self, id, **kwargs):
        if id in self.env_specs:
            raise error.Error("Cannot re-register id: {}".format(id))
        self.env_specs[id] = EnvSpec(id, **kwargs)

    def spec(self, id):
        try:
            return self.env_specs[id]
        except KeyError:
            raise error.UnregisteredEnv("No registered env with id: {}".format(id))

    def make(self, id):
        spec = self.spec(id)
        env = spec.make()
        if env.spec.max_episode_steps is not None:
            from ..wrappers.time_limit import TimeLimit
            env = TimeLimit(env, max_episode_steps=env.spec.max_episode_steps)
        return env


registry = EnvRegistry()


def register(id, **kwargs):
    return registry.register(id, **kwargs)


def make(id):
    return registry.make(id)


def spec(id):
    return registry.spec(id)
