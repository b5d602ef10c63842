"""Stub of the 2018-era `gym` (0.10.x) API surface the reference touches [upstream, recalled].
TEST INFRASTRUCTURE: see oracle/refharness/stubs/README.md."""
from . import error, spaces, utils  # noqa: F401
from .core import Env, GoalEnv, Wrapper  # noqa: F401
from .envs.registration import make, register, spec  # noqa: F401

__version__ = "0.10.5-stub"
