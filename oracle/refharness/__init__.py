"""Replay harness: the UNMODIFIED reference `gym_blocks` run under stub packages (TEST INFRASTRUCTURE).

This is how the CPU oracle (oracle/blockphys_oracle.c, oracle/gym_blocks_oracle.py) and the committed
golden fixtures (tests/golden/ref_*.npz) are pinned to the reference's own code:

  * the reference package is imported as it lies under /root/reference (or, where that does not exist --
    the GPU box --, from the byte-compiled copy oracle/build_ref.py leaves in oracle/_ref/);
  * `gym`, `mujoco_py`, `baselines`, `tensorflow`, `mpi4py` resolve to oracle/refharness/stubs (see its
    README): the fake MjSim puts BlockPhys in the `sim.step()` slot, `seeding.np_random` hands out the
    Philox replay generator, and `fetch_env.np` is replaced by a proxy whose `.random` routes the
    reference's *global* np.random draws (fetch_env.py:390,392,490,492,507,509,734,736) to the calling
    env's Philox stream 1;
  * `make(env_id)` = `gym.make(env_id)` (the reference's own registration, gym_blocks/__init__.py:6-53)
    inside a thin wrapper that tells the Philox generators where an episode starts.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import os
import sys

import numpy as np

from .. import coracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = os.path.dirname(_HERE)
STUBS = os.path.join(_HERE, "stubs")
REF_SOURCE = "/root/reference"
REF_COMPILED = os.path.join(_ORACLE, "_ref")

_state = {"root": None}


def reference_root():
    """Where the reference package is taken from: its sources in this container, else the compiled copy
    (BP_REF_ROOT overrides, e.g. to exercise oracle/_ref here)."""
    forced = os.environ.get("BP_REF_ROOT")
    if forced:
        return forced if os.path.isdir(os.path.join(forced, "gym_blocks")) else None
    if os.path.isdir(os.path.join(REF_SOURCE, "gym_blocks")):
        return REF_SOURCE
    if os.path.exists(os.path.join(REF_COMPILED, "gym_blocks", "__init__" + ".pyb")):
        return REF_COMPILED
    return None


def available():
    return reference_root() is not None


class _GlobalRandomRouter(object):
    """Stands in for the numpy.random module inside fetch_env: the reference's process-global draws are
    routed to the Philox stream 1 of the env whose method is making the call."""

    @staticmethod
    def _stream():
        env = sys._getframe(2).f_locals["self"]
        return env.np_random.global_stream

    def normal(self, loc=0.0, scale=1.0, size=None):
        return self._stream().normal(loc, scale, size)

    def uniform(self, low=0.0, high=1.0, size=None):
        return self._stream().uniform(low, high, size)

    def shuffle(self, x):
        raise NotImplementedError("np.random.shuffle is commented out in the reference (fetch_env.py:438,583,636)")


class _NumpyProxy(object):
    def __init__(self, real):
        self._real = real
        self.random = _GlobalRandomRouter()

    def __getattr__(self, name):
        return getattr(self._real, name)


COMPILED_EXT = ".pyb"


class _CompiledReferenceFinder(object):
    """Import finder for oracle/_ref: `gym_blocks[.x.y]` -> the sourceless bytecode files oracle/build_ref.py wrote."""

    def __init__(self, root):
        self.root = root

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery
        import importlib.util
        if name != "gym_blocks" and not name.startswith("gym_blocks."):
            return None
        rel = os.path.join(self.root, *name.split("."))
        pkg, mod = os.path.join(rel, "__init__" + COMPILED_EXT), rel + COMPILED_EXT
        if os.path.exists(pkg):
            return importlib.util.spec_from_file_location(name, pkg, loader=importlib.machinery.SourcelessFileLoader(name, pkg),
                                                          submodule_search_locations=[rel])
        if os.path.exists(mod):
            return importlib.util.spec_from_file_location(name, mod, loader=importlib.machinery.SourcelessFileLoader(name, mod))
        return None


def activate(root=None):
    """Make `import gym_blocks` resolve to the unmodified reference under the stub packages."""
    if _state["root"] is not None:
        return _state["root"]
    root = root or reference_root()
    if root is None:
        raise RuntimeError("neither /root/reference nor oracle/_ref is present: run oracle/build_ref.py where the reference exists")
    repo = os.path.dirname(_ORACLE)
    compiled = not os.path.exists(os.path.join(root, "gym_blocks", "__init__.py"))
    for p in (repo, STUBS) + (() if compiled else (root,)):
        if p not in sys.path:
            sys.path.append(p)
    if compiled and not any(isinstance(f, _CompiledReferenceFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _CompiledReferenceFinder(root))   # ahead of PathFinder, which would see namespace packages
    import gym
    assert gym.__version__.endswith("-stub"), "a real gym shadows the replay stubs"
    import gym_blocks  # noqa: F401  registers the seven ids (gym_blocks/__init__.py:6-53)
    from gym_blocks.envs import fetch_env
    assert os.path.abspath(fetch_env.__file__).startswith(os.path.abspath(root)), fetch_env.__file__
    if not isinstance(fetch_env.np, _NumpyProxy):
        fetch_env.np = _NumpyProxy(np)
    _state["root"] = root
    return root


def modules():
    """(gym, gym_blocks, fetch_env) of the activated harness."""
    activate()
    import gym
    import gym_blocks
    from gym_blocks.envs import fetch_env
    return gym, gym_blocks, fetch_env


def _replay_env_class():
    import gym

    class ReplayEnv(gym.Wrapper):
        """gym.make(env_id) of the reference + the episode bookkeeping of the Philox replay: reset() number n
        draws under episode counter n with the draw counters restarted."""

        def __init__(self, env):
            super(ReplayEnv, self).__init__(env)
            self._episodes = 0

        def __getattr__(self, name):
            return getattr(self.env, name)

        def seed(self, seed=None):
            self._episodes = 0
            return self.env.seed(seed)

        def reset(self):
            self.env.unwrapped.np_random.begin_episode(self._episodes)
            obs = self.env.reset()
            self._episodes += 1
            return obs

    return ReplayEnv


def make(env_id, seed=None):
    """The reference env `gym.make(env_id)` returns (TimeLimit(50) around the tasks.py class), replay-wrapped."""
    gym, _, _ = modules()
    env = _replay_env_class()(gym.make(env_id))
    if seed is not None:
        env.seed(seed)
    return env


def state_record(env):
    """The canonical per-env state record (coracle.STATE_DTYPE = bp_env_state) of a replay-wrapped reference env."""
    u = env.unwrapped
    s = u.sim.s
    rec = np.zeros((), dtype=coracle.STATE_DTYPE)
    rec["grip_pos"] = s.g[:]; rec["grip_vel"] = s.gv[:]
    rec["finger_q"] = s.q[:]; rec["finger_qv"] = s.qv[:]
    for i in range(s.nblocks):
        b = s.blk[i]
        rec["blk_pos"][i] = b.pos[:]; rec["blk_cs"][i] = (b.c, b.s); rec["blk_vel"][i] = b.vel[:]; rec["blk_w"][i] = b.w
    ag = np.full(36, -1, np.int8)
    flat = np.asarray(u.achieved_goal).ravel()
    assert np.all(flat == np.round(flat))
    ag[:flat.size] = flat.astype(np.int8)
    rec["ag"] = ag
    rec["num_objs"] = u.num_objs
    rec["has_succeeded"] = int(bool(u.has_succeeded))
    rec["t"] = env.env._elapsed_steps
    rec["episode"] = env._episodes
    rec["draws"] = (u.np_random.draws, u.np_random.global_stream.draws)
    return rec


def callers():
    """The reference's callers of the env, imported unmodified: (gym_blocks.rollout, gym_blocks.config,
    gym_blocks/policy_gradient/rollout.py).  The policy-gradient directory has no __init__.py upstream (its
    files are run as scripts), so that module is loaded by path."""
    import importlib.util
    root = activate()
    import gym_blocks.rollout as ro
    import gym_blocks.config as cfg
    name = "gym_blocks_policy_gradient_rollout"
    if name not in sys.modules:
        import importlib.machinery
        base = os.path.join(root, "gym_blocks", "policy_gradient", "rollout")
        if os.path.exists(base + ".py"):
            spec = importlib.util.spec_from_file_location(name, base + ".py")
        else:
            spec = importlib.util.spec_from_file_location(name, base + COMPILED_EXT,
                                                          loader=importlib.machinery.SourcelessFileLoader(name, base + COMPILED_EXT))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules[name] = mod
    return ro, cfg, sys.modules[name]
