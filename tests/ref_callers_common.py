"""Shared pieces of the "reference callers, unmodified" tests (CPU: on the reference's own envs; GPU: on the drop-in)."""
import types

import numpy as np


class QuantisedPolicy(object):
    """A seeded closed-loop stand-in for DDPG.get_actions (ddpg.py:122-156) / PGGD.get_actions: reads o, ag and g,
    returns actions on a 1/8 grid in [-1.25, 1.25] (some outside the action box: the env clips).  Quantising makes
    the action insensitive to the <= 1e-6 float64-vs-binary32 differences of the observations it reads, so the
    same policy drives bit-identical episodes on the reference env and on the CUDA drop-in."""

    def __init__(self, dimo, dimg, env_name, seed=3):
        rng = np.random.RandomState(seed)
        self.Wo = rng.normal(size=(dimo, 4)) * 2.0
        self.Wg = rng.normal(size=(dimg, 4)) * 0.3
        self.kwargs = {"info": {"env_name": env_name}}     # what policy_gradient/rollout.py:49,119,148 read
        self.calls = 0

    def _u(self, o, ag, g):
        x = np.asarray(o, np.float64) @ self.Wo + (np.asarray(ag, np.float64) - np.asarray(g, np.float64)) @ self.Wg
        self.calls += 1
        return (np.round(np.tanh(x) * 10.0) / 8.0).astype(np.float32)

    # gym_blocks/rollout.py:92-97
    def get_actions(self, o, ag, g, compute_Q=False, noise_eps=0., random_eps=0., use_target_net=False, exploit=None):
        u = self._u(o, ag, g)
        if exploit is not None:                              # policy_gradient/rollout.py:199-201: (u, raw, sigma)
            return u, u.copy(), np.zeros_like(u)
        return u


def dims_of(env):
    """config.configure_dims (config.py:159-183) without the cached env: o / u / g / info_is_success."""
    obs = env.reset()
    return {"o": obs["observation"].shape[0], "u": 4, "g": obs["desired_goal"].shape[0], "info_is_success": 1}


class QuietLogger(object):
    def info(self, *a):
        pass

    warning = warn = info


def fake_pg_self(env_name):
    """`self` for calling policy_gradient RolloutStudent.trim (rollout.py:105-171) straight from its source."""
    return types.SimpleNamespace(kwargs={"info": {"env_name": env_name}})


# ---- BlockPhys v2 channel events: scripted episodes that make a finger land on a cube (the z
# channel leaves the propagator: "lift") and close the fingers on a cube (the finger channels leave it: "closing undone")
def grasp_and_land_actions(ref, steps=36):
    """Closed-loop on an OracleVecEnv (BlocksTouch-v0): even envs descend with CLOSED fingers onto cube 0 (the fingers
    land on its top face), odd envs open the fingers, descend around cube 0 and close them on it.  Returns the actions
    [steps][B][4] and the per-step oracle outputs."""
    import numpy as np
    B = ref.n
    acts, outs = [], []
    for t in range(steps):
        st = ref.get_state()
        gp, cube = st["grip_pos"], st["blk_pos"][:, 0]
        a = np.zeros((B, 4), np.float32)
        dxy = (cube[:, :2] - gp[:, :2]) / 0.05
        a[:, :2] = dxy / np.maximum(1.0, np.abs(dxy).max(axis=1, keepdims=True))
        near = np.abs(cube[:, :2] - gp[:, :2]).max(axis=1) < 0.004
        land = (np.arange(B) % 2) == 0
        # land: fingers closed, go down once above the cube.  grasp: fingers open on the way, closed from step 22 on
        a[:, 2] = np.where(near, -1.0, np.clip((0.60 - gp[:, 2]) / 0.05, -1, 1))
        a[:, 3] = np.where(land, -1.0, 1.0 if t < 22 else -1.0)
        acts.append(a)
        outs.append(ref.step(a))
    return np.stack(acts), outs


def homing_actions(ref, steps=100, seed=0):
    """Contact-heavy closed loop on an OracleVecEnv of any id: the gripper homes in on a (per-env, per-phase) cube with noisy
    steps, dives to table level and rams it, while the fingers open and close at random -- lifts, closing-undone events,
    pushes off the table, cube-cube and tower contacts all occur.  Auto-reset at the TimeLimit.  Returns actions [steps][B][4]."""
    import numpy as np
    rng = np.random.RandomState(seed)
    B = ref.n
    acts = []
    for t in range(steps):
        st = ref.get_state()
        nb = np.maximum(st["num_objs"] - 2, 1)
        target = (np.arange(B) + t // 17) % nb
        cube = st["blk_pos"][np.arange(B), target]
        gp = st["grip_pos"]
        a = rng.uniform(-0.35, 0.35, size=(B, 4)).astype(np.float32)
        d = (cube - gp) / 0.05
        d[:, 2] = (np.where(t % 17 < 6, 0.56, 0.47) - gp[:, 2]) / 0.05          # hover, then dive
        a[:, :3] += np.clip(d, -1, 1).astype(np.float32)
        a[:, 3] = np.where(rng.rand(B) < 0.5, 1.0, -1.0)
        a[rng.rand(B) < 0.02] *= 3.0                                               # some out-of-range actions (the clip)
        acts.append(a)
        ref.step(a, auto_reset=True)
    return np.stack(acts)
