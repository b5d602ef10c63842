"""Writes tests/golden/ref_<id>.npz: traces of the UNMODIFIED reference (matthew9671/BlockPuzzle-gym).

The reference's own `gym_blocks` package is imported as it lies under /root/reference and run under the stub
packages of oracle/refharness (fake mujoco_py with BlockPhys in the `sim.step()` slot, Philox behind the two
numpy RNGs).  Every number in these fixtures was returned by the reference's own `reset` / `step` / `set_test` /
`increase_difficulty` (tasks.py classes behind gym.make + TimeLimit), or read from its attributes
(`achieved_goal`, `has_succeeded`, `num_objs`, draw counters) -- see tests/golden/ref_scenario.py for the
programme.  Also writes ref_calls.npz: direct calls of reference functions (compute_reward, _sample_goal,
out_of_table, configure_her's reward_fun, convert_episode_to_batch_major, policy_gradient trim).

Run from the repo root in the build container (needs /root/reference):  python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_scenario as sc  # noqa: E402
from oracle import coracle, refharness as rh  # noqa: E402


def philox_actions_for(name):
    """Seeded random actions: the action stream of a side oracle instance (any deterministic source would do;
    the actions are stored in the fixture)."""
    side = coracle.OracleVecEnv(name, sc.NUM_ENVS, seed=sc.SEED + 7)
    side.reset()

    def draw(n):
        a = side.random_actions()
        side.step(a, auto_reset=True)
        return a
    return draw


def main():
    assert rh.reference_root() == rh.REF_SOURCE, "fixtures are recorded from the reference sources"
    for name in sc.ENV_IDS:
        drv = sc.RefDriver(name)
        trace = sc.record(name, drv, philox_actions_for(name))
        z = sc.pack(trace)
        succ = int(z["succ"].max(axis=0).sum())
        touched = int((z["ag"] == 1).any(axis=(0, 2)).sum())
        np.savez_compressed(os.path.join(HERE, "ref_%s.npz" % name), env_name=name, num_envs=sc.NUM_ENVS, seed=sc.SEED,
                            goal_dtype=drv.goal_dtype, **z)
        print("%-32s %4d events, %3d steps; envs that touched anything: %d, that succeeded: %d; goal dtype %s" % (
            name, len(trace), int((z["op"] == sc.OP_STEP).sum()), touched, succ, drv.goal_dtype))


if __name__ == "__main__":
    main()
