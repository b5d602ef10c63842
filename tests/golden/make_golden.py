"""Generates the auto-reset / HER fixtures of tests/golden/ from the C oracle (oracle/blockphys_oracle.c).

These vectors (<id>.npz, her_relabel.npz) pin the batched conveniences the reference does not have --
auto-reset inside a fused launch, the Philox HER sampler -- at the commit that froze them; they are
oracle outputs.  The fixtures recorded from the REFERENCE ITSELF are ref_<id>.npz (make_ref_golden.py),
and the oracle is pinned to those.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import coracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def env_fixture(name, num_envs=6, seed=1234, steps=60):
    env = coracle.OracleVecEnv(name, num_envs, seed=seed)
    reset_obs, _, goal = env.reset()
    rng = np.random.RandomState(99)
    acts, obs, ag, rew, suc = [], [], [], [], []
    for k in range(steps):
        # half Philox actions, half host actions that leave [-1, 1] (exercise the clip)
        a = env.random_actions() if k % 2 == 0 else rng.uniform(-1.5, 1.5, size=(num_envs, 4)).astype(np.float32)
        o, g, r, s, _, _ = env.step(a, auto_reset=True)
        acts.append(a); obs.append(o); ag.append(g); rew.append(r); suc.append(s)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), num_envs=num_envs, seed=seed, reset_obs=reset_obs, goal=goal,
                        actions=np.stack(acts), obs=np.stack(obs), ag=np.stack(ag).astype(np.int8).astype(np.float32),
                        reward=np.stack(rew), success=np.stack(suc), final_state=env.get_state())


def her_fixture():
    B, T = 12, 50
    env = coracle.OracleVecEnv("BlocksTouch-v0", B, seed=7)
    _, ag0, g = env.reset()
    ags = [ag0]
    for _ in range(T):
        _, ag, _, _, _, _ = env.step(env.random_actions())
        ags.append(ag)
    ep_ag = np.stack(ags, 1).astype(np.float32)
    ep_g = np.repeat(g[:, None, :], T, 1).astype(np.float32)
    n, fp, seed, off = 512, 0.8, 31, 4096
    out = coracle.her_relabel(ep_ag, ep_g, n, fp, seed, off)
    np.savez_compressed(os.path.join(OUT, "her_relabel.npz"), ep_ag=ep_ag.astype(np.int8), ep_g=ep_g.astype(np.int8), n=n,
                        future_p=fp, seed=seed, offset=off, ep_idx=out["ep_idx"], t=out["t"], future_t=out["future_t"],
                        g=out["g"].astype(np.int8), r=out["r"])


if __name__ == "__main__":
    for name in coracle.ENV_IDS:
        env_fixture(name)
    her_fixture()
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))
