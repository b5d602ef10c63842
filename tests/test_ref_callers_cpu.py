"""The reference's CALLERS run unmodified (CPU): pins the caller-side restatements of oracle/callers_oracle.py.

  * gym_blocks/rollout.py RolloutStudent (:9-172) on the reference's own envs (gym.make under the stub packages)
    with a closed-loop policy -> the episode batch `convert_episode_to_batch_major` (util.py:118-128) returns; the
    C oracle stepped with the same policy must produce the same batch (layout, dtypes, values);
  * policy_gradient/rollout.py RolloutStudent (:28-300): its O(T^2) return accumulation (:255-258) against
    callers_oracle.discounted_returns bit for bit, and its `trim` (:105-171) executed from source against
    callers_oracle.trim;
  * config.configure_her (config.py:107-123): the reward_fun closure it builds around env.compute_reward.

Needs the reference (sources here, oracle/_ref elsewhere); skipped where neither exists.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from oracle import callers_oracle as co, coracle, refharness as rh  # noqa: E402
from ref_callers_common import QuantisedPolicy, QuietLogger, dims_of, fake_pg_self  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.available(), reason="neither /root/reference nor oracle/_ref is present")
T = 50
IDS = ["GripperTouch-v0", "BlocksTouch-v0", "ToppleTower-v0", "BlocksTouchChooseCurriculum-v0", "BlocksTouchVariation-v0"]


def oracle_episode(env, policy, test=False):
    """RolloutStudent.generate_rollouts (rollout.py:75-172) restated over a C-oracle vec env: reset (+ set_test), T
    closed-loop steps, batch-major float32 episode as convert_episode_to_batch_major (util.py:118-128) lays it out."""
    o, ag, g = env.reset()
    if test:
        o, ag, g = env.set_test()
    ep = dict(o=[o], ag=[ag], u=[], g=[], info_is_success=[], r=[])
    for _ in range(T):
        u = policy.get_actions(o, ag, g)
        o, ag, r, s, _, _ = env.step(u)
        ep["o"].append(o); ep["ag"].append(ag); ep["u"].append(u); ep["g"].append(g); ep["info_is_success"].append(s[:, None]); ep["r"].append(r)
    return {k: np.stack(v).swapaxes(0, 1).astype(np.float32) for k, v in ep.items()}


def close(a, b):
    return np.allclose(a, b, rtol=1e-6, atol=1e-6)


def has_set_test(name):
    return name not in ("GripperTouch-v0", "ToppleTower-v0")                # fetch_env.py:99-101 raises for these


@pytest.mark.parametrize("name", IDS)
def test_rollout_student_unmodified_on_reference_envs_equals_oracle_rollout(name):
    ro, _, _ = rh.callers()
    B, seed = 2, 40
    dims = dims_of(rh.make(name, seed=0))
    pol = QuantisedPolicy(dims["o"], dims["g"], name)
    worker = ro.RolloutStudent(lambda: rh.make(name), pol, dims, QuietLogger(), T, rollout_batch_size=B)
    worker.seed(seed)                                                       # env i <- seed + 1000 i (rollout.py:206-210)
    orc = coracle.OracleVecEnv(name, B, seed=seed)
    orc_pol = QuantisedPolicy(dims["o"], dims["g"], name)
    for test in (False, False, True) if has_set_test(name) else (False, False):
        ep = worker.generate_rollouts(test=test)
        want = oracle_episode(orc, orc_pol, test=test)
        assert ep["o"].shape == (B, T + 1, dims["o"]) and ep["ag"].shape == (B, T + 1, dims["g"])
        assert ep["u"].shape == (B, T, 4) and ep["g"].shape == (B, T, dims["g"]) and ep["info_is_success"].shape == (B, T, 1)
        assert np.array_equal(ep["u"], want["u"]), "closed-loop actions diverged"
        assert np.array_equal(ep["ag"], want["ag"]) and np.array_equal(ep["g"], want["g"])
        assert np.array_equal(ep["info_is_success"], want["info_is_success"])
        assert close(ep["o"], want["o"])
        assert worker.success_history[-1] == want["info_is_success"][:, -1, 0].mean()   # rollout.py:163-167


@pytest.mark.parametrize("name", ["BlocksTouch-v0", "BlocksTouchVariation-v0"])
def test_policy_gradient_rollout_student_returns_accumulation(name):
    """policy_gradient/rollout.py run unmodified: its G (the O(T^2) loop of :255-258, gamma = 1 - 1/T, config.py:83)
    equals callers_oracle.discounted_returns on the rewards step() returned, bit for bit in float64."""
    _, _, pg = rh.callers()
    B, seed = 3, 11
    dims = dims_of(rh.make(name, seed=0))
    pol = QuantisedPolicy(dims["o"], dims["g"], name)
    gamma = 1. - 1. / T
    worker = pg.RolloutStudent(lambda: rh.make(name), pol, None, dims, QuietLogger(), T, rollout_batch_size=B, gamma=gamma)
    worker.seed(seed)
    ep = worker.generate_rollouts(exploit=True)
    orc = coracle.OracleVecEnv(name, B, seed=seed)
    want = oracle_episode(orc, QuantisedPolicy(dims["o"], dims["g"], name))
    assert np.array_equal(ep["u"], want["u"]) and np.array_equal(ep["ag"], want["ag"])
    G = co.discounted_returns(want["r"].T, gamma).T                         # oracle: time-major in, like the loop
    assert ep["G"].shape == (B, T) and ep["G"].dtype == np.float64
    assert np.array_equal(ep["G"].view(np.uint64), np.ascontiguousarray(G).view(np.uint64))


def test_policy_gradient_trim_executed_from_source():
    """RolloutStudent.trim (policy_gradient/rollout.py:105-171) called on the reference's own method object."""
    _, _, pg = rh.callers()
    name = "BlocksTouchVariation-v0"
    orc = coracle.OracleVecEnv(name, 24, seed=5)
    dims = dict(o=87, g=36)
    ep = oracle_episode(orc, QuantisedPolicy(87, 36, name))
    for t in (0, 13, 50):
        o, g, ag = ep["o"][:, t], ep["g"][:, min(t, T - 1)], ep["ag"][:, t]
        ro_, rg_, rag_ = pg.RolloutStudent.trim(fake_pg_self(name), o, g, ag, 40, 16)
        o_, g_, ag_ = co.trim(o, g, ag, 40, 16, name)
        assert np.array_equal(np.asarray(ro_, np.float32), np.asarray(o_, np.float32))
        assert np.array_equal(rg_, g_) and np.array_equal(rag_, ag_)
    tower = coracle.OracleVecEnv("ToppleTower-v0", 4, seed=5)
    ep4 = oracle_episode(tower, QuantisedPolicy(70, 36, "ToppleTower-v0"))
    o, g, ag = ep4["o"][:, 9], ep4["g"][:, 9], ep4["ag"][:, 9]
    ro_, rg_, rag_ = pg.RolloutStudent.trim(fake_pg_self("ToppleTower-v0"), o, g, ag, 40, 16)
    o_, g_, ag_ = co.trim(o, g, ag, 40, 16, "ToppleTower-v0")
    assert np.array_equal(ro_, o_) and np.array_equal(rg_, g_) and np.array_equal(rag_, ag_)
    same = pg.RolloutStudent.trim(fake_pg_self(name), o_, g_, ag_, 40, 16)
    assert same[0] is o_                                                    # :107-108 nothing to trim


def test_configure_her_reward_fun_and_sampler():
    """config.configure_her (config.py:107-123) unmodified: reward_fun(ag_2, g, info) forwards to the env's
    compute_reward through the TimeLimit wrapper with the keyword `desired_goal`; the sampler it returns is built
    with replay_strategy / replay_k of config.py:49-50."""
    _, cfg, _ = rh.callers()
    name = "BlocksTouch-v0"
    assert cfg.DEFAULT_PARAMS["replay_strategy"] == "none" and cfg.DEFAULT_PARAMS["replay_k"] == 4   # config.py:49-50
    params = dict(make_env=lambda: rh.make(name, seed=0), replay_strategy="future", replay_k=cfg.DEFAULT_PARAMS["replay_k"])
    sampler = cfg.configure_her(params)
    assert params["_replay_k"] == 4 and "replay_k" not in params            # config.py:117-120
    orc = coracle.OracleVecEnv(name, 16, seed=2)
    ep = oracle_episode(orc, QuantisedPolicy(40, 16, name))
    batch = {k: ep[k] for k in ("o", "u", "g", "ag", "info_is_success")}
    batch["o_2"], batch["ag_2"] = batch["o"][:, 1:], batch["ag"][:, 1:]
    tr = sampler(batch, 256)
    assert tr["r"].shape == (256,) and tr["r"].dtype == np.float32
    assert np.array_equal(tr["r"].view(np.uint32), co.compute_reward(tr["ag_2"], tr["g"]).view(np.uint32))
    assert np.array_equal(tr["r"].view(np.uint32), coracle.compute_reward(tr["ag_2"], tr["g"]).view(np.uint32))
