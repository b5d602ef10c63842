"""Pins the CPU oracle to the reference's OWN code.

Two layers:
  * always: the C oracle replays the committed traces tests/golden/ref_<id>.npz -- recorded from the unmodified
    reference by tests/golden/make_ref_golden.py -- touch matrix, goal, reward bits, success latch, done,
    num_objs, draw counters and the binary32 sim state bit for bit, float64 observations within 1e-6;
  * live, wherever the reference can be imported (its sources under /root/reference in the build container, the
    compiled copy oracle/_ref elsewhere): the fixtures are re-recorded and must come out byte-identical, the
    oracle is compared with the reference over 10 000 resets per sampler (rejection decisions / draw counters),
    and reference functions are called directly (compute_reward, _sample_goal, out_of_table).
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_scenario as sc  # noqa: E402
from oracle import coracle, refharness as rh  # noqa: E402

GOLDEN = os.path.join(HERE, "golden")
live = pytest.mark.skipif(not rh.available(), reason="neither /root/reference nor oracle/_ref is present")


def load_trace(name):
    z = np.load(os.path.join(GOLDEN, "ref_%s.npz" % name))
    assert int(z["num_envs"]) == sc.NUM_ENVS and int(z["seed"]) == sc.SEED
    return z, sc.unpack(z)


@pytest.mark.parametrize("name", sc.ENV_IDS)
def test_c_oracle_replays_the_reference_trace(name):
    _, want = load_trace(name)
    got = sc.replay(sc.OracleDriver(name), want)
    sc.compare(got, want, who="C oracle vs reference, " + name)


@pytest.mark.parametrize("name", sc.ENV_IDS)
def test_reference_traces_cover_the_contact_branches(name):
    """The scripted episodes drive what random actions rarely reach: finger/cube, cube/cube and cube/table touches,
    successes, and a cube leaving the table."""
    z, _ = load_trace(name)
    ag = z["ag"]
    N = int(round(np.sqrt(ag.shape[-1])))
    m = ag.reshape(len(ag), sc.NUM_ENVS, N, N)
    assert (m[:, :, 0, 2] == 1).any(), "gripper never touched cube 0"
    assert (m[:, :, 1, 2] == 1).any(), "cube 0 never touched the table"
    if name not in ("GripperTouch-v0",):
        assert (m[:, :, 2, 3] == 1).any(), "cubes 0 and 1 never touched"
    if name != "ToppleTower-v0":
        r = z["r"][z["op"] == sc.OP_STEP]
        assert z["succ"].any() and (r == 0).any()
        assert np.signbit(r[r == 0]).all()                                  # success reward is -0.0 (fetch_env.py:143)
    fell = z["state"]["blk_pos"][:, :, 0, 2] < 0.3
    assert fell.any(), "no cube left the table"
    assert (z["done"].sum(axis=0) >= 4).all()                               # TimeLimit fired in every full episode


@live
@pytest.mark.parametrize("name", sc.ENV_IDS)
def test_fixture_is_what_the_reference_produces_here(name):
    """Re-records the trace from the reference (sources or oracle/_ref) and compares it byte for byte."""
    z, want = load_trace(name)
    drv = sc.RefDriver(name)
    got = sc.replay(drv, want)
    packed = sc.pack(got)
    assert drv.goal_dtype == str(z["goal_dtype"])
    for k in sc.FIELDS:
        assert packed[k].tobytes() == np.ascontiguousarray(z[k]).tobytes(), "%s: field %s differs from the committed fixture" % (name, k)


@live
@pytest.mark.parametrize("name,resets,max_level", [
    ("GripperTouch-v0", 3000, False), ("BlocksTouch-v0", 10000, False), ("ToppleTower-v0", 2000, False),
    ("BlocksTouchCurriculum-v0", 5000, True), ("BlocksTouchChoose-v0", 5000, False),
    ("BlocksTouchChooseCurriculum-v0", 5000, True), ("BlocksTouchVariation-v0", 5000, True)])
def test_spawn_samplers_make_the_reference_rejection_decisions(name, resets, max_level):
    """R1-R5: the binary64 spawn samplers of the oracle take the reference's accept/reject decisions -- the draw
    counters of both RNG streams and the spawned positions agree on every one of `resets` resets (mismatch
    rate 0; with the binary32 samplers of BlockPhys <= v1.2 a boundary case could consume a different number
    of draws).  Curriculum ids are checked at the widest range, where the table edge rejects most often."""
    ref = rh.make(name, seed=77)
    orc = coracle.OracleVecEnv(name, 1, seed=77)
    if max_level:
        while not ref.unwrapped.increase_difficulty():
            pass
        while not orc.increase_difficulty():
            pass
    mismatches, rejected = 0, 0
    for _ in range(resets):
        ref.reset()
        orc.reset()
        a, b = rh.state_record(ref), orc.get_state()[0]
        if a.tobytes() != b.tobytes():
            mismatches += 1
        base = {"GripperTouch-v0": (1, 0), "ToppleTower-v0": (1, 0), "BlocksTouch-v0": (1, 2), "BlocksTouchCurriculum-v0": (1, 2)}.get(name)
        if base is not None and tuple(b["draws"]) != base:
            rejected += 1
    print("%s: %d resets, %d with at least one rejected draw, %d mismatches" % (name, resets, rejected, mismatches))
    assert mismatches == 0
    if name in ("GripperTouch-v0", "BlocksTouchCurriculum-v0"):
        assert rejected > 0                                                 # the rejection path was exercised


@live
@pytest.mark.parametrize("name", ["BlocksTouchChoose-v0", "BlocksTouchChooseCurriculum-v0"])
def test_choose_env_with_challenge_true_matches_the_reference(name):
    """R3 with `challenge=True` (fetch_env.py:403,416,452-463): no tasks.py class passes it, so the reference's own
    `_randomize_objects` is run with the attribute its constructor would have set.  2000 resets: state records (spawned
    positions, both draw counters) of reference, C oracle and Python oracle agree byte for byte, the blue / green pair is
    >= 0.15 apart and the wrong block within 0.04 of their centre.  Other env classes do not take the argument."""
    from oracle import gym_blocks_oracle as pyo
    ref = rh.make(name, seed=5)
    assert ref.unwrapped.challenge is False
    ref.unwrapped.challenge = True
    orc = coracle.OracleVecEnv(name, 1, seed=5)
    orc.set_challenge(True)
    py = pyo.make(name, challenge=True)
    py.seed(5)
    for i in range(2000):
        ref.reset(); orc.reset()
        a, b = rh.state_record(ref), orc.get_state()[0]
        assert a.tobytes() == b.tobytes(), i
        if i < 50:
            py.reset()
            for k in range(3):
                assert np.array_equal(np.asarray(py.sim.obj_pos(k), np.float32), b["blk_pos"][k]), (i, k)
        green, blue, wrong = b["blk_pos"][0][:2].astype(np.float64), b["blk_pos"][1][:2].astype(np.float64), b["blk_pos"][2][:2].astype(np.float64)
        assert np.linalg.norm(green - blue) >= 0.15 - 1e-6
        assert np.linalg.norm(wrong - (green + blue) / 2) <= 0.04 + 1e-6
    with pytest.raises(TypeError):
        coracle.OracleVecEnv("BlocksTouch-v0", 1, seed=5).set_challenge(True)
    with pytest.raises(TypeError):
        pyo.make("BlocksTouch-v0", challenge=True)


@live
def test_reference_compute_reward_goal_and_table_test_called_directly():
    """S6 / S7 / G0 straight from the reference module: BlocksEnv.compute_reward (fetch_env.py:135-143) against the
    oracle's bpo_compute_reward on random touch matrices, every id's _sample_goal (:260-273, :682-695) against
    the oracle's goal rows, out_of_table and the module constants (:19-32)."""
    _, _, fe = rh.modules()
    rng = np.random.RandomState(5)
    for name in sc.ENV_IDS:
        env = rh.make(name, seed=1)
        o = env.reset()
        dimg = o["achieved_goal"].size
        g = np.asarray(o["desired_goal"])
        _, _, og = coracle.OracleVecEnv(name, 1, seed=1).reset()
        assert np.array_equal(og[0], g.astype(np.float32))
        ag = rng.randint(-1, 2, size=(4, 33, dimg)).astype(np.float64)
        sat = np.where(g != 0, g, ag[0, 0])                                 # rows that satisfy the goal exactly
        ag[1, :5] = sat
        r_ref = env.unwrapped.compute_reward(ag, g, None)
        r_orc = coracle.compute_reward(ag, np.broadcast_to(g, ag.shape))
        assert r_ref.dtype == np.float32 and r_ref.shape == (4, 33)
        assert np.array_equal(r_ref.view(np.uint32), r_orc.view(np.uint32))
        assert np.signbit(r_ref[1, :5]).all() and (r_ref[1, :5] == 0).all()
        # the TimeLimit wrapper forwards compute_reward(achieved_goal, desired_goal, info) (config.py:110-111)
        assert np.array_equal(env.compute_reward(achieved_goal=ag, desired_goal=g, info={}), r_ref)
    assert fe.TABLE_H == 0.32499999999999996 and fe.TABLE_W == 0.225 and fe.MIN_BLOCK_DIST == 0.07500000000000001
    assert fe.out_of_table([1.3 + 0.2250001, 0.75]) and not fe.out_of_table([1.3 + 0.2249, 0.75 - 0.3249])
