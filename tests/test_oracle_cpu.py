"""CPU tests (no GPU): the oracle against the known answers derivable from the reference tree,
the Python restatement against the C restatement, golden fixtures, and physics invariants."""
import ctypes as C
import math
import os
import sys

import numpy as np
import pytest

from oracle import coracle
from oracle import gym_blocks_oracle as pyo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------- known answers
def test_goal_and_reward_known_answers():
    env = pyo.make("BlocksTouch-v0")
    goal = env._sample_goal()
    exp = np.zeros(16, int); exp[2 * 4 + 3] = exp[3 * 4 + 2] = 1
    assert np.array_equal(goal, exp) and goal.dtype.kind == "i"
    r = env.compute_reward(-np.ones(16), goal, None)
    assert r.dtype == np.float32 and r == -1.0
    ag = -np.ones(16); ag[2 * 4 + 3] = ag[3 * 4 + 2] = 1
    r = env.compute_reward(ag, goal, None)
    assert r == 0 and np.signbit(r)                      # -0.0 on success (appendix A7)
    r3 = env.compute_reward(np.stack([ag] * 3), goal, None)
    assert r3.shape == (3,) and np.signbit(r3).all() and (r3 == 0).all()
    tower = pyo.make("ToppleTower-v0")
    g = tower._sample_goal().reshape(6, 6)
    assert g[0, 5] == g[5, 0] == -1 and g[1, 5] == g[5, 1] == 1 and np.count_nonzero(g) == 4
    grip = pyo.make("GripperTouch-v0")._sample_goal().reshape(3, 3)
    assert grip[0, 2] == grip[2, 0] == 1 and np.count_nonzero(grip) == 2
    var = pyo.make("BlocksTouchVariation-v0")._sample_goal()
    assert var.shape == (36,) and var.reshape(6, 6)[2, 3] == 1 and np.count_nonzero(var) == 2


def test_c_compute_reward_equals_numpy_formula():
    rng = np.random.RandomState(0)
    for dimg in (9, 16, 25, 36):
        ag = rng.randint(-1, 2, size=(500, dimg)).astype(np.float32)
        g = rng.randint(-1, 2, size=(500, dimg)).astype(np.float32)
        ag[::4] = g[::4]
        d = np.sum(ag * g, axis=-1); c = np.count_nonzero(g, axis=-1)     # fetch_env.py:141-143 verbatim semantics
        ref = -(d != c).astype(np.float32)
        assert np.array_equal(coracle.compute_reward(ag, g).view(np.uint32), ref.view(np.uint32))


def test_env_dims_table():
    for name, (dimo, dimg) in zip(coracle.ENV_IDS, zip(coracle.DIMO, coracle.DIMG)):
        env = pyo.make(name)
        env.seed(0)
        obs = env.reset()
        assert obs["observation"].shape == (dimo,), name      # 10 + 15 n  (87 for Variation)
        assert obs["achieved_goal"].shape == (dimg,) and obs["desired_goal"].shape == (dimg,)
        assert env._max_episode_steps == 50


def test_curriculum_levels_follow_python_float_arithmetic():
    env = pyo.make("BlocksTouchCurriculum-v0")
    seen = []
    for _ in range(7):
        seen.append((env.increase_difficulty(), env.obj_range, env.get_difficulty()))
    r = 0.08
    exp = []
    lvl = 0
    for _ in range(7):
        r += 0.025
        if r > 0.2:
            r = 0.2; exp.append((True, r, lvl))
        else:
            lvl += 1; exp.append((False, r, lvl))
    assert seen == exp
    with pytest.raises(AttributeError):
        pyo.make("BlocksTouchChoose-v0").increase_difficulty()
    with pytest.raises(NotImplementedError):
        pyo.make("GripperTouch-v0").increase_difficulty()
    with pytest.raises(NotImplementedError):
        pyo.make("ToppleTower-v0").set_test()
    nc = pyo.make("BlocksTouch-v0")                           # step 0: never reports max (fetch_env.py:346-358)
    assert nc.increase_difficulty() is False and nc.get_difficulty() == 1


# ---------------------------------------------------------------- spec'd functions: Python vs C, and accuracy
def test_philox_known_answer_and_python_equals_c():
    # Random123 known-answer test for philox4x32-10
    assert coracle.philox4x32(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert coracle.philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert coracle.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    rng = np.random.RandomState(1)
    for _ in range(50):
        a = [int(x) for x in rng.randint(0, 2 ** 32, size=6, dtype=np.uint64)]
        assert list(pyo.philox4x32(*a)) == coracle.philox4x32(*a)


def test_elementary_functions_python_equals_c_and_are_accurate():
    L = coracle.lib()
    rng = np.random.RandomState(2)
    s, c = C.c_float(), C.c_float()
    for w in rng.randint(0, 2 ** 32, size=300, dtype=np.uint64):
        w = int(w)
        u = pyo.u01(w); uo = pyo.u01_open(w)
        assert u == np.float32(L.bpo_u01(w)) and uo == np.float32(L.bpo_u01_open(w))
        assert 0.0 <= u < 1.0 and 0.0 < uo < 1.0
        lg = pyo.bp_log(uo)
        assert lg == np.float32(L.bpo_log(float(uo)))
        assert abs(float(lg) - math.log(float(uo))) < 2e-6 * max(1.0, abs(math.log(float(uo))))
        ps, pc = pyo.bp_sincos2pi(u)
        L.bpo_sincos2pi(float(u), C.byref(s), C.byref(c))
        assert ps == np.float32(s.value) and pc == np.float32(c.value)
        assert abs(float(ps) - math.sin(2 * math.pi * float(u))) < 1e-6
        th = float(u) * 6.2 - 3.1
        sn, cs = np.float32(math.sin(th)), np.float32(math.cos(th))
        at = pyo.bp_atan2(sn, cs)
        assert at == np.float32(L.bpo_atan2(float(sn), float(cs)))
        assert abs(float(at) - th) < 1e-6
    assert pyo.bp_atan2(0.0, 1.0) == 0.0


# ---------------------------------------------------------------- Python restatement == C restatement
@pytest.mark.parametrize("name", coracle.ENV_IDS)
def test_python_oracle_equals_c_oracle(name):
    """Two independent restatements of the env logic (Python class hierarchy vs C), sharing only the
    BlockPhys sim.step(): every output bit-identical over 2 episodes with set_test / curriculum in between."""
    n = 3
    cenv = coracle.OracleVecEnv(name, n, seed=17)
    penvs = [pyo.make(name) for _ in range(n)]
    for i, e in enumerate(penvs):
        e.seed(17 + 1000 * i)                               # rollout.py:206-210
    for ep in range(2):
        co, cag, cg = cenv.reset()
        for i, e in enumerate(penvs):
            o = e.reset()
            assert np.array_equal(o["observation"], co[i]), (name, "reset obs")
            assert np.array_equal(o["achieved_goal"].astype(np.float32), cag[i])
            assert np.array_equal(o["desired_goal"].astype(np.float32), cg[i])
        for t in range(50):
            acts = cenv.random_actions()
            co, cag, cr, cs, _, _ = cenv.step(acts)
            for i, e in enumerate(penvs):
                assert np.array_equal(e.random_action(), acts[i])
                o, r, done, info = e.step(acts[i])
                assert np.array_equal(o["observation"], co[i]), (name, ep, t, i)
                assert np.array_equal(o["achieved_goal"].astype(np.float32), cag[i]), (name, ep, t, i)
                assert r.view(np.uint32) == cr[i].view(np.uint32)
                assert float(info["is_success"]) == cs[i]
                assert done == (t == 49)
        if name not in ("GripperTouch-v0", "ToppleTower-v0"):
            co, cag, cg = cenv.set_test()
            for i, e in enumerate(penvs):
                o = e.set_test()
                assert np.array_equal(o["observation"], co[i]), (name, "set_test")
            if name != "BlocksTouchChoose-v0":
                assert cenv.increase_difficulty() == penvs[0].increase_difficulty()
                for e in penvs[1:]:
                    e.increase_difficulty()


# ---------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("name", coracle.ENV_IDS)
def test_golden_fixture(name):
    """Fixtures written by tests/golden/make_golden.py from the oracle at the commit that froze
    BlockPhys: guards the spec against silent drift (the reference itself has no vectors)."""
    path = os.path.join(GOLDEN, name + ".npz")
    z = np.load(path)
    env = coracle.OracleVecEnv(name, int(z["num_envs"]), seed=int(z["seed"]))
    o, ag, g = env.reset()
    assert np.array_equal(o, z["reset_obs"]) and np.array_equal(g, z["goal"])
    for k in range(z["actions"].shape[0]):
        o, ag, r, s, _, _ = env.step(z["actions"][k], auto_reset=True)
        assert np.array_equal(o, z["obs"][k]), (name, k)
        assert np.array_equal(ag, z["ag"][k])
        assert np.array_equal(r.view(np.uint32), z["reward"][k].view(np.uint32))
        assert np.array_equal(s, z["success"][k])
    assert env.get_state().tobytes() == z["final_state"].tobytes()


def test_golden_her_fixture():
    z = np.load(os.path.join(GOLDEN, "her_relabel.npz"))
    out = coracle.her_relabel(z["ep_ag"], z["ep_g"], int(z["n"]), float(z["future_p"]), int(z["seed"]), int(z["offset"]))
    for k in ("ep_idx", "t", "future_t", "g", "r"):
        assert np.array_equal(out[k], z[k]), k


# ---------------------------------------------------------------- HER sampler semantics
def test_her_relabel_semantics():
    rng = np.random.RandomState(0)
    B, T, dimg = 40, 50, 16
    ag = rng.randint(-1, 2, size=(B, T + 1, dimg)).astype(np.float32)
    g = rng.randint(-1, 2, size=(B, T, dimg)).astype(np.float32)
    n = 5000
    out = coracle.her_relabel(ag, g, n, 0.8, 5)
    e, t, ft = out["ep_idx"], out["t"], out["future_t"]
    assert e.min() >= 0 and e.max() < B and t.min() >= 0 and t.max() < T
    her = ft >= 0
    assert abs(her.mean() - 0.8) < 0.03                                   # future_p = 1 - 1/(1+4)
    assert (ft[her] > t[her]).all() and (ft[her] <= T).all()            # t+1 .. T
    assert np.array_equal(out["ag_2"], ag[e, t + 1])                     # ag_2 = ag[:, 1:]
    assert np.array_equal(out["g"][her], ag[e[her], ft[her]])
    assert np.array_equal(out["g"][~her], g[e[~her], t[~her]])
    d = (out["ag_2"] * out["g"]).sum(-1); c = np.count_nonzero(out["g"], axis=-1)
    assert np.array_equal(out["r"], -(d != c).astype(np.float32))
    none = coracle.her_relabel(ag, g, n, 0.0, 5)
    assert (none["future_t"] == -1).all() and np.array_equal(none["g"], g[none["ep_idx"], none["t"]])
    # different counter offsets give different, reproducible samples
    a = coracle.her_relabel(ag, g, 100, 0.8, 5, index_offset=100)
    assert np.array_equal(a["ep_idx"], out["ep_idx"][100:200])


# ---------------------------------------------------------------- physics invariants of BlockPhys
@pytest.mark.parametrize("name", coracle.ENV_IDS)
def test_physics_invariants_under_random_actions(name):
    env = coracle.OracleVecEnv(name, 128, seed=5)
    o, ag, g = env.reset()
    dimg = env.dimg
    n_obj = int(round(math.sqrt(dimg)))
    for t in range(120):
        o, ag, r, s, _, _ = env.step(env.random_actions(), auto_reset=True)
        assert np.isfinite(o).all()
        m = ag.reshape(-1, n_obj, n_obj)
        assert np.array_equal(m, m.transpose(0, 2, 1))                    # _check_goal, fetch_env.py:119-124
        assert set(np.unique(ag)).issubset({-1.0, 0.0, 1.0})
        assert (np.diagonal(m, axis1=1, axis2=2) == -1).all()
        assert set(np.unique(r)).issubset({0.0, -1.0})
        st = env.get_state()
        nb = st["num_objs"] - 2
        for b in range(4):
            live = nb > b
            cs = st["blk_cs"][live, b]
            # BlockPhys v1.2 renormalises with one Newton step: |c^2 + s^2 - 1| <= 3/4 dth^4 <= 1.6e-4 at the 60 rad/s cap
            assert np.allclose((cs ** 2).sum(-1), 1.0, atol=2e-4)
            assert (st["blk_pos"][live, b, 2] >= 0.025 - 1e-7).all()      # never below the floor
        assert (st["grip_pos"][:, 2] >= 0.4785).all()
        assert (np.abs(st["blk_vel"]) <= 5.0).all() and (np.abs(st["blk_w"]) <= 60.0).all()
        assert ((st["finger_q"] >= 0) & (st["finger_q"] <= 0.05)).all()


def test_rest_state_is_an_exact_fixed_point():
    """Cubes at rest (incl. the ToppleTower stack) must not drift by a single ulp under zero actions."""
    for name in ("BlocksTouch-v0", "ToppleTower-v0", "BlocksTouchVariation-v0"):
        env = coracle.OracleVecEnv(name, 16, seed=1)
        env.reset()
        s0 = env.get_state()
        for _ in range(5):
            env.step(np.zeros((16, 4), np.float32))
        s1 = env.get_state()
        for f in ("blk_pos", "blk_cs", "blk_vel", "blk_w"):
            assert np.array_equal(s0[f], s1[f]), (name, f)


def test_success_latch_and_touch_persistence():
    """Appendix A1/A2: is_success is sticky within an episode; the touch matrix is NOT cleared by reset()."""
    env = coracle.OracleVecEnv("BlocksTouch-v0", 1, seed=3)
    env.reset()
    st = env.get_state()
    st["ag"][0][2 * 4 + 3] = st["ag"][0][3 * 4 + 2] = 1                    # pretend the cubes touch now
    env.set_state(st)
    # first step: cubes are apart, so 1 -> 0 downgrade; not a success
    o, ag, r, s, _, _ = env.step(np.zeros((1, 4), np.float32))
    assert ag[0][2 * 4 + 3] == 0 and r[0] == -1 and s[0] == 0
    st = env.get_state(); st["has_succeeded"][0] = 1; env.set_state(st)
    o, ag, r, s, _, _ = env.step(np.zeros((1, 4), np.float32))
    assert r[0] == -1 and s[0] == 1                                        # latch survives a failing step
    o, ag2, g = env.reset()
    assert ag2[0][2 * 4 + 3] == 0 and env.get_state()["has_succeeded"][0] == 0   # 0 (touched before) survives reset
    var = coracle.OracleVecEnv("BlocksTouchVariation-v0", 1, seed=3)
    var.reset(); st = var.get_state(); st["ag"][0][:] = 0; var.set_state(st)
    o, agv, g = var.reset()
    assert (agv == -1).all()                                               # Variation does clear it (:663)


def test_set_test_returns_stale_observation():
    """Appendix A4: set_test() re-spawns but returns obs computed from the pre-spawn site positions."""
    env = coracle.OracleVecEnv("BlocksTouchCurriculum-v0", 4, seed=9)
    o0, _, _ = env.reset()
    o1, _, _ = env.set_test()
    assert np.array_equal(o0, o1)
    o2, *_ = env.step(np.zeros((4, 4), np.float32))
    assert not np.array_equal(o2[:, 10:12], o0[:, 10:12])                  # new cube positions visible after a step


def test_scripted_push_makes_blocks_touch():
    """The task is solvable in BlockPhys: a scripted straight push of cube 0 into cube 1 latches
    is_success (reward -0.0 at the touching step) for most spawns."""

    def run(seed):
        env = coracle.OracleVecEnv("BlocksTouch-v0", 1, seed=seed)
        env.reset()
        st = env.get_state()[0]
        b0, b1 = st["blk_pos"][0][:2].copy(), st["blk_pos"][1][:2].copy()
        d = (b1 - b0) / np.linalg.norm(b1 - b0)
        res = []

        def goto(xy, z, steps):
            for _ in range(steps):
                gp = env.get_state()[0]["grip_pos"]
                a = np.zeros(4, np.float32)
                dxy = (xy - gp[:2]) / 0.05
                a[:2] = dxy / max(1.0, np.abs(dxy).max())                  # straight-line approach
                a[2] = np.clip((z - gp[2]) / 0.05, -1, 1); a[3] = -1
                res.append(env.step(a[None]))

        goto(b0 - d * 0.07, 0.55, 6)
        goto(b0 - d * 0.07, 0.48, 4)
        goto(b1 + d * 0.1, 0.48, 12)
        touched = any(r[2][0] == 0 and np.signbit(r[2][0]) for r in res)
        assert touched == bool(res[-1][3][0])                              # latch <=> some step had r == -0.0
        return touched

    assert sum(run(s) for s in range(16)) >= 12


def test_propagator_tables_are_generated_and_identical_on_both_sides():
    """BlockPhys v2: oracle/blockphys_tables.h and blockpuzzle_gym_b200/csrc/bp_tables.cuh are the exact powers of the
    substep map, rounded to binary32, written by tools/gen_blockphys_tables.py -- both files are re-rendered and
    compared byte for byte, and the 20th power is checked against 20 applications of the float64 recurrence."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_tables", os.path.join(root, "tools", "gen_blockphys_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    for kind, path in gen.OUT.items():
        assert open(path).read() == gen.render(kind), path
    t = gen.tables()
    for (K, B, pre) in ((2500.0, 100.0, "G"), (30000.0 / 104.0, 1000.0 / 104.0, "F")):
        h = 0.002
        for d0, v0 in ((1.0, 0.0), (0.0, 1.0)):
            d, v = d0, v0
            for _ in range(20):
                v = v + h * (-K * d - B * v)
                d = d + h * v
            a, c = (t[pre + "A"][20], t[pre + "C"][20]) if d0 else (t[pre + "B"][20], t[pre + "D"][20])
            assert abs(d - a) <= 1e-6 * max(1.0, abs(a)) and abs(v - c) <= 1e-5 * max(1.0, abs(c))


def test_v2_channel_events_finger_lands_on_a_cube_and_fingers_close_on_a_cube():
    """BlockPhys v2: a gripper / finger channel follows the tabulated propagator until a contact acts on it.  Scripted
    episodes drive both events: closed fingers descending onto cube 0 come to rest on its top face (grip z = cube z +
    0.0385 + 0.025 - 0.02 while the target is below), open fingers lowered around cube 0 and closed stall on its faces
    (q = 0.025 + 0.007 - 0.0079) -- and the cube stays where it was."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import ref_callers_common as rc
    ref = coracle.OracleVecEnv("BlocksTouch-v0", 64, seed=11)
    ref.reset()
    cube0 = ref.get_state()["blk_pos"][:, 0].copy()
    rc.grasp_and_land_actions(ref)
    st = ref.get_state()
    land = (np.arange(64) % 2) == 0
    assert np.mean(np.abs(st["grip_pos"][land, 2] - 0.5285) < 1e-6) > 0.9
    assert np.mean(np.abs(st["finger_q"][~land] - 0.0241).max(axis=1) < 2e-4) > 0.9
    assert np.mean(np.abs(st["blk_pos"][:, 0] - cube0).max(axis=1) < 2e-3) > 0.9


def test_quiet_path_interval_bound_contains_every_substep():
    """The CUDA quiet path (quiet_gripper_step, bp_device.cuh) bounds the gripper / finger state over the 20 substeps of an
    env-step by an interval product: state_n = target + A[n] d0 + B[n] v0 with A[n] in [Amin, Amax], B[n] in [Bmin, Bmax]
    (all positive).  Checked here on the tables themselves, in the kernel's binary32 arithmetic, for 2*10^5 random start
    states: every table-driven position lies inside [lo, hi] up to a few ulp -- far inside the kernel's 1e-5 guard."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_tables", os.path.join(root, "tools", "gen_blockphys_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    t = {k: np.array(v, np.float32) for k, v in gen.tables().items()}
    f32 = np.float32
    rng = np.random.RandomState(3)
    n = 200000
    for pre, span_d, span_v, tgt in (("G", 0.06, 1.0, 1.3), ("F", 0.2, 2.0, 0.05)):
        A, Bt = t[pre + "A"][1:], t[pre + "B"][1:]
        assert (A > 0).all() and (Bt > 0).all()
        amin, amax, bmin, bmax = A.min(), A.max(), Bt.min(), Bt.max()
        d0 = rng.uniform(-span_d, span_d, n).astype(f32)
        v0 = rng.uniform(-span_v, span_v, n).astype(f32)
        v0[: n // 10] = 0
        d0[n // 10: n // 5] = 0
        m = np.full(n, tgt, f32)
        lo = m + (np.minimum(amin * d0, amax * d0) + np.minimum(bmin * v0, bmax * v0))
        hi = m + (np.maximum(amin * d0, amax * d0) + np.maximum(bmin * v0, bmax * v0))
        assert lo.dtype == f32 and hi.dtype == f32
        for k in range(20):
            # the kernel's F(A, d0, F(B, v0, m)), evaluated in float64 and rounded: within an ulp of the binary32 FMA chain
            g = (A[k].astype(np.float64) * d0 + (Bt[k].astype(np.float64) * v0 + m)).astype(f32)
            assert (g >= lo - f32(1e-6)).all() and (g <= hi + f32(1e-6)).all(), (pre, k)
