"""Single-env Python restatement of the reference step loop -- CPU ORACLE, test infrastructure only.

PARITY STATUS: the reference-owned logic restated here is PINNED to the reference's own code through the C
restatement: oracle/refharness runs the unmodified reference (stub gym / mujoco_py, BlockPhys in the MjSim
slot), tests/test_ref_pin.py holds the C oracle to its traces bit for bit, and tests/test_oracle_cpu.py holds this
module to the C oracle bit for bit.  The dynamics slot itself stays "own spec" (MuJoCo is absent from the
reference tree and from this image).

Structure follows the reference one function at a time (paths relative to
/root/reference/gym_blocks), with its own numpy-float32 arithmetic, and plugs the
BlockPhys v2 C model (oracle/blockphys_oracle.c, via ctypes -- the role mujoco_py
plays in the reference) into the `sim` slot:

    RobotEnv.seed / step / reset        envs/robot_env.py:53-82
    BlocksEnv.compute_reward            envs/fetch_env.py:135-143
    BlocksEnv._step_callback            envs/fetch_env.py:148-167
    BlocksEnv._set_action               envs/fetch_env.py:170-185
    BlocksEnv._get_obs                  envs/fetch_env.py:187-228 (Variation :567-621)
    BlocksEnv._reset_sim/_sample_goal/_is_success   envs/fetch_env.py:247-281
    *_randomize_objects                 envs/fetch_env.py:328-336, 370-399, 448-517, 697-764, 777-787
    increase_difficulty / set_test      envs/fetch_env.py:351-368, 419-446, 623-644
    constructor constants               envs/tasks.py:4-144, __init__.py:6-53

(bench.py's CPU legs time the reference itself -- oracle/refharness -- and fall back to this module only where
neither /root/reference nor oracle/_ref exists.)
The env-level C oracle (bpo_env_*) restates the same logic independently;
tests/test_oracle_cpu.py checks the two agree bit for bit.
"""
import ctypes as C

import numpy as np

from . import coracle

f32 = np.float32
GREY, RED, GREEN, BLUE = 0, 1, 2, 3          # fetch_env.py:11-15
NUM_COLORS = 4
BLOCK_SIZE = f32(0.05)                         # fetch_env.py:19
MIN_BLOCK_DIST = f32(0.075)                    # 1.5 * BLOCK_SIZE, fetch_env.py:20
TABLE_X, TABLE_Y = f32(1.3), f32(0.75)         # fetch_env.py:24-25
TABLE_W, TABLE_H = f32(0.225), f32(0.325)      # fetch_env.py:27-28
GRIP0 = np.array([1.3419, 0.7491, 0.5347], f32)  # initial_gripper_xpos (pinned; fetch_env.py:300-301)
MAX_SPAWN_ATTEMPTS = 10000


# the same constants as python floats, evaluated by the reference's own expressions (fetch_env.py:19-28)
BLOCK_SIZE64 = 0.05
MIN_BLOCK_DIST64 = 1.5 * BLOCK_SIZE64
TABLE_X64, TABLE_Y64 = 1.05 + 0.25, 0.40 + 0.35
TABLE_W64, TABLE_H64 = 0.25 - BLOCK_SIZE64 / 2, 0.35 - BLOCK_SIZE64 / 2


def out_of_table64(pos):                        # fetch_env.py:30-32
    return bool(abs(pos[0] - TABLE_X64) > TABLE_W64 or abs(pos[1] - TABLE_Y64) > TABLE_H64)


def one_hot_color(c):                           # fetch_env.py:34-37
    r = np.zeros(NUM_COLORS, f32)
    r[c] = 1
    return r


# ---------------------------------------------------------------- Philox4x32-10 and the spec'd functions
M32 = 0xFFFFFFFF


def philox4x32(c0, c1, c2, c3, k0, k1):
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def u01(x):
    return f32(f32(x >> 8) * f32(5.9604644775390625e-08))


def u01_open(x):
    return f32(f32(f32(x >> 9) + f32(0.5)) * f32(1.1920928955078125e-07))


def bp_log(x):
    ix = int(np.array(x, f32).view(np.uint32))
    ix = (ix + 0x3f800000 - 0x3f3504f3) & M32
    e = (ix >> 23) - 127
    ix = (ix & 0x007fffff) + 0x3f3504f3
    f = f32(np.array(ix, np.uint32).view(f32) - f32(1.0))
    s = f32(f / f32(f32(2.0) + f))
    z = f32(s * s)
    w = f32(z * z)
    t1 = f32(w * f32(f32(0.40000972152) + f32(w * f32(0.24279078841))))
    t2 = f32(z * f32(f32(0.66666662693) + f32(w * f32(0.28498786688))))
    R = f32(t2 + t1)
    hfsq = f32(f32(f32(0.5) * f) * f)
    dk = f32(e)
    a = f32(f32(s * f32(hfsq + R)) + f32(dk * f32(9.0580006145e-06)))
    return f32(f32(f32(a - hfsq) + f) + f32(dk * f32(6.9313812256e-01)))


def bp_sincos2pi(u):
    t = f32(u * f32(4.0))
    k = int(f32(t + f32(0.5)))
    f = f32(t - f32(k))
    x = f32(f * f32(1.57079637))
    x2 = f32(x * x)
    ps = f32(f32(-0.16666667) + f32(x2 * f32(f32(0.0083333338) + f32(x2 * f32(-0.00019841270)))))
    sp = f32(x + f32(f32(x * x2) * ps))
    pc = f32(f32(-0.5) + f32(x2 * f32(f32(0.041666668) + f32(x2 * f32(f32(-0.0013888889) + f32(x2 * f32(2.4801588e-05)))))))
    cp = f32(f32(1.0) + f32(x2 * pc))
    k &= 3
    if k == 0:
        return sp, cp
    if k == 1:
        return cp, f32(-sp)
    if k == 2:
        return f32(-sp), f32(-cp)
    return f32(-cp), sp


def bp_atan2(s, c):
    s, c = f32(s), f32(c)
    a_s, a_c = abs(s), abs(c)
    mx, mn = (a_s, a_c) if a_s > a_c else (a_c, a_s)
    if mx == 0:
        return f32(0.0)
    a = f32(mn / mx)
    off = f32(0.0)
    if a > f32(0.41421357):
        a = f32(f32(a - f32(1.0)) / f32(a + f32(1.0)))
        off = f32(0.78539819)
    a2 = f32(a * a)
    p = f32(0.076923080)
    for coef in (-0.090909094, 0.11111111, -0.14285715, 0.2, -0.33333334):
        p = f32(f32(coef) + f32(a2 * p))
    r = f32(off + f32(a + f32(f32(a * a2) * p)))
    if a_s > a_c:
        r = f32(f32(1.57079637) - r)
    if c < 0:
        r = f32(f32(3.14159274) - r)
    if s < 0:
        r = f32(-r)
    return r


def bp_normal2(w0, w1):
    u1, u2 = u01_open(w0), u01(w1)
    r = np.sqrt(f32(f32(-2.0) * bp_log(u1)))
    sn, cs = bp_sincos2pi(u2)
    return np.array([f32(r * cs), f32(r * sn)], f32)


class PhiloxRandomState:
    """Replays one RNG stream of the reference on Philox draws: stream 0 stands in for the env's
    self.np_random (robot_env.py:54), stream 1 for the process-global np.random used at
    fetch_env.py:390,392,490,492,507,509,734,736.  One call = one Philox block."""

    def __init__(self, stream):
        self.stream = stream
        self.seed_value = 0
        self.episode = 0
        self.draws = 0

    def _block(self):
        w = philox4x32(self.draws, self.episode, self.stream, 0, self.seed_value & M32, (self.seed_value >> 32) & M32)
        self.draws += 1
        return w

    def uniform(self, low=0.0, high=1.0, size=None):        # numpy: low + (high - low) * u, float64
        w = self._block()
        low, high = float(low), float(high)
        if size is None:
            return low + (high - low) * float(u01(w[0]))
        assert size == 2
        return np.array([low + (high - low) * float(u01(w[0])), low + (high - low) * float(u01(w[1]))], np.float64)

    def normal(self, size=2):
        assert size == 2
        w = self._block()
        return bp_normal2(w[0], w[1]).astype(np.float64)

    def randint(self, n):
        return (self._block()[0] * n) >> 32


# ---------------------------------------------------------------- the `sim` slot (mujoco_py.MjSim stand-in)
class _CBlock(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("c", C.c_float), ("s", C.c_float), ("vel", C.c_float * 3), ("w", C.c_float)]


class _CSim(C.Structure):
    _fields_ = [("g", C.c_float * 3), ("gv", C.c_float * 3), ("q", C.c_float * 2), ("qv", C.c_float * 2),
                ("m", C.c_float * 3), ("ctrl", C.c_float * 2), ("blk", _CBlock * 4),
                ("nblocks", C.c_int32), ("block_gripper", C.c_int32), ("contacts", C.c_uint32)]


class _Contact:
    def __init__(self, g1, g2):
        self.geom1, self.geom2 = g1, g2


class BlockSim:
    """BlockPhys behind the handful of mujoco_py calls the reference makes."""

    GEOMS = ["floor0", "robot0:r_gripper_finger_link", "robot0:l_gripper_finger_link", "table",
             "object0", "object1", "object2", "object3"]
    nsubsteps = 20            # tasks.py:16
    timestep = 0.002          # 2blocks.xml:4

    def __init__(self, env_id, nblocks):
        self.L = coracle.lib()
        self.env_id = env_id
        self.s = _CSim()
        self.L.bpo_sim_init(C.byref(self.s), env_id)
        self.s.nblocks = nblocks
        self.ngeom = 4 + nblocks

    def geom_id2name(self, i):
        return self.GEOMS[i]

    def set_state_initial(self):                    # sim.set_state(self.initial_state), fetch_env.py:248
        nb = self.s.nblocks
        self.L.bpo_sim_init(C.byref(self.s), self.env_id)
        self.s.nblocks = nb

    def set_action(self, a):                        # utils.ctrl_set_action + mocap_set_action [upstream]
        arr = (C.c_float * 4)(*[float(x) for x in a])
        self.L.bpo_sim_set_action(C.byref(self.s), arr)

    def step(self):                                 # robot_env.py:60
        self.L.bpo_sim_step(C.byref(self.s))

    def forward(self):
        pass

    @property
    def ncon(self):
        return len(self.contacts())

    def contacts(self):
        """The contact list of the last substep as (geom1, geom2) pairs, like d.contact[i]."""
        out = []
        obj_geom = {0: 1, 1: 3}                     # object id -> a representative geom id
        for o1 in range(6):
            for o2 in range(o1 + 1, 6):
                bit = o1 * (2 * 6 - o1 - 1) // 2 + (o2 - o1 - 1)
                if self.s.contacts >> bit & 1:
                    g1 = obj_geom.get(o1, o1 + 2)
                    g2 = obj_geom.get(o2, o2 + 2)
                    out.append(_Contact(g1, g2))
        return out

    # accessors in the role of sim.data.get_site_xpos / xvelp / xmat / xvelr, robot_get_obs
    def grip_pos(self):
        return np.array(self.s.g[:], f32)

    def grip_velp(self):
        return np.array(self.s.gv[:], f32)

    def finger_qpos(self):
        return np.array(self.s.q[:], f32)

    def finger_qvel(self):
        return np.array(self.s.qv[:], f32)

    def obj_pos(self, i):
        return np.array(self.s.blk[i].pos[:], f32)

    def obj_yaw_cs(self, i):
        return f32(self.s.blk[i].c), f32(self.s.blk[i].s)

    def obj_velp(self, i):
        return np.array(self.s.blk[i].vel[:], f32)

    def obj_velr(self, i):
        return np.array([0.0, 0.0, self.s.blk[i].w], f32)

    def set_obj_xy(self, i, xy):                    # object_qpos[:2] = xy; set_joint_qpos(...)
        self.s.blk[i].pos[0] = float(xy[0])
        self.s.blk[i].pos[1] = float(xy[1])


# ---------------------------------------------------------------- per-id constants (tasks.py, __init__.py)
TASKS = {
    "GripperTouch-v0": dict(eid=0, nblocks=1, block_gripper=False, kind="gripper"),
    "BlocksTouch-v0": dict(eid=1, nblocks=2, block_gripper=False, kind="touch", curriculum=False),
    "ToppleTower-v0": dict(eid=2, nblocks=4, block_gripper=False, kind="tower"),
    "BlocksTouchCurriculum-v0": dict(eid=3, nblocks=2, block_gripper=True, kind="touch", curriculum=True),
    "BlocksTouchChoose-v0": dict(eid=4, nblocks=3, block_gripper=True, kind="choose", curriculum=False),
    "BlocksTouchChooseCurriculum-v0": dict(eid=5, nblocks=3, block_gripper=True, kind="choose", curriculum=True),
    "BlocksTouchVariation-v0": dict(eid=6, nblocks=4, block_gripper=True, kind="variation"),
}
MAX_EPISODE_STEPS = 50  # __init__.py:10


class BlocksEnvOracle:
    """One env of any registered id: BlocksEnv + its task subclass + the TimeLimit wrapper."""

    def __init__(self, env_name, challenge=False):
        cfg = TASKS[env_name]
        self.cfg, self.kind, self.eid = cfg, cfg["kind"], cfg["eid"]
        if challenge and cfg["kind"] != "choose":
            raise TypeError("challenge is an argument of BlocksTouchChooseEnv only")   # fetch_env.py:403
        self.challenge = challenge                              # fetch_env.py:416
        self.block_gripper = cfg["block_gripper"]
        self.max_num_blocks = cfg["nblocks"]
        self.obj_range = 0.15                                   # tasks.py:18
        self.num_objs = cfg["nblocks"] + 2                      # fetch_env.py:75
        self.obj_colors = self._sample_colors()
        n = self.num_objs
        self.achieved_goal = -1 * np.ones([n, n])               # fetch_env.py:78
        self.has_succeeded = False
        self.difficulty = 0
        self.sim = BlockSim(self.eid, cfg["nblocks"])
        self.id2obj = [self._geom2objid(i) for i in range(self.sim.ngeom)]  # fetch_env.py:284
        self.initial_gripper_xpos = GRIP0.copy()
        # subclass constructors: fetch_env.py:340-348, 404-415, 561-563
        if self.kind == "touch":
            if cfg["curriculum"]:
                self.obj_range, self.obj_range_step, self.max_obj_range = 0.08, 0.025, 0.2
            else:
                self.max_obj_range, self.obj_range_step = self.obj_range, 0
        elif self.kind == "choose":
            if cfg["curriculum"]:
                self.obj_range, self.obj_range_step = 0.08, 0.025
                self.wrong_obj_range, self.wrong_obj_range_step, self.max_obj_range = 0.2, 0.02, 0.3
            else:
                self.wrong_obj_range, self.max_obj_range = 0, 0.2
        elif self.kind == "variation":
            self.obj_range, self.obj_range_step, self.max_obj_range = 0.08, 0.025, 0.2
        self.np_random = PhiloxRandomState(0)                   # robot_env.py:54
        self.global_random = PhiloxRandomState(1)               # the process-global np.random
        self.episode = 0
        self.goal = self._sample_goal()                         # robot_env.py:37
        self._elapsed_steps = None                              # TimeLimit
        self._max_episode_steps = MAX_EPISODE_STEPS

    # ---- RobotEnv ----
    def seed(self, seed=None):                                  # robot_env.py:53-55
        seed = 0 if seed is None else int(seed)
        for rs in (self.np_random, self.global_random):
            rs.seed_value, rs.episode, rs.draws = seed, 0, 0
        self._seed, self.episode = seed, 0
        return [seed]

    def step(self, action):                                     # robot_env.py:57-69
        action = np.asarray(action, f32)
        action = np.where(np.isnan(action), f32(0), action)     # flagged, not raised (BlockPhys)
        action = np.clip(action, f32(-1), f32(1))
        self._set_action(action)
        self.sim.step()
        self._step_callback()
        obs = self._get_obs()
        info = {"is_success": self._is_success(obs["achieved_goal"], self.goal)}
        reward = self.compute_reward(obs["achieved_goal"], self.goal, info)
        self._elapsed_steps += 1                                # TimeLimit.step [upstream]
        done = self._elapsed_steps >= self._max_episode_steps
        return obs, reward, done, info

    def reset(self):                                            # robot_env.py:71-82
        for rs in (self.np_random, self.global_random):
            rs.episode, rs.draws = self.episode, 0
        did = False
        while not did:
            did = self._reset_sim()
        self.goal = self._sample_goal().copy()
        self.episode += 1
        self._elapsed_steps = 0
        return self._get_obs()

    # ---- BlocksEnv ----
    def _geom2objid(self, i):                                   # fetch_env.py:106-117
        name = self.sim.geom_id2name(i)
        if name is not None:
            if "finger" in name:
                return 0
            if name == "table":
                return 1
            if "object" in name:
                return int(name[6:]) + 2
        return None

    def _check_goal(self):                                      # fetch_env.py:119-124
        return bool((self.achieved_goal == self.achieved_goal.T).all())

    def compute_reward(self, achieved_goal, goal, info):        # fetch_env.py:135-143
        d = np.sum(achieved_goal * goal, axis=-1)
        c = np.count_nonzero(goal, axis=-1)
        return -(d != c).astype(np.float32)

    def _step_callback(self):                                   # fetch_env.py:148-167
        n = self.num_objs
        sub = self.achieved_goal[:n, :n]
        sub[sub == 1] = 0
        for con in self.sim.contacts():
            o1, o2 = self.id2obj[con.geom1], self.id2obj[con.geom2]
            if o1 is not None and o2 is not None:
                self.achieved_goal[o1][o2] = 1
                self.achieved_goal[o2][o1] = 1

    def _set_action(self, action):                              # fetch_env.py:170-185
        assert action.shape == (4,)
        self.sim.set_action(action.copy())

    def _get_obs(self):                                         # fetch_env.py:187-228 / :567-621
        sim = self.sim
        grip_pos = sim.grip_pos()
        dt = f32(0.04)                                          # nsubsteps * timestep, :190
        grip_velp = (sim.grip_velp() * dt).astype(f32)
        gripper_state = sim.finger_qpos()
        gripper_vel = (sim.finger_qvel() * dt).astype(f32)
        num_blocks = self.num_objs - 2
        var = self.kind == "variation"
        parts = [[f32(num_blocks)]] if var else []
        parts += [grip_pos, gripper_state, grip_velp, gripper_vel]
        for i in range(num_blocks):
            pos = sim.obj_pos(i)
            c, s = sim.obj_yaw_cs(i)
            rot = np.array([-0.0, 0.0, bp_atan2(s, c)], f32)    # mat2euler of a pure yaw (roll = -arctan2(0, 1))
            velp = ((sim.obj_velp(i) * dt).astype(f32) - grip_velp).astype(f32)
            velr = (sim.obj_velr(i) * dt).astype(f32)
            parts += [pos, (pos - grip_pos).astype(f32), rot, velp, velr]
            if var:
                parts.append(one_hot_color(self.obj_colors[i + 2]))
        obs = np.concatenate(parts).astype(f32)
        if var:
            obs = np.concatenate([obs, np.zeros(19 * (self.max_num_blocks - num_blocks), f32)])
        assert self._check_goal()
        return {"observation": obs.copy(), "achieved_goal": self.achieved_goal.copy().ravel(),
                "desired_goal": self.goal.copy()}

    def _reset_sim(self, test=False):                           # fetch_env.py:247-255 / :646-679
        if self.kind == "variation":
            num_grey = self.np_random.randint(3)
            self.num_objs = 4 + num_grey
            self.sim = BlockSim(self.eid, 2 + num_grey)
            self.id2obj = [self._geom2objid(i) for i in range(self.sim.ngeom)]
            self.achieved_goal = -1 * np.ones([6, 6])
        self.sim.set_state_initial()
        self.obj_colors = self._sample_colors()
        self._randomize_objects(test)
        self.sim.forward()
        self.has_succeeded = False
        return True

    def _sample_goal(self):                                     # fetch_env.py:260-273 / :682-695
        C_ = self.obj_colors
        n = 6 if self.kind == "variation" else self.num_objs
        goal = []
        for i in range(n):
            for j in range(n):
                pair = {C_[i], C_[j]}
                goal.append(-1 if pair == {RED, BLUE} else (1 if pair == {GREEN, BLUE} else 0))
        return np.asarray(goal)

    def _is_success(self, achieved_goal, desired_goal):         # fetch_env.py:275-281
        r = self.compute_reward(self.achieved_goal.ravel(), self.goal, None)
        if r == 0:
            self.has_succeeded = True
        return self.has_succeeded

    def _sample_colors(self):                                   # fetch_env.py:323-326,360-363,434-441,632-639,772-775
        return {"gripper": [BLUE, GREY, GREEN], "touch": [GREY, GREY, GREEN, BLUE],
                "tower": [RED, GREEN, GREY, GREY, GREY, BLUE], "choose": [GREY, GREY, GREEN, BLUE, GREY],
                "variation": [GREY, GREY, GREEN, BLUE, GREY, GREY]}[self.kind]

    # The spawn samplers run in python floats / float64 arrays exactly like the reference (v1.3): only
    # set_joint_qpos narrows object_xpos into the sim's binary32 state.
    def _sample_from_table(self):                               # fetch_env.py:88-90
        return np.asarray([TABLE_X64 + self.np_random.uniform(-TABLE_W64, TABLE_W64),
                           TABLE_Y64 + self.np_random.uniform(-TABLE_H64, TABLE_H64)])

    def _around(self, base, lo, hi):                            # fetch_env.py:390-393 and siblings
        direction = self.global_random.normal(size=2)
        direction = direction / np.linalg.norm(direction)
        mag = self.global_random.uniform(lo, hi)
        return base + direction * mag

    def _randomize_objects(self, test=False):
        g0 = self.initial_gripper_xpos[:2].astype(np.float64)
        if self.kind in ("gripper", "tower"):                   # fetch_env.py:328-336, 777-787
            xy, it = g0, 0
            while np.linalg.norm(xy - g0) < 0.1 and it < MAX_SPAWN_ATTEMPTS:
                it += 1
                xy = g0 + self.np_random.uniform(-self.obj_range, self.obj_range, size=2)
            for i in range(self.max_num_blocks):
                self.sim.set_obj_xy(i, xy)
        elif self.kind == "touch":                              # fetch_env.py:370-399
            r = self.max_obj_range if test else self.obj_range
            p0 = g0 + self.np_random.uniform(-r / 2, r / 2, size=2)
            self.sim.set_obj_xy(0, p0)
            it = 0
            while True:
                xy = self._around(p0, MIN_BLOCK_DIST64, r)
                it += 1
                if not out_of_table64(xy) or it >= MAX_SPAWN_ATTEMPTS:
                    break
            self.sim.set_obj_xy(1, xy)
        elif self.kind == "choose":                             # fetch_env.py:448-517
            if test or self.challenge:                          # :452-463
                r, wrong_r = self.max_obj_range, 0
            else:
                r, wrong_r = self.obj_range, self.wrong_obj_range
            if self.challenge:
                max_wrong_r, min_r = 0.04, 0.15
            else:
                min_r, max_wrong_r = MIN_BLOCK_DIST64, self.max_obj_range
            blocks = self.obj_colors[2:5]
            blue, green = blocks.index(BLUE), blocks.index(GREEN)
            wrong = [i for i in range(3) if i not in (blue, green)][0]
            pb = self._blue(r)
            self.sim.set_obj_xy(blue, pb)
            pg = self._green(pb, r, min_r)
            self.sim.set_obj_xy(green, pg)
            centre = (pb + pg) / 2.0
            it = 0
            while True:
                xy = self._around(centre, wrong_r, max_wrong_r)
                it += 1
                bad = (out_of_table64(xy) or np.linalg.norm(xy - pb) < MIN_BLOCK_DIST64
                       or np.linalg.norm(xy - pg) < MIN_BLOCK_DIST64)
                if not bad or it >= MAX_SPAWN_ATTEMPTS:
                    break
            self.sim.set_obj_xy(wrong, xy)
        else:                                                   # variation, fetch_env.py:697-764
            num_blocks = self.num_objs - 2
            r = self.max_obj_range if test else self.obj_range
            blocks = self.obj_colors[2:2 + num_blocks]
            blue, green = blocks.index(BLUE), blocks.index(GREEN)
            pb = self._blue(r)
            self.sim.set_obj_xy(blue, pb)
            pg = self._green(pb, r)
            self.sim.set_obj_xy(green, pg)
            placed = [pb, pg]
            for i in range(num_blocks):
                if i in (blue, green):
                    continue
                it = 0
                while True:
                    xy = self._sample_from_table()
                    it += 1
                    for p in placed:
                        if np.linalg.norm(xy - p) < MIN_BLOCK_DIST64:
                            again = True
                            break
                    else:
                        again = out_of_table64(xy)
                    if not again or it >= MAX_SPAWN_ATTEMPTS:
                        break
                self.sim.set_obj_xy(i, xy)
                placed.append(xy)

    def _blue(self, r):                                         # fetch_env.py:475-480, 719-724
        g0 = self.initial_gripper_xpos[:2].astype(np.float64)
        it = 0
        while True:
            xy = g0 + self.np_random.uniform(-r / 2, r / 2, size=2)
            it += 1
            if not out_of_table64(xy) or it >= MAX_SPAWN_ATTEMPTS:
                return xy

    def _green(self, pb, r, min_r=MIN_BLOCK_DIST64):            # fetch_env.py:488-494, 732-738
        it = 0
        while True:
            xy = self._around(pb, min_r, r)
            it += 1
            if not out_of_table64(xy) or it >= MAX_SPAWN_ATTEMPTS:
                return xy

    # ---- curriculum ----
    def get_difficulty(self):                                   # fetch_env.py:96-97
        return self.difficulty

    def increase_difficulty(self):                              # fetch_env.py:351-358, 419-432, 623-630
        if self.kind in ("touch", "variation"):
            self.obj_range += self.obj_range_step
            if self.obj_range > self.max_obj_range:
                self.obj_range = self.max_obj_range
                return True
            self.difficulty += 1
            return False
        if self.kind == "choose":
            if not self.cfg["curriculum"]:
                raise AttributeError("obj_range_step")          # fetch_env.py:413-415 never sets it
            self.obj_range += self.obj_range_step
            self.wrong_obj_range -= self.wrong_obj_range_step
            if self.obj_range > self.max_obj_range:
                self.obj_range = self.max_obj_range
                if self.wrong_obj_range < 0:
                    self.wrong_obj_range = 0
                    return True
            elif self.wrong_obj_range < 0:
                self.wrong_obj_range = 0
            self.difficulty += 1
            return False
        raise NotImplementedError()                             # fetch_env.py:93-94

    def set_test(self):                                         # fetch_env.py:365-368, 443-446, 641-644
        if self.kind in ("gripper", "tower"):
            raise NotImplementedError()                         # fetch_env.py:100-101
        if self.kind == "variation":
            return self._get_obs()
        stale = self._get_obs()                                 # site positions are not refreshed (appendix A4)
        for rs in (self.np_random, self.global_random):
            rs.episode = self.episode - 1
        self._randomize_objects(True)
        self.goal = self._sample_goal().copy()
        stale["desired_goal"] = self.goal.copy()
        return stale

    @property
    def unwrapped(self):
        return self

    def random_action(self):
        """Philox action stream 2: counter (t, episode-1)."""
        w = philox4x32(self._elapsed_steps, (self.episode - 1) & M32, 2, 0, self._seed & M32, (self._seed >> 32) & M32)
        return np.array([f32(f32(f32(2.0) * u01(x)) - f32(1.0)) for x in w], f32)


def make(env_name, challenge=False):
    return BlocksEnvOracle(env_name, challenge=challenge)
