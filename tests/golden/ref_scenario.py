"""The replay scenario behind tests/golden/ref_<id>.npz (shared by the fixture maker and the parity tests).

A scenario is a fixed programme of env calls -- reset / step / set_test / increase_difficulty -- run on
E envs of one registered id (env i seeded seed + 1000*i, rollout.py:206-210).  `record()` runs it on a
driver and returns the trace; make_ref_golden.py records it on the UNMODIFIED reference (oracle/refharness)
and commits the trace; the tests replay the committed actions on the C oracle (CPU) and on the CUDA
library (GPU) and `compare()` the traces: integer state, draw counters and the binary32 sim state bit for
bit, float64 observations within FLOAT_TOL.

Episodes of the programme (T = 50 steps each, __init__.py:10):
  0  Philox actions alternating with host actions outside [-1, 1]   (clip, robot_env.py:58)
  1  scripted push: cube 0 straight into cube 1 / the gripper onto the cube / into the tower
  2  scripted push of cube 0 off the table edge (+x), fingers closed
  3  scripted grasp: open fingers, descend over cube 0, close, lift, drag
  then set_test() + 20 steps, then every curriculum level: increase_difficulty(), reset(), 5 steps.
"""
import numpy as np

OP_RESET, OP_STEP, OP_SET_TEST, OP_INC = 0, 1, 2, 3
RC_NOT_IMPLEMENTED, RC_ATTRIBUTE_ERROR = -1, -2
T = 50
NUM_ENVS = 3
SEED = 20260
# float64 reference observation vs the binary32 arithmetic of the oracle / kernels (north_star: 1e-6 relative;
# positions are O(1) so the absolute floor equals the relative bound on the operands of the differences)
FLOAT_TOL = dict(rtol=1e-6, atol=1e-6)

ENV_IDS = ["GripperTouch-v0", "BlocksTouch-v0", "ToppleTower-v0", "BlocksTouchCurriculum-v0", "BlocksTouchChoose-v0",
           "BlocksTouchChooseCurriculum-v0", "BlocksTouchVariation-v0"]


# ------------------------------------------------------------------------------------------------ policies
def _block_pos(name, obs, i):
    if name == "BlocksTouchVariation-v0":
        return obs[11 + 19 * i: 14 + 19 * i]
    return obs[10 + 15 * i: 13 + 15 * i]


def _grip(name, obs):
    return obs[1:4] if name == "BlocksTouchVariation-v0" else obs[0:3]


class Waypoints(object):
    """Proportional waypoint follower: [(xy, z, a3, steps)], straight-line approach in xy."""

    def __init__(self, plan):
        self.plan = list(plan)
        self.k = 0
        self.left = self.plan[0][3] if self.plan else 0

    def act(self, grip):
        a = np.zeros(4, np.float32)
        if self.k >= len(self.plan):
            return a
        xy, z, a3, _ = self.plan[self.k]
        dxy = (np.asarray(xy) - grip[:2]) / 0.05
        a[:2] = dxy / max(1.0, np.abs(dxy).max())
        a[2] = np.clip((z - grip[2]) / 0.05, -1, 1)
        a[3] = a3
        self.left -= 1
        if self.left <= 0:
            self.k += 1
            self.left = self.plan[self.k][3] if self.k < len(self.plan) else 0
        return a


def plan_push(name, obs):
    b0 = _block_pos(name, obs, 0)[:2].copy()
    if name == "GripperTouch-v0":
        return [(b0, 0.56, -1, 5), (b0, 0.47, -1, 6), (b0 + [0.1, 0.0], 0.47, -1, 39)]
    if name == "ToppleTower-v0":
        side = b0 - [0.1, 0.0]
        return [(side, 0.56, -1, 5), (side, 0.50, -1, 4), (b0 + [0.15, 0.02], 0.50, -1, 41)]
    b1 = _block_pos(name, obs, 1)[:2].copy()
    d = (b1 - b0) / np.linalg.norm(b1 - b0)
    return [(b0 - d * 0.07, 0.55, -1, 6), (b0 - d * 0.07, 0.48, -1, 4), (b1 + d * 0.1, 0.48, -1, 40)]


def plan_off_table(name, obs):
    b0 = _block_pos(name, obs, 0)[:2].copy()
    back = b0 - [0.07, 0.0]
    return [(back, 0.55, -1, 6), (back, 0.48, -1, 4), (np.array([1.62, b0[1]]), 0.48, -1, 40)]


def plan_grasp(name, obs):
    b0 = _block_pos(name, obs, 0).copy()
    top = b0[2] if name != "ToppleTower-v0" else 0.56
    return [(b0[:2], 0.58, 1, 6), (b0[:2], top, 1, 6), (b0[:2], top, -1, 8), (b0[:2] + [0.0, 0.08], top + 0.05, -1, 12),
            (b0[:2] + [0.1, -0.1], top, 1, 18)]


# ------------------------------------------------------------------------------------------------ recording
def _event(op, E, dimo, dimg):
    from oracle import coracle
    return dict(op=op, action=np.zeros((E, 4), np.float32), obs=np.zeros((E, dimo), np.float64), ag=np.zeros((E, dimg), np.int8),
                g=np.zeros((E, dimg), np.int8), r=np.zeros(E, np.float32), succ=np.zeros(E, np.uint8), done=np.zeros(E, np.uint8),
                state=np.zeros(E, coracle.STATE_DTYPE), rc=0, difficulty=0)


def _call(driver, ev, actions=None):
    """Run one event on a driver and fill in what it returned."""
    op = ev["op"]
    if op == OP_RESET:
        ev["obs"], ev["ag"], ev["g"] = driver.reset()
    elif op == OP_STEP:
        ev["action"] = np.asarray(actions, np.float32)
        ev["obs"], ev["ag"], ev["r"], ev["succ"], ev["done"] = driver.step(ev["action"])
    elif op == OP_SET_TEST:
        try:
            ev["obs"], ev["ag"], ev["g"] = driver.set_test()
        except NotImplementedError:
            ev["rc"] = RC_NOT_IMPLEMENTED
    elif op == OP_INC:
        try:
            ev["rc"] = int(bool(driver.increase_difficulty()))
        except NotImplementedError:
            ev["rc"] = RC_NOT_IMPLEMENTED
        except AttributeError:
            ev["rc"] = RC_ATTRIBUTE_ERROR
    ev["difficulty"] = driver.get_difficulty()
    ev["state"] = driver.state()
    ev["obs"] = np.asarray(ev["obs"], np.float64)
    ev["ag"] = np.asarray(ev["ag"]).astype(np.int8)
    ev["g"] = np.asarray(ev["g"]).astype(np.int8)
    ev["succ"] = (np.asarray(ev["succ"]) != 0).astype(np.uint8)
    ev["done"] = (np.asarray(ev["done"]) != 0).astype(np.uint8)
    ev["r"] = np.asarray(ev["r"], np.float32)
    return ev


def record(name, driver, philox_actions, num_envs=NUM_ENVS):
    """Run the programme on `driver`, choosing actions from its observations.  `philox_actions(t_global)` supplies
    the seeded random actions (any deterministic source; they are stored in the trace)."""
    E, dimo, dimg = num_envs, driver.dimo, driver.dimg
    trace = []
    rng = np.random.RandomState(99)
    n_rand = [0]

    def ev(op, actions=None):
        e = _call(driver, _event(op, E, dimo, dimg), actions)
        trace.append(e)
        return e

    def random_action(k):
        n_rand[0] += 1
        if k % 2 == 0:
            return philox_actions(n_rand[0])
        return rng.uniform(-1.5, 1.5, size=(E, 4)).astype(np.float32)

    # episode 0: random + out-of-range actions
    ev(OP_RESET)
    for k in range(T):
        ev(OP_STEP, random_action(k))
    # episodes 1-3: scripted contact scenarios, one controller per env
    for plan in (plan_push, plan_off_table, plan_grasp):
        e0 = ev(OP_RESET)
        ctl = [Waypoints(plan(name, e0["obs"][i])) for i in range(E)]
        obs = e0["obs"]
        for k in range(T):
            a = np.stack([ctl[i].act(_grip(name, obs[i])) for i in range(E)])
            obs = ev(OP_STEP, a)["obs"]
    # set_test (raises for the ids that do not override it, fetch_env.py:99-101) + 20 steps
    ev(OP_SET_TEST)
    for k in range(20):
        ev(OP_STEP, random_action(k))
    # the curriculum: every level, then two calls beyond the maximum
    beyond = 0
    for level in range(16):
        e = ev(OP_INC)
        if e["rc"] < 0:
            break
        ev(OP_RESET)
        for k in range(5):
            ev(OP_STEP, random_action(2 * k))
        if e["rc"] == 1 or level >= 13 or (name in ("BlocksTouch-v0",) and level >= 1):
            beyond += 1
            if beyond == 2:
                break
    return trace


def replay(driver, trace, num_envs=NUM_ENVS):
    """Run the committed programme (ops + actions of `trace`) on another driver."""
    out = []
    for e in trace:
        out.append(_call(driver, _event(int(e["op"]), num_envs, driver.dimo, driver.dimg), e["action"]))
    return out


FIELDS = ("op", "action", "obs", "ag", "g", "r", "succ", "done", "state", "rc", "difficulty")


def pack(trace):
    return {k: np.stack([np.asarray(e[k]) for e in trace]) for k in FIELDS}


def unpack(z):
    n = len(z["op"])
    return [{k: z[k][i] for k in FIELDS} for i in range(n)]


def compare(got, want, who="", float_tol=FLOAT_TOL, exact_floats=False):
    """`got` against the reference trace `want`; raises AssertionError naming the first differing event."""
    assert len(got) == len(want), (len(got), len(want))
    for n, (a, b) in enumerate(zip(got, want)):
        tag = "%s event %d (op %d)" % (who, n, int(b["op"]))
        assert int(a["op"]) == int(b["op"]), tag
        assert int(a["rc"]) == int(b["rc"]), "%s: rc %d != %d" % (tag, int(a["rc"]), int(b["rc"]))
        assert int(a["difficulty"]) == int(b["difficulty"]), tag + ": difficulty"
        sa, sb = np.asarray(a["state"]), np.asarray(b["state"])
        if sa.tobytes() != sb.tobytes():
            for f in sa.dtype.names:
                assert np.array_equal(sa[f].view(np.uint8), sb[f].view(np.uint8)), "%s: state field %s differs\n%r\n%r" % (tag, f, sa[f], sb[f])
        if int(b["rc"]) < 0:
            continue
        assert np.array_equal(a["ag"], b["ag"]), tag + ": touch matrix"
        assert np.array_equal(a["g"], b["g"]), tag + ": goal"
        assert np.array_equal(np.asarray(a["r"]).view(np.uint32), np.asarray(b["r"]).view(np.uint32)), tag + ": reward (incl. the sign of -0.0)"
        assert np.array_equal(a["succ"], b["succ"]), tag + ": is_success latch"
        assert np.array_equal(a["done"], b["done"]), tag + ": done"
        if exact_floats:
            assert np.array_equal(np.asarray(a["obs"], np.float32).view(np.uint32), np.asarray(b["obs"], np.float32).view(np.uint32)), tag + ": observation bits"
        else:
            err = np.abs(np.asarray(a["obs"], np.float64) - b["obs"])
            lim = float_tol["atol"] + float_tol["rtol"] * np.abs(b["obs"])
            assert np.all(err <= lim), "%s: observation off by %.3g" % (tag, float((err - lim).max()))


# ------------------------------------------------------------------------------------------------ drivers
class RefDriver(object):
    """E unmodified reference envs (gym.make under the stub packages, oracle/refharness)."""

    def __init__(self, name, num_envs=NUM_ENVS, seed=SEED):
        from oracle import refharness as rh
        self.rh = rh
        self.envs = [rh.make(name, seed=seed + 1000 * i) for i in range(num_envs)]
        o = self.envs[0].unwrapped._get_obs()
        self.dimo, self.dimg = o["observation"].size, o["achieved_goal"].size
        self.goal_dtype = str(np.asarray(o["desired_goal"]).dtype)

    @staticmethod
    def _stack(obs):
        return (np.stack([o["observation"] for o in obs]), np.stack([o["achieved_goal"] for o in obs]),
                np.stack([o["desired_goal"] for o in obs]))

    def reset(self):
        return self._stack([e.reset() for e in self.envs])

    def step(self, a):
        res = [e.step(a[i]) for i, e in enumerate(self.envs)]
        o, ag, _ = self._stack([r[0] for r in res])
        return (o, ag, np.array([r[1] for r in res], np.float32), np.array([r[3]["is_success"] for r in res]),
                np.array([r[2] for r in res]))

    def set_test(self):
        return self._stack([e.unwrapped.set_test() for e in self.envs])

    def increase_difficulty(self):
        rcs = [e.unwrapped.increase_difficulty() for e in self.envs]
        assert len(set(rcs)) == 1
        return rcs[0]

    def get_difficulty(self):
        return int(self.envs[0].unwrapped.get_difficulty())

    def state(self):
        return np.stack([self.rh.state_record(e) for e in self.envs])


class OracleDriver(object):
    """The C restatement (oracle/blockphys_oracle.c)."""

    def __init__(self, name, num_envs=NUM_ENVS, seed=SEED):
        from oracle import coracle
        self.env = coracle.OracleVecEnv(name, num_envs, seed=seed)
        self.dimo, self.dimg = self.env.dimo, self.env.dimg
        self.L = self.env.L

    def reset(self):
        return self.env.reset()

    def step(self, a):
        o, ag, r, s, _, _ = self.env.step(a, auto_reset=False)
        return o, ag, r, s, self.env.get_state()["t"] >= T

    def set_test(self):
        return self.env.set_test()

    def increase_difficulty(self):
        return self.env.increase_difficulty()

    def get_difficulty(self):
        return self.env.get_difficulty()

    def state(self):
        return self.env.get_state()
