"""ctypes binding of the C oracle (oracle/blockphys_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from blockpuzzle_gym_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libblockphys_oracle.so")

ENV_IDS = [
    "GripperTouch-v0",
    "BlocksTouch-v0",
    "ToppleTower-v0",
    "BlocksTouchCurriculum-v0",
    "BlocksTouchChoose-v0",
    "BlocksTouchChooseCurriculum-v0",
    "BlocksTouchVariation-v0",
]
NBLOCKS = [1, 2, 4, 2, 3, 3, 4]
DIMO = [25, 40, 70, 40, 55, 55, 87]
DIMG = [9, 16, 36, 16, 25, 25, 36]
T = 50

# canonical per-env state record: identical to bp_env_state in include/blockpuzzle_b200.h
STATE_DTYPE = np.dtype(
    [
        ("grip_pos", "<f4", (3,)),
        ("grip_vel", "<f4", (3,)),
        ("finger_q", "<f4", (2,)),
        ("finger_qv", "<f4", (2,)),
        ("blk_pos", "<f4", (4, 3)),
        ("blk_cs", "<f4", (4, 2)),
        ("blk_vel", "<f4", (4, 3)),
        ("blk_w", "<f4", (4,)),
        ("ag", "i1", (36,)),
        ("num_objs", "<i4"),
        ("has_succeeded", "<i4"),
        ("t", "<i4"),
        ("episode", "<u4"),
        ("draws", "<u4", (2,)),
    ]
)


def build(force=False):
    """Compile the C oracle with the committed Makefile."""
    src = os.path.join(_HERE, "blockphys_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        vp = C.c_void_p
        L.bpo_sizeof_env.restype = C.c_int64
        L.bpo_sizeof_state.restype = C.c_int64
        L.bpo_philox4x32.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        L.bpo_u01.restype = C.c_float
        L.bpo_u01.argtypes = [C.c_uint32]
        L.bpo_u01_open.restype = C.c_float
        L.bpo_u01_open.argtypes = [C.c_uint32]
        L.bpo_log.restype = C.c_float
        L.bpo_log.argtypes = [C.c_float]
        L.bpo_atan2.restype = C.c_float
        L.bpo_atan2.argtypes = [C.c_float, C.c_float]
        L.bpo_sincos2pi.argtypes = [C.c_float, fp, fp]
        L.bpo_normal2.argtypes = [C.c_uint32, C.c_uint32, fp, fp]
        L.bpo_env_init.argtypes = [vp, C.c_int]
        L.bpo_env_seed.argtypes = [vp, C.c_uint64]
        L.bpo_env_reset.argtypes = [vp, vp, vp, vp]
        L.bpo_env_step.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.bpo_env_step.restype = C.c_int
        L.bpo_env_set_test.argtypes = [vp, vp, vp, vp]
        L.bpo_env_set_test.restype = C.c_int
        L.bpo_env_increase_difficulty.argtypes = [vp]
        L.bpo_env_increase_difficulty.restype = C.c_int
        L.bpo_env_get_obs.argtypes = [vp, vp, vp, vp]
        L.bpo_env_get_difficulty.argtypes = [vp]
        L.bpo_env_get_difficulty.restype = C.c_int
        L.bpo_env_get_obj_range.argtypes = [vp]
        L.bpo_env_get_obj_range.restype = C.c_double
        L.bpo_env_set_ranges.argtypes = [vp, C.c_double, C.c_double]
        L.bpo_env_set_challenge.argtypes = [vp, C.c_int]
        L.bpo_env_set_challenge.restype = C.c_int
        L.bpo_env_get_state.argtypes = [vp, vp]
        L.bpo_env_set_state.argtypes = [vp, vp]
        L.bpo_env_random_action.argtypes = [vp, vp]
        L.bpo_compute_reward.argtypes = [vp, vp, C.c_int64, C.c_int, vp]
        L.bpo_her_relabel.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_float,
                                      C.c_uint64, C.c_int64, vp, vp, vp, vp, vp, vp]
        L.bpo_vec_init.argtypes = [vp, C.c_int64, C.c_int, C.c_uint64, C.c_uint64]
        L.bpo_vec_reset.argtypes = [vp, C.c_int64, vp, vp, vp]
        L.bpo_vec_step.argtypes = [vp, C.c_int64, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
        L.bpo_vec_step_random.argtypes = [vp, C.c_int64, C.c_int, C.c_int, vp]
        L.bpo_sim_init.argtypes = [vp, C.c_int]
        L.bpo_sim_set_action.argtypes = [vp, vp]
        L.bpo_sim_set_targets.argtypes = [vp, vp, vp]
        L.bpo_sim_step.argtypes = [vp]
        assert L.bpo_sizeof_state() == STATE_DTYPE.itemsize, (L.bpo_sizeof_state(), STATE_DTYPE.itemsize)
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox4x32(c0, c1, c2, c3, k0, k1):
    out = (C.c_uint32 * 4)()
    lib().bpo_philox4x32(c0, c1, c2, c3, k0, k1, out)
    return [int(x) for x in out]


class OracleVecEnv:
    """n independent C-oracle envs; env i is seeded seed + 1000*(offset+i) (rollout.py:206-210)."""

    def __init__(self, env_name, num_envs, seed=0, env_index_offset=0):
        self.L = lib()
        self.env_id = ENV_IDS.index(env_name) if isinstance(env_name, str) else int(env_name)
        self.n = int(num_envs)
        self.dimo, self.dimg = DIMO[self.env_id], DIMG[self.env_id]
        self._esz = int(self.L.bpo_sizeof_env())
        self._buf = np.zeros(self.n * self._esz, dtype=np.uint8)
        self.L.bpo_vec_init(_p(self._buf), self.n, self.env_id, seed, env_index_offset)
        self.stats = np.zeros(4, dtype=np.int32)

    def _env_ptr(self, i):
        return C.c_void_p(self._buf.ctypes.data + i * self._esz)

    def reset(self):
        obs = np.zeros((self.n, self.dimo), np.float32)
        ag = np.zeros((self.n, self.dimg), np.float32)
        g = np.zeros((self.n, self.dimg), np.float32)
        self.L.bpo_vec_reset(_p(self._buf), self.n, _p(obs), _p(ag), _p(g))
        return obs, ag, g

    def reset_one(self, i):
        obs = np.zeros(self.dimo, np.float32)
        ag = np.zeros(self.dimg, np.float32)
        g = np.zeros(self.dimg, np.float32)
        self.L.bpo_env_reset(self._env_ptr(i), _p(obs), _p(ag), _p(g))
        return obs, ag, g

    def step(self, actions, auto_reset=False):
        actions = np.ascontiguousarray(actions, np.float32).reshape(self.n, 4)
        obs = np.zeros((self.n, self.dimo), np.float32)
        ag = np.zeros((self.n, self.dimg), np.float32)
        r = np.zeros(self.n, np.float32)
        succ = np.zeros(self.n, np.float32)
        robs = np.zeros((self.n, self.dimo), np.float32)
        rag = np.zeros((self.n, self.dimg), np.float32)
        self.L.bpo_vec_step(_p(self._buf), self.n, _p(actions), int(auto_reset), _p(obs), _p(ag),
                            _p(r), _p(succ), _p(robs), _p(rag), _p(self.stats))
        return obs, ag, r, succ, robs, rag

    def set_test(self):
        obs = np.zeros((self.n, self.dimo), np.float32)
        ag = np.zeros((self.n, self.dimg), np.float32)
        g = np.zeros((self.n, self.dimg), np.float32)
        for i in range(self.n):
            rc = self.L.bpo_env_set_test(self._env_ptr(i), _p(obs[i]), _p(ag[i]), _p(g[i]))
            if rc != 0:
                raise NotImplementedError()
        return obs, ag, g

    def increase_difficulty(self):
        rc = 0
        for i in range(self.n):
            rc = self.L.bpo_env_increase_difficulty(self._env_ptr(i))
            if rc == -2:
                raise AttributeError("'BlocksTouchChooseEnv' object has no attribute 'obj_range_step'")  # fetch_env.py:413-420
            if rc < 0:
                raise NotImplementedError()
        return bool(rc)

    def get_difficulty(self):
        return int(self.L.bpo_env_get_difficulty(self._env_ptr(0)))

    def get_obj_range(self):
        return float(self.L.bpo_env_get_obj_range(self._env_ptr(0)))

    def set_ranges(self, obj_range, wrong_obj_range=0.0):
        for i in range(self.n):
            self.L.bpo_env_set_ranges(self._env_ptr(i), float(obj_range), float(wrong_obj_range))

    def set_challenge(self, challenge=True):
        """BlocksTouchChooseEnv(challenge=True) (fetch_env.py:403,416): the two Choose ids only."""
        for i in range(self.n):
            if self.L.bpo_env_set_challenge(self._env_ptr(i), int(bool(challenge))) != 0:
                raise TypeError("challenge is an argument of BlocksTouchChooseEnv only")

    def random_actions(self):
        a = np.zeros((self.n, 4), np.float32)
        for i in range(self.n):
            self.L.bpo_env_random_action(self._env_ptr(i), _p(a[i]))
        return a

    def get_state(self):
        st = np.zeros(self.n, dtype=STATE_DTYPE)
        for i in range(self.n):
            self.L.bpo_env_get_state(self._env_ptr(i), C.c_void_p(st.ctypes.data + i * STATE_DTYPE.itemsize))
        return st

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=STATE_DTYPE)
        for i in range(self.n):
            self.L.bpo_env_set_state(self._env_ptr(i), C.c_void_p(st.ctypes.data + i * STATE_DTYPE.itemsize))

    def run_random(self, steps, lo=0, hi=None):
        """Throughput loop over envs [lo, hi): Philox actions, auto-reset; returns [episodes, successes, steps]."""
        hi = self.n if hi is None else hi
        stats = np.zeros(3, dtype=np.int64)
        self.L.bpo_vec_step_random(C.c_void_p(self._buf.ctypes.data + lo * self._esz), hi - lo, steps, 0, _p(stats))
        return stats


def compute_reward(ag, g):
    ag = np.ascontiguousarray(ag, np.float32)
    g = np.ascontiguousarray(g, np.float32)
    dimg = ag.shape[-1]
    n = ag.size // dimg
    r = np.zeros(ag.shape[:-1], np.float32)
    lib().bpo_compute_reward(_p(ag), _p(g), n, dimg, _p(r))
    return r


def her_relabel(ep_ag, ep_g, n, future_p, seed, index_offset=0):
    ep_ag = np.ascontiguousarray(ep_ag, np.float32)
    ep_g = np.ascontiguousarray(ep_g, np.float32)
    B, T1, dimg = ep_ag.shape
    Tn = T1 - 1
    assert ep_g.shape == (B, Tn, dimg)
    e = np.zeros(n, np.int32)
    t = np.zeros(n, np.int32)
    ft = np.zeros(n, np.int32)
    ag2 = np.zeros((n, dimg), np.float32)
    g = np.zeros((n, dimg), np.float32)
    r = np.zeros(n, np.float32)
    lib().bpo_her_relabel(_p(ep_ag), _p(ep_g), B, Tn, dimg, n, float(future_p), seed, index_offset,
                          _p(e), _p(t), _p(ft), _p(ag2), _p(g), _p(r))
    return dict(ep_idx=e, t=t, future_t=ft, ag_2=ag2, g=g, r=r)
