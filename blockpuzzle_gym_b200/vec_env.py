"""Host-side mirror of the reference's env interface over the CUDA C-ABI.

`VecBlocksEnv` is the batched environment (B envs on one GPU, torch CUDA tensors
in and out); `GymBlocksEnv` is a single-env view with the exact gym GoalEnv
surface the reference's callers use, so `RolloutStudent` (gym_blocks/rollout.py)
runs unchanged on it:

    reset()                    robot_env.py:71-82
    step(u)                    robot_env.py:57-69 under TimeLimit(50)  (__init__.py:10)
    compute_reward(ag, g, info)  fetch_env.py:135-143, called as env.compute_reward(
                               achieved_goal=, desired_goal=, info=) at config.py:110-111
    seed(s)                    robot_env.py:53-55, rollout.py:206-210
    unwrapped.set_test() / increase_difficulty() / get_difficulty()
                               fetch_env.py:93-101, 351-368, 419-446, 623-644
    _max_episode_steps         read at config.py:79-80
    action_space / observation_space   robot_env.py:39-44
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ENV_IDS, check

MAX_EPISODE_STEPS = 50  # __init__.py:10

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
assert STATE_DTYPE.itemsize == _lib.STATE_BYTES


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Box:
    """Minimal stand-in for gym.spaces.Box (robot_env.py:39)."""

    def __init__(self, low, high, shape, dtype="float32"):
        self.low = np.full(shape, low, dtype=dtype)
        self.high = np.full(shape, high, dtype=dtype)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self._rng = np.random.RandomState()

    def seed(self, s=None):
        self._rng = np.random.RandomState(s)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class DictSpace:
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]


class VecBlocksEnv:
    """B independent gym_blocks envs of one registered id, resident on one B200."""

    def __init__(self, env_name, num_envs, device=None, seed=0, env_index_offset=0, challenge=False):
        self.L = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.BlockPuzzleError("blockpuzzle_gym_b200 needs a CUDA device (no CPU fallback)")
        if env_name not in ENV_IDS:
            raise KeyError(f"No registered env with id: {env_name}")
        self.env_name = env_name
        self.env_id = ENV_IDS.index(env_name)
        self.num_envs = int(num_envs)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.dimo, self.dimg, self.nblocks = _lib.env_dims(self.env_id)
        self.dimu = 4  # fetch_env.py:85 n_actions
        self._max_episode_steps = MAX_EPISODE_STEPS
        self.env_index_offset = int(env_index_offset)
        h = C.c_void_p()
        check(self.L.bp_create(self.env_id, self.num_envs, self.device.index, self.env_index_offset, C.byref(h)))
        self._h = h
        self.action_space = Box(-1.0, 1.0, (4,), "float32")
        self.observation_space = DictSpace(dict(
            desired_goal=Box(-np.inf, np.inf, (self.dimg,)),
            achieved_goal=Box(-np.inf, np.inf, (self.dimg,)),
            observation=Box(-np.inf, np.inf, (self.dimo,)),
        ))
        sp = C.c_void_p()
        check(self.L.bp_stats_ptr(self._h, C.byref(sp)))
        self._stats_ptr = sp.value
        self._goal = None
        # `challenge`: the constructor argument of BlocksTouchChooseEnv (fetch_env.py:403,416); any other class of the
        # reference would raise TypeError on the unexpected keyword
        self.challenge = bool(challenge)
        if self.challenge:
            if "Choose" not in env_name:
                raise TypeError("__init__() got an unexpected keyword argument 'challenge'")
            self.set_option("challenge", 1)
        self.seed(seed)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.L.bp_destroy(self._h)
                self._h = None
        except Exception:
            pass

    close = __del__

    # ------------------------------------------------------------------ tensors
    def _empty(self, *shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ gym-like API (batched)
    def seed(self, seed=None):
        seed = 0 if seed is None else int(seed)
        check(self.L.bp_seed(self._h, seed & 0xFFFFFFFFFFFFFFFF, _stream(self.device)))
        return [seed]

    def reset(self, mask=None):
        """Reset all envs (or those with mask != 0).  Returns a dict of [B, dim] CUDA tensors."""
        obs, ag, g = self._empty(self.num_envs, self.dimo), self._empty(self.num_envs, self.dimg), self._empty(self.num_envs, self.dimg)
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            obs.zero_(); ag.zero_(); g.zero_()
        check(self.L.bp_reset(self._h, _ptr(mask), _ptr(obs), _ptr(ag), _ptr(g), _stream(self.device)))
        self._goal = g if mask is None else self._goal
        return dict(observation=obs, achieved_goal=ag, desired_goal=g)

    def step(self, actions):
        """One env step for every env.  actions: [B, 4] CUDA float tensor."""
        out = self.step_fused(actions.reshape(1, self.num_envs, 4), auto_reset=False, want_done=True)
        if self._goal is None:
            self._goal = self.goal()
        obs = dict(observation=out["observation"][0], achieved_goal=out["achieved_goal"][0], desired_goal=self._goal)
        return obs, out["reward"][0], out["done"][0].bool(), dict(is_success=out["is_success"][0])

    def step_fused(self, actions=None, K=None, auto_reset=True, out=None, want_done=False, want_reset_obs=False,
                   want_actions=False, outputs=("observation", "achieved_goal", "reward", "is_success")):
        """K fused steps in one launch.

        actions: [K, B, 4] float32 CUDA tensor, or None to draw them in-kernel from the
        env's Philox action stream (the replay harness).  Returns a dict of time-major
        tensors: observation [K,B,dimo], achieved_goal [K,B,dimg], reward [K,B], is_success [K,B].
        """
        B = self.num_envs
        if actions is not None:
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
            K = actions.shape[0]
            assert actions.shape == (K, B, 4), actions.shape
        assert K is not None and K > 0
        out = {} if out is None else out

        def buf(name, *shape, dtype=torch.float32):
            t = out.get(name)
            # a reused buffer is handed to the kernel as a raw pointer: it must be exactly what the kernel will write
            if (t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != self.device
                    or not t.is_contiguous()):
                out[name] = self._empty(*shape, dtype=dtype)
            return out[name]

        obs = buf("observation", K, B, self.dimo) if "observation" in outputs else None
        ag = buf("achieved_goal", K, B, self.dimg) if "achieved_goal" in outputs else None
        rew = buf("reward", K, B) if "reward" in outputs else None
        suc = buf("is_success", K, B) if "is_success" in outputs else None
        done = buf("done", K, B, dtype=torch.uint8) if want_done else None
        robs = buf("reset_observation", B, self.dimo) if want_reset_obs else None
        rag = buf("reset_achieved_goal", B, self.dimg) if want_reset_obs else None
        aout = buf("actions", K, B, 4) if want_actions else None
        check(self.L.bp_step(self._h, _ptr(actions), K, _ptr(obs), _ptr(ag), _ptr(rew), _ptr(suc), _ptr(done),
                             int(bool(auto_reset)), _ptr(robs), _ptr(rag), _ptr(aout), _stream(self.device)))
        return out

    def step_host(self, actions, auto_reset=True, out=None):
        """End-to-end path with HOST arrays: actions [K,B,4] numpy float32 (ideally pinned)."""
        a = np.ascontiguousarray(actions, dtype=np.float32)
        K, B = a.shape[0], self.num_envs
        assert a.shape == (K, B, 4)
        if out is None:
            out = dict(observation=np.empty((K, B, self.dimo), np.float32), achieved_goal=np.empty((K, B, self.dimg), np.float32),
                       reward=np.empty((K, B), np.float32), is_success=np.empty((K, B), np.float32))
        p = lambda x: C.c_void_p(x.ctypes.data)
        for k, shp in (("observation", (K, B, self.dimo)), ("achieved_goal", (K, B, self.dimg)), ("reward", (K, B)), ("is_success", (K, B))):
            x = out[k]
            assert x.shape == shp and x.dtype == np.float32 and x.flags.c_contiguous, (k, x.shape, x.dtype)
        check(self.L.bp_step_host(self._h, p(a), K, p(out["observation"]), p(out["achieved_goal"]), p(out["reward"]),
                                  p(out["is_success"]), int(bool(auto_reset)), _stream(self.device)))
        return out

    def step_host_ptrs(self, a_ptr, K, obs_ptr, ag_ptr, r_ptr, s_ptr, auto_reset=True):
        """bp_step_host on raw host addresses (e.g. pinned torch tensors' data_ptr())."""
        check(self.L.bp_step_host(self._h, C.c_void_p(a_ptr), K, C.c_void_p(obs_ptr), C.c_void_p(ag_ptr),
                                  C.c_void_p(r_ptr), C.c_void_p(s_ptr), int(bool(auto_reset)), _stream(self.device)))

    def generate_rollouts(self, actions=None, test=False):
        """RolloutStudent.generate_rollouts (rollout.py:75-172) for open-loop actions [T, B, 4] (None: the env's
        Philox action stream): reset (+ set_test), T = 50 steps, and the episode returned batch-major exactly as
        convert_episode_to_batch_major (util.py:118-128) lays it out: o [B,T+1,dimo], u [B,T,4], g [B,T,dimg],
        ag [B,T+1,dimg], info_is_success [B,T,1] (+ r [B,T]).  One reset launch + one fused step launch."""
        B, T = self.num_envs, MAX_EPISODE_STEPS
        if actions is not None:
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
            assert actions.shape == (T, B, 4), actions.shape
        ep = self._episode_tensors()
        check(self.L.bp_rollout(self._h, _ptr(actions), int(bool(test)), _ptr(ep["o"]), _ptr(ep["ag"]), _ptr(ep["g"]),
                                _ptr(ep["u"]), _ptr(ep["info_is_success"]), _ptr(ep["r"]), _stream(self.device)))
        return ep

    def _episode_tensors(self):
        B, T = self.num_envs, MAX_EPISODE_STEPS
        return dict(o=self._empty(B, T + 1, self.dimo), u=self._empty(B, T, 4), g=self._empty(B, T, self.dimg),
                    ag=self._empty(B, T + 1, self.dimg), info_is_success=self._empty(B, T, 1), r=self._empty(B, T))

    def rollout_begin(self, ep, test=False, g0=None):
        """reset_all_rollouts (rollout.py:48-64) into slot 0 of the episode tensors `ep` (see generate_rollouts)."""
        check(self.L.bp_rollout_begin(self._h, int(bool(test)), _ptr(ep["o"]), _ptr(ep["ag"]), _ptr(g0), _stream(self.device)))

    def rollout_step(self, ep, t, actions):
        """One step of every env on `actions` [B, 4] (the policy's u_t): slot t + 1 of o / ag, slot t of g / u / ..."""
        assert actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous() and tuple(actions.shape) == (self.num_envs, 4)
        check(self.L.bp_rollout_step(self._h, int(t), _ptr(actions), _ptr(ep["o"]), _ptr(ep["ag"]), _ptr(ep["g"]), _ptr(ep["u"]),
                                     _ptr(ep["info_is_success"]), _ptr(ep["r"]), _stream(self.device)))

    def collect_rollouts(self, policy, test=False, graph=False):
        """RolloutStudent.generate_rollouts (rollout.py:75-172) CLOSED-LOOP for the whole batch: `policy(o, ag, g)`
        -- the stand-in for policy.get_actions (rollout.py:92-97) -- maps CUDA tensors o [B, dimo], ag [B, dimg],
        g [B, dimg] to actions [B, 4] at every one of the T = 50 steps; the episode comes back batch-major exactly as
        convert_episode_to_batch_major (util.py:118-128) lays it out.  One reset launch + one step launch per step,
        no host synchronisation; graph=True captures the 50-step loop (policy included) in a CUDA graph on first use
        and replays it afterwards (the policy must then be capturable: static shapes, no host syncs)."""
        T = MAX_EPISODE_STEPS
        if graph:
            key = (id(policy), bool(test))
            cache = self.__dict__.setdefault("_graphs", {})
            if key not in cache:
                ep = self._episode_tensors()
                g0 = self._empty(self.num_envs, self.dimg)
                state = self.get_state()                                   # warm-up and capture must not consume episodes
                side = torch.cuda.Stream(self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    self._collect(policy, ep, g0, test)                    # warm-up (lazy module init, cudaFuncSetAttribute)
                torch.cuda.current_stream(self.device).wait_stream(side)
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    self._collect(policy, ep, g0, test)
                self.set_state(state)
                cache[key] = (gr, ep)
            gr, ep = cache[key]
            gr.replay()
            return {k: v.clone() for k, v in ep.items()}
        ep = self._episode_tensors()
        self._collect(policy, ep, self._empty(self.num_envs, self.dimg), test)
        return ep

    def _collect(self, policy, ep, g0, test):
        self.rollout_begin(ep, test=test, g0=g0)
        for t in range(MAX_EPISODE_STEPS):
            u = policy(ep["o"][:, t], ep["ag"][:, t], g0)
            self.rollout_step(ep, t, u.to(torch.float32).contiguous())

    def goal(self):
        """desired_goal rows [B, dimg]: fixed per env id (colours are fixed lists, fetch_env.py:260-273)."""
        g = _GOALS.get(self.env_id)
        return torch.from_numpy(g).to(self.device).expand(self.num_envs, self.dimg)

    def compute_reward(self, achieved_goal, desired_goal, info=None):
        return compute_reward(achieved_goal, desired_goal, info)

    def set_test(self):
        obs, ag, g = self._empty(self.num_envs, self.dimo), self._empty(self.num_envs, self.dimg), self._empty(self.num_envs, self.dimg)
        check(self.L.bp_set_test(self._h, _ptr(obs), _ptr(ag), _ptr(g), _stream(self.device)))
        return dict(observation=obs, achieved_goal=ag, desired_goal=g)

    def increase_difficulty(self):
        r = C.c_int()
        check(self.L.bp_increase_difficulty(self._h, C.byref(r)))
        return bool(r.value)

    def get_difficulty(self):
        r = C.c_int()
        check(self.L.bp_get_difficulty(self._h, C.byref(r)))
        return r.value

    def set_option(self, name, value):
        """Measurement knobs of the library (bp_set_option), e.g. "force_full_physics"."""
        check(self.L.bp_set_option(self._h, name.encode(), int(value)))

    def get_ranges(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        check(self.L.bp_get_ranges(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(obj_range=a.value, wrong_obj_range=b.value, max_obj_range=c.value)

    def set_ranges(self, obj_range, wrong_obj_range=0.0):
        check(self.L.bp_set_ranges(self._h, float(obj_range), float(wrong_obj_range)))

    # ------------------------------------------------------------------ replay harness / stats
    def get_state(self):
        buf = torch.empty(self.num_envs * _lib.STATE_BYTES, dtype=torch.uint8, device=self.device)
        check(self.L.bp_get_state(self._h, _ptr(buf), _stream(self.device)))
        return buf.cpu().numpy().view(STATE_DTYPE)

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=STATE_DTYPE)
        buf = torch.from_numpy(st.view(np.uint8).copy()).to(self.device)
        check(self.L.bp_set_state(self._h, _ptr(buf), _stream(self.device)))
        torch.cuda.current_stream(self.device).synchronize()

    def stats_tensor(self):
        """float64[8] device tensor ALIASING the handle's statistics vector (all-reduce it in place).  It is a view
        into memory the handle owns: it dangles after close(); take .clone() to keep the numbers."""
        if not getattr(self, "_h", None):
            raise _lib.BlockPuzzleError("the env has been closed")
        return _wrap_device_f64(self._stats_ptr, _lib.BP_NUM_STATS, self.device)

    def stats(self):
        v = self.stats_tensor().cpu().numpy()          # .cpu() copies: nothing aliasing the handle escapes
        return {k: float(v[i]) for i, k in enumerate(_lib.STAT_NAMES)}

    def stats_reset(self):
        check(self.L.bp_stats_reset(self._h, _stream(self.device)))


class _CudaArrayView:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr=typestr, data=(ptr, False), version=3, strides=None)


def _wrap_device_f64(ptr, n, device):
    return torch.as_tensor(_CudaArrayView(ptr, n, "<f8"), device=device)


def _goal_matrix(env_id):
    # _sample_goal, fetch_env.py:260-273 (:682-695 Variation) on the fixed colour lists
    GREY, RED, GREEN, BLUE = 0, 1, 2, 3
    colours = {
        0: [BLUE, GREY, GREEN],
        1: [GREY, GREY, GREEN, BLUE],
        2: [RED, GREEN, GREY, GREY, GREY, BLUE],
        3: [GREY, GREY, GREEN, BLUE],
        4: [GREY, GREY, GREEN, BLUE, GREY],
        5: [GREY, GREY, GREEN, BLUE, GREY],
        6: [GREY, GREY, GREEN, BLUE, GREY, GREY],
    }[env_id]
    n = len(colours)
    g = np.zeros((n, n), np.float32)
    for i in range(n):
        for j in range(n):
            a, b = colours[i], colours[j]
            if {a, b} == {RED, BLUE}:
                g[i, j] = -1
            elif {a, b} == {GREEN, BLUE}:
                g[i, j] = 1
    return g.ravel()


_GOALS = {i: _goal_matrix(i) for i in range(len(ENV_IDS))}


def compute_reward(achieved_goal, desired_goal, info=None):
    """BlocksEnv.compute_reward (fetch_env.py:135-143) on the GPU for any leading batch shape.

    Accepts torch tensors (any device) or numpy arrays; returns the same kind, float32.
    """
    L = _lib.load()
    is_np = not torch.is_tensor(achieved_goal)
    dev = torch.device("cuda", torch.cuda.current_device())
    ag = torch.as_tensor(np.asarray(achieved_goal, dtype=np.float32)) if is_np else achieved_goal
    g = torch.as_tensor(np.asarray(desired_goal, dtype=np.float32)) if not torch.is_tensor(desired_goal) else desired_goal
    src_dev = ag.device
    ag = ag.to(device=dev if ag.device.type != "cuda" else ag.device, dtype=torch.float32)
    g = g.to(device=ag.device, dtype=torch.float32)
    ag, g = torch.broadcast_tensors(ag, g)
    ag, g = ag.contiguous(), g.contiguous()
    dimg = ag.shape[-1]
    n = ag.numel() // dimg if dimg else 0
    r = torch.empty(ag.shape[:-1], dtype=torch.float32, device=ag.device)
    check(L.bp_compute_reward(_ptr(ag), _ptr(g), n, dimg, _ptr(r), _stream(ag.device)))
    if is_np:
        return r.cpu().numpy()
    return r if src_dev.type == "cuda" else r.to(src_dev)


class GymBlocksEnv:
    """Single-env gym GoalEnv surface (the object `gym.make(env_name)` returns in the reference),
    backed by a 1-env VecBlocksEnv.  numpy float64 observations like the reference (fetch_env.py:224-228)."""

    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 25}
    reward_range = (-float("inf"), float("inf"))

    def __init__(self, env_name, device=None, seed=0, reward_type="sparse", challenge=False):
        # reward_type: the registered kwarg (gym_blocks/__init__.py:9 ...); the reference stores it and never reads
        # it (fetch_env.py:72, appendix A8)
        self.reward_type = reward_type
        self._vec = VecBlocksEnv(env_name, 1, device=device, seed=seed, challenge=challenge)
        self.spec_id = env_name
        self._max_episode_steps = MAX_EPISODE_STEPS
        self._elapsed_steps = None
        self.action_space = self._vec.action_space
        self.observation_space = self._vec.observation_space
        self.unwrapped = self

    def _obs(self, d):
        o = {k: v[0].double().cpu().numpy() for k, v in d.items()}
        if self._vec.env_name != "BlocksTouchVariation-v0":
            # _sample_goal builds the goal from python ints (fetch_env.py:260-273): int64; Variation pads into a
            # float array (:682-695)
            o["desired_goal"] = o["desired_goal"].astype(np.int64)
        return o

    def seed(self, seed=None):
        return self._vec.seed(seed)

    def reset(self):
        self._elapsed_steps = 0
        return self._obs(self._vec.reset())

    def step(self, action):
        assert self._elapsed_steps is not None, "Cannot call env.step() before calling reset()"
        a = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(1, 4), device=self._vec.device)
        obs, r, done, info = self._vec.step(a)
        self._elapsed_steps += 1
        return (self._obs(obs), np.float32(r[0].item()), bool(self._elapsed_steps >= self._max_episode_steps),
                {"is_success": bool(info["is_success"][0].item() != 0)})

    def compute_reward(self, achieved_goal, desired_goal, info=None):
        return compute_reward(achieved_goal, desired_goal, info)

    def set_test(self):
        return self._obs(self._vec.set_test())

    def increase_difficulty(self):
        return self._vec.increase_difficulty()

    def get_difficulty(self):
        return self._vec.get_difficulty()

    def render(self, mode="human"):
        return None  # no viewer in a batched backend (robot_env.py:89-99)

    def close(self):
        self._vec.close()
