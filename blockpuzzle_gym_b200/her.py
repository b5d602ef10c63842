"""Device-side mirrors of the replay / HER / normaliser pieces the reference's DDPG trainer drives
(SURVEY.md section 8 rows H0, (f)2, (f)3).  The trainer itself (ddpg.py) stays out of scope; these are
the objects it constructs and calls, with the same names, arguments and return layout, over torch CUDA
tensors and the C-ABI kernels of csrc/bp_replay.cu:

    make_sample_her_transitions(replay_strategy, replay_k, reward_fun)
                      baselines.her.her [upstream]; built at config.py:107-123 (replay_strategy /
                      replay_k of config.py:49-50), called from ddpg.py:171 and through the buffer
    ReplayBuffer      baselines.her.replay_buffer [upstream]; built at ddpg.py:100-106, used at
                      ddpg.py:164 (store_episode), :192 (get_current_size), :215 (sample)
    Normalizer        baselines.her.normalizer [upstream]; o_stats at ddpg.py:185-188
    update_normalizer the update_stats branch of DDPG.store_episode, ddpg.py:166-190, in ONE kernel
                      launch (sampling, relabelling, reward, clip and the column sums fused)

Nothing here imports oracle/; without the CUDA library every entry point raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check

CLIP_OBS = 200.0  # config.py:35


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _f32c(x):
    return x.to(torch.float32).contiguous()


def make_sample_her_transitions(replay_strategy, replay_k, reward_fun=None, seed=0, clip_obs=0.0):
    """replay_strategy in {'future', 'none'} (config.py:49); replay_k=4 -> future_p=0.8 (config.py:50).

    reward_fun is accepted for signature compatibility: the kernel applies compute_reward's arithmetic
    (fetch_env.py:135-143) to (ag_2, relabelled g), which is what the reference's closure
    (config.py:110-111) evaluates.  clip_obs > 0 additionally applies DDPG._preprocess_og
    (ddpg.py:111-120) to o, o_2 and g, i.e. returns what DDPG.sample_batch (ddpg.py:214-222) stages.

    The returned sampler takes an episode batch of CUDA tensors {o [B,T+1,dimo], u [B,T,dimu],
    g [B,T,dimg], ag [B,T+1,dimg], info_is_success [B,T,1] (optional)} and returns the transition
    dict of upstream (o, o_2, u, g, ag, ag_2, r, info_is_success) plus the replay bookkeeping
    ep_idx / t / future_t.  `stats` (float64 [2*dimo+1] CUDA tensor) receives the normaliser sums.
    """
    if replay_strategy == "future":
        future_p = 1 - (1.0 / (1 + replay_k))
    else:
        future_p = 0
    state = dict(calls=0)

    def _sample_her_transitions(episode_batch, batch_size_in_transitions, index_offset=None, stats=None,
                                keys=("o", "o_2", "u", "g", "ag", "ag_2", "r", "info_is_success"), clip=None, out=None):
        """out: a dict returned by an earlier call with the same shapes -- its tensors are overwritten
        instead of allocating new ones (a trainer's staging buffers)."""
        L = _lib.load()
        clip = clip_obs if clip is None else clip
        ag = episode_batch["ag"]
        assert torch.is_tensor(ag) and ag.is_cuda, "episode_batch must hold CUDA tensors"
        ag = _f32c(ag)
        g = _f32c(episode_batch["g"])
        B, T1, dimg = ag.shape
        T = T1 - 1
        assert g.shape == (B, T, dimg)
        o = _f32c(episode_batch["o"]) if "o" in episode_batch else None
        u = _f32c(episode_batch["u"]) if "u" in episode_batch else None
        succ = _f32c(episode_batch["info_is_success"]) if "info_is_success" in episode_batch else None
        dimo = o.shape[-1] if o is not None else 1
        dimu = u.shape[-1] if u is not None else 0
        if o is not None:
            assert o.shape == (B, T + 1, dimo)
        if u is not None:
            assert u.shape == (B, T, dimu)
        n = int(batch_size_in_transitions)
        dev = ag.device
        off = state["calls"] * (1 << 40) if index_offset is None else int(index_offset)
        state["calls"] += 1
        want = lambda k, have=True: have and k in keys
        if out is None:
            new = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype, device=dev)
            out = dict(ep_idx=new(n, dtype=torch.int32), t=new(n, dtype=torch.int32), future_t=new(n, dtype=torch.int32))
            if want("o", o is not None): out["o"] = new(n, dimo)
            if want("o_2", o is not None): out["o_2"] = new(n, dimo)
            if want("u", u is not None): out["u"] = new(n, dimu)
            if want("g"): out["g"] = new(n, dimg)
            if want("ag"): out["ag"] = new(n, dimg)
            if want("ag_2"): out["ag_2"] = new(n, dimg)
            if want("r"): out["r"] = new(n)
            if want("info_is_success", succ is not None): out["info_is_success"] = new(n, 1)
        else:
            assert out["ep_idx"].shape == (n,) and all(v.is_cuda and v.is_contiguous() for v in out.values())
        if stats is not None:
            assert stats.dtype == torch.float64 and stats.numel() == 2 * dimo + 1 and stats.is_cuda and "o" in out
        check(L.bp_her_sample(_ptr(o), _ptr(u), _ptr(g), _ptr(ag), _ptr(succ), B, T, dimo, dimu, dimg, n,
                              float(future_p), float(clip), int(seed) & 0xFFFFFFFFFFFFFFFF, off,
                              _ptr(out["ep_idx"]), _ptr(out["t"]), _ptr(out["future_t"]), _ptr(out.get("o")), _ptr(out.get("o_2")),
                              _ptr(out.get("u")), _ptr(out.get("g")), _ptr(out.get("ag")), _ptr(out.get("ag_2")), _ptr(out.get("r")),
                              _ptr(out.get("info_is_success")), _ptr(stats), _stream(dev)))
        return out

    _sample_her_transitions.future_p = future_p
    return _sample_her_transitions


class ReplayBuffer:
    """baselines.her.replay_buffer.ReplayBuffer [upstream] with the episode store resident in HBM.

    buffer_shapes: {key: (T or T+1, dim)} as built at ddpg.py:100-103; size_in_transitions is rounded
    down to whole episodes (ddpg.py:105).  At the BASELINE.json configs[3] size (20 000 episodes of
    BlocksTouch-v0) the store is 0.3 GB of the 180 GB.
    """

    def __init__(self, buffer_shapes, size_in_transitions, T, sample_transitions, device=None, rng=None):
        self.buffer_shapes = dict(buffer_shapes)
        self.size = size_in_transitions // T
        self.T = T
        self.sample_transitions = sample_transitions
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.buffers = {key: torch.empty((self.size, *shape), dtype=torch.float32, device=self.device)
                        for key, shape in self.buffer_shapes.items()}
        self.current_size = 0
        self.n_transitions_stored = 0
        self._rng = rng or np.random.RandomState()   # upstream draws overwrite slots from the global np.random

    @property
    def full(self):
        return self.current_size == self.size

    def sample(self, batch_size, **kw):
        """Returns {key: tensor [batch_size, dim]} (o, o_2, u, g, ag, ag_2, r, ...)."""
        assert self.current_size > 0
        buffers = {key: buf[:self.current_size] for key, buf in self.buffers.items()}
        # o_2 = o[:, 1:], ag_2 = ag[:, 1:] of upstream are not materialised: the kernel reads row t + 1
        transitions = self.sample_transitions(buffers, batch_size, **kw)
        for key in (["r", "o_2", "ag_2"] + list(self.buffers.keys())):
            assert key in transitions, "key %s missing from transitions" % key
        return transitions

    def store_episode(self, episode_batch):
        """episode_batch: {key: [rollout_batch_size, T or T+1, dim]} CUDA tensors (or numpy arrays)."""
        batch_sizes = [len(episode_batch[key]) for key in episode_batch.keys()]
        assert all(b == batch_sizes[0] for b in batch_sizes)
        batch_size = batch_sizes[0]
        idxs = self._get_storage_idx(batch_size)
        contiguous = isinstance(idxs, slice)
        if not contiguous:
            # numpy's buffers[key][idxs] = batch lets the LAST of duplicate slots win; a parallel scatter has
            # no order, so keep only each slot's last writer
            idxs = np.atleast_1d(idxs)
            _, first_rev = np.unique(idxs[::-1], return_index=True)
            rows = np.sort(len(idxs) - 1 - first_rev)
            slots = torch.as_tensor(idxs[rows], device=self.device, dtype=torch.long)
            rows = torch.as_tensor(rows, device=self.device, dtype=torch.long)
        for key in self.buffers.keys():
            src = torch.as_tensor(episode_batch[key], device=self.device, dtype=torch.float32)
            if contiguous:
                self.buffers[key][idxs].copy_(src)
            else:
                self.buffers[key].index_copy_(0, slots, src.index_select(0, rows))
        self.n_transitions_stored += batch_size * self.T

    def get_current_episode_size(self):
        return self.current_size

    def get_current_size(self):
        return self.current_size * self.T

    def get_transitions_stored(self):
        return self.n_transitions_stored

    def clear_buffer(self):
        self.current_size = 0

    def _get_storage_idx(self, inc=None):
        inc = inc or 1
        assert inc <= self.size, "Batch committed to replay is too large!"
        if self.current_size + inc <= self.size:      # fill consecutively ...
            idx = slice(self.current_size, self.current_size + inc)
        elif self.current_size < self.size:           # ... then the tail plus random slots ...
            overflow = inc - (self.size - self.current_size)
            idx = np.concatenate([np.arange(self.current_size, self.size), self._rng.randint(0, self.current_size, overflow)])
        else:                                         # ... then random slots only
            idx = self._rng.randint(0, self.size, inc)
        self.current_size = min(self.size, self.current_size + inc)
        return idx


class Normalizer:
    """baselines.her.normalizer.Normalizer [upstream] (o_stats, ddpg.py:185-188) with device-resident sums.

    update() adds sum, sum of squares and count with one reduction kernel (bp_moments); recompute_stats()
    folds the local sums into the totals -- summing them over ranks with ONE all-reduce of the
    float64 [2*size+1] vector when torch.distributed is initialised (upstream averages the same three
    quantities over MPI ranks, which yields the same mean and std) -- and refreshes mean / std.
    """

    def __init__(self, size, eps=1e-2, default_clip_range=np.inf, device=None):
        self.size = int(size)
        self.eps = eps
        self.default_clip_range = default_clip_range
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.local = torch.zeros(2 * self.size + 1, dtype=torch.float64, device=self.device)   # sum | sumsq | count
        self.total = torch.zeros(2 * self.size + 1, dtype=torch.float64, device=self.device)
        self.total[-1] = 1.0                                                                  # total_count starts at 1 upstream
        self.mean = torch.zeros(self.size, dtype=torch.float32, device=self.device)
        self.std = torch.ones(self.size, dtype=torch.float32, device=self.device)

    def update(self, v, col0=0, clip=0.0):
        """v: [..., size + col0] CUDA float tensor; columns col0.. are accumulated (col0 = 1 is the
        Variation rule o[:, 1:] of ddpg.py:180-181)."""
        v = _f32c(v)
        ld = v.shape[-1]
        assert ld == self.size + col0
        n = v.numel() // ld
        check(_lib.load().bp_moments(_ptr(v), n, self.size, ld, col0, float(clip), _ptr(self.local), _stream(self.device)))

    def add_sums(self, sums, col0=0):
        """Fold sums produced by the fused sampler (`stats=` of the HER sampler: float64 [2*dimo+1])."""
        dimo = (sums.numel() - 1) // 2
        assert dimo == self.size + col0
        self.local[:self.size] += sums[col0:dimo]
        self.local[self.size:2 * self.size] += sums[dimo + col0:2 * dimo]
        self.local[-1] += sums[-1]

    def recompute_stats(self):
        local = self.local.clone()
        self.local.zero_()
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(local)
        self.total += local
        s, q, cnt = self.total[:self.size], self.total[self.size:2 * self.size], self.total[-1]
        mean = s / cnt
        self.mean = mean.to(torch.float32)
        self.std = torch.sqrt(torch.clamp(q / cnt - mean * mean, min=self.eps ** 2)).to(torch.float32)

    def normalize(self, v, clip_range=None):
        if clip_range is None:
            clip_range = self.default_clip_range
        return torch.clamp((v - self.mean) / self.std, -clip_range, clip_range)

    def denormalize(self, v):
        return self.mean + v * self.std


def update_normalizer(episode_batch, sample_transitions, o_stats, env_name, clip_obs=CLIP_OBS, index_offset=None):
    """The update_stats branch of DDPG.store_episode (ddpg.py:166-190): sample as many HER transitions
    as the episode batch holds, clip them (_preprocess_og), and add the o rows to o_stats -- one launch."""
    B, T = episode_batch["u"].shape[0], episode_batch["u"].shape[1]
    dimo = episode_batch["o"].shape[-1]
    sums = torch.zeros(2 * dimo + 1, dtype=torch.float64, device=o_stats.device)
    tr = sample_transitions(episode_batch, B * T, index_offset=index_offset, stats=sums, clip=clip_obs)   # ddpg.py:171-174
    o_stats.add_sums(sums, col0=1 if "Variation" in env_name else 0)   # ddpg.py:180-183
    o_stats.recompute_stats()                                          # ddpg.py:188
    return tr
