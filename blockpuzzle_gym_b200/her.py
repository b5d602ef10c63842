"""Device-side mirror of baselines.her.her.make_sample_her_transitions [upstream], the HER
sampler the reference wires up in gym_blocks/config.py:107-123 and calls from
ddpg.py:106,214-215.  Same factory signature; the returned sampler relabels goals and
recomputes rewards (BlocksEnv.compute_reward, fetch_env.py:135-143) in one CUDA kernel.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def make_sample_her_transitions(replay_strategy, replay_k, reward_fun=None, seed=0):
    """replay_strategy in {'future', 'none'} (config.py:49); replay_k=4 -> future_p=0.8 (config.py:50).

    reward_fun is accepted for signature compatibility; rewards are computed by the
    fused kernel with compute_reward's arithmetic.
    """
    if replay_strategy == "future":
        future_p = 1 - (1.0 / (1 + replay_k))
    else:
        future_p = 0
    state = dict(calls=0)

    def _sample_her_transitions(episode_batch, batch_size_in_transitions, index_offset=None):
        L = _lib.load()
        ag = episode_batch["ag"]
        g = episode_batch["g"]
        assert torch.is_tensor(ag) and ag.is_cuda, "episode_batch must hold CUDA tensors"
        ag = ag.to(torch.float32).contiguous()
        g = g.to(torch.float32).contiguous()
        B, T1, dimg = ag.shape
        T = T1 - 1
        assert g.shape == (B, T, dimg)
        n = int(batch_size_in_transitions)
        dev = ag.device
        off = state["calls"] * (1 << 40) if index_offset is None else int(index_offset)
        state["calls"] += 1
        out = dict(
            ep_idx=torch.empty(n, dtype=torch.int32, device=dev),
            t=torch.empty(n, dtype=torch.int32, device=dev),
            future_t=torch.empty(n, dtype=torch.int32, device=dev),
            ag_2=torch.empty(n, dimg, device=dev),
            g=torch.empty(n, dimg, device=dev),
            r=torch.empty(n, device=dev),
        )
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(L.bp_her_relabel(_ptr(ag), _ptr(g), B, T, dimg, n, float(future_p), int(seed), off,
                               _ptr(out["ep_idx"]), _ptr(out["t"]), _ptr(out["future_t"]), _ptr(out["ag_2"]),
                               _ptr(out["g"]), _ptr(out["r"]), stream))
        # the remaining transition keys are plain gathers at (ep_idx, t)
        e, t = out["ep_idx"].long(), out["t"].long()
        for key, val in episode_batch.items():
            if key in ("g", "ag_2", "r"):  # relabelled goal, its reward and ag_2 come from the kernel
                continue
            if key == "o_2":
                out[key] = val[e, t]
            elif key == "ag":
                out[key] = ag[e, t]
            elif torch.is_tensor(val) and val.dim() >= 2 and val.shape[0] == B:
                out[key] = val[e, t]
        return out

    _sample_her_transitions.future_p = future_p
    return _sample_her_transitions
