"""Device-side mirrors of the two batch helpers of the policy-gradient rollout worker
(policy_gradient/rollout.py), SURVEY.md section 8(f) row 4.  The trainer (pggd.py) stays out of scope.

    discounted_returns(r, gamma)   the `returns` accumulation of policy_gradient/rollout.py:255-258
                                   (gamma = 1 - 1/T, policy_gradient/config.py:84) for a whole batch
    trim(o, g, ag, dimo, dimg, ...) RolloutStudent.trim, policy_gradient/rollout.py:105-171
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def discounted_returns(r, gamma):
    """r: [B, T] CUDA float32 rewards of an episode batch (batch-major, as VecBlocksEnv.generate_rollouts
    returns them) -> G [B, T] float64 with G[b, t] = r[b, t] + sum_j gamma**j * r[b, t + j], accumulated in
    increasing j in float64 exactly like the reference's nested loop (rollout.py:255-258)."""
    assert r.is_cuda and r.dim() == 2
    r = r.to(torch.float32).contiguous()
    B, T = r.shape
    pw = torch.tensor([float(gamma) ** j for j in range(T)], dtype=torch.float64, device=r.device)   # python float powers, :258
    G = torch.empty(B, T, dtype=torch.float64, device=r.device)
    check(_lib.load().bp_discounted_returns(_ptr(r), B, T, _ptr(pw), _ptr(G), _stream(r.device)))
    return G


def trim(o, g, ag, dimo, dimg, env_name="", num_objs=4):
    """Batched RolloutStudent.trim: o [n, dimo_in], g / ag [n, dimg_in] CUDA tensors -> rows cut down to an
    expert policy's dimo / dimg = num_objs**2.  Returns the inputs untouched when nothing has to be trimmed
    (rollout.py:107-108)."""
    if o.shape[-1] == dimo:
        return o, g, ag
    assert o.is_cuda and o.dim() == 2 and dimg == num_objs * num_objs
    o, g, ag = (x.to(torch.float32).contiguous() for x in (o, g, ag))
    n = o.shape[0]
    o_ = torch.empty(n, dimo, device=o.device)
    g_ = torch.empty(n, dimg, device=o.device)
    ag_ = torch.empty(n, dimg, device=o.device)
    check(_lib.load().bp_trim(_ptr(o), _ptr(g), _ptr(ag), n, o.shape[1], g.shape[1], dimo, num_objs,
                              int("Variation" in env_name), _ptr(o_), _ptr(g_), _ptr(ag_), _stream(o.device)))
    return o_, g_, ag_
