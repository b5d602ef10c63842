#!/usr/bin/env python
"""BASELINE.json configs[3]: HER relabel + batched compute_reward over a 1 Mi-transition replay batch
(future_p = 0.8 from replay_k = 4, config.py:49-50; and 0.0 for replay_strategy='none'), plus the plain
compute_reward kernel.  Prints one JSON line per measurement (CUDA events, L2 flushed between runs)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import blockpuzzle_gym_b200 as bpg  # noqa: E402
from blockpuzzle_gym_b200 import _lib  # noqa: E402
import ctypes as C  # noqa: E402


def main():
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    B_ep, T, dimg, n = 20000, 50, 16, 1 << 20
    env = bpg.make_vec("BlocksTouch-v0", B_ep, device=0, seed=0)
    o0 = env.reset()
    out = env.step_fused(None, K=T, auto_reset=False, outputs=("achieved_goal",))   # a real rollout
    ag = torch.cat([o0["achieved_goal"][None], out["achieved_goal"]], 0).transpose(0, 1).contiguous()
    g = env.goal()[:, None, :].expand(B_ep, T, dimg).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    L = _lib.load()
    p = lambda t: C.c_void_p(t.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {k: torch.empty(n, dtype=torch.int32, device=dev) for k in ("e", "t", "ft")}
    ag2 = torch.empty(n, dimg, device=dev); gout = torch.empty(n, dimg, device=dev); r = torch.empty(n, device=dev)

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    for fp in (0.8, 0.0):
        ms = timed(lambda: L.bp_her_relabel(p(ag), p(g), B_ep, T, dimg, n, fp, 0, 0, p(res["e"]), p(res["t"]), p(res["ft"]), p(ag2), p(gout), p(r), stream))
        byt = n * (12 * dimg + 12)
        print(json.dumps({"metric": "her_transitions_per_sec", "value": n / (ms * 1e-3), "unit": "transitions/s", "future_p": fp,
                          "ms": ms, "config": {"workload": "HER relabel + compute_reward, 1Mi transitions, 20000x50 episode store, dimg 16"},
                          "roofline": {"bound": "hbm", "achieved": byt / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": byt / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_transition": 12 * dimg + 12}}))
    a = ag2; b = gout
    ms = timed(lambda: L.bp_compute_reward(p(a), p(b), n, dimg, p(r), stream))
    byt = n * (8 * dimg + 4)
    print(json.dumps({"metric": "compute_reward_rows_per_sec", "value": n / (ms * 1e-3), "unit": "rows/s", "ms": ms,
                      "roofline": {"bound": "hbm", "achieved": byt / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": byt / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_row": 8 * dimg + 4}}))
    # CPU comparison: the numpy formula of fetch_env.py:141-143 on the same rows
    an, bn = a.cpu().numpy(), b.cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(5):
        d = np.sum(an * bn, axis=-1); c = np.count_nonzero(bn, axis=-1); rr = -(d != c).astype(np.float32)
    cpu = 5 * n / (time.perf_counter() - t0)
    assert np.array_equal(rr, r.cpu().numpy())
    print(json.dumps({"metric": "compute_reward_rows_per_sec", "impl": "reference numpy formula, 1 core", "value": cpu, "unit": "rows/s"}))


if __name__ == "__main__":
    main()
