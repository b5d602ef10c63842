#!/usr/bin/env python
"""BASELINE.json configs[3]: HER relabel + batched compute_reward over a 1 Mi-transition replay batch
(future_p = 0.8 from replay_k = 4, config.py:49-50; and 0.0 for replay_strategy='none'), plus the plain
compute_reward kernel.  Prints one JSON line per measurement (CUDA events, L2 flushed between runs: a 256 MB write followed by a 256 MB read sweep)."""
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


def measure(emit=print, device=0, cpu_legs=True):
    """Runs every replay-side measurement; returns the list of result dicts (each also passed to `emit` as JSON)."""
    results = []

    def out_line(d):
        results.append(d)
        if emit is not None:
            emit(json.dumps(d))

    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    B_ep, T, dimg, n = 20000, 50, 16, 1 << 20
    env = bpg.make_vec("BlocksTouch-v0", B_ep, device=device, seed=0)
    o0 = env.reset()
    out = env.step_fused(None, K=T, auto_reset=False, outputs=("achieved_goal",))   # a real rollout
    ag = torch.cat([o0["achieved_goal"][None], out["achieved_goal"]], 0).transpose(0, 1).contiguous()
    g = env.goal()[:, None, :].expand(B_ep, T, dimg).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sweep = torch.zeros(64 << 20, dtype=torch.float32, device=dev)   # 256 MB read after the flush: see timed()
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
            # L2 flush: write a 256 MB buffer (2x the 126 MB L2), then read another 256 MB so that the lines the timed kernel
            # evicts are clean -- otherwise it also pays for writing the flush's own dirty lines back to HBM (measured on
            # Normalizer.update: 47 us after the write alone, 41 us after write + read sweep, 39 us after a read-only flush)
            flush.zero_()
            sweep.sum()
            torch.cuda._sleep(400000)   # ~0.2 ms of GPU spin: the host enqueues the timed launch before the GPU gets there,
                                        # so the events bracket device time, not Python / ctypes call latency
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    for fp in (0.8, 0.0):
        # read ag_2, read g or the future ag, write g', write r, the (episode, t) indices -- ag_2 itself is
        # only an input of the reward here (the full sampler line below gathers it too), so it is not written out
        ms = timed(lambda: L.bp_her_relabel(p(ag), p(g), B_ep, T, dimg, n, fp, 0, 0, p(res["e"]), p(res["t"]), None, None, p(gout), p(r), stream))
        byt = n * (12 * dimg + 12)
        out_line(({"metric": "her_transitions_per_sec", "value": n / (ms * 1e-3), "unit": "transitions/s", "future_p": fp,
                          "ms": ms, "config": {"workload": "HER relabel + compute_reward, 1Mi transitions, 20000x50 episode store, dimg 16"},
                          "roofline": {"bound": "hbm", "achieved": byt / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": byt / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_transition": 12 * dimg + 12}}))
    L.bp_her_relabel(p(ag), p(g), B_ep, T, dimg, n, 0.8, 0, 0, p(res["e"]), p(res["t"]), p(res["ft"]), p(ag2), p(gout), p(r), stream)
    a = ag2; b = gout
    ms = timed(lambda: L.bp_compute_reward(p(a), p(b), n, dimg, p(r), stream))
    byt = n * (8 * dimg + 4)
    out_line(({"metric": "compute_reward_rows_per_sec", "value": n / (ms * 1e-3), "unit": "rows/s", "ms": ms,
                      "roofline": {"bound": "hbm", "achieved": byt / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": byt / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_row": 8 * dimg + 4}}))
    if not cpu_legs:
        return _device_only_tail(results, out_line, env, bpg, timed, peak, dev, n, T, dimg, device)
    # CPU comparison: the numpy formula of fetch_env.py:141-143 on the same rows
    an, bn = a.cpu().numpy(), b.cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(5):
        d = np.sum(an * bn, axis=-1); c = np.count_nonzero(bn, axis=-1); rr = -(d != c).astype(np.float32)
    cpu = 5 * n / (time.perf_counter() - t0)
    assert np.array_equal(rr, r.cpu().numpy())
    out_line(({"metric": "compute_reward_rows_per_sec", "impl": "reference numpy formula, 1 core", "value": cpu, "unit": "rows/s"}))

    def line(metric, unit, units, ms, byt, per_key, per, workload):
        out_line(({"metric": metric, "value": units / (ms * 1e-3), "unit": unit, "ms": ms, "config": {"workload": workload},
                          "roofline": {"bound": "hbm", "achieved": byt / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": byt / (ms * 1e-3) / 1e9 / peak, per_key: per}}))

    # ---- the full transition sampler (ReplayBuffer.sample + HER + _preprocess_og), with and
    # without the fused normaliser sums, on a real episode store from the batch-major rollout collector
    del ag, g, out
    ep = env.generate_rollouts(None)
    dimo, dimu = env.dimo, 4
    sampler = bpg.make_sample_her_transitions("future", 4, None, seed=0, clip_obs=200.0)
    per = 4 * (2 * (2 * dimo + dimu + 3 * dimg) + 1 + 1 + 1) + 12          # gathered rows in + out, succ in, r + succ out, 3 int32 indices
    stats = torch.zeros(2 * dimo + 1, dtype=torch.float64, device=dev)
    keep = {}
    keep["tr"] = sampler(ep, n, index_offset=0)
    def run_sampler(with_stats):
        sampler(ep, n, index_offset=0, stats=stats if with_stats else None, out=keep["tr"])   # staging buffers reused
    for ws in (False, True):
        ms = timed(lambda: run_sampler(ws))
        line("her_sample_transitions_per_sec", "transitions/s", n, ms, n * per, "algorithmic_bytes_per_transition", per,
             "bp_her_sample: 1Mi full transitions (o,o_2,u,g,ag,ag_2,r,info) from a 20000x50 BlocksTouch-v0 episode store, clip 200, future_p 0.8"
             + (", normaliser sums fused" if ws else ""))
    # CPU comparison: the numpy restatement of the upstream sampler on 1 core (bounded sample)
    from oracle import callers_oracle as co
    hp = {k: v.cpu().numpy() for k, v in ep.items() if k != "r"}
    ref_s = co.make_sample_her_transitions("future", 4, lambda ag_2, g, info: co.compute_reward(ag_2, g, info), seed=0)
    batch = dict(hp, o_2=hp["o"][:, 1:], ag_2=hp["ag"][:, 1:])
    t0 = time.perf_counter(); ref = ref_s(batch, 1 << 18, index_offset=0); ro, rg = co.preprocess_og(ref["o"], ref["ag"], ref["g"]); dt = time.perf_counter() - t0
    assert np.array_equal(ro, keep["tr"]["o"][:1 << 18].cpu().numpy()) and np.array_equal(ref["r"], keep["tr"]["r"][:1 << 18].cpu().numpy())
    out_line(({"metric": "her_sample_transitions_per_sec", "impl": "numpy restatement of the upstream sampler, 1 core, 256Ki transitions", "value": (1 << 18) / dt, "unit": "transitions/s"}))

    # ---- Normalizer.update: column sums of a [1Mi, dimo] matrix
    x = keep["tr"]["o"]
    nz = bpg.Normalizer(dimo)
    ms = timed(lambda: nz.update(x))
    line("normalizer_rows_per_sec", "rows/s", n, ms, n * dimo * 4, "algorithmic_bytes_per_row", dimo * 4, "bp_moments: sum / sum of squares over [1Mi, 40] float32")
    xn = x[:1 << 18].cpu().numpy()
    t0 = time.perf_counter(); xn.sum(axis=0); np.square(xn).sum(axis=0); dt = time.perf_counter() - t0
    out_line(({"metric": "normalizer_rows_per_sec", "impl": "numpy (upstream Normalizer.update), 1 core, 256Ki rows", "value": (1 << 18) / dt, "unit": "rows/s"}))

    # ---- policy-gradient returns and trim
    rr = -(torch.rand(n, T, device=dev) < 0.9).float()
    ms = timed(lambda: bpg.discounted_returns(rr, 1 - 1 / T))   # includes the host-side power table and the output allocation
    line("discounted_return_episodes_per_sec", "episodes/s", n, ms, n * T * 12, "algorithmic_bytes_per_episode", T * 12, "bp_discounted_returns: 1Mi episodes x T=50, float32 r -> float64 G")
    t0 = time.perf_counter(); co.discounted_returns(rr[:4096].cpu().numpy().T, 1 - 1 / T); dt = time.perf_counter() - t0
    out_line(({"metric": "discounted_return_episodes_per_sec", "impl": "reference O(T^2) loop (numpy), 1 core, 4096 episodes", "value": 4096 / dt, "unit": "episodes/s"}))
    del rr, keep, x, ep
    venv = bpg.make_vec("BlocksTouchVariation-v0", 1 << 20, device=device, seed=0)
    o0 = venv.reset()
    ov, agv, gv = o0["observation"], o0["achieved_goal"], o0["desired_goal"]
    ms = timed(lambda: bpg.trim(ov, gv, agv, 40, 16, "BlocksTouchVariation-v0"))
    per = 4 * (87 + 2 * 36 + 40 + 2 * 16)
    line("trim_rows_per_sec", "rows/s", n, ms, n * per, "algorithmic_bytes_per_row", per, "bp_trim: 1Mi BlocksTouchVariation-v0 rows (87/36/36 -> 40/16/16)")
    t0 = time.perf_counter(); co.trim(ov[:8192].cpu().numpy(), gv[:8192].cpu().numpy(), agv[:8192].cpu().numpy(), 40, 16, "BlocksTouchVariation-v0"); dt = time.perf_counter() - t0
    out_line(({"metric": "trim_rows_per_sec", "impl": "reference trim loops (numpy), 1 core, 8192 rows", "value": 8192 / dt, "unit": "rows/s"}))


    return results


def _device_only_tail(results, out_line, env, bpg, timed, peak, dev, n, T, dimg, device):
    """The device measurements of the second half of measure() without the CPU comparison legs (bench.py's "her" block)."""
    def line(metric, unit, units, ms, byt, per_key, per, workload):
        out_line({"metric": metric, "value": units / (ms * 1e-3), "unit": unit, "ms": ms, "config": {"workload": workload},
                  "roofline": {"bound": "hbm", "achieved": byt / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": byt / (ms * 1e-3) / 1e9 / peak, per_key: per}})
    ep = env.generate_rollouts(None)
    dimo, dimu = env.dimo, 4
    sampler = bpg.make_sample_her_transitions("future", 4, None, seed=0, clip_obs=200.0)
    per = 4 * (2 * (2 * dimo + dimu + 3 * dimg) + 1 + 1 + 1) + 12
    stats = torch.zeros(2 * dimo + 1, dtype=torch.float64, device=dev)
    tr = sampler(ep, n, index_offset=0)
    for ws in (False, True):
        ms = timed(lambda: sampler(ep, n, index_offset=0, stats=stats if ws else None, out=tr))
        line("her_sample_transitions_per_sec" + ("_with_stats" if ws else ""), "transitions/s", n, ms, n * per, "algorithmic_bytes_per_transition", per,
             "bp_her_sample: 1Mi full transitions from a 20000x50 BlocksTouch-v0 episode store, clip 200, future_p 0.8" + (", normaliser sums fused" if ws else ""))
    x = tr["o"]
    nz = bpg.Normalizer(dimo)
    ms = timed(lambda: nz.update(x))
    line("normalizer_rows_per_sec", "rows/s", n, ms, n * dimo * 4, "algorithmic_bytes_per_row", dimo * 4, "bp_moments: sum / sum of squares over [1Mi, 40] float32")
    rr = -(torch.rand(n, T, device=dev) < 0.9).float()
    ms = timed(lambda: bpg.discounted_returns(rr, 1 - 1 / T))
    line("discounted_return_episodes_per_sec", "episodes/s", n, ms, n * T * 12, "algorithmic_bytes_per_episode", T * 12, "bp_discounted_returns: 1Mi episodes x T=50")
    del rr, tr, x, ep
    venv = bpg.make_vec("BlocksTouchVariation-v0", 1 << 20, device=device, seed=0)
    o0 = venv.reset()
    ov, agv, gv = o0["observation"], o0["achieved_goal"], o0["desired_goal"]
    ms = timed(lambda: bpg.trim(ov, gv, agv, 40, 16, "BlocksTouchVariation-v0"))
    per = 4 * (87 + 2 * 36 + 40 + 2 * 16)
    line("trim_rows_per_sec", "rows/s", n, ms, n * per, "algorithmic_bytes_per_row", per, "bp_trim: 1Mi BlocksTouchVariation-v0 rows (87/36/36 -> 40/16/16)")
    return results


def main():
    measure()


if __name__ == "__main__":
    main()
