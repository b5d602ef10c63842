#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched gym_blocks hot path on N B200s (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[2]/[4] -- 1,048,576 BlocksTouch-v0 envs per GPU,
K = 64 fused env steps per launch, random actions resident in HBM, outputs obs/ag/reward/success
written to HBM every step, auto-reset at T = 50.  One bench "step" = one launch = 64 Mi env-steps
per GPU.  Multi-GPU: envs shard by global index (weak scaling), one NCCL all-reduce of the
8-element statistics vector per launch.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C-ABI)
  python bench.py --impl reference ...                     the reference-style CPU step loop
                                                           (oracle port: Python env logic + C sim),
                                                           one process per host core
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_NAME = "BlocksTouch-v0"
ENVS_PER_GPU = 1 << 20
FUSED = 64
DIMU, DIMO, DIMG = 4, 40, 16
STATE_BYTES = 36 * 4  # device state words per env (18 + 9 * nblocks) * 4, bp_create
# 4*(dimu + dimo + dimg + 2) + 2*state/K bytes per env-step
BYTES_PER_ENV_STEP = 4 * (DIMU + DIMO + DIMG + 2) + 2.0 * STATE_BYTES / FUSED
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"



def _set_env(name):
    """Rebind the workload constants to another registered env id."""
    global ENV_NAME, DIMO, DIMG, STATE_BYTES, BYTES_PER_ENV_STEP
    table = {"GripperTouch-v0": (25, 9, 1), "BlocksTouch-v0": (40, 16, 2), "ToppleTower-v0": (70, 36, 4), "BlocksTouchCurriculum-v0": (40, 16, 2),
             "BlocksTouchChoose-v0": (55, 25, 3), "BlocksTouchChooseCurriculum-v0": (55, 25, 3), "BlocksTouchVariation-v0": (87, 36, 4)}
    ENV_NAME = name
    DIMO, DIMG, nb = table[name]
    STATE_BYTES = (18 + 9 * nb) * 4
    BYTES_PER_ENV_STEP = 4 * (DIMU + DIMO + DIMG + 2) + 2.0 * STATE_BYTES / FUSED


def _peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic():
    """dram bytes per launch of the step kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows = []
        self.p = None
        self.index = index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "25"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """Index of the next sample: brackets the timed region."""
        return len(self.rows)

    def stop(self, first=0, last=None):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)   # let the sample that covers the end of the timed region arrive
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        last = len(self.rows) if last is None else min(len(self.rows), last + 2)
        rows = self.rows[first:last]
        window = "timed region"
        if not rows:   # region shorter than one sampling period: the samples taken under the warm-up load just before it
            rows, window = self.rows[max(0, first - 8):first], "warm-up launches right before the timed region"
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- CPU legs
def _cpu_worker(args):
    """One reference-style single-env loop: Python env logic, native sim.step() (like mujoco_py)."""
    seed, nsteps = args
    from oracle import gym_blocks_oracle as pyo
    env = pyo.make(ENV_NAME)
    env.seed(seed)
    env.reset()
    t0 = time.perf_counter()
    for t in range(nsteps):
        env.step(env.random_action())
        if (t + 1) % 50 == 0:
            env.reset()
    return nsteps, time.perf_counter() - t0


def cpu_baseline_single(nsteps=3000, repeats=5):
    """BlocksTouch-v0, seed 0, Philox actions, reset every 50; median of 5."""
    _cpu_worker((0, 200))
    rates = []
    for _ in range(repeats):
        n, dt = _cpu_worker((0, nsteps))
        rates.append(n / dt)
    return statistics.median(rates)


def c_oracle_rate(n_envs=256, steps=200):
    from oracle import coracle
    env = coracle.OracleVecEnv(ENV_NAME, n_envs, seed=0)
    env.run_random(10)
    t0 = time.perf_counter()
    st = env.run_random(steps)
    return float(st[2]) / (time.perf_counter() - t0)


def run_reference(args):
    """--impl reference: the oracle port on all host cores, one process per core
    (the reference's own parallelism: mpirun -np N -bind-to core, util.py:102-112)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_proc = 10000  # env steps per process per bench step: a bounded sample (~2 s of wall clock per step)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(max(1, args.warmup)):
            pool.map(_cpu_worker, [(s, 200) for s in range(cores)])
        times = []
        total = 0
        for k in range(args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(1000 * s + k, per_proc) for s in range(cores)])
            times.append(time.perf_counter() - t0)
            total += sum(r[0] for r in res)
    value = total / sum(times)
    sample = f"{cores} processes x {per_proc} env-steps per bench step, {args.steps} steps; Python env logic + C BlockPhys sim (oracle port; MuJoCo cost not included)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BlocksTouch-v0 single-env Python step loop per process, Philox random actions, reset every 50 steps",
                       "env_id": ENV_NAME},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import blockpuzzle_gym_b200 as bpg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU loop"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, K = args.envs, args.fused
    env = bpg.make_vec(ENV_NAME, B, device=local, seed=0, env_index_offset=rank * B)
    env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    actions = torch.rand(K, B, 4, device=dev, generator=gen) * 2 - 1
    out = {}
    stats = env.stats_tensor()

    gstats = torch.zeros_like(stats)

    def launch():
        env.step_fused(actions, auto_reset=True, out=out)
        gstats.copy_(stats)
        if world > 1:
            dist.all_reduce(gstats)  # replaces mpi_moments (train.py:21-26); one tiny all-reduce per launch

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        launch()
        env.stats_reset()
    sync()
    # nvidia-smi needs about a second before its first sample: keep the GPU under the same load (untimed
    # launches, rank 0 decides for all ranks) until the sampler delivers, so that the clocks line describes the
    # loaded state even when the timed region is shorter than that
    for _ in range(150):
        flag = torch.tensor([1 if (rank != 0 or sampler.p is None or sampler.mark() >= 4) else 0], device=dev)
        if world > 1:
            dist.broadcast(flag, 0)
        if int(flag.item()):
            break
        launch()
        env.stats_reset()
        torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    sync()
    m0 = sampler.mark()
    ev[0].record()
    for k in range(args.steps):
        launch()
        ev[k + 1].record()
    sync()
    clocks = sampler.stop(m0, sampler.mark()) if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    from blockpuzzle_gym_b200.dist import stats_dict
    st = stats_dict(gstats)  # global (all-rank) statistics of the timed launches

    # ---- e2e: the same metric through the C-ABI with HOST buffers (pinned), copies inside the timed region
    Ke, Be = args.e2e_fused, B
    h_act = torch.empty(Ke, Be, 4, dtype=torch.float32).pin_memory()
    h_act.uniform_(-1, 1)
    h_obs = torch.empty(Ke, Be, DIMO, dtype=torch.float32).pin_memory()
    h_ag = torch.empty(Ke, Be, DIMG, dtype=torch.float32).pin_memory()
    h_r = torch.empty(Ke, Be, dtype=torch.float32).pin_memory()
    h_s = torch.empty(Ke, Be, dtype=torch.float32).pin_memory()

    def e2e_call():
        env.step_host_ptrs(h_act.data_ptr(), Ke, h_obs.data_ptr(), h_ag.data_ptr(), h_r.data_ptr(), h_s.data_ptr(), True)

    e2e_call()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_call()
    sync()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_value = world * Be * Ke * args.e2e_steps / e2e_s

    if rank == 0:
        value = world * B * K * args.steps / (total_ms * 1e-3)
        peak, peak_src = _peak_hbm()
        launch_ms = statistics.mean(kern_ms)
        achieved = B * K * BYTES_PER_ENV_STEP / (launch_ms * 1e-3) / 1e9
        tr = _traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("BASELINE.json configs[2]: " if ENV_NAME == "BlocksTouch-v0" else "per-id table: ") + "1Mi batched %s envs per GPU, K=64 fused steps per launch, uniform random actions, auto-reset at T=50" % ENV_NAME,
                       "env_id": ENV_NAME, "envs_per_gpu": B, "fused_steps_per_launch": K, "env_steps_per_bench_step": world * B * K,
                       "l2": "no flush needed: per-launch inputs (actions %.2f GB) and outputs (%.2f GB) are far larger than the 126 MB L2"
                             % (B * K * 16 / 1e9, B * K * 4 * (DIMO + DIMG + 2) / 1e9),
                       "sharding": "envs by global index, one NCCL all-reduce of float64[8] stats per launch" if world > 1 else "single GPU"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (tr or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP, "kernel": "step_kernel (dominant; 1 launch per bench step)",
                         "launch_ms": launch_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": Be * Ke * 16, "d2h_bytes_per_step": Be * Ke * 4 * (DIMO + DIMG + 2),
                    "fused_steps_per_call": Ke, "calls": args.e2e_steps, "api": "bp_step_host (pinned host buffers, chunked double-buffered copies)"},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "episode_stats": st,
        }
        if world == 1 and not args.no_cpu_baseline:
            v = cpu_baseline_single()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "BlocksTouch-v0, seed 0, 3000 env-steps x 5 runs (median), reset every 50; Python env logic + C BlockPhys sim.step() "
                                              "(oracle port of the reference step loop; MuJoCo's solver cost is not in this number)",
                                    "host_cpus": os.cpu_count(), "c_oracle_single_core_steps_per_s": c_oracle_rate()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--fused", type=int, default=FUSED)
    ap.add_argument("--e2e-fused", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--env", default=ENV_NAME, help="env id (default: the BASELINE workload BlocksTouch-v0); other ids are parity-suite configs, benched for the per-id table of profiles/README.md")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.env != ENV_NAME:
        _set_env(args.env)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
