#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched gym_blocks hot path on N B200s (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[2]/[4] -- 1,048,576 BlocksTouch-v0 envs per GPU,
K = 64 fused env steps per launch, random actions resident in HBM, outputs obs/ag/reward/success
written to HBM every step, auto-reset at T = 50.  One bench "step" = one launch = 64 Mi env-steps
per GPU.  Multi-GPU: envs shard by global index (weak scaling), one NCCL all-reduce of the
8-element statistics vector per launch.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C-ABI)
  python bench.py --impl reference ...                     the reference's OWN step loop on the host cores: the
                                                           unmodified gym_blocks env (oracle/_ref under the stub
                                                           packages of oracle/refharness, BlockPhys in the MjSim
                                                           slot), one process per core like `mpirun -np N`

Besides the headline line the N = 1 run reports `workloads` (uniform random actions / a scripted push policy / the
all-full-physics floor, each with its full-physics fraction and roofline fraction) and `her` (BASELINE.json
configs[3]: the replay-side kernels with a roofline each).
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
    """One single-env step loop of the REFERENCE (gym.make(id) of the unmodified gym_blocks under the replay stubs;
    oracle port only where neither /root/reference nor oracle/_ref exists): random actions, reset at done."""
    seed, nsteps = args
    import numpy as np
    from oracle import refharness as rh
    if rh.available():
        env = rh.make(ENV_NAME, seed=seed)
        rng = np.random.RandomState(seed)
        env.reset()
        acts = rng.uniform(-1, 1, size=(nsteps, 4)).astype(np.float32)
        t0 = time.perf_counter()
        for t in range(nsteps):
            _, _, done, _ = env.step(acts[t])
            if done:
                env.reset()
        return nsteps, time.perf_counter() - t0
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


def _cpu_kind():
    from oracle import refharness as rh
    root = rh.reference_root()
    if root is None:
        return "port", "Python restatement of the env logic (oracle/gym_blocks_oracle.py) + C BlockPhys sim.step()"
    what = "sources under /root/reference" if root == rh.REF_SOURCE else "compiled copy oracle/_ref"
    return "reference", ("the reference's unmodified gym_blocks env (%s) run under the stub gym / mujoco_py packages of oracle/refharness "
                         "with the C BlockPhys model in the MjSim slot (MuJoCo's own solver cost is not in this number)" % what)


def cpu_baseline_single(nsteps=3000, repeats=5):
    """BlocksTouch-v0, seed 0, random actions, reset every 50; median of 5."""
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
    """--impl reference: the reference's step loop on all host cores, one process per core
    (the reference's own parallelism: mpirun -np N -bind-to core, util.py:102-112)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_proc = 3000  # env steps per process per bench step: a bounded sample (~1 s of wall clock per step)
    kind, what = _cpu_kind()
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
    sample = f"{cores} processes x {per_proc} env-steps per bench step, {args.steps} steps; {what}"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s single-env Python step loop per process (gym.make + TimeLimit), uniform random actions, reset at done (every 50 steps)" % ENV_NAME,
                       "env_id": ENV_NAME},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local):
    """Pin this rank to the CPUs of its GPU's NUMA node BEFORE any pinned host buffer is allocated: eight ranks that
    stage through one socket's DRAM / root ports are what capped the 8-GPU end-to-end number of round 1."""
    info = {}
    try:
        bus = subprocess.check_output(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"], text=True).strip().lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]                                     # nvidia-smi prints an 8-digit PCI domain, sysfs a 4-digit one
        info["pci"] = bus
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes"] = len(nodes)
        if node < 0 or len(nodes) < 2:
            info["bound"] = False                             # single-node host (or no affinity reported): nothing to bind
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["bound"], info["cpus"] = True, len(cpus)
        else:
            info["bound"] = False
    except Exception as e:
        info["bound"], info["error"] = False, "%s: %s" % (type(e).__name__, e)
    return info


def _mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return None


def push_policy(step_counter):
    """A scripted "move behind cube 0, come down, push it into cube 1" policy on the device (BlocksTouch-v0 observation
    layout, fetch_env.py:187-222): the closed-loop stand-in for a trained policy that actually pushes cubes."""
    import torch

    def policy(o, ag, g):
        t = step_counter[0]
        step_counter[0] += 1
        grip, b0, b1 = o[:, 0:3], o[:, 10:13], o[:, 25:28]
        d = b1[:, :2] - b0[:, :2]
        d = d / d.norm(dim=1, keepdim=True).clamp_min(1e-6)
        behind = b0[:, :2] - 0.07 * d
        if t < 6:
            xy, z = behind, 0.55
        elif t < 10:
            xy, z = behind, 0.48
        else:
            xy, z = b1[:, :2] + 0.1 * d, 0.48
        dxy = (xy - grip[:, :2]) / 0.05
        dxy = dxy / dxy.abs().amax(dim=1, keepdim=True).clamp_min(1.0)
        az = ((z - grip[:, 2]) / 0.05).clamp(-1, 1)
        return torch.cat([dxy, az[:, None], torch.full_like(az[:, None], -1.0)], 1).contiguous()
    return policy


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU loop"
    numa = bind_to_gpu_numa_node(local)       # before the first pinned allocation
    import blockpuzzle_gym_b200 as bpg
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
    peak, peak_src = _peak_hbm()

    # ---- the other workloads (N = 1 only): what a policy that really pushes cubes costs, and the floor
    workloads = None
    if world == 1 and not args.no_workloads:
        def timed_launches(fn, reps, before=None):
            ms = []
            for _ in range(reps):
                if before is not None:
                    before()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            return statistics.mean(ms)

        def entry(ms, steps_per_launch, note):
            s_ = env.stats()
            v = B * steps_per_launch / (ms * 1e-3)
            return {"value": v, "unit": UNIT, "ms_per_launch": ms, "fused_steps_per_launch": steps_per_launch,
                    "full_physics_frac": s_["worker_steps"] / max(s_["steps"], 1.0), "success_rate": s_["successes"] / max(s_["episodes"], 1.0) if s_["episodes"] else None,
                    "frac": v * BYTES_PER_ENV_STEP / 1e9 / peak, "note": note}

        workloads = {"uniform_random": {"value": B * K * args.steps / (ev[0].elapsed_time(ev[-1]) * 1e-3), "unit": UNIT,
                                        "ms_per_launch": statistics.mean(kern_ms), "fused_steps_per_launch": K,
                                        "full_physics_frac": st["worker_steps"] / max(st["steps"], 1.0), "success_rate": st.get("success_rate"),
                                        "frac": B * K * BYTES_PER_ENV_STEP / (statistics.mean(kern_ms) * 1e-3) / 1e9 / peak,
                                        "note": "the headline workload: actions ~ U(-1, 1)^4"}}
        if ENV_NAME in ("BlocksTouch-v0", "BlocksTouchCurriculum-v0"):
            # (ii) scripted push: one closed-loop episode with the policy on the device records its actions
            # (bp_rollout_begin + 50 x bp_rollout_step), then the same episode is replayed as ONE fused launch of
            # T = 50 steps on those actions (identical trajectories: the env is deterministic) and timed
            env.seed(7); 
            saved = None
            counter = [0]
            ep = env.collect_rollouts(push_policy(counter))
            u = ep["u"].transpose(0, 1).contiguous()                 # [T, B, 4]
            del ep
            env.seed(7); env.reset()
            saved = env.get_state()
            out50 = {}
            def replay():
                env.step_fused(u, auto_reset=False, out=out50)
            def restore():
                env.set_state(saved); env.stats_reset()
            restore(); replay(); torch.cuda.synchronize()
            ms = timed_launches(replay, 5, before=restore)
            workloads["scripted_push"] = entry(ms, 50, "closed-loop 'move behind cube 0, descend, push it into cube 1' policy: recorded once with the "
                                                       "closed-loop collector, replayed as one fused launch of a whole episode (T = 50, no auto-reset)")
            del out50, u
        # (iii) the floor: every env-step takes the complete BlockPhys step (no quiet path), same actions as the headline
        env.set_option("force_full_physics", 1)
        env.seed(0); env.reset(); env.stats_reset()
        launch(); torch.cuda.synchronize(); env.stats_reset()
        ms = timed_launches(launch, 3)
        workloads["all_full_physics"] = entry(ms, K, "bp_set_option(force_full_physics): uniform random actions, but no env-step may take the quiet path")
        env.set_option("force_full_physics", 0)
        env.seed(0); env.reset(); env.stats_reset()

    # ---- e2e: the same metric through the C-ABI with HOST buffers (pinned), copies inside the timed region
    del out
    out = {}
    torch.cuda.empty_cache()
    Ke, Be = args.e2e_fused, B
    per_call_gb = Be * Ke * 4 * (4 + DIMO + DIMG + 2) / 1e9
    avail = _mem_available_gb()
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    e2e_note = None
    while Ke > 8 and avail is not None and per_call_gb * local_world > 0.45 * avail:   # never pin more than ~half of the host's free memory
        Ke //= 2                                                        # not enough host memory to pin K = 64 on every rank
        per_call_gb = Be * Ke * 4 * (4 + DIMO + DIMG + 2) / 1e9
        e2e_note = "host memory (%.0f GB available) does not hold K = %d pinned buffers for %d ranks: K reduced" % (avail, args.e2e_fused, local_world)
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
    del h_act, h_obs, h_ag, h_r, h_s

    if rank == 0:
        value = world * B * K * args.steps / (total_ms * 1e-3)
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
                         "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP, "kernel": "step_kernel_async (dominant; 1 launch per bench step)",
                         "launch_ms": launch_ms,
                         "note": "nominally HBM-bound; measured limiter is instruction issue / instruction delivery (DESIGN.md section 6)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": Be * Ke * 16, "d2h_bytes_per_step": Be * Ke * 4 * (DIMO + DIMG + 2),
                    "fused_steps_per_call": Ke, "calls": args.e2e_steps, "api": "bp_step_host (pinned host buffers, chunked double-buffered copies)",
                    "numa_binding": numa, "note": e2e_note},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "episode_stats": st,
        }
        if workloads is not None:
            line["workloads"] = workloads
        if world == 1 and not args.no_her:
            import bench_her
            del env, actions
            torch.cuda.empty_cache()
            her = {}
            for d in bench_her.measure(emit=None, device=local, cpu_legs=False):
                key = d["metric"] + ("_future_p_%g" % d["future_p"] if "future_p" in d else "")
                her[key] = {"value": d["value"], "unit": d["unit"], "ms": d["ms"], "roofline": d["roofline"], "workload": d.get("config", {}).get("workload")}
            her["_method"] = ("CUDA events around each launch, median of 20; L2 flushed before every timed launch by writing a 256 MB "
                              "buffer and then reading another 256 MB (so the evicted lines are clean)")
            line["her"] = her
        if world == 1 and not args.no_cpu_baseline:
            v = cpu_baseline_single()
            kind, what = _cpu_kind()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": "%s, seed 0, 3000 env-steps x 5 runs (median), reset every 50; %s" % (ENV_NAME, what),
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
    ap.add_argument("--e2e-fused", type=int, default=FUSED, help="fused steps per bp_step_host call (default: the headline K = 64)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-workloads", action="store_true")
    ap.add_argument("--no-her", action="store_true")
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
