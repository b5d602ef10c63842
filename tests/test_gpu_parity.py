"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on the
same seeded inputs.  Bar: bit-exact for integer state (touch matrix, latch, counters, Philox
draw counters) AND bit-identical fp32 for every float (the BlockPhys specification fixes the op order)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import coracle  # noqa: E402


def _make(name, B, seed=0, offset=0):
    import blockpuzzle_gym_b200 as bpg
    env = bpg.make_vec(name, B, device=0, seed=seed, env_index_offset=offset)
    ref = coracle.OracleVecEnv(name, B, seed=seed, env_index_offset=offset)
    return env, ref


def _assert_state_equal(env, ref, ctx=""):
    gs, rs = env.get_state(), ref.get_state()
    for f in gs.dtype.names:
        a, b = np.ascontiguousarray(gs[f]), np.ascontiguousarray(rs[f])
        assert a.tobytes() == b.tobytes(), f"state field {f} differs {ctx}"


@pytest.mark.parametrize("name", coracle.ENV_IDS)
def test_reset_matches_oracle(name):
    env, ref = _make(name, 257, seed=7)
    for _ in range(3):
        o = env.reset()
        ro, rag, rg = ref.reset()
        assert np.array_equal(o["observation"].cpu().numpy(), ro)
        assert np.array_equal(o["achieved_goal"].cpu().numpy(), rag)
        assert np.array_equal(o["desired_goal"].cpu().numpy(), rg)
        _assert_state_equal(env, ref)


@pytest.mark.parametrize("name", coracle.ENV_IDS)
def test_step_by_step_replay(name):
    """2 episodes of host-provided random actions, one launch per step, every output compared."""
    B = 192
    env, ref = _make(name, B, seed=11)
    rng = np.random.RandomState(5)
    env.reset(); ref.reset()
    for ep in range(2):
        for t in range(50):
            a = rng.uniform(-1.3, 1.3, size=(B, 4)).astype(np.float32)  # some outside [-1,1]: exercises the clip
            obs, r, done, info = env.step(torch.from_numpy(a).cuda())
            o2, ag2, r2, s2, _, _ = ref.step(a)
            assert np.array_equal(obs["observation"].cpu().numpy(), o2), (name, ep, t)
            assert np.array_equal(obs["achieved_goal"].cpu().numpy(), ag2), (name, ep, t)
            assert np.array_equal(r.cpu().numpy().view(np.uint32), r2.view(np.uint32)), "reward incl. sign of -0.0"
            assert np.array_equal(info["is_success"].cpu().numpy(), s2)
            assert bool(done.all().item()) == (t == 49)
        _assert_state_equal(env, ref, f"after episode {ep}")
        env.reset(); ref.reset()


def test_4096_envs_fused_philox_replay():
    """BASELINE config 2: 4096 batched envs, Philox-replayed bit-exact check (K fused steps, auto-reset)."""
    B, K = 4096, 64
    env, ref = _make("BlocksTouch-v0", B, seed=0)
    env.reset(); ref.reset()
    for launch in range(2):
        out = env.step_fused(None, K=K, auto_reset=True, want_actions=True, want_reset_obs=True, want_done=True)
        acts = out["actions"].cpu().numpy()
        obs = out["observation"].cpu().numpy(); ag = out["achieved_goal"].cpu().numpy()
        rew = out["reward"].cpu().numpy(); suc = out["is_success"].cpu().numpy()
        done = out["done"].cpu().numpy()
        robs = out["reset_observation"].cpu().numpy(); rag = out["reset_achieved_goal"].cpu().numpy()
        n_done = 0
        for k in range(K):
            a = ref.random_actions()
            assert np.array_equal(a, acts[k])
            t_before = ref.get_state()["t"]
            o2, ag2, r2, s2, ro2, ra2 = ref.step(a, auto_reset=True)
            assert np.array_equal(obs[k], o2), (launch, k)
            assert np.array_equal(ag[k], ag2), (launch, k)
            assert np.array_equal(rew[k].view(np.uint32), r2.view(np.uint32))
            assert np.array_equal(suc[k], s2)
            d2 = t_before == 49                                     # TimeLimit: done on the 50th step
            assert np.array_equal(done[k].astype(bool), d2)
            if d2.any():                                            # fresh observation of the in-kernel reset
                assert np.array_equal(robs[d2], ro2[d2]) and np.array_equal(rag[d2], ra2[d2])
                n_done += int(d2.sum())
        assert n_done == B                                          # every env finishes exactly one episode per 64-step launch here
        _assert_state_equal(env, ref, f"launch {launch}")
    st = env.stats()
    assert st["steps"] == 2 * K * B
    assert st["episodes"] == ref.stats[0] and st["successes"] == ref.stats[1]


@pytest.mark.parametrize("name", ["GripperTouch-v0", "ToppleTower-v0", "BlocksTouchChooseCurriculum-v0", "BlocksTouchVariation-v0"])
def test_fused_replay_other_envs(name):
    B, K = 512, 120
    env, ref = _make(name, B, seed=3)
    env.reset(); ref.reset()
    out = env.step_fused(None, K=K, auto_reset=True, want_actions=True)
    acts = out["actions"].cpu().numpy()
    obs = out["observation"].cpu().numpy(); ag = out["achieved_goal"].cpu().numpy()
    for k in range(K):
        o2, ag2, r2, s2, _, _ = ref.step(acts[k], auto_reset=True)
        assert np.array_equal(obs[k], o2), (name, k)
        assert np.array_equal(ag[k], ag2), (name, k)
    _assert_state_equal(env, ref)


def test_10k_seeded_replay_episodes_state_hash():
    """North-star bar: bit-exact state versus the oracle on 10k seeded replay episodes
    (2048 envs x 5 episodes = 10240), compared through the canonical state records."""
    B, episodes = 2048, 5
    env, ref = _make("BlocksTouch-v0", B, seed=2024)
    env.reset(); ref.reset()
    for ep in range(episodes):
        env.step_fused(None, K=50, auto_reset=True, outputs=())
        ref.run_random(50)
        _assert_state_equal(env, ref, f"episode {ep}")
    assert env.stats()["episodes"] == B * episodes


def test_sharding_is_invariant_to_env_offset():
    """Multi-GPU contract: env i of a shard with offset o equals env o+i of the unsharded batch."""
    import blockpuzzle_gym_b200 as bpg
    full = bpg.make_vec("BlocksTouch-v0", 256, device=0, seed=9)
    lo = bpg.make_vec("BlocksTouch-v0", 128, device=0, seed=9, env_index_offset=0)
    hi = bpg.make_vec("BlocksTouch-v0", 128, device=0, seed=9, env_index_offset=128)
    for e in (full, lo, hi):
        e.reset()
        e.step_fused(None, K=70, auto_reset=True, outputs=())
    sf = full.get_state()
    assert sf[:128].tobytes() == lo.get_state().tobytes()
    assert sf[128:].tobytes() == hi.get_state().tobytes()


@pytest.mark.parametrize("name", ["BlocksTouch-v0", "BlocksTouchCurriculum-v0", "BlocksTouchChooseCurriculum-v0", "BlocksTouchVariation-v0"])
def test_set_test_and_curriculum(name):
    env, ref = _make(name, 130, seed=21)
    env.reset(); ref.reset()
    o = env.set_test(); ro, rag, rg = ref.set_test()
    assert np.array_equal(o["observation"].cpu().numpy(), ro)
    assert np.array_equal(o["desired_goal"].cpu().numpy(), rg)
    _assert_state_equal(env, ref, "after set_test")
    for level in range(7):
        assert env.increase_difficulty() == ref.increase_difficulty()
        env.reset(); ref.reset()
        _assert_state_equal(env, ref, f"level {level}")
    assert env.get_difficulty() == ref.get_difficulty()
    assert env.get_ranges()["obj_range"] == ref.get_obj_range()


def test_not_implemented_paths():
    import blockpuzzle_gym_b200 as bpg
    for name in ("GripperTouch-v0", "ToppleTower-v0"):
        env = bpg.make_vec(name, 8, device=0)
        env.reset()
        with pytest.raises(NotImplementedError):
            env.set_test()
        with pytest.raises(NotImplementedError):
            env.increase_difficulty()
    env = bpg.make_vec("BlocksTouchChoose-v0", 8, device=0)
    with pytest.raises(AttributeError):       # the reference never sets obj_range_step (fetch_env.py:413-415,420)
        env.increase_difficulty()


def test_partial_reset_mask():
    env, ref = _make("BlocksTouch-v0", 64, seed=4)
    env.reset(); ref.reset()
    a = np.zeros((64, 4), np.float32)
    env.step(torch.from_numpy(a).cuda()); ref.step(a)
    mask = torch.zeros(64, dtype=torch.uint8)
    mask[::3] = 1
    env.reset(mask=mask.cuda())
    for i in range(0, 64, 3):
        ref.reset_one(i)
    _assert_state_equal(env, ref)


def test_nan_actions_are_flagged_not_fatal():
    env, ref = _make("BlocksTouch-v0", 32, seed=1)
    env.reset(); ref.reset()
    a = np.zeros((32, 4), np.float32)
    a[3, 1] = np.nan
    a[5, :] = np.inf
    obs, r, done, info = env.step(torch.from_numpy(a).cuda())
    o2, ag2, r2, s2, _, _ = ref.step(a)
    assert np.isfinite(obs["observation"].cpu().numpy()).all()
    assert np.array_equal(obs["observation"].cpu().numpy(), o2)
    assert env.stats()["invalid"] == 1 == ref.stats[3]


def test_compute_reward_matches_oracle_and_kats():
    import blockpuzzle_gym_b200 as bpg
    rng = np.random.RandomState(0)
    for dimg in (9, 16, 25, 36):
        ag = rng.randint(-1, 2, size=(1000, dimg)).astype(np.float32)
        g = rng.randint(-1, 2, size=(1000, dimg)).astype(np.float32)
        g[::7] = 0  # c == 0 rows
        ag[::5] = g[::5] * g[::5] * g[::5]  # satisfied rows (ag == g where g != 0)
        r = bpg.compute_reward(torch.from_numpy(ag).cuda(), torch.from_numpy(g).cuda(), None).cpu().numpy()
        assert np.array_equal(r.view(np.uint32), coracle.compute_reward(ag, g).view(np.uint32))
    # known answers derived from the in-tree code
    goal = np.zeros(16, np.float32); goal[2 * 4 + 3] = goal[3 * 4 + 2] = 1
    ag = -np.ones(16, np.float32)
    assert bpg.compute_reward(ag, goal, None) == np.float32(-1.0)
    ag[2 * 4 + 3] = ag[3 * 4 + 2] = 1
    r = bpg.compute_reward(ag, goal, None)
    assert r == 0 and np.signbit(r)
    r3 = bpg.compute_reward(np.stack([ag] * 3), goal, None)  # broadcast goal like config.py:110-111
    assert r3.shape == (3,) and (r3 == 0).all() and np.signbit(r3).all()
    assert bpg.compute_reward(np.zeros((0, 16), np.float32), np.zeros((0, 16), np.float32), None).shape == (0,)


@pytest.mark.parametrize("name,dimg,n", [("BlocksTouch-v0", 16, 20000), ("BlocksTouch-v0", 16, 20001), ("ToppleTower-v0", 36, 5001), ("GripperTouch-v0", 9, 5000)])
def test_her_relabel_matches_oracle(name, dimg, n):
    """bp_her_relabel against the oracle: rows of 4 float4 chunks (dimg 16) take the lane-cooperative kernel
    (her_relabel_coop_kernel; n = 20001 leaves a partial last block), dimg 36 / 9 the generic sampler kernel."""
    import blockpuzzle_gym_b200 as bpg
    B, T = 300, 50
    env = bpg.make_vec(name, B, device=0, seed=5)
    o0 = env.reset()
    out = env.step_fused(None, K=T, auto_reset=False)
    ag = torch.cat([o0["achieved_goal"][None], out["achieved_goal"]], 0).transpose(0, 1).contiguous()  # [B,T+1,dimg]
    g = env.goal()[:, None, :].expand(B, T, dimg).contiguous()
    for strategy, fp in (("future", 0.8), ("none", 0.0)):
        sampler = bpg.make_sample_her_transitions(strategy, 4, None, seed=77)
        assert abs(sampler.future_p - fp) < 1e-12
        tr = sampler(dict(ag=ag, g=g), n, index_offset=1000)
        ref = coracle.her_relabel(ag.cpu().numpy(), g.cpu().numpy(), n, fp, 77, 1000)
        for k in ("ep_idx", "t", "future_t", "ag_2", "g", "r"):
            assert np.array_equal(tr[k].cpu().numpy(), ref[k]), (strategy, k)
        if fp:
            frac = float((tr["future_t"] >= 0).float().mean())
            assert abs(frac - 0.8) < 0.02
            ft = tr["future_t"][tr["future_t"] >= 0]
            assert int(ft.max()) <= T and int((ft - tr["t"][tr["future_t"] >= 0]).min()) >= 1


def test_zero_and_tiny_actions_through_subnormal_velocities():
    """With (near-)zero actions the gripper velocity decays by 0.8x per substep into the subnormal range and on to
    zero within ~25 env-steps: the packed f32x2 integrators must keep subnormals exactly like the scalar oracle."""
    B, K = 96, 45
    env, ref = _make("BlocksTouch-v0", B, seed=17)
    env.reset(); ref.reset()
    a = np.zeros((K, B, 4), np.float32)
    a[:3] = np.random.RandomState(0).uniform(-1, 1, size=(3, B, 4)).astype(np.float32)   # get it moving first
    a[3:, B // 2:, :3] = 1e-30                                                          # subnormal-sized targets for half the envs
    out = env.step_fused(torch.from_numpy(a).cuda(), auto_reset=False)
    obs = out["observation"].cpu().numpy()
    tiny = 0
    for k in range(K):
        o, ag, r, s, _, _ = ref.step(a[k])
        assert np.array_equal(obs[k].view(np.uint32), o.view(np.uint32)), k
        v = np.abs(o[:, 5:8])                                                          # grip_velp * dt
        tiny += int(((v > 0) & (v < 1.2e-38)).sum())
    assert tiny > 0, "the test never reached subnormal velocities"
    _assert_state_equal(env, ref)


def test_step_host_matches_device_path():
    import blockpuzzle_gym_b200 as bpg
    B, K = 1000, 7
    a = np.random.RandomState(3).uniform(-1, 1, size=(K, B, 4)).astype(np.float32)
    e1 = bpg.make_vec("BlocksTouch-v0", B, device=0, seed=8); e1.reset()
    e2 = bpg.make_vec("BlocksTouch-v0", B, device=0, seed=8); e2.reset()
    d = e1.step_fused(torch.from_numpy(a).cuda(), auto_reset=True)
    h = e2.step_host(a, auto_reset=True)
    for k in ("observation", "achieved_goal", "reward", "is_success"):
        assert np.array_equal(d[k].cpu().numpy(), h[k]), k
    assert e1.get_state().tobytes() == e2.get_state().tobytes()


@pytest.mark.parametrize("name,B,K", [("ToppleTower-v0", 1, 1), ("ToppleTower-v0", 1001, 3), ("BlocksTouchVariation-v0", 1, 1),
                                      ("BlocksTouchVariation-v0", 1003, 5), ("GripperTouch-v0", 131, 3), ("BlocksTouchChoose-v0", 257, 1)])
def test_step_host_odd_sizes_keep_the_row_stores_aligned(name, B, K):
    """the staged sub-buffers of bp_step_host start on 16 bytes for every (id, B, K) -- ToppleTower /
    Variation rows (dimo 70 / 87, dimg 36) used to leave d_ag misaligned for odd K*n and fault in the float4 stores."""
    import blockpuzzle_gym_b200 as bpg
    a = np.random.RandomState(4).uniform(-1, 1, size=(K, B, 4)).astype(np.float32)
    e1 = bpg.make_vec(name, B, device=0, seed=9); e1.reset()
    e2 = bpg.make_vec(name, B, device=0, seed=9); e2.reset()
    d = e1.step_fused(torch.from_numpy(a).cuda(), auto_reset=True)
    h = e2.step_host(a, auto_reset=True)
    for k in ("observation", "achieved_goal", "reward", "is_success"):
        assert np.array_equal(d[k].cpu().numpy(), h[k]), k
    assert e1.get_state().tobytes() == e2.get_state().tobytes()


def test_step_host_is_ordered_after_the_callers_stream():
    """reset() queued on the caller's stream, step_host() right behind it on the library's private
    streams -- the step kernel must see the reset state (no synchronising call in between)."""
    import blockpuzzle_gym_b200 as bpg
    B, K = 200000, 2
    a = np.random.RandomState(5).uniform(-1, 1, size=(K, B, 4)).astype(np.float32)
    e1 = bpg.make_vec("BlocksTouch-v0", B, device=0, seed=2)
    e2 = bpg.make_vec("BlocksTouch-v0", B, device=0, seed=2)
    da = torch.from_numpy(a).cuda()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            e2.reset()                      # asynchronous launches on `side`; the last one defines the state
        h = e2.step_host(a, auto_reset=True)
    for _ in range(3):
        e1.reset()
    d = e1.step_fused(da, auto_reset=True)
    for k in ("observation", "achieved_goal", "reward", "is_success"):
        assert np.array_equal(d[k].cpu().numpy(), h[k]), k


def test_step_counter_saturates_instead_of_wrapping():
    """with auto_reset = 0 the packed 8-bit step counter used to wrap at 256 into the success bit."""
    import blockpuzzle_gym_b200 as bpg
    env = bpg.make_vec("BlocksTouch-v0", 64, device=0, seed=1); env.reset()
    a = torch.zeros(300, 64, 4, device="cuda")
    out = env.step_fused(a, auto_reset=False, want_done=True)
    done = out["done"].cpu().numpy()
    assert not done[:49].any() and done[49:].all()              # TimeLimit keeps done = True past T
    assert (out["is_success"].cpu().numpy() == 0).all()         # nothing touched: the latch never sets
    assert (env.get_state()["t"] == 255).all()


def test_entry_points_leave_the_current_device_alone():
    import blockpuzzle_gym_b200 as bpg
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    env = bpg.make_vec("BlocksTouch-v0", 256, device=1, seed=1)
    env.reset(); env.step_fused(None, K=3); env.get_state()
    assert torch.cuda.current_device() == 0
    assert torch.empty(1, device="cuda").device.index == 0


def test_step_fused_reallocates_mismatched_output_buffers():
    import blockpuzzle_gym_b200 as bpg
    env = bpg.make_vec("BlocksTouch-v0", 100, device=0, seed=1); env.reset()
    out = env.step_fused(None, K=4)
    keep = out["observation"]
    out2 = env.step_fused(None, K=4, out=out)
    assert out2["observation"] is keep                          # matching buffers are reused
    out3 = env.step_fused(None, K=6, out=out)                   # a different K: fresh buffers, no out-of-bounds write
    assert out3["observation"].shape == (6, 100, 40)
    out3["reward"] = out3["reward"][:, ::2]                     # a sliced view is never handed to the kernel
    out4 = env.step_fused(None, K=6, out=out3)
    assert out4["reward"].shape == (6, 100) and out4["reward"].is_contiguous()


@pytest.mark.parametrize("name", ["BlocksTouch-v0", "GripperTouch-v0"])
def test_step_fused_into_output_tensors_that_are_only_16_byte_aligned(name):
    """The lean kernel writes 32-byte-multiple rows with unchecked 256-bit stores, so launch_step only takes it for 32-byte
    aligned output tensors (bp_kernels.cu: launch_step).  Views that start 16 bytes into an allocation must fall back to the
    general instantiation (run-time alignment test, 128-bit stores) and give the same bits."""
    B, K = 1001, 7
    env, _ = _make(name, B, seed=11)
    env2, _ = _make(name, B, seed=11)
    env.reset(); env2.reset()
    g = torch.Generator().manual_seed(5)
    a = (torch.rand(K, B, 4, generator=g) * 2 - 1).cuda()
    want = env.step_fused(a, auto_reset=True)
    out = {}
    for k, shp in (("observation", (K, B, env2.dimo)), ("achieved_goal", (K, B, env2.dimg))):
        n = K * B * shp[2]
        raw = torch.empty(n + 8, dtype=torch.float32, device="cuda")
        v = raw[4:4 + n].view(*shp)
        assert v.data_ptr() % 32 == 16 and v.is_contiguous()
        out[k] = v
    got = env2.step_fused(a, auto_reset=True, out=out)
    assert got["observation"].data_ptr() == out["observation"].data_ptr()
    for k in ("observation", "achieved_goal", "reward", "is_success"):
        assert torch.equal(got[k].view(torch.int32), want[k].view(torch.int32)), k
    # 4 bytes into an allocation: refused (the rows are written with 128-bit stores), not a device fault
    from blockpuzzle_gym_b200._lib import BlockPuzzleError
    raw = torch.empty(K * B * env2.dimo + 8, dtype=torch.float32, device="cuda")
    bad = {"observation": raw[1:1 + K * B * env2.dimo].view(K, B, env2.dimo)}
    with pytest.raises(BlockPuzzleError, match="16-byte aligned"):
        env2.step_fused(a, auto_reset=True, out=bad)
    torch.cuda.synchronize()
    env2.step_fused(a, auto_reset=True)   # the handle and the context are still usable

@pytest.mark.parametrize("name,test", [("BlocksTouch-v0", False), ("BlocksTouchCurriculum-v0", True), ("ToppleTower-v0", False),
                                       ("BlocksTouchVariation-v0", False)])
def test_closed_loop_collector_matches_stepwise_oracle(name, test):
    """SURVEY 8(f)1, closed loop (rollout.py:91-150): bp_rollout_begin + 50 x bp_rollout_step with a policy that reads
    o_t / ag_t / g, against the oracle stepped with the same policy; and the CUDA-graph replay of the same loop."""
    B, T = 300, 50
    env, ref = _make(name, B, seed=21)
    W = torch.from_numpy(np.random.RandomState(1).normal(size=(env.dimo, 4)).astype(np.float32)).cuda()

    def policy(o, ag, g):
        # quantised to a 1/8 grid so that device / host matmul rounding cannot change an action
        return torch.round(torch.tanh(o @ W * 0.7 + (ag - g).sum(1, keepdim=True) * 0.05) * 10.0) / 8.0

    for rep, graph in enumerate((False, True, True)):
        ep = env.collect_rollouts(policy, test=test, graph=graph)
        o, ag, g = ref.reset()
        if test:
            o, ag, g = ref.set_test()
        assert np.array_equal(ep["o"][:, 0].cpu().numpy(), o) and np.array_equal(ep["ag"][:, 0].cpu().numpy(), ag)
        for t in range(T):
            u = policy(torch.from_numpy(o).cuda(), torch.from_numpy(ag).cuda(), torch.from_numpy(g).cuda()).cpu().numpy()
            assert np.array_equal(ep["u"][:, t].cpu().numpy(), u), (rep, t)
            o, ag, r, s, _, _ = ref.step(u)
            assert np.array_equal(ep["o"][:, t + 1].cpu().numpy(), o), (rep, t)
            assert np.array_equal(ep["ag"][:, t + 1].cpu().numpy(), ag)
            assert np.array_equal(ep["g"][:, t].cpu().numpy(), g)
            assert np.array_equal(ep["r"][:, t].cpu().numpy().view(np.uint32), r.view(np.uint32))
            assert np.array_equal(ep["info_is_success"][:, t, 0].cpu().numpy(), s)
        _assert_state_equal(env, ref, f"after rollout {rep}")


def test_scripted_push_policy_at_scale_matches_oracle():
    """the GPU parity suite must also see a policy that really pushes cubes.  4096 envs
    driven closed-loop by bench.py's scripted push policy (move behind cube 0, descend, push it into cube 1); the C
    oracle replays the recorded actions: every observation, touch matrix, reward and latch bit-identical, and the
    workload is contact-heavy (most episodes succeed, the full-physics share is well above the random-action one)."""
    import bench
    B, T = 4096, 50
    env, ref = _make("BlocksTouch-v0", B, seed=77)
    env.stats_reset()
    ep = env.collect_rollouts(bench.push_policy([0]))
    st = env.stats()
    u = ep["u"].cpu().numpy()
    o, ag, g = ref.reset()
    assert np.array_equal(ep["o"][:, 0].cpu().numpy(), o)
    succ_any = np.zeros(B, bool)
    for t in range(T):
        o, ag, r, s, _, _ = ref.step(u[:, t])
        assert np.array_equal(ep["o"][:, t + 1].cpu().numpy(), o), t
        assert np.array_equal(ep["ag"][:, t + 1].cpu().numpy(), ag), t
        assert np.array_equal(ep["r"][:, t].cpu().numpy().view(np.uint32), r.view(np.uint32)), t
        assert np.array_equal(ep["info_is_success"][:, t, 0].cpu().numpy(), s), t
        succ_any |= s != 0
    _assert_state_equal(env, ref, "after the scripted episode")
    assert succ_any.mean() > 0.6, succ_any.mean()                   # the policy does make the cubes touch
    assert st["worker_steps"] / st["steps"] > 0.18                  # and needs the full physics far more often than random actions (0.13)
    # the same episode replayed as one fused launch (what bench.py times) gives the same final state
    env2, _ = _make("BlocksTouch-v0", B, seed=77)
    env2.reset()
    out = env2.step_fused(ep["u"].transpose(0, 1).contiguous(), auto_reset=False)
    assert np.array_equal(out["observation"][-1].cpu().numpy(), o)
    assert env2.get_state().tobytes() == env.get_state().tobytes()
    # forcing the full physics everywhere changes nothing but the speed
    env3, _ = _make("BlocksTouch-v0", B, seed=77)
    env3.set_option("force_full_physics", 1)
    env3.reset(); env3.stats_reset()
    out3 = env3.step_fused(ep["u"].transpose(0, 1).contiguous(), auto_reset=False)
    assert env3.stats()["worker_steps"] == env3.stats()["steps"] == B * T
    for k in ("observation", "achieved_goal", "reward", "is_success"):
        assert torch.equal(out[k], out3[k]), k


def test_v2_channel_events_match_oracle_on_every_step_kernel():
    """BlockPhys v2: the gripper z channel and the finger channels follow the tabulated propagator
    until a contact acts on them.  tests/ref_callers_common.grasp_and_land_actions drives both events on the oracle (a
    finger landing on a cube; fingers closing on a cube -- asserted there and in tests/test_oracle_cpu.py); the recorded
    actions replayed by the async, split and simple kernels and with the quiet path disabled: every output bit-identical."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import ref_callers_common as rc
    B = 256
    env, ref = _make("BlocksTouch-v0", B, seed=11)
    ref.reset()
    acts, outs = rc.grasp_and_land_actions(ref)
    land = (np.arange(B) % 2) == 0
    st = ref.get_state()
    assert np.mean(np.abs(st["grip_pos"][land, 2] - 0.5285) < 1e-6) > 0.9 and np.mean(np.abs(st["finger_q"][~land] - 0.0241).max(axis=1) < 2e-4) > 0.9
    a = torch.from_numpy(acts).cuda()
    for opt in ({}, {"force_full_physics": 1}, {"step_kernel": 4}, {"step_kernel": 2}):
        e2, _ = _make("BlocksTouch-v0", B, seed=11)
        for k, v in opt.items():
            e2.set_option(k, v)
        e2.reset()
        out = e2.step_fused(a, auto_reset=False)
        for t in range(acts.shape[0]):
            o, ag, r, s, _, _ = outs[t]
            assert np.array_equal(out["observation"][t].cpu().numpy().view(np.uint32), o.view(np.uint32)), (opt, t)
            assert np.array_equal(out["achieved_goal"][t].cpu().numpy(), ag), (opt, t)
            assert np.array_equal(out["reward"][t].cpu().numpy().view(np.uint32), r.view(np.uint32)), (opt, t)
        _assert_state_equal(e2, ref, f"after the land / grasp episode with {opt}")


@pytest.mark.parametrize("name", ["BlocksTouch-v0", "GripperTouch-v0", "ToppleTower-v0", "BlocksTouchChoose-v0", "BlocksTouchVariation-v0"])
def test_contact_heavy_homing_policy_matches_oracle(name):
    """A contact-heavy closed loop on every pass form (register pass: 1-2 cubes; column pass: 3-4 cubes; fingers free and
    blocked): the gripper homes in on a cube, dives and rams it while the fingers open and close at random
    (tests/ref_callers_common.homing_actions: by step 90 every env has a turned cube and 20-70 % have pushed one off the
    table).  120 fused steps with auto-reset: every observation, touch matrix, reward and latch and the final state records
    are bit-identical to the oracle's; BlocksTouch-v0 also with the quiet path disabled and on the split kernels."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import ref_callers_common as rc
    B, K = 512, 120
    rec = coracle.OracleVecEnv(name, B, seed=3)
    rec.reset()
    acts = rc.homing_actions(rec, K)
    a = torch.from_numpy(acts).cuda()
    for opt in ({}, {"force_full_physics": 1}, {"step_kernel": 4}) if name == "BlocksTouch-v0" else ({},):
        env, ref = _make(name, B, seed=3)
        for k, v in opt.items():
            env.set_option(k, v)
        env.reset(); ref.reset(); env.stats_reset()
        out = env.step_fused(a, auto_reset=True)
        for t in range(K):
            o, ag, r, s, _, _ = ref.step(acts[t], auto_reset=True)
            assert np.array_equal(out["observation"][t].cpu().numpy().view(np.uint32), o.view(np.uint32)), (opt, t)
            assert np.array_equal(out["achieved_goal"][t].cpu().numpy(), ag), (opt, t)
            assert np.array_equal(out["reward"][t].cpu().numpy().view(np.uint32), r.view(np.uint32)), (opt, t)
            assert np.array_equal(out["is_success"][t].cpu().numpy(), s), (opt, t)
        _assert_state_equal(env, ref, f"after the homing episodes with {opt}")
        st = env.stats()
        assert st["episodes"] == 2 * B and st["worker_steps"] / st["steps"] > 0.3     # far more full-physics steps than random actions (0.1)


@pytest.mark.parametrize("name,obj_range", [("GripperTouch-v0", 0.05), ("ToppleTower-v0", 0.06)])
def test_spawn_rejection_loop_cap(name, obj_range):
    """the 10 000-attempt cap of the spawn loops.  With obj_range < 0.1 / sqrt(2) the reference's
    `while norm(xy - gripper) < 0.1` (fetch_env.py:330, 780) can never accept -- it would spin forever; BlockPhys caps
    the loop.  Kernel and oracle must stop at the same attempt with the same last draw."""
    env, ref = _make(name, 96, seed=5)
    env.set_ranges(obj_range); ref.set_ranges(obj_range)
    o = env.reset(); ro, rag, rg = ref.reset()
    assert np.array_equal(o["observation"].cpu().numpy(), ro)
    _assert_state_equal(env, ref, "after a capped spawn")
    assert (env.get_state()["draws"][:, 0] >= 10000).all()
    a = torch.zeros(3, 96, 4, device="cuda")
    out = env.step_fused(a, auto_reset=False)
    for k in range(3):
        o2, ag2, r2, s2, _, _ = ref.step(np.zeros((96, 4), np.float32))
        assert np.array_equal(out["observation"][k].cpu().numpy(), o2)
    _assert_state_equal(env, ref, "stepping from a capped spawn")


@pytest.mark.parametrize("name", ["BlocksTouchChoose-v0", "BlocksTouchChooseCurriculum-v0"])
def test_choose_env_challenge_argument(name):
    """BlocksTouchChooseEnv(challenge=True) (fetch_env.py:403,416,452-463; pinned to the reference's own sampler in
    tests/test_ref_pin.py): resets and auto-resets inside a fused launch spawn exactly like the oracle's."""
    import blockpuzzle_gym_b200 as bpg
    B = 512
    env = bpg.make_vec(name, B, device=0, seed=9, challenge=True)
    ref = coracle.OracleVecEnv(name, B, seed=9)
    ref.set_challenge(True)
    o = env.reset(); ro, rag, rg = ref.reset()
    assert np.array_equal(o["observation"].cpu().numpy(), ro)
    _assert_state_equal(env, ref, "after a challenge reset")
    st = ref.get_state()
    green, blue, wrong = (st["blk_pos"][:, k, :2].astype(np.float64) for k in range(3))
    assert (np.linalg.norm(green - blue, axis=1) >= 0.15 - 1e-6).all()
    assert (np.linalg.norm(wrong - (green + blue) / 2, axis=1) <= 0.04 + 1e-6).all()
    out = env.step_fused(None, K=60, auto_reset=True, want_actions=True)       # crosses the auto-reset at step 50
    acts = out["actions"].cpu().numpy()
    for k in range(60):
        obs, ag, r, s, _, _ = ref.step(acts[k], auto_reset=True)
        assert np.array_equal(out["observation"][k].cpu().numpy(), obs), k
    _assert_state_equal(env, ref, "after the fused launch")
    with pytest.raises(TypeError):
        bpg.make_vec("BlocksTouch-v0", 4, device=0, challenge=True)


def test_gym_single_env_surface():
    """The object the reference gets from gym.make(env_name): reset/step/compute_reward/seed + TimeLimit."""
    import blockpuzzle_gym_b200 as bpg
    env = bpg.make("BlocksTouch-v0")
    assert env._max_episode_steps == 50
    env.seed(42)
    ref = coracle.OracleVecEnv("BlocksTouch-v0", 1, seed=42)
    obs = env.reset(); ro, rag, rg = ref.reset()
    assert obs["observation"].dtype == np.float64 and obs["observation"].shape == (40,)
    assert np.array_equal(obs["observation"].astype(np.float32), ro[0])
    for t in range(50):
        u = env.action_space.sample()
        o, r, done, info = env.step(u)
        o2, ag2, r2, s2, _, _ = ref.step(u[None])
        assert np.array_equal(o["observation"].astype(np.float32), o2[0])
        assert isinstance(r, np.float32) and r == r2[0]
        assert info["is_success"] == bool(s2[0])
        assert done == (t == 49)
        assert env.compute_reward(o["achieved_goal"], o["desired_goal"], info) == r


def test_all_step_kernels_agree():
    """The quiet path and the schedulers must be result-neutral: BP_STEP_KERNEL=simple runs the full physics (the
    shared-memory column form, sim_step_col) for every env-step in order; the step-synchronous kernels (split: quiet
    kernel + compacted full-physics kernel + reset kernel per step, the default) and the warp-autonomous slab kernel
    (async) must leave byte-identical state and outputs (checked via hashes)."""
    import hashlib
    import subprocess
    import sys
    code = (
        "import hashlib, torch, blockpuzzle_gym_b200 as bpg\n"
        "h = hashlib.sha256()\n"
        "for name in ('BlocksTouch-v0', 'ToppleTower-v0', 'BlocksTouchVariation-v0', 'GripperTouch-v0'):\n"
        "    env = bpg.make_vec(name, 1000, device=0, seed=13); env.reset()\n"
        "    out = env.step_fused(None, K=130, auto_reset=True)\n"
        "    for k in ('observation', 'achieved_goal', 'reward', 'is_success'): h.update(out[k].cpu().numpy().tobytes())\n"
        "    h.update(env.get_state().tobytes())\n"
        "print(h.hexdigest())\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for mode in ("split", "async", "simple"):
        env = dict(os.environ, BP_STEP_KERNEL=mode, PYTHONPATH=root)
        outs.append(subprocess.check_output([sys.executable, "-c", code], env=env, cwd=root).decode().strip().splitlines()[-1])
    assert outs[0] == outs[1] == outs[2]


@pytest.mark.parametrize("name,test", [("BlocksTouch-v0", False), ("BlocksTouchCurriculum-v0", True), ("BlocksTouchVariation-v0", False),
                                       ("ToppleTower-v0", False)])
def test_batch_major_rollout_collector(name, test):
    """generate_rollouts with the env loop and convert_episode_to_batch_major fused.  The oracle
    follows rollout.py:48-64 + :91-172 step by step; the episode tensors must match bit for bit."""
    B, T = 300, 50
    env, ref = _make(name, B, seed=31)
    rng = np.random.RandomState(8)
    for episode in range(2):
        a = rng.uniform(-1, 1, size=(T, B, 4)).astype(np.float32)
        ep = env.generate_rollouts(torch.from_numpy(a).cuda(), test=test)
        o0, ag0, g0 = ref.reset()
        if test:
            o0, ag0, g0 = ref.set_test()
        obs, ags, succ, rew = [o0], [ag0], [], []
        for t in range(T):
            o, ag, r, s, _, _ = ref.step(a[t])
            obs.append(o); ags.append(ag); succ.append(s); rew.append(r)
        swap = lambda x: np.stack(x).swapaxes(0, 1)                      # util.py:118-128
        assert np.array_equal(ep["o"].cpu().numpy(), swap(obs))
        assert np.array_equal(ep["ag"].cpu().numpy(), swap(ags))
        assert np.array_equal(ep["u"].cpu().numpy(), a.swapaxes(0, 1))
        assert np.array_equal(ep["g"].cpu().numpy(), np.repeat(g0[:, None, :], T, 1))
        assert np.array_equal(ep["info_is_success"].cpu().numpy()[..., 0], swap(succ))
        assert np.array_equal(ep["r"].cpu().numpy().view(np.uint32), swap(rew).view(np.uint32))
        _assert_state_equal(env, ref, f"episode {episode}")
    # the HER sampler consumes the collector's output directly
    import blockpuzzle_gym_b200 as bpg
    tr = bpg.make_sample_her_transitions("future", 4, None, seed=3)(ep, 4096, index_offset=0)
    ref_tr = coracle.her_relabel(ep["ag"].cpu().numpy(), ep["g"].cpu().numpy(), 4096, 0.8, 3, 0)
    for k in ("ep_idx", "t", "g", "r"):
        assert np.array_equal(tr[k].cpu().numpy(), ref_tr[k]), k
    assert tr["o"].shape == (4096, env.dimo) and tr["u"].shape == (4096, 4)
