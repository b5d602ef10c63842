"""GPU tests at BASELINE.json's full sizes (configs[2] and [3]): 1 Mi envs x K = 64 fused steps and a
1 Mi-transition HER batch are beyond what the CPU oracle replays in seconds, so they are checked through
size-independent properties of the domain, plus an oracle-verified slice: every env (transition) is a pure
function of its global index and seed, so the first rows of the big run must equal a small run that the
parity suite ties to the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import callers_oracle as co  # noqa: E402
from oracle import coracle  # noqa: E402


def test_one_million_envs_64_fused_steps():
    import blockpuzzle_gym_b200 as bpg
    B, K, S = 1 << 20, 64, 2048
    env = bpg.make_vec("BlocksTouch-v0", B, device=0, seed=5)
    o0 = env.reset()
    out = env.step_fused(None, K=K, auto_reset=True, want_actions=True, want_done=True)   # Philox actions
    torch.cuda.synchronize()
    obs, ag, r, succ, done = out["observation"], out["achieved_goal"], out["reward"], out["is_success"], out["done"]
    # -- touch matrix: values in {-1, 0, 1}, symmetric (_check_goal, fetch_env.py:119-124), diagonal never set
    m = ag.view(K, B, 4, 4)
    assert bool(((ag == -1) | (ag == 0) | (ag == 1)).all())
    assert bool((m == m.transpose(2, 3)).all()) and bool((torch.diagonal(m, dim1=2, dim2=3) == -1).all())
    # -- reward is compute_reward(ag, goal) (fetch_env.py:135-143) and only takes the values -0.0 / -1.0
    rr = bpg.compute_reward(ag, env.goal().expand(K, B, 16), None)
    assert bool((rr.view(torch.int32) == r.view(torch.int32)).all())
    assert bool(((r == 0) | (r == -1)).all()) and bool(torch.signbit(r).all())
    # -- TimeLimit: every env finishes exactly one episode in 64 steps (t = 50), at step index 49
    assert bool((done[49] == 1).all()) and int(done.sum()) == B
    # -- the latch: is_success is monotone inside an episode, set exactly by the first success reward, cleared by the reset
    for lo, hi in ((0, 50), (50, 64)):
        s, ok = succ[lo:hi], (r[lo:hi] == 0)
        assert bool((s[1:] >= s[:-1]).all())
        assert bool((s == (torch.cumsum(ok.int(), 0) > 0).float()).all())
    st = env.stats()
    assert st["steps"] == B * K and st["episodes"] == B and st["invalid"] == 0
    assert st["successes"] == float(succ[49].sum()) and st["reward_sum"] == float(r.sum())
    assert bool(torch.isfinite(obs).all())
    # -- the oracle-verified slice: envs 0..S-1 of the big run == the C oracle on the same seeds
    ref = coracle.OracleVecEnv("BlocksTouch-v0", S, seed=5)
    ro, _, _ = ref.reset()
    assert np.array_equal(o0["observation"][:S].cpu().numpy(), ro)
    acts = out["actions"][:, :S].cpu().numpy()
    for k in range(K):
        o, g, rw, sc, _, _ = ref.step(acts[k], auto_reset=True)
        assert np.array_equal(obs[k, :S].cpu().numpy(), o), k
        assert np.array_equal(ag[k, :S].cpu().numpy(), g), k
        assert np.array_equal(r[k, :S].cpu().numpy().view(np.uint32), rw.view(np.uint32)), k
        assert np.array_equal(succ[k, :S].cpu().numpy(), sc), k
    # -- and the last envs of the batch == a small handle created at that global offset (sharding arithmetic)
    tail = bpg.make_vec("BlocksTouch-v0", S, device=0, seed=5, env_index_offset=B - S)
    tail.reset()
    t_out = tail.step_fused(None, K=K, auto_reset=True)
    for key in ("observation", "achieved_goal", "reward", "is_success"):
        assert bool((t_out[key].view(torch.int32) == out[key][:, B - S:].view(torch.int32)).all()), key


def test_one_million_her_transitions():
    import blockpuzzle_gym_b200 as bpg
    B_ep, T, n, S = 20000, 50, 1 << 20, 4096
    env = bpg.make_vec("BlocksTouch-v0", B_ep, device=0, seed=2)
    ep = env.generate_rollouts(None)
    sampler = bpg.make_sample_her_transitions("future", 4, None, seed=6, clip_obs=200.0)
    tr = sampler(ep, n, index_offset=0)
    e, t, ft = tr["ep_idx"].long(), tr["t"].long(), tr["future_t"].long()
    assert int(e.min()) >= 0 and int(e.max()) < B_ep and int(t.min()) >= 0 and int(t.max()) < T
    her = ft >= 0
    assert abs(float(her.float().mean()) - 0.8) < 0.005                      # future_p = 1 - 1/(1 + replay_k), config.py:50
    assert bool((ft[her] > t[her]).all()) and int(ft.max()) <= T
    # gathers are exact copies of the store rows (clip at 200 is the identity on these observations)
    assert bool((tr["o"] == ep["o"][e, t]).all()) and bool((tr["o_2"] == ep["o"][e, t + 1]).all())
    assert bool((tr["u"] == ep["u"][e, t]).all()) and bool((tr["ag"] == ep["ag"][e, t]).all()) and bool((tr["ag_2"] == ep["ag"][e, t + 1]).all())
    assert bool((tr["g"][~her] == ep["g"][e[~her], t[~her]]).all()) and bool((tr["g"][her] == ep["ag"][e[her], ft[her]]).all())
    assert bool((tr["info_is_success"] == ep["info_is_success"][e, t]).all())
    # reward = compute_reward(ag_2, relabelled g); relabelling can only help
    assert bool((bpg.compute_reward(tr["ag_2"], tr["g"], None).view(torch.int32) == tr["r"].view(torch.int32)).all())
    # the oracle-verified slice: transition i is a pure function of (seed, i)
    hp = {k: v.cpu().numpy() for k, v in ep.items() if k != "r"}
    ref = co.make_sample_her_transitions("future", 4, lambda ag_2, g, info: co.compute_reward(ag_2, g, info), seed=6)(
        dict(hp, o_2=hp["o"][:, 1:], ag_2=hp["ag"][:, 1:]), S, index_offset=0)
    for k in ("o", "o_2", "u", "g", "ag", "ag_2", "r"):
        assert np.array_equal(tr[k][:S].cpu().numpy().view(np.uint32), np.ascontiguousarray(ref[k], dtype=np.float32).view(np.uint32)), k
    # sharding by transition range: two half-batches with offsets == the whole batch
    a = sampler(ep, n // 2, index_offset=0, keys=("g", "r"))
    b = sampler(ep, n // 2, index_offset=n // 2, keys=("g", "r"))
    assert bool((torch.cat([a["r"], b["r"]]).view(torch.int32) == tr["r"].view(torch.int32)).all())
    assert bool((torch.cat([a["g"], b["g"]]) == tr["g"]).all())
