"""CPU tests of oracle/callers_oracle.py: the numpy restatements of the HER
transition sampler, replay buffer, normaliser, discounted returns and trim are pinned against what
can be pinned without the (absent) upstream packages: the C oracle's independent restatement of the
same HER draw, closed forms, and the env oracle's own observation layout."""
import numpy as np
import pytest

from oracle import callers_oracle as co
from oracle import coracle


def _episodes(name="BlocksTouch-v0", B=24, T=50, seed=3):
    """A real episode batch from the C oracle, laid out batch-major like convert_episode_to_batch_major."""
    ref = coracle.OracleVecEnv(name, B, seed=seed)
    o0, ag0, g0 = ref.reset()
    rng = np.random.RandomState(seed)
    obs, ags, us, succ, rew = [o0], [ag0], [], [], []
    for t in range(T):
        a = rng.uniform(-1, 1, size=(B, 4)).astype(np.float32)
        o, ag, r, s, _, _ = ref.step(a)
        obs.append(o); ags.append(ag); us.append(a); succ.append(s); rew.append(r)
    sw = lambda x: np.ascontiguousarray(np.stack(x).swapaxes(0, 1))
    return dict(o=sw(obs), u=sw(us), g=np.repeat(g0[:, None, :], T, 1), ag=sw(ags), info_is_success=sw(succ)[..., None]), sw(rew)


def test_philox_vec_matches_c_oracle():
    c = np.arange(5, dtype=np.uint64) * 977 + 3
    w = co.philox4x32_vec(c, c >> np.uint64(3), 3, 0, 0x1234, 0x9)
    for i in range(5):
        assert tuple(int(x[i]) for x in w) == tuple(coracle.philox4x32(int(c[i]), int(c[i] >> np.uint64(3)), 3, 0, 0x1234, 0x9))


@pytest.mark.parametrize("strategy,fp", [("future", 0.8), ("none", 0.0)])
def test_her_sampler_agrees_with_c_restatement(strategy, fp):
    ep, _ = _episodes()
    n = 5000
    sampler = co.make_sample_her_transitions(strategy, 4, lambda ag_2, g, info: co.compute_reward(ag_2, g, info), seed=77)
    batch = dict(ep, o_2=ep["o"][:, 1:], ag_2=ep["ag"][:, 1:])
    tr = sampler(batch, n, index_offset=1000)
    ref = coracle.her_relabel(ep["ag"], ep["g"], n, fp, 77, 1000)
    assert np.array_equal(tr["_ep_idx"], ref["ep_idx"]) and np.array_equal(tr["_t"], ref["t"])
    assert np.array_equal(tr["_future_t"], ref["future_t"])
    assert np.array_equal(tr["g"], ref["g"]) and np.array_equal(tr["ag_2"], ref["ag_2"])
    assert np.array_equal(tr["r"].view(np.uint32), ref["r"].view(np.uint32))          # incl. the sign of -0.0
    e, t = tr["_ep_idx"], tr["_t"]
    assert np.array_equal(tr["o"], ep["o"][e, t]) and np.array_equal(tr["o_2"], ep["o"][e, t + 1])
    assert np.array_equal(tr["u"], ep["u"][e, t]) and np.array_equal(tr["ag"], ep["ag"][e, t])
    her = tr["_future_t"] >= 0
    assert abs(her.mean() - fp) < 0.03
    assert np.array_equal(tr["g"][~her], ep["g"][e[~her], t[~her]])
    if fp:
        assert (tr["_future_t"][her] > t[her]).all() and tr["_future_t"].max() <= 50


def test_replay_buffer_storage_policy():
    T = 5
    shapes = dict(o=(T + 1, 3), u=(T, 2), g=(T, 4), ag=(T + 1, 4))
    buf = co.ReplayBufferOracle(shapes, 10 * T, T, None, rng=np.random.RandomState(1))
    mk = lambda n, v: {k: np.full((n, *s), v, np.float32) for k, s in shapes.items()}
    buf.store_episode(mk(4, 1)); assert buf.get_current_episode_size() == 4 and not buf.full
    buf.store_episode(mk(4, 2)); assert buf.get_current_size() == 8 * T
    buf.store_episode(mk(4, 3))                       # 2 fill the tail, 2 overwrite random earlier slots
    assert buf.full and buf.get_transitions_stored() == 12 * T
    assert (buf.buffers["o"][8:10] == 3).all() and (buf.buffers["o"][:8] == 3).reshape(8, -1).all(1).sum() == 2
    buf.store_episode(mk(3, 4))                       # full: random slots only
    assert buf.current_size == 10 and (buf.buffers["u"] == 4).reshape(10, -1).all(1).sum() in (2, 3)
    buf.clear_buffer(); assert buf.get_current_size() == 0


def test_normalizer_matches_plain_moments():
    rng = np.random.RandomState(0)
    x = rng.normal(2.0, 3.0, size=(1000, 7)).astype(np.float32)
    nz = co.NormalizerOracle(7)
    nz.update(x[:400]); nz.update(x[400:]); nz.recompute_stats()
    # total_count starts at 1 upstream, so the mean is sum / (n + 1)
    assert np.allclose(nz.mean, x.astype(np.float64).sum(0) / 1001, rtol=1e-6)
    var = (x.astype(np.float64) ** 2).sum(0) / 1001 - (x.astype(np.float64).sum(0) / 1001) ** 2
    assert np.allclose(nz.std, np.sqrt(var), rtol=1e-6)
    assert np.all(co.NormalizerOracle(3).std == 1) and co.NormalizerOracle(3, eps=0.5).eps == 0.5


def test_store_episode_stats_variation_drops_block_count_column():
    ep, _ = _episodes("BlocksTouchVariation-v0", B=8)
    sampler = co.make_sample_her_transitions("future", 4, lambda ag_2, g, info: co.compute_reward(ag_2, g, info), seed=5)
    nz = co.NormalizerOracle(86)
    tr = co.store_episode_stats(ep, sampler, nz, "BlocksTouchVariation-v0")
    assert tr["o"].shape == (8 * 50, 87)
    assert np.allclose(nz.mean, tr["o"][:, 1:].astype(np.float64).sum(0) / (8 * 50 + 1), rtol=1e-6, atol=1e-9)


def test_discounted_returns_closed_form_and_recurrence():
    T, B = 50, 6
    gamma = 1. - 1. / T                                                    # policy_gradient/config.py:84
    _, rew = _episodes(B=B)
    r = rew.T                                                              # time-major [T][B]
    G = co.discounted_returns(r, gamma)
    assert G.dtype == np.float64 and G.shape == (T, B)
    direct = np.array([[sum(gamma ** (t - t0) * float(r[t, b]) for t in range(t0, T)) for b in range(B)] for t0 in range(T)])
    assert np.allclose(G, direct, rtol=1e-12)
    assert np.allclose(G[:-1] - gamma * G[1:], r[:-1], atol=1e-12)         # G_t = r_t + gamma * G_{t+1}
    assert np.array_equal(G[-1], r[-1].astype(np.float64))
    allfail = co.discounted_returns(-np.ones((T, 1), np.float32), gamma)   # -(1 - gamma^T) / (1 - gamma)
    assert abs(allfail[0, 0] + (1 - gamma ** T) / (1 - gamma)) < 1e-9


def test_trim_variation_gives_the_two_block_layout():
    ep, _ = _episodes("BlocksTouchVariation-v0", B=16)
    o, g, ag = ep["o"][:, 7], ep["g"][:, 7], ep["ag"][:, 7]
    o_, g_, ag_ = co.trim(o, g, ag, 40, 16, "BlocksTouchVariation-v0")
    assert o_.shape == (16, 40) and g_.shape == (16, 16) and ag_.shape == (16, 16)
    assert np.array_equal(o_[:, :10], o[:, 1:11])                          # block count dropped
    assert np.array_equal(o_[:, 10:25], o[:, 11:26]) and np.array_equal(o_[:, 25:40], o[:, 30:45])   # GREEN, BLUE = blocks 0, 1
    keep = [i * 6 + j for i in range(4) for j in range(4)]
    assert np.array_equal(g_, g[:, keep]) and np.array_equal(ag_, ag[:, keep])
    # the trimmed goal is BlocksTouch's goal matrix (fetch_env.py:260-273 on [GREY,GREY,GREEN,BLUE])
    want = np.zeros((4, 4)); want[2, 3] = want[3, 2] = 1
    assert np.array_equal(g_[0].reshape(4, 4), want)
    same = co.trim(o_, g_, ag_, 40, 16, "BlocksTouchVariation-v0")
    assert same[0] is o_ or np.array_equal(same[0], o_)                    # nothing to trim (rollout.py:107-108)
    # a non-Variation env with more blocks: plain column cut (rollout.py:169)
    ep4, _ = _episodes("ToppleTower-v0", B=4)
    o4_, g4_, _ = co.trim(ep4["o"][:, 0], ep4["g"][:, 0], ep4["ag"][:, 0], 40, 16, "ToppleTower-v0")
    assert np.array_equal(o4_, ep4["o"][:, 0, :40]) and g4_.shape == (4, 16)
