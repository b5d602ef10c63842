"""GPU parity tests of the callers' data path through the C-ABI:
bp_her_sample / ReplayBuffer / Normalizer / update_normalizer / bp_discounted_returns / bp_trim
against oracle/callers_oracle.py.  Gathers, indices, rewards, returns and trims are bit-exact; the
float64 column sums of the normaliser are order-dependent and compared at 1e-12 relative (the north
star's float tolerance is 1e-6)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import callers_oracle as co  # noqa: E402


def _reward_fun(ag_2, g, info):
    return co.compute_reward(ag_2, g, info)


def _rollout(name, B, seed=11):
    import blockpuzzle_gym_b200 as bpg
    env = bpg.make_vec(name, B, device=0, seed=seed)
    ep = env.generate_rollouts(None)          # o, u, g, ag, info_is_success, r  (batch-major CUDA tensors)
    return env, ep


def _np(ep):
    return {k: v.cpu().numpy() for k, v in ep.items()}


def _bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


@pytest.mark.parametrize("name", ["BlocksTouch-v0", "GripperTouch-v0", "BlocksTouchChoose-v0", "BlocksTouchVariation-v0", "ToppleTower-v0"])
@pytest.mark.parametrize("strategy", ["future", "none"])
def test_her_sample_matches_oracle(name, strategy):
    import blockpuzzle_gym_b200 as bpg
    env, ep = _rollout(name, 200)
    n = 5000 + 37                                            # ragged last block
    hp = {k: v for k, v in _np(ep).items() if k != "r"}
    ref_s = co.make_sample_her_transitions(strategy, 4, _reward_fun, seed=21)
    ref = ref_s(dict(hp, o_2=hp["o"][:, 1:], ag_2=hp["ag"][:, 1:]), n, index_offset=5)
    for clip in (0.0, 200.0):
        tr = bpg.make_sample_her_transitions(strategy, 4, None, seed=21, clip_obs=clip)(ep, n, index_offset=5)
        assert np.array_equal(tr["ep_idx"].cpu().numpy(), ref["_ep_idx"]) and np.array_equal(tr["t"].cpu().numpy(), ref["_t"])
        assert np.array_equal(tr["future_t"].cpu().numpy(), ref["_future_t"])
        ro, rg = (ref["o"], ref["g"]) if clip == 0 else co.preprocess_og(ref["o"], ref["ag"], ref["g"], clip)
        ro2 = ref["o_2"] if clip == 0 else co.preprocess_og(ref["o_2"], ref["ag_2"], ref["g"], clip)[0]
        for k, want in (("o", ro), ("o_2", ro2), ("g", rg), ("u", ref["u"]), ("ag", ref["ag"]), ("ag_2", ref["ag_2"]),
                        ("r", ref["r"]), ("info_is_success", ref["info_is_success"])):
            assert np.array_equal(_bits(tr[k].cpu().numpy()), _bits(want.astype(np.float32))), (name, strategy, clip, k)


def test_her_sample_clip_nan_and_empty():
    import blockpuzzle_gym_b200 as bpg
    B, T, dimo, dimg = 7, 5, 12, 9
    rng = np.random.RandomState(0)
    o = (rng.normal(size=(B, T + 1, dimo)) * 300).astype(np.float32)       # values beyond +-200
    o[0, 0, 0] = np.nan
    ep = dict(o=o, u=rng.normal(size=(B, T, 3)).astype(np.float32), g=rng.randint(-1, 2, size=(B, T, dimg)).astype(np.float32),
              ag=rng.randint(-1, 2, size=(B, T + 1, dimg)).astype(np.float32))
    dev = {k: torch.from_numpy(v).cuda() for k, v in ep.items()}
    tr = bpg.make_sample_her_transitions("future", 4, None, seed=1, clip_obs=200.0)(dev, 999, index_offset=0)
    ref = co.make_sample_her_transitions("future", 4, _reward_fun, seed=1)(dict(ep, o_2=o[:, 1:], ag_2=ep["ag"][:, 1:]), 999, index_offset=0)
    want_o, want_g = co.preprocess_og(ref["o"], ref["ag"], ref["g"])
    assert np.array_equal(_bits(tr["o"].cpu().numpy()), _bits(want_o))      # NaN passes through np.clip unchanged
    assert np.array_equal(tr["g"].cpu().numpy(), want_g) and np.array_equal(_bits(tr["r"].cpu().numpy()), _bits(ref["r"]))
    assert float(tr["o"][~torch.isnan(tr["o"])].abs().max()) == 200.0
    empty = bpg.make_sample_her_transitions("future", 4, None)(dev, 0)
    assert empty["o"].shape == (0, dimo) and empty["r"].shape == (0,)


def test_replay_buffer_matches_oracle():
    import blockpuzzle_gym_b200 as bpg
    env, ep = _rollout("BlocksTouch-v0", 64)
    T = 50
    shapes = dict(o=(T + 1, env.dimo), u=(T, 4), g=(T, env.dimg), ag=(T + 1, env.dimg), info_is_success=(T, 1))
    dev_buf = bpg.ReplayBuffer(shapes, 150 * T, T, bpg.make_sample_her_transitions("future", 4, None, seed=9), rng=np.random.RandomState(4))
    ref_buf = co.ReplayBufferOracle(shapes, 150 * T, T, co.make_sample_her_transitions("future", 4, _reward_fun, seed=9), rng=np.random.RandomState(4))
    for rnd in range(4):                                    # 64, 128, then overflow (tail + random), then random slots only
        ep = env.generate_rollouts(None)
        batch = {k: v for k, v in ep.items() if k != "r"}
        dev_buf.store_episode(batch); ref_buf.store_episode(_np(batch))
        assert dev_buf.get_current_size() == ref_buf.get_current_size() and dev_buf.full == ref_buf.full
        for k in shapes:
            assert np.array_equal(dev_buf.buffers[k][:dev_buf.current_size].cpu().numpy(), ref_buf.buffers[k][:ref_buf.current_size]), (rnd, k)
        tr = dev_buf.sample(256, index_offset=rnd * 1000)   # batch_size = 256, config.py:42
        ref = ref_buf.sample(256, index_offset=rnd * 1000)
        for k in ("o", "o_2", "u", "g", "ag", "ag_2", "r", "info_is_success"):
            assert np.array_equal(_bits(tr[k].cpu().numpy()), _bits(ref[k])), (rnd, k)
    assert dev_buf.get_transitions_stored() == ref_buf.get_transitions_stored() == 4 * 64 * T


@pytest.mark.parametrize("name", ["BlocksTouch-v0", "BlocksTouchVariation-v0"])
def test_normalizer_and_fused_store_episode_stats(name):
    import blockpuzzle_gym_b200 as bpg
    env, ep = _rollout(name, 300)
    hp = {k: v for k, v in _np(ep).items() if k != "r"}
    size = env.dimo - (1 if "Variation" in name else 0)
    nz, ref_nz = bpg.Normalizer(size), co.NormalizerOracle(size)
    ref_s = co.make_sample_her_transitions("future", 4, _reward_fun, seed=2)
    dev_s = bpg.make_sample_her_transitions("future", 4, None, seed=2)
    for rnd in range(2):
        ref_tr = co.store_episode_stats(hp, ref_s, ref_nz, name, index_offset=rnd * 77)
        tr = bpg.update_normalizer(ep, dev_s, nz, name, index_offset=rnd * 77)
        assert np.array_equal(_bits(tr["o"].cpu().numpy()), _bits(ref_tr["o"].astype(np.float32)))
        assert np.allclose(nz.total.cpu().numpy()[:size], ref_nz.total_sum, rtol=1e-12, atol=1e-9)
        assert np.allclose(nz.total.cpu().numpy()[size:2 * size], ref_nz.total_sumsq, rtol=1e-12, atol=1e-9)
        assert nz.total[-1].item() == ref_nz.total_count[0]
        assert np.allclose(nz.mean.cpu().numpy(), ref_nz.mean, rtol=1e-6, atol=1e-7) and np.allclose(nz.std.cpu().numpy(), ref_nz.std, rtol=1e-6, atol=1e-7)
    # the stand-alone reduction (Normalizer.update) over a big ragged matrix, incl. the col0 = 1 rule
    x = torch.randn(100003, size + 1, device="cuda") * 5
    nz2, ref2 = bpg.Normalizer(size), co.NormalizerOracle(size)
    nz2.update(x, col0=1); nz2.update(x[:17], col0=1); nz2.recompute_stats()
    xn = x.cpu().numpy()[:, 1:]
    ref2.update(xn); ref2.update(xn[:17]); ref2.recompute_stats()
    assert np.allclose(nz2.mean.cpu().numpy(), ref2.mean, rtol=1e-6, atol=1e-7) and np.allclose(nz2.std.cpu().numpy(), ref2.std, rtol=1e-6)
    v = x[:5, 1:]
    assert np.allclose(nz2.normalize(v, 5.0).cpu().numpy(), ref2.normalize(v.cpu().numpy(), 5.0), rtol=1e-5, atol=1e-6)


def test_discounted_returns_bit_exact():
    import blockpuzzle_gym_b200 as bpg
    env, ep = _rollout("BlocksTouch-v0", 500)
    T = 50
    gamma = 1. - 1. / T
    G = bpg.discounted_returns(ep["r"], gamma)
    ref = co.discounted_returns(ep["r"].cpu().numpy().T, gamma).T          # the oracle is time-major like the reference loop
    assert G.dtype == torch.float64 and np.array_equal(G.cpu().numpy().view(np.uint64), np.ascontiguousarray(ref).view(np.uint64))
    # arbitrary rewards; odd and even T (the kernel pairs outputs t and T - 1 - t: an odd T has a self-paired middle),
    # T = 1, 2, the maximum T = 128, and episode counts that leave a partial last block
    for Bn, Tn in ((33, 7), (5, 1), (9, 2), (130, 50), (67, 128), (1, 51)):
        r = torch.randn(Bn, Tn, device="cuda")
        assert np.array_equal(bpg.discounted_returns(r, 0.9).cpu().numpy().view(np.uint64),
                              np.ascontiguousarray(co.discounted_returns(r.cpu().numpy().T, 0.9).T).view(np.uint64)), (Bn, Tn)


def test_trim_bit_exact():
    import blockpuzzle_gym_b200 as bpg
    env, ep = _rollout("BlocksTouchVariation-v0", 300)
    for t in (0, 17, 50):
        o, g, ag = ep["o"][:, t].contiguous(), ep["g"][:, min(t, 49)].contiguous(), ep["ag"][:, t].contiguous()
        o_, g_, ag_ = bpg.trim(o, g, ag, 40, 16, "BlocksTouchVariation-v0")
        ro, rg, rag = co.trim(o.cpu().numpy(), g.cpu().numpy(), ag.cpu().numpy(), 40, 16, "BlocksTouchVariation-v0")
        assert np.array_equal(_bits(o_.cpu().numpy()), _bits(ro.astype(np.float32)))
        assert np.array_equal(g_.cpu().numpy(), rg) and np.array_equal(ag_.cpu().numpy(), rag)
    env4, ep4 = _rollout("ToppleTower-v0", 50)
    o, g, ag = ep4["o"][:, 3].contiguous(), ep4["g"][:, 3].contiguous(), ep4["ag"][:, 3].contiguous()
    o_, g_, ag_ = bpg.trim(o, g, ag, 40, 16, "ToppleTower-v0")
    ro, rg, rag = co.trim(o.cpu().numpy(), g.cpu().numpy(), ag.cpu().numpy(), 40, 16, "ToppleTower-v0")
    assert np.array_equal(o_.cpu().numpy(), ro) and np.array_equal(g_.cpu().numpy(), rg) and np.array_equal(ag_.cpu().numpy(), rag)
    same = bpg.trim(o_, g_, ag_, 40, 16, "ToppleTower-v0")
    assert same[0] is o_                                                     # nothing to trim (rollout.py:107-108)
