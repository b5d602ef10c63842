"""The reference's CALLERS, unmodified, on the CUDA drop-in.

`gym_blocks/rollout.py` RolloutStudent, `gym_blocks/policy_gradient/rollout.py` RolloutStudent and
`gym_blocks/config.py` configure_her are imported from the reference itself (its sources in the build container,
the compiled copy oracle/_ref on the GPU box) under the stub packages of oracle/refharness and handed
`make_env = lambda: blockpuzzle_gym_b200.make(env_id)`.  What they return is compared with
  (a) the same caller running on the reference's own envs (gym.make under the stubs),
  (b) the batched collectors of this repo: bp_rollout on the recorded actions and the closed-loop
      bp_rollout_begin / bp_rollout_step collector driven by the same policy.
"""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from oracle import refharness as rh  # noqa: E402
from ref_callers_common import QuantisedPolicy, QuietLogger, dims_of, fake_pg_self  # noqa: E402

needs_ref = pytest.mark.skipif(not rh.available(), reason="neither /root/reference nor oracle/_ref is present")
T = 50
IDS = ["GripperTouch-v0", "BlocksTouch-v0", "ToppleTower-v0", "BlocksTouchChooseCurriculum-v0", "BlocksTouchVariation-v0"]


def close(a, b):
    return np.allclose(a, b, rtol=1e-6, atol=1e-6)


def has_set_test(name):
    return name not in ("GripperTouch-v0", "ToppleTower-v0")


@needs_ref
@pytest.mark.parametrize("name", IDS)
def test_reference_rollout_student_runs_unmodified_on_the_drop_in(name):
    import blockpuzzle_gym_b200 as bpg
    ro, _, _ = rh.callers()
    B, seed = 2, 40
    dims = dims_of(rh.make(name, seed=0))
    mk = lambda make_env: ro.RolloutStudent(make_env, QuantisedPolicy(dims["o"], dims["g"], name), dims, QuietLogger(), T, rollout_batch_size=B)
    on_ref = mk(lambda: rh.make(name))                  # the reference's own envs
    on_gpu = mk(lambda: bpg.make(name))                 # the drop-in: same constructor call site (rollout.py:33)
    on_ref.seed(seed); on_gpu.seed(seed)
    vec = bpg.make_vec(name, B, device=0, seed=seed)    # the batched collectors, env i <- seed + 1000 i as well
    vec2 = bpg.make_vec(name, B, device=0, seed=seed)
    pol = QuantisedPolicy(dims["o"], dims["g"], name)

    def torch_policy(o, ag, g):
        return torch.from_numpy(pol.get_actions(o.cpu().numpy(), ag.cpu().numpy(), g.cpu().numpy())).cuda()

    for test in (False, False, True) if has_set_test(name) else (False, False):
        a, b = on_ref.generate_rollouts(test=test), on_gpu.generate_rollouts(test=test)
        assert set(a) == set(b) == {"o", "u", "g", "ag", "info_is_success"}
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
        assert np.array_equal(a["u"], b["u"]), "closed-loop actions diverged"
        assert np.array_equal(a["ag"], b["ag"]) and np.array_equal(a["g"], b["g"])
        assert np.array_equal(a["info_is_success"], b["info_is_success"])
        assert close(b["o"], a["o"])
        assert on_ref.success_history[-1] == on_gpu.success_history[-1]
        # (b) the batched open-loop collector on the recorded actions == what the reference's worker collected
        u = torch.from_numpy(np.ascontiguousarray(b["u"].swapaxes(0, 1))).cuda()
        ep = vec.generate_rollouts(u, test=test)
        for k in ("o", "u", "g", "ag", "info_is_success"):
            assert np.array_equal(ep[k].cpu().numpy(), b[k]), k
        # ... and the closed-loop batched collector driven by the same policy
        ep2 = vec2.collect_rollouts(torch_policy, test=test)
        for k in ("o", "u", "g", "ag", "info_is_success"):
            assert np.array_equal(ep2[k].cpu().numpy(), b[k]), k
    if has_set_test(name) and "Curriculum" in name:
        assert on_ref.increase_difficulty() == on_gpu.increase_difficulty() == 1     # rollout.py:66-73


@needs_ref
@pytest.mark.parametrize("name", ["BlocksTouch-v0", "BlocksTouchVariation-v0"])
def test_reference_policy_gradient_worker_on_the_drop_in(name):
    """policy_gradient/rollout.py consumes step()'s reward (:225-231) and accumulates returns (:255-258)."""
    import blockpuzzle_gym_b200 as bpg
    _, _, pg = rh.callers()
    B, seed = 3, 11
    dims = dims_of(rh.make(name, seed=0))
    gamma = 1. - 1. / T
    mk = lambda make_env: pg.RolloutStudent(make_env, QuantisedPolicy(dims["o"], dims["g"], name), None, dims, QuietLogger(), T,
                                            rollout_batch_size=B, gamma=gamma)
    on_ref, on_gpu = mk(lambda: rh.make(name)), mk(lambda: bpg.make(name))
    on_ref.seed(seed); on_gpu.seed(seed)
    a, b = on_ref.generate_rollouts(exploit=True), on_gpu.generate_rollouts(exploit=True)
    assert np.array_equal(a["u"], b["u"]) and np.array_equal(a["ag"], b["ag"])
    assert a["G"].dtype == b["G"].dtype == np.float64
    assert np.array_equal(a["G"].view(np.uint64), b["G"].view(np.uint64))
    # the device mirror of that accumulation on the batched collector's rewards
    vec = bpg.make_vec(name, B, device=0, seed=seed)
    ep = vec.generate_rollouts(torch.from_numpy(np.ascontiguousarray(b["u"].swapaxes(0, 1))).cuda())
    G = bpg.discounted_returns(ep["r"], gamma)
    assert np.array_equal(G.cpu().numpy().view(np.uint64), b["G"].view(np.uint64))
    # RolloutStudent.trim executed from the reference's source against bp_trim
    if name == "BlocksTouchVariation-v0":
        o, g, ag = ep["o"][:, 7].contiguous(), ep["g"][:, 7].contiguous(), ep["ag"][:, 7].contiguous()
        ro_, rg_, rag_ = pg.RolloutStudent.trim(fake_pg_self(name), o.cpu().numpy(), g.cpu().numpy(), ag.cpu().numpy(), 40, 16)
        o_, g_, ag_ = bpg.trim(o, g, ag, 40, 16, name)
        assert np.array_equal(o_.cpu().numpy(), np.asarray(ro_, np.float32))
        assert np.array_equal(g_.cpu().numpy(), rg_) and np.array_equal(ag_.cpu().numpy(), rag_)


@needs_ref
def test_reference_configure_her_on_the_drop_in():
    """config.configure_her (config.py:107-123) builds reward_fun around env.compute_reward(achieved_goal=,
    desired_goal=, info=): with make_env -> the drop-in that is bp_compute_reward; the sampler's rewards must
    equal the reference env's compute_reward on the same rows, and the device sampler's on the same draws."""
    import blockpuzzle_gym_b200 as bpg
    from oracle import callers_oracle as co
    _, cfg, _ = rh.callers()
    name = "BlocksTouch-v0"
    params = dict(make_env=lambda: bpg.make(name), replay_strategy="future", replay_k=4)
    sampler = cfg.configure_her(params)
    vec = bpg.make_vec(name, 64, device=0, seed=3)
    ep = vec.generate_rollouts(None)
    batch = {k: ep[k].cpu().numpy() for k in ("o", "u", "g", "ag", "info_is_success")}
    batch["o_2"], batch["ag_2"] = batch["o"][:, 1:], batch["ag"][:, 1:]
    tr = sampler(batch, 4096)
    ref_env = rh.make(name, seed=0)
    assert tr["r"].dtype == np.float32 and tr["r"].shape == (4096,)
    want = ref_env.compute_reward(achieved_goal=tr["ag_2"], desired_goal=tr["g"], info={})
    assert np.array_equal(tr["r"].view(np.uint32), np.asarray(want, np.float32).view(np.uint32))
    assert (tr["r"] == 0).any() or True
    # same draws on the device sampler (bp_her_sample): identical transitions
    dev = bpg.make_sample_her_transitions("future", 4, None, seed=0)(ep, 4096, index_offset=0)
    for k in ("o", "o_2", "u", "g", "ag", "ag_2", "r"):
        assert np.array_equal(dev[k].cpu().numpy().view(np.uint32), np.ascontiguousarray(tr[k], np.float32).view(np.uint32)), k
    assert co.compute_reward(tr["ag_2"], tr["g"]).tobytes() == tr["r"].tobytes()
