"""The CUDA path (through the C-ABI) against the REFERENCE'S OWN outputs.

tests/golden/ref_<id>.npz are traces recorded from the unmodified reference (tests/golden/make_ref_golden.py).
The CUDA library replays the committed programme -- reset / step / set_test / increase_difficulty, scripted
contact episodes included -- and must reproduce the reference's touch matrix, goal, reward bits, success latch,
done flags, num_objs, Philox draw counters and binary32 sim state bit for bit, its float64 observations within
1e-6.  Where the reference itself is importable (oracle/_ref on the GPU box) its callers are run UNMODIFIED on
the drop-in env in tests/test_gpu_ref_callers.py.
"""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_scenario as sc  # noqa: E402


class CudaDriver(object):
    """VecBlocksEnv behind the scenario's driver interface (every call goes through libblockpuzzle_b200.so)."""

    def __init__(self, name, num_envs=sc.NUM_ENVS, seed=sc.SEED):
        import blockpuzzle_gym_b200 as bpg
        self.env = bpg.make_vec(name, num_envs, device=0, seed=seed)
        self.dimo, self.dimg = self.env.dimo, self.env.dimg

    @staticmethod
    def _np(d):
        return d["observation"].cpu().numpy(), d["achieved_goal"].cpu().numpy(), d["desired_goal"].cpu().numpy()

    def reset(self):
        return self._np(self.env.reset())

    def step(self, a):
        obs, r, done, info = self.env.step(torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda())
        return (obs["observation"].cpu().numpy(), obs["achieved_goal"].cpu().numpy(), r.cpu().numpy(),
                info["is_success"].cpu().numpy(), done.cpu().numpy())

    def set_test(self):
        return self._np(self.env.set_test())

    def increase_difficulty(self):
        return self.env.increase_difficulty()

    def get_difficulty(self):
        return self.env.get_difficulty()

    def state(self):
        return self.env.get_state()


@pytest.mark.parametrize("name", sc.ENV_IDS)
def test_cuda_replays_the_reference_trace(name):
    z = np.load(os.path.join(HERE, "golden", "ref_%s.npz" % name))
    want = sc.unpack(z)
    got = sc.replay(CudaDriver(name), want)
    sc.compare(got, want, who="CUDA vs reference, " + name)


@pytest.mark.parametrize("name", sc.ENV_IDS)
def test_cuda_equals_c_oracle_on_the_reference_programme(name):
    """Same programme, CUDA against the C restatement: every float bit-identical (the oracle is binary32 too)."""
    z = np.load(os.path.join(HERE, "golden", "ref_%s.npz" % name))
    want = sc.unpack(z)
    got = sc.replay(CudaDriver(name), want)
    orc = sc.replay(sc.OracleDriver(name), want)
    sc.compare(got, orc, who="CUDA vs C oracle, " + name, exact_floats=True)
