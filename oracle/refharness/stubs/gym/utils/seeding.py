"""gym.utils.seeding [upstream]: np_random(seed) -> (RandomState, seed).

Upstream builds a Mersenne-Twister RandomState from a hashed seed.  The replay harness instead hands
out the Philox4x32-10 generator of oracle/refharness/philox_rng.py (stream 0 = this env's
self.np_random), keyed by the seed itself, so that the reference, the CPU
oracle and the CUDA kernels consume the same draws."""
from oracle.refharness.philox_rng import PhiloxRandomState


def np_random(seed=None):
    seed = 0 if seed is None else int(seed)
    return PhiloxRandomState(seed, stream=0), seed
