"""Philox4x32-10 replay of the two numpy RNGs the reference draws from (TEST INFRASTRUCTURE).

The reference spawns blocks from the env's `self.np_random` (robot_env.py:53-55; fetch_env.py:89-90,
332,380,479,647,723,782) and from the process-global `np.random` (fetch_env.py:390,392,490,492,507,509,
734,736).  Both are replaced by counter-based generators so that the reference,
the CPU oracle and the CUDA kernels consume identical draws:

    words = Philox4x32-10(counter = (draw, episode, stream, 0), key = the env's 64-bit seed)

stream 0 = self.np_random, stream 1 = the global np.random (per env here).  One RandomState call = one
Philox block.  `episode` is the number of reset() calls that came before (set by begin_episode(), which
the harness' gym.Wrapper calls at every reset()); `draw` restarts at 0 with every episode.

Values are returned as numpy would return them -- float64 -- built from the 24-bit fractions / the
binary32 Box-Muller pair of the BlockPhys spec (oracle/blockphys_oracle.c: bpo_u01, bpo_normal2):
uniform(low, high) = low + (high - low) * u exactly as numpy's legacy uniform computes it.
"""
import ctypes as C

import numpy as np

from .. import coracle

_M32 = 0xFFFFFFFF


class PhiloxRandomState(object):
    def __init__(self, seed=0, stream=0, parent=None):
        self.L = coracle.lib()
        self.seed_value = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.stream = stream
        self.parent = parent          # the stream-0 generator that owns the episode counter
        self.episode = 0
        self.draws = 0
        self._out = (C.c_uint32 * 4)()
        # the env's "global np.random" stand-in: same key, stream 1, same episode counter
        self.global_stream = PhiloxRandomState(seed, 1, parent=self) if stream == 0 else None

    # ---- episode bookkeeping (the harness wrapper calls this at the start of every reset())
    def begin_episode(self, episode):
        self.episode = int(episode)
        self.draws = 0
        if self.global_stream is not None:
            self.global_stream.draws = 0

    def _block(self):
        ep = self.parent.episode if self.parent is not None else self.episode
        self.L.bpo_philox4x32(self.draws & _M32, ep & _M32, self.stream, 0,
                              self.seed_value & _M32, (self.seed_value >> 32) & _M32, self._out)
        self.draws += 1
        return self._out

    @staticmethod
    def _u01(x):
        return float(x >> 8) * 5.9604644775390625e-08   # exact: 24-bit fraction

    # ---- the numpy.random.RandomState calls the reference makes
    def uniform(self, low=0.0, high=1.0, size=None):
        w = self._block()
        low, high = float(low), float(high)
        if size is None:
            return low + (high - low) * self._u01(w[0])
        assert size == 2 or tuple(np.atleast_1d(size)) == (2,), "the reference only draws scalars and pairs"
        return np.array([low + (high - low) * self._u01(w[0]), low + (high - low) * self._u01(w[1])], dtype=np.float64)

    def normal(self, loc=0.0, scale=1.0, size=None):
        assert size == 2 and loc == 0.0 and scale == 1.0, "the reference only calls normal(size=2)"
        w = self._block()
        z0, z1 = C.c_float(), C.c_float()
        self.L.bpo_normal2(w[0], w[1], C.byref(z0), C.byref(z1))
        return np.array([z0.value, z1.value], dtype=np.float64)

    def randint(self, low, high=None, size=None):
        assert high is None and size is None
        w = self._block()
        return (int(w[0]) * int(low)) >> 32

    def seed(self, seed=None):
        self.__init__(0 if seed is None else seed, self.stream, self.parent)
