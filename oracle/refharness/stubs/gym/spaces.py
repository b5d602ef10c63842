"""gym.spaces [upstream, recalled]: the Box / Dict subset RobotEnv.__init__ builds (robot_env.py:39-44)."""
import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)


class Box(Space):
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        if shape is None:
            assert low.shape == high.shape
            shape = low.shape
        else:
            assert np.isscalar(low) and np.isscalar(high)
            low = low + np.zeros(shape)
            high = high + np.zeros(shape)
        if dtype is None:
            dtype = np.uint8 if (high == 255).all() else np.float32
        self.low = low.astype(dtype)
        self.high = high.astype(dtype)
        Space.__init__(self, shape, dtype)
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)

    def sample(self):
        return self._rng.uniform(low=self.low, high=self.high + (0 if self.dtype.kind == "f" else 1), size=self.low.shape).astype(self.dtype)

    def contains(self, x):
        return x.shape == self.shape and (x >= self.low).all() and (x <= self.high).all()


class Dict(Space):
    def __init__(self, spaces):
        if isinstance(spaces, dict):
            spaces = dict(sorted(list(spaces.items())))
        self.spaces = spaces
        Space.__init__(self, None, None)

    def __getitem__(self, k):
        return self.spaces[k]

    def sample(self):
        return dict((k, space.sample()) for k, space in self.spaces.items())
