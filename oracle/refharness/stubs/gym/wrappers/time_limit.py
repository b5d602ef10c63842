"""gym.wrappers.TimeLimit [upstream, recalled]: counts steps, sets done at max_episode_steps; exposes
_max_episode_steps (read at config.py:79-80)."""
from ..core import Wrapper


class TimeLimit(Wrapper):
    def __init__(self, env, max_episode_seconds=None, max_episode_steps=None):
        super(TimeLimit, self).__init__(env)
        self._max_episode_seconds = max_episode_seconds
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0
        self._episode_started_at = None

    def _past_limit(self):
        return self._max_episode_steps is not None and self._max_episode_steps <= self._elapsed_steps

    def step(self, action):
        assert self._episode_started_at is not None, "Cannot call env.step() before calling reset()"
        observation, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._past_limit():
            done = True
        return observation, reward, done, info

    def reset(self):
        self._episode_started_at = 0
        self._elapsed_steps = 0
        return self.env.reset()
