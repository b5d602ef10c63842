"""Import shim: gym_blocks/config.py does `from ddpg import DDPG` at module level (the trainer is out of scope);
configure_her (config.py:107-123), the only function the harness calls, never touches it."""


class DDPG(object):
    def __init__(self, *a, **kw):
        raise NotImplementedError("the DDPG trainer is out of scope of the env hot path")
