"""blockpuzzle_gym_b200 -- B200-native batched implementation of the gym_blocks env hot path.

Keeps the ids registered by the reference (gym_blocks/__init__.py:6-53, all with
max_episode_steps=50) and the reset/step/compute_reward signatures.

    make(env_id)            -> single-env gym-style object (what gym.make returns in the reference)
    make_vec(env_id, B)     -> VecBlocksEnv, B envs on one GPU, torch CUDA tensors
    compute_reward(ag, g, info)           batched BlocksEnv.compute_reward
    make_sample_her_transitions(...)      HER sampler mirror (her.py)

Importing this package never touches the GPU; creating an env without a CUDA
device or without the built extension raises (there is no CPU fallback).
"""
from ._lib import ENV_IDS, BlockPuzzleError

REGISTRY = {name: dict(max_episode_steps=50, kwargs={"reward_type": "sparse"}) for name in ENV_IDS}


def make(env_id, **kw):
    from .vec_env import GymBlocksEnv
    if env_id not in REGISTRY:
        raise KeyError(f"No registered env with id: {env_id}")
    return GymBlocksEnv(env_id, **kw)


def make_vec(env_id, num_envs, **kw):
    from .vec_env import VecBlocksEnv
    if env_id not in REGISTRY:
        raise KeyError(f"No registered env with id: {env_id}")
    return VecBlocksEnv(env_id, num_envs, **kw)


def compute_reward(achieved_goal, desired_goal, info=None):
    from .vec_env import compute_reward as _cr
    return _cr(achieved_goal, desired_goal, info)


def register_with_gym():
    """Register the seven ids with an installed `gym` so gym.make(env_id) resolves here."""
    from gym.envs.registration import register
    for name, spec in REGISTRY.items():
        register(id=name, entry_point="blockpuzzle_gym_b200.vec_env:GymBlocksEnv",
                 kwargs={"env_name": name}, max_episode_steps=spec["max_episode_steps"])


def make_sample_her_transitions(*a, **kw):
    from .her import make_sample_her_transitions as _m
    return _m(*a, **kw)


def __getattr__(name):
    # lazy: ReplayBuffer / Normalizer / update_normalizer (her.py), discounted_returns / trim (pg.py)
    if name in ("ReplayBuffer", "Normalizer", "update_normalizer"):
        from . import her
        return getattr(her, name)
    if name in ("discounted_returns", "trim"):
        from . import pg
        return getattr(pg, name)
    raise AttributeError(name)


__all__ = ["ENV_IDS", "REGISTRY", "BlockPuzzleError", "make", "make_vec", "compute_reward",
           "register_with_gym", "make_sample_her_transitions", "ReplayBuffer", "Normalizer",
           "update_normalizer", "discounted_returns", "trim"]
