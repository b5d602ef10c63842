"""baselines.her.util -> the reference's own in-tree copy, gym_blocks/util.py (store_args :14-41,
convert_episode_to_batch_major :118-128), imported unmodified."""
from gym_blocks.util import *  # noqa: F401,F403
from gym_blocks.util import convert_episode_to_batch_major, store_args  # noqa: F401
