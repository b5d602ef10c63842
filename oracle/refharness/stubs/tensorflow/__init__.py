"""Import shim: gym_blocks/util.py does `import tensorflow as tf` at module level; none of the functions the
harness calls (store_args, convert_episode_to_batch_major) touches it."""
