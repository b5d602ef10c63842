"""placeholder: gym_blocks/util.py imports it at module level and only uses it inside TensorFlow helpers"""
