from . import tf_util  # noqa: F401


def set_global_seeds(seed):
    return None
