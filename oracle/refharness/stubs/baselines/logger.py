"""baselines.logger subset: the reference's rollout / config modules only log through it."""
import sys

_kv = {}


def info(*args):
    print(*args, file=sys.stderr)


warn = warning = info


def record_tabular(key, val):
    _kv[key] = val


def dump_tabular():
    _kv.clear()


def get_dir():
    return None
