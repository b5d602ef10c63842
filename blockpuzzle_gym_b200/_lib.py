"""ctypes binding of libblockpuzzle_b200.so (include/blockpuzzle_b200.h).

There is no CPU fallback: if the CUDA library is missing or no GPU is present the
product path raises.  Nothing here imports or calls oracle/.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BP_LIB_PATH: an alternative build of the same library (A/B measurements of kernel variants on one box)
LIB_PATH = os.environ.get("BP_LIB_PATH") or os.path.join(_HERE, "libblockpuzzle_b200.so")

ENV_IDS = (
    "GripperTouch-v0",
    "BlocksTouch-v0",
    "ToppleTower-v0",
    "BlocksTouchCurriculum-v0",
    "BlocksTouchChoose-v0",
    "BlocksTouchChooseCurriculum-v0",
    "BlocksTouchVariation-v0",
)

BP_OK = 0
BP_ERR_INVALID_ARG = -1
BP_ERR_CUDA = -2
BP_ERR_NOT_IMPLEMENTED = -3
BP_ERR_NO_DEVICE = -4
BP_ERR_NO_ATTRIBUTE = -5
BP_NUM_STATS = 8
STAT_NAMES = ("episodes", "successes", "steps", "invalid", "reward_sum", "worker_steps", "sched_iterations", "sched_passes")
STATE_BYTES = 244

# every symbol include/blockpuzzle_b200.h declares: (name, restype, argtypes)
_vp, _i, _i64, _u64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float
SYMBOLS = {
    "bp_abi_version": (_i, []),
    "bp_last_error": (C.c_char_p, []),
    "bp_env_dims": (_i, [_i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "bp_env_id_from_name": (_i, [C.c_char_p]),
    "bp_env_name": (C.c_char_p, [_i]),
    "bp_create": (_i, [_i, _i64, _i, _u64, C.POINTER(_vp)]),
    "bp_destroy": (_i, [_vp]),
    "bp_num_envs": (_i64, [_vp]),
    "bp_seed": (_i, [_vp, _u64, _vp]),
    "bp_reset": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "bp_step": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "bp_step_host": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "bp_rollout": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bp_rollout_begin": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "bp_rollout_step": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bp_set_test": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "bp_increase_difficulty": (_i, [_vp, C.POINTER(_i)]),
    "bp_get_difficulty": (_i, [_vp, C.POINTER(_i)]),
    "bp_set_option": (_i, [_vp, C.c_char_p, _i]),
    "bp_get_ranges": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "bp_set_ranges": (_i, [_vp, C.c_double, C.c_double]),
    "bp_get_state": (_i, [_vp, _vp, _vp]),
    "bp_set_state": (_i, [_vp, _vp, _vp]),
    "bp_stats_ptr": (_i, [_vp, C.POINTER(_vp)]),
    "bp_stats_reset": (_i, [_vp, _vp]),
    "bp_compute_reward": (_i, [_vp, _vp, _i64, _i, _vp, _vp]),
    "bp_her_relabel": (_i, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, _i64, _f, _u64, _i64,
                            _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bp_her_sample": (_i, [_vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _i64, _f, _f, _u64, _i64,
                           _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bp_moments": (_i, [_vp, _i64, C.c_int32, C.c_int32, C.c_int32, _f, _vp, _vp]),
    "bp_discounted_returns": (_i, [_vp, _i64, C.c_int32, _vp, _vp, _vp]),
    "bp_trim": (_i, [_vp, _vp, _vp, _i64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]),
}

_lib = None


class BlockPuzzleError(RuntimeError):
    pass


def load():
    """Load the CUDA extension; raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BlockPuzzleError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(blockpuzzle_gym_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.bp_abi_version() != 2:
            raise BlockPuzzleError("ABI version mismatch")
        _lib = L
    return _lib


def check(rc):
    if rc == BP_OK:
        return
    msg = load().bp_last_error().decode()
    if rc == BP_ERR_NOT_IMPLEMENTED:
        raise NotImplementedError(msg)
    if rc == BP_ERR_NO_ATTRIBUTE:
        raise AttributeError(msg)
    raise BlockPuzzleError(f"blockpuzzle_b200 error {rc}: {msg}")


def env_dims(env_id):
    o, g, n = C.c_int(), C.c_int(), C.c_int()
    check(load().bp_env_dims(env_id, C.byref(o), C.byref(g), C.byref(n)))
    return o.value, g.value, n.value
