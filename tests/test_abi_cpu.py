"""CPU tests of the boundary: the C-ABI library loads without a GPU and exports every symbol the
header declares; host-side logic that needs no device; the product never routes through oracle/."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from blockpuzzle_gym_b200 import _lib
    header = open(os.path.join(ROOT, "include", "blockpuzzle_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(bp_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.load()  # resolves each symbol; raises if one is missing
    for name in declared:
        assert hasattr(L, name)
    assert L.bp_abi_version() == 2


def test_env_table_through_the_abi():
    from blockpuzzle_gym_b200 import _lib
    from oracle import coracle
    L = _lib.load()
    for i, name in enumerate(coracle.ENV_IDS):
        assert L.bp_env_id_from_name(name.encode()) == i
        assert L.bp_env_name(i).decode() == name
        assert _lib.env_dims(i) == (coracle.DIMO[i], coracle.DIMG[i], coracle.NBLOCKS[i])
    assert L.bp_env_id_from_name(b"haha-v0") < 0          # the README's id is not registered
    assert b"haha-v0" in L.bp_last_error()
    assert L.bp_env_name(99) is None


def test_state_record_layout_matches_oracle():
    from blockpuzzle_gym_b200 import _lib
    from oracle import coracle
    assert coracle.lib().bpo_sizeof_state() == _lib.STATE_BYTES == coracle.STATE_DTYPE.itemsize == 244


def test_no_cpu_fallback_without_gpu():
    import torch
    import blockpuzzle_gym_b200 as bpg
    from blockpuzzle_gym_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.BlockPuzzleError):
        bpg.make_vec("BlocksTouch-v0", 4)
    h = C.c_void_p()
    rc = _lib.load().bp_create(1, 4, 0, 0, C.byref(h))
    assert rc == _lib.BP_ERR_NO_DEVICE and b"no CPU fallback" in _lib.load().bp_last_error()
    with pytest.raises(KeyError):
        bpg.make_vec("haha-v0", 4)


def test_registry_keeps_reference_ids():
    import blockpuzzle_gym_b200 as bpg
    ids = re.findall(r"id='([^']+)'", """
        id='GripperTouch-v0' id='BlocksTouch-v0' id='ToppleTower-v0' id='BlocksTouchCurriculum-v0'
        id='BlocksTouchChoose-v0' id='BlocksTouchChooseCurriculum-v0' id='BlocksTouchVariation-v0'""")
    assert list(bpg.ENV_IDS) == ids                       # registration order of gym_blocks/__init__.py:6-53
    assert all(v["max_episode_steps"] == 50 for v in bpg.REGISTRY.values())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "blockpuzzle_gym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("CPU oracle", "").replace("the oracle", "").replace("oracle/", "ORACLE_DIR"), f
                assert "import oracle" not in txt and "from oracle" not in txt and "blockphys_oracle" not in txt, f


def test_her_future_p_matches_config():
    import blockpuzzle_gym_b200 as bpg
    assert abs(bpg.make_sample_her_transitions("future", 4, None).future_p - 0.8) < 1e-15   # config.py:49-50
    assert bpg.make_sample_her_transitions("none", 4, None).future_p == 0                    # CLI default train.py:217
