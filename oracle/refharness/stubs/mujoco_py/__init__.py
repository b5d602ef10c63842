"""Stub `mujoco_py`: the fake simulator of oracle/refharness/fake_mujoco.py (BlockPhys, the C model of
oracle/blockphys_oracle.c, behind the MjSim calls the reference makes).  TEST INFRASTRUCTURE."""
from oracle.refharness.fake_mujoco import (MjSim, MjSimState, MujocoException, MjViewer,  # noqa: F401
                                           load_model_from_path)
from . import modder  # noqa: F401
