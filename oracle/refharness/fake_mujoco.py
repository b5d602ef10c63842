"""Fake `mujoco_py` for running the unmodified reference (TEST INFRASTRUCTURE).

`MjSim` puts BlockPhys -- the normative C model of oracle/blockphys_oracle.c, i.e. exactly what the CPU
oracle and the CUDA kernels implement in the reference's `sim.step()` slot (robot_env.py:60) -- behind
the `mujoco_py` calls gym_blocks makes:

    load_model_from_path / MjSim(model, nsubsteps)          robot_env.py:24-25, fetch_env.py:553-554
    sim.step / forward / get_state / set_state              robot_env.py:35,60; fetch_env.py:152,248,253,289,297,670,673
    sim.model.geom_id2name / ngeom / opt.timestep           fetch_env.py:107,190,284; robot_env.py:47-48
    sim.data.get_site_xpos / xvelp / xmat / xvelr           fetch_env.py:189-208,292,301-302
    sim.data.get_joint_qpos / set_joint_qpos                fetch_env.py:150-151,287,333-336,...
    sim.data.set_mocap_pos / set_mocap_quat                 fetch_env.py:294-295
    sim.data.ncon / contact[i].geom1 / geom2                fetch_env.py:158-162
    (+ what the stubbed gym.envs.robotics.utils walk: ctrl, mocap_pos, mocap_quat, qpos by joint name)

MuJoCo semantics that the reference's quirks depend on are kept:
  * site positions / rotation matrices come from the kinematics cache of the last forward() or step();
    set_joint_qpos / set_state do not refresh it (set_test() returns a stale observation, appendix A4);
    velocities are computed from the live qvel, as mujoco_py's get_site_xvelp does;
  * all values handed out are float64 arrays (copies), the state inside is BlockPhys' binary32.

E0 (`_env_setup`, fetch_env.py:283-302) is the one place where MuJoCo's own transient matters: the arm
is driven to the mocap target and the cubes are pushed out of the table by ten settling steps.  The fake
defines the outcome instead of simulating the transient: until the first get_state(), set_mocap_pos
places the welded gripper on the mocap (rigid weld) and step() projects every cube onto its support
(table top, or the cube below when its centre is over that cube's footprint).  The result is compared
bit for bit with the oracle's pinned `bpo_sim_init` in tests/test_ref_pin.py.
"""
import ctypes as C
import os

import numpy as np

from .. import coracle
from . import mjcf


class MujocoException(Exception):
    pass


class MjViewer(object):
    def __init__(self, sim):
        raise MujocoException("no viewer in the replay harness")


# ---- BlockPhys constants the fake needs on the Python side (checked against the MJCF at load time)
_HB = 0.025
_Z_REST = np.float32(0.485)
_Z_FLOOR = np.float32(0.025)
_TBL = (1.3, 0.75, 0.25, 0.35)
# the arm's home pose is not modelled: the grip site starts where `_env_setup`'s offset
# (fetch_env.py:292) leads to the pinned initial_gripper_xpos (1.3419, 0.7491, 0.5347)f
_GRIP0 = np.array([np.float32(1.3419), np.float32(0.7491), np.float32(0.5347)], dtype=np.float64)
_SETUP_OFFSET = np.array([-0.498, 0.005, -0.431 + 0.2])


class _CBlock(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("c", C.c_float), ("s", C.c_float), ("vel", C.c_float * 3), ("w", C.c_float)]


class _CSim(C.Structure):   # bpo_sim, oracle/blockphys_oracle.h
    _fields_ = [("g", C.c_float * 3), ("gv", C.c_float * 3), ("q", C.c_float * 2), ("qv", C.c_float * 2),
                ("m", C.c_float * 3), ("ctrl", C.c_float * 2), ("blk", _CBlock * 4),
                ("nblocks", C.c_int32), ("block_gripper", C.c_int32), ("contacts", C.c_uint32)]


class _Opt(object):
    def __init__(self, timestep):
        self.timestep = timestep


class PyMjModel(object):
    def __init__(self, digest, path):
        self.digest = digest
        self.path = path
        self.opt = _Opt(float(digest["timestep"]))
        self.geom_names = list(digest["geom_names"])
        self.ngeom = len(self.geom_names)
        self.nblocks = int(digest["nblocks"])
        self.joint_names = list(digest["joint_names"])
        self.nmocap = 1
        # finger position actuators, in actuator order (2blocks.xml:39-42): l then r
        self.actuator_joint = [a["joint"] for a in digest["actuators"]]
        self.actuator_biastype = [1] * len(self.actuator_joint)          # mjBIAS_AFFINE: position servos
        self.weld_pairs = [(0, "robot0:gripper_link")]                    # shared.xml:48-50
        # geometry BlockPhys pins must be what the scene says
        assert digest["cube_half"] == [0.025, 0.025, 0.025], digest["cube_half"]
        assert digest["table_half"] == [0.25, 0.35, 0.23] and digest["table_body_pos"] == [0.25, 0.35, 0.23]
        assert digest["finger_half"] == [0.0385, 0.007, 0.0135]
        assert abs(self.opt.timestep - 0.002) < 1e-15

    def geom_id2name(self, i):
        return self.geom_names[i]

    def geom_name2id(self, name):
        return self.geom_names.index(name)

    def body_name2id(self, name):
        return self.digest["body_names"].index(name)

    def site_name2id(self, name):
        return self.digest["site_names"].index(name)

    def reset_weld_relpose(self):
        return None


def load_model_from_path(path):
    return PyMjModel(mjcf.load_digest(path), path)


class MjSimState(object):
    """sim.get_state(): time + a copy of the binary32 BlockPhys state."""

    def __init__(self, time, raw):
        self.time = time
        self.raw = bytes(raw)


class _Contact(object):
    __slots__ = ("geom1", "geom2")

    def __init__(self, g1, g2):
        self.geom1, self.geom2 = g1, g2


class PyMjData(object):
    def __init__(self, sim):
        self._sim = sim
        self.mocap_pos = np.zeros((1, 3))
        self.mocap_quat = np.array([[1.0, 0.0, 0.0, 0.0]])
        self.ctrl = np.zeros(2)
        self.qpos = True       # upstream robot_get_obs only tests `is not None`
        self.contact = []
        self.ncon = 0
        self.time = 0.0

    # ---- joints (live state)
    def _finger(self, name):
        return 0 if "r_gripper" in name else 1

    def get_joint_qpos(self, name):
        s = self._sim.s
        if name.startswith("object"):
            b = s.blk[int(name[6:name.find(":")])]
            half = 0.5 * np.arctan2(float(b.s), float(b.c))
            return np.array([b.pos[0], b.pos[1], b.pos[2], np.cos(half), 0.0, 0.0, np.sin(half)], dtype=np.float64)
        if "gripper_finger_joint" in name:
            return float(s.q[self._finger(name)])
        return float(self._sim._other_qpos.get(name, 0.0))

    def get_joint_qvel(self, name):
        s = self._sim.s
        if "gripper_finger_joint" in name:
            return float(s.qv[self._finger(name)])
        return 0.0

    def set_joint_qpos(self, name, value):
        sim, s = self._sim, self._sim.s
        if name.startswith("object"):
            i = int(name[6:name.find(":")])
            v = np.asarray(value, dtype=np.float64)
            assert v.shape == (7,)
            if i >= s.nblocks:
                return
            b = s.blk[i]
            cur = self.get_joint_qpos(name)
            b.pos[0], b.pos[1], b.pos[2] = np.float32(v[0]), np.float32(v[1]), np.float32(v[2])
            # the cubes of BlockPhys only yaw: quaternion (w, 0, 0, z) -> (cos, sin) of the yaw angle.  MuJoCo stores
            # the quaternion itself, so writing back the one just read (all the reference ever does: it only edits
            # qpos[:2], fetch_env.py:335,383,398,...) must leave the orientation bit for bit as it was
            if not np.array_equal(v[3:], cur[3:]):
                w, z = v[3], v[6]
                b.c, b.s = np.float32(w * w - z * z), np.float32(2.0 * w * z)
        elif "gripper_finger_joint" in name:
            s.q[self._finger(name)] = np.float32(value)
        else:
            sim._other_qpos[name] = float(value)
            if name.startswith("table0:slide"):
                k = int(name[-1])
                assert abs(sim.model.digest["table_body_pos"][k] + float(value) - (1.3, 0.75, 0.23)[k]) < 1e-12, \
                    "BlockPhys pins the table at (1.3, 0.75, 0.23)"

    # ---- kinematics cache (refreshed by forward() / step() only)
    def get_site_xpos(self, name):
        k = self._sim._kin
        if name == "robot0:grip":
            return k["grip"].copy()
        if name.startswith("object"):
            return k["obj_pos"][int(name[6:])].copy()
        raise KeyError(name)

    def get_body_xpos(self, name):
        assert name == "robot0:gripper_link"
        return self._sim._kin["grip"].copy()   # the fake has one frame for the welded gripper

    def get_body_xquat(self, name):
        return np.array([1.0, 0.0, 0.0, 0.0])

    def get_site_xmat(self, name):
        return self._sim._kin["obj_mat"][int(name[6:])].copy()

    def get_site_xvelp(self, name):
        s = self._sim.s
        if name == "robot0:grip":
            return np.array([s.gv[0], s.gv[1], s.gv[2]], dtype=np.float64)
        b = s.blk[int(name[6:])]
        return np.array([b.vel[0], b.vel[1], b.vel[2]], dtype=np.float64)

    def get_site_xvelr(self, name):
        b = self._sim.s.blk[int(name[6:])]
        return np.array([0.0, 0.0, b.w], dtype=np.float64)

    # ---- mocap
    def set_mocap_pos(self, name, value):
        self.mocap_pos[0][:] = value
        if self._sim._settling:   # E0: rigid weld while `_env_setup` positions the end effector
            s = self._sim.s
            for d in range(3):
                s.g[d] = np.float32(value[d]); s.gv[d] = 0.0; s.m[d] = s.g[d]

    def set_mocap_quat(self, name, value):
        self.mocap_quat[0][:] = value

    @property
    def body_xpos(self):
        return np.tile(self._sim._kin["grip"], (len(self._sim.model.digest["body_names"]), 1))


class MjSim(object):
    def __init__(self, model, nsubsteps=1):
        self.L = coracle.lib()
        self.model = model
        self.nsubsteps = nsubsteps
        self.s = _CSim()
        self.s.nblocks = model.nblocks
        for i in range(4):
            self.s.blk[i].c = 1.0
        # arm home pose (see _GRIP0 above); fingers closed
        home = _GRIP0 - _SETUP_OFFSET
        self._home = home
        for d in range(3):
            self.s.g[d] = np.float32(home[d])
            self.s.m[d] = self.s.g[d]
        self._other_qpos = {}
        self._settling = True
        self._kin = None
        self.data = PyMjData(self)
        # geom ids per object for the contact list: finger r, table, objectK
        self._geom_of_obj = [model.geom_name2id("robot0:r_gripper_finger_link"), model.geom_name2id("table")] + \
                            [model.geom_name2id("object%d" % i) for i in range(model.nblocks)]
        self.forward()
        if self._settling:
            self._kin["grip"] = home.copy()   # exact float64 home so that offset + home == the pinned float32 values

    # ---- state
    def get_state(self):
        self._settling = False
        return MjSimState(self.data.time, bytes(self.s))

    def set_state(self, state):
        C.memmove(C.byref(self.s), state.raw, C.sizeof(_CSim))
        self.data.time = state.time

    # ---- kinematics
    def forward(self):
        s = self.s
        kin = dict(grip=np.array([s.g[0], s.g[1], s.g[2]], dtype=np.float64), obj_pos=[], obj_mat=[])
        for i in range(s.nblocks):
            b = s.blk[i]
            kin["obj_pos"].append(np.array([b.pos[0], b.pos[1], b.pos[2]], dtype=np.float64))
            c, sn = float(b.c), float(b.s)
            kin["obj_mat"].append(np.array([[c, -sn, 0.0], [sn, c, 0.0], [0.0, 0.0, 1.0]]))
        self._kin = kin

    def _settle(self):
        """E0: the statically feasible configuration `_env_setup`'s settling steps lead to."""
        s = self.s
        order = sorted(range(s.nblocks), key=lambda i: (float(s.blk[i].pos[2]), i))
        placed = []
        for i in order:
            b = s.blk[i]
            x, y = float(b.pos[0]), float(b.pos[1])
            over = abs(np.float32(x) - np.float32(_TBL[0])) <= np.float32(_TBL[2]) and abs(np.float32(y) - np.float32(_TBL[1])) <= np.float32(_TBL[3])
            z = _Z_REST if over else _Z_FLOOR
            for j in placed:
                o = s.blk[j]
                if abs(x - float(o.pos[0])) <= _HB and abs(y - float(o.pos[1])) <= _HB:
                    z = max(z, np.float32(np.float32(o.pos[2]) + np.float32(0.05)))
            b.pos[2] = z
            b.vel[0] = b.vel[1] = b.vel[2] = 0.0
            b.w = 0.0
            placed.append(i)

    def step(self):
        s = self.s
        if self._settling:
            self._settle()
        else:
            m = (C.c_double * 3)(*[float(x) for x in self.data.mocap_pos[0]])
            ctrl_by_joint = dict(zip(self.model.actuator_joint, self.data.ctrl))
            ctrl = (C.c_double * 2)(float(ctrl_by_joint["robot0:r_gripper_finger_joint"]), float(ctrl_by_joint["robot0:l_gripper_finger_joint"]))
            self.L.bpo_sim_set_targets(C.byref(s), m, ctrl)
            # BlockPhys v2 propagates a whole env-step (its tables cover the 20 substeps tasks.py:16 asks for)
            assert self.nsubsteps == 20, self.nsubsteps
            self.L.bpo_sim_step(C.byref(s))
        self.data.time += self.nsubsteps * self.model.opt.timestep
        self.forward()
        # the contact list of the last substep (sim.data.contact / ncon)
        cons = []
        bits = int(s.contacts)
        n = 2 + s.nblocks
        for o1 in range(n):
            for o2 in range(o1 + 1, n):
                if bits >> (o1 * (2 * 6 - o1 - 1) // 2 + (o2 - o1 - 1)) & 1:
                    cons.append(_Contact(self._geom_of_obj[o1], self._geom_of_obj[o2]))
        self.data.contact = cons
        self.data.ncon = len(cons)
