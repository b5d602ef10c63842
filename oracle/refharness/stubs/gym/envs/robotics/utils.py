"""gym.envs.robotics.utils [upstream, recalled] over the fake MjSim (fetch_env.py:184-185,192,288).

Same control flow as upstream; the MuJoCo model tables upstream walks (eq_type / body_mocapid /
actuator_biastype / jnt_qposadr) are replaced by the fake's named accessors."""
import numpy as np


def robot_get_obs(sim):
    """(qpos, qvel) of the joints whose names start with 'robot'; the reference reads the last two,
    the finger joints r, l (fetch_env.py:192-197)."""
    if sim.data.qpos is not None and sim.model.joint_names:
        names = [n for n in sim.model.joint_names if n.startswith("robot")]
        return (np.array([sim.data.get_joint_qpos(name) for name in names]),
                np.array([sim.data.get_joint_qvel(name) for name in names]))
    return np.zeros(0), np.zeros(0)


def ctrl_set_action(sim, action):
    """Position actuators (biastype affine): ctrl = qpos[joint] + action."""
    if sim.model.nmocap > 0:
        _, action = np.split(action, (sim.model.nmocap * 7,))
    if sim.data.ctrl is not None:
        for i in range(action.shape[0]):
            if sim.model.actuator_biastype[i] == 0:
                sim.data.ctrl[i] = action[i]
            else:
                sim.data.ctrl[i] = sim.data.get_joint_qpos(sim.model.actuator_joint[i]) + action[i]


def reset_mocap2body_xpos(sim):
    """Snap every mocap body to the body it is welded to."""
    for mocap_id, body_name in sim.model.weld_pairs:
        sim.data.mocap_pos[mocap_id][:] = sim.data.get_body_xpos(body_name)
        sim.data.mocap_quat[mocap_id][:] = sim.data.get_body_xquat(body_name)


def mocap_set_action(sim, action):
    if sim.model.nmocap > 0:
        action, _ = np.split(action, (sim.model.nmocap * 7,))
        action = action.reshape(sim.model.nmocap, 7)
        pos_delta = action[:, :3]
        quat_delta = action[:, 3:]
        reset_mocap2body_xpos(sim)
        sim.data.mocap_pos[:] = sim.data.mocap_pos + pos_delta
        sim.data.mocap_quat[:] = sim.data.mocap_quat + quat_delta


def reset_mocap_welds(sim):
    """Reset every weld's relative pose to identity, then forward()."""
    sim.model.reset_weld_relpose()
    sim.forward()
