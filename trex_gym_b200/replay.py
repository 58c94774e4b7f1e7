"""State export for visual replay.

The reference renders through the pybullet GUI / ``getCameraImage`` (trex_gym/trex_env.py:156-181) and writes movies
from a live simulation (trex_train.py:126-136).  Rendering stays off the GPU path here: a rollout's environment records
are exported as frames -- base COM-frame pose (what ``resetBasePositionAndOrientation`` takes, trex_robot.py:322-328) and
the 25 joint angles by URDF joint name -- that a pybullet GUI (or any mesh viewer over ``assets/``) can play back offline.
"""
from __future__ import annotations

import json

import numpy as np

from .model_compiler import CompiledModel, load_builtin

# environment-record layout (include/trex_b200.h)
_POS, _QUAT, _Q = slice(0, 3), slice(3, 7), slice(13, 38)


def joint_names_in_state_order(model: CompiledModel | None = None) -> list:
    """URDF joint names in the order of the record's joint block (pybullet link order of the revolute joints)."""
    model = model if model is not None else load_builtin()
    return list(model.meta["body_joint_names"][1:])


class ReplayRecorder:
    """Collects frames of selected environments from a :class:`TrexBatchSim` while it is stepped."""

    def __init__(self, sim, env_indices=(0,), model: CompiledModel | None = None):
        import torch

        self.sim = sim
        self.idx = torch.as_tensor(list(env_indices), device=sim.device, dtype=torch.long)
        self.names = joint_names_in_state_order(model if model is not None else sim.model)
        self.frames = []  # each [n_selected, 160] float32 on the host
        self.dt = 0.01    # trex_env.py:54: one env step

    def capture(self):
        """Append the current state of the selected environments (one small device-to-host copy)."""
        self.frames.append(self.sim.get_state().index_select(0, self.idx).cpu().numpy())
        return self

    def as_dict(self, which: int = 0) -> dict:
        s = np.stack([f[which] for f in self.frames]) if self.frames else np.zeros((0, 160), np.float32)
        return {
            "dt": self.dt,
            "joint_names": self.names,
            "base_position": s[:, _POS].astype(float).tolist(),          # COM frame of the base link, world
            "base_orientation_xyzw": s[:, _QUAT].astype(float).tolist(),  # base -> world
            "joint_positions": s[:, _Q].astype(float).tolist(),
        }

    def save(self, path: str, which: int = 0) -> str:
        with open(path, "w") as f:
            json.dump(self.as_dict(which), f)
        return path


def link_world_position(model: CompiledModel, base_position, base_orientation_xyzw, joint_positions, body: int | None = None,
                        point=None) -> np.ndarray:
    """Host-side forward kinematics of one exported frame on the merged model: world position of ``point`` (body
    coordinates) of rigid body ``body``.  Defaults: the head link's COM -- what ``getLinkState(head)[0]`` returns in the
    reference (trex_robot.py:330-335) and what the kernels report in ``aux[:, 0:3]``.  Lets a consumer of a replay file
    check / place things without the GPU library."""
    from scipy.spatial.transform import Rotation

    S = model.sections
    nb = int(S["mb_n_bodies"][0])
    parent = S["mb_parent"]
    E0 = S["mb_E0"].reshape(nb, 3, 3)
    r0 = S["mb_r0"].reshape(nb, 3)
    body = int(S["mb_head_body"][0]) if body is None else int(body)
    point = S["mb_head_p"] if point is None else np.asarray(point, float)
    q = np.concatenate([[0.0], np.asarray(joint_positions, float)])  # joint k moves body k + 1 (state order)
    chain = []
    b = body
    while b > 0:
        chain.append(b)
        b = int(parent[b])
    R = Rotation.from_quat(np.asarray(base_orientation_xyzw, float)).as_matrix()
    x = np.asarray(base_position, float)
    for b in reversed(chain):
        c, s = np.cos(q[b]), np.sin(q[b])
        x = x + R @ r0[b]
        R = R @ E0[b] @ np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    return x + R @ point


def play_in_pybullet(path: str, urdf_path: str, realtime: bool = True, pb=None, gui: bool = True):
    """Offline playback on the reference side: drives the reference's URDF in pybullet frame by frame (GUI by default;
    ``gui=False`` for a headless check).  ``pb``: the pybullet module (injected by the tests, which have no pybullet).
    Returns the number of frames played."""
    import time

    if pb is None:  # pragma: no cover - needs pybullet
        import pybullet as pb

    rec = json.load(open(path))
    pb.connect(pb.GUI if gui else pb.DIRECT)
    body = pb.loadURDF(urdf_path, flags=pb.URDF_USE_INERTIA_FROM_FILE)
    name_to_index = {pb.getJointInfo(body, i)[1].decode(): i for i in range(pb.getNumJoints(body))}
    ids = [name_to_index[n] for n in rec["joint_names"]]
    for p, q, js in zip(rec["base_position"], rec["base_orientation_xyzw"], rec["joint_positions"]):
        pb.resetBasePositionAndOrientation(body, p, q)
        for i, a in zip(ids, js):
            pb.resetJointState(body, i, a)
        if realtime:
            time.sleep(rec["dt"])
    return len(rec["base_position"])
