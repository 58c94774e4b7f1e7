"""The pybullet golden recorder (tests/golden/record_pybullet_golden.py) cannot run here -- pybullet is not installed in
this image nor on the GPU box -- so its control flow is exercised against a minimal fake of the pybullet calls it makes:
the call sequence must be the reference's (trex_env.py:98-122,128-154; trex_robot.py:39-65,300-320,404-411)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


class FakePyBullet:
    DIRECT, POSITION_CONTROL, JOINT_REVOLUTE, JOINT_FIXED, URDF_USE_INERTIA_FROM_FILE = 2, 2, 0, 4, 2

    def __init__(self, model):
        self.calls = []
        names = model.meta["joint_names"]  # pybullet link order
        links = model.meta["link_names"][1:]
        rev = set(model.meta["obs_joint_names"])
        lo, hi = model["full_lower"], model["full_upper"]
        self.info = [(i, n.encode(), self.JOINT_REVOLUTE if n in rev else self.JOINT_FIXED, 7, 6, 1, 1.0, 0.0, float(lo[i]), float(hi[i]),
                      100.0, 1.0, links[i].encode(), (0, 0, 1), (0, 0, 0), (0, 0, 0, 1), -1) for i, n in enumerate(names)]
        self.q = np.zeros(len(names))
        self.mass = model["full_mass"]
        self.steps = 0

    def _log(self, name, *a, **k):
        self.calls.append((name, a, k))

    def connect(self, mode):
        self._log("connect", mode)
        return 0

    def resetSimulation(self):
        self._log("resetSimulation")

    def loadURDF(self, path, flags=0):
        self._log("loadURDF", os.path.basename(path), flags)
        return 0 if path.endswith("floor.urdf") else 1

    def getNumJoints(self, body):
        return len(self.info)

    def getJointInfo(self, body, i):
        return self.info[i]

    def getQuaternionFromEuler(self, rpy):
        return (0.0, 0.0, 0.0, 1.0)

    def resetBasePositionAndOrientation(self, body, p, q):
        self._log("resetBase", tuple(p), tuple(q))

    def resetBaseVelocity(self, body, v, w):
        self._log("resetBaseVelocity")

    def resetJointState(self, body, i, targetValue=0, targetVelocity=0):
        self.q[i] = targetValue

    def setJointMotorControlArray(self, body, idx, mode, **kw):
        self._log("motors", tuple(idx), mode, {k: tuple(v) for k, v in kw.items()})
        self.targets = (tuple(idx), tuple(kw["targetPositions"]))

    def getDynamicsInfo(self, body, i):
        return (float(self.mass[i + 1]) if body == 1 else 0.0, 0.5, (1.0, 1.0, 1.0))

    def setPhysicsEngineParameter(self, **kw):
        self._log("setPhysicsEngineParameter", kw)

    def setTimeStep(self, dt):
        self._log("setTimeStep", dt)

    def setGravity(self, x, y, z):
        self._log("setGravity", (x, y, z))

    def stepSimulation(self):
        self.steps += 1
        idx, tgt = self.targets
        for i, t in zip(idx, tgt):  # a trivial servo so the trajectory is not constant
            self.q[i] += 0.1 * (t - self.q[i])
        self._log("stepSimulation")

    def getPhysicsEngineParameters(self):
        return {"fixedTimeStep": 0.002, "numSolverIterations": 60}

    def getJointStates(self, body, idx):
        return [(float(self.q[i]), 0.0, (0.0,) * 6, 1.5) for i in idx]

    def getLinkState(self, body, i, computeLinkVelocity=0, computeForwardKinematics=0):
        return ((0.0, 3.1, 3.2),)

    def getBasePositionAndOrientation(self, body):
        return (0.0, 0.0, 3.0), (0.0, 0.0, 0.0, 1.0)

    def getBaseVelocity(self, body):
        return (0.0, 0.0, -0.02), (0.0, 0.0, 0.0)


def test_recorder_follows_the_reference_call_sequence(model, tmp_path):
    import record_pybullet_golden as rec

    pb = FakePyBullet(model)
    out = rec.record(pb, "/nonexistent/reference", str(tmp_path / "c1_fake"), n_steps=3)
    names = [c[0] for c in pb.calls]
    # trex_env.py:102-120: resetSimulation, floor, robot (no flags), base pose [0,0,3], zero-gain motors on all joints,
    # solver iterations 60, dt 0.002, gravity, ONE physics step
    assert names[:4] == ["connect", "resetSimulation", "loadURDF", "loadURDF"]
    assert pb.calls[2][1] == ("floor.urdf", 0) and pb.calls[3][1] == ("trex.urdf", 0)
    assert ("resetBase", ((0, 0, 3), (0.0, 0.0, 0.0, 1.0)), {}) in pb.calls
    first_motor = next(c for c in pb.calls if c[0] == "motors")
    assert len(first_motor[1][0]) == 132 and set(first_motor[1][2]["forces"]) == {0}
    i_param, i_dt, i_g = names.index("setPhysicsEngineParameter"), names.index("setTimeStep"), names.index("setGravity")
    assert pb.calls[i_param][1][0] == {"numSolverIterations": 60} and abs(pb.calls[i_dt][1][0] - 0.002) < 1e-15
    assert pb.calls[i_g][1][0] == (0, 0, -9.81) and names[i_g + 1] == "stepSimulation"
    # trex_env.py:148-150: five (set motors; step) pairs per env step on the 25 name-sorted revolute joints
    tail = pb.calls[i_g + 2:]
    assert [c[0] for c in tail] == ["motors", "stepSimulation"] * 15
    idx, mode, kw = tail[0][1]
    assert list(idx) == model.meta["obs_pybullet_link_index"] and mode == pb.POSITION_CONTROL
    assert set(kw["forces"]) == {300000.0} and set(kw["positionGains"]) == {0.005}
    assert all(abs(v - 0.1) < 1e-12 for v in kw["velocityGains"]) and set(kw["targetVelocities"]) == {0.0}
    g = np.load(out)
    assert g["obs"].shape == (3, 75) and g["reward"].shape == (3,) and g["base"].shape == (3, 13)
    assert list(g["joint_names"]) == model.meta["obs_joint_names"]
    lo = model["mb_lower"][1:][model["obs_dof"]]
    hi = model["mb_upper"][1:][model["obs_dof"]]
    assert np.allclose(kw["targetPositions"], np.clip(g["actions"][0], lo, hi))
    meta = json.load(open(str(tmp_path / "c1_fake.json")))
    assert len(meta["dynamics_info"]) == 132 and meta["engine_parameters"]["numSolverIterations"] == 60
    assert abs(meta["total_mass_links"] - 4834.866376) < 1e-3  # trex_robot.py:318-320
    # the action stream is the one the oracle fixture was made with
    assert np.array_equal(g["actions"], np.load(os.path.join(HERE, "golden", "c1_oracle_trajectory.npz"))["actions"][:3])


def test_recorder_reports_missing_pybullet():
    import record_pybullet_golden as rec

    try:
        import pybullet  # noqa: F401
    except ImportError:
        assert rec.main(["--steps", "1"]) == 2
