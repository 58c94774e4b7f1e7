#!/usr/bin/env python
"""Record BASELINE.json configs[0] with a REAL pybullet: the recorder that turns "parity unpinned" into pinned.

pybullet is not vendored in the reference, is unpinned (setup.py:12), is not installed in this image nor on the
GPU box (probed in round 2: `import pybullet` -> ModuleNotFoundError, `pip download pybullet` -> no matching
distribution, no baseline/_ref).  This script is therefore committed unrun; run it on any machine that has pybullet
and a checkout of bingjeff/trex-gym:

    python tests/golden/record_pybullet_golden.py --reference /path/to/trex-gym [--derived-urdf] [--inertia-from-file]

It drives pybullet DIRECT exactly as the reference does (trex_gym/trex_env.py:98-122,128-154 and
trex_gym/trex_robot.py:39-65,300-320,359-422, call by call), with name mapping N1 applied (the
checked-in env addresses joints by the names of an older URDF and raises KeyError against assets/trex.urdf), on the
action stream of tests/golden/make_c1_golden.py, and writes

    tests/golden/c1_pybullet.npz   1,000 steps x {obs (q | qd | tau), reward, head xyz, base pose + velocity}, actions
    tests/golden/c1_pybullet.json  getPhysicsEngineParameters(), per-link getDynamicsInfo(), getJointInfo(), versions

so that every [RECALL] constant of oracle/trex_oracle.c (ERP, contact ERP, warm start, split impulse, link damping,
default inertia H6, link order) becomes checkable.  `--compare` also steps the CPU oracle on the same actions and prints
per-step deltas.

    --derived-urdf       load the derived URDF with <collision> contact points (model_compiler.emit_derived_urdf, N2)
                         instead of the literal collision-less assets/trex.urdf (which free-falls through the floor)
    --inertia-from-file  pass URDF_USE_INERTIA_FROM_FILE (the reference passes no flags, trex_robot.py:56: pybullet then
                         recomputes inertia from the collision shapes)
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NUM_SUBSTEPS = 5                       # trex_env.py:18
GRAVITY = 9.81                         # trex_env.py:20
MAX_JOINT_TORQUE = 300000.0            # trex_robot.py:260
STARTING_CONFIGURATION = {             # trex_env.py:81-87 with N1 applied
    "joint_femur_left": -0.6, "joint_tibia_left": 0.4, "joint_tarsometatarsus_left": -1.2,
    "joint_femur_right": -0.6, "joint_tibia_right": 0.4, "joint_tarsometatarsus_right": -1.2,
}
HEAD_LINK = b"link_atlas_axis"         # trex_robot.py:316 with N1 applied
REWARD_WEIGHTS = (200.0, 1e-6, 1.0)    # distance, energy, drift: trex_train.py:66 (the weights of the C1 fixture)


def golden_actions(n_steps=1000):
    """The action stream of tests/golden/make_c1_golden.py (name-sorted joint order, U(lower, upper))."""
    from trex_gym_b200.model_compiler import load_builtin

    model = load_builtin()
    lo = model["mb_lower"][1:][model["obs_dof"]]
    hi = model["mb_upper"][1:][model["obs_dof"]]
    rng = np.random.Generator(np.random.Philox(key=20261018))
    return rng.uniform(lo, hi, size=(n_steps, 25))


class ReferenceLoop:
    """The reference's TrexBulletEnv + TrexRobot call sequence on a pybullet module `pb` (DIRECT mode)."""

    def __init__(self, pb, floor_urdf, robot_urdf, inertia_from_file=False):
        self.pb = pb
        self.client = pb.connect(pb.DIRECT)                                    # trex_env.py:78
        self.dt = 0.01 / NUM_SUBSTEPS                                          # trex_env.py:54,71
        self.iters = 300 / NUM_SUBSTEPS                                        # trex_env.py:57,72
        pb.resetSimulation()                                                   # trex_env.py:102
        self.floor = pb.loadURDF(floor_urdf)                                   # trex_env.py:103
        flags = pb.URDF_USE_INERTIA_FROM_FILE if inertia_from_file else 0
        self.body = pb.loadURDF(robot_urdf, flags=flags) if flags else pb.loadURDF(robot_urdf)  # trex_robot.py:56
        self.n_joints = pb.getNumJoints(self.body)
        self.joint_info = [pb.getJointInfo(self.body, i) for i in range(self.n_joints)]
        name_to_index = {info[1]: info[0] for info in self.joint_info}        # trex_robot.py:86-98 (bytes keys)
        link_to_index = {info[12]: info[0] for info in self.joint_info}
        pb.resetBasePositionAndOrientation(self.body, [0, 0, 3], pb.getQuaternionFromEuler([0, 0, 0]))  # :57-63
        pb.resetBaseVelocity(self.body, [0, 0, 0], [0, 0, 0])                  # trex_robot.py:64
        for i in range(self.n_joints):                                         # trex_robot.py:304
            pb.resetJointState(self.body, i, targetValue=0, targetVelocity=0)
        for name, value in STARTING_CONFIGURATION.items():                     # trex_robot.py:305-308
            pb.resetJointState(self.body, name_to_index[name.encode()], targetValue=value, targetVelocity=0)
        n = self.n_joints                                                      # trex_robot.py:309 -> :234-245
        pb.setJointMotorControlArray(self.body, list(range(n)), pb.POSITION_CONTROL, targetPositions=[0] * n,
                                     targetVelocities=[0] * n, forces=[0] * n, positionGains=[0] * n, velocityGains=[0] * n)
        revolute = [(info[1], info[0]) for info in self.joint_info if info[2] == pb.JOINT_REVOLUTE]
        revolute.sort(key=lambda t: t[0])                                      # trex_robot.py:311-314 (bytes-wise)
        self.revolute = [i for _, i in revolute]
        self.revolute_names = [n.decode() for n, _ in revolute]
        self.head = link_to_index[HEAD_LINK]                                   # trex_robot.py:316
        self.total_mass = sum(pb.getDynamicsInfo(self.body, i)[0] for i in range(self.n_joints))  # :318-320
        pb.setPhysicsEngineParameter(numSolverIterations=int(self.iters))      # trex_env.py:115
        pb.setTimeStep(self.dt)                                                # trex_env.py:116
        pb.setGravity(0, 0, -GRAVITY)                                          # trex_env.py:117
        pb.stepSimulation()                                                    # trex_env.py:120
        self.engine_parameters = pb.getPhysicsEngineParameters()               # trex_env.py:121 prints this
        lims = [(self.joint_info[i][8], self.joint_info[i][9]) for i in self.revolute]
        self.low = np.array([a for a, _ in lims])
        self.high = np.array([b for _, b in lims])

    def observations(self):                                                    # trex_robot.py:359-365
        st = self.pb.getJointStates(self.body, self.revolute)
        return np.array([s[0] for s in st] + [s[1] for s in st] + [s[3] for s in st])

    def head_position(self):                                                   # trex_robot.py:330-335
        return np.array(self.pb.getLinkState(self.body, self.head, computeLinkVelocity=1, computeForwardKinematics=1)[0])

    def base(self):
        p, q = self.pb.getBasePositionAndOrientation(self.body)
        v, w = self.pb.getBaseVelocity(self.body)
        return np.array(list(p) + list(q) + list(w) + list(v))                 # the oracle's state layout [0:13]

    def step(self, action):                                                    # trex_env.py:128-154
        pb = self.pb
        a = np.clip(action, self.low, self.high)                               # trex_env.py:147
        n = len(self.revolute)
        kp = [0.005] * n                                                       # trex_robot.py:421
        kd = [math.sqrt(2.0 * 1.0 * k) for k in kp]                            # trex_robot.py:398
        for _ in range(NUM_SUBSTEPS):                                          # trex_env.py:148 (action_repeat 1 x 5)
            pb.setJointMotorControlArray(self.body, self.revolute, pb.POSITION_CONTROL, targetPositions=list(a[:n]),
                                         targetVelocities=[0.0] * n, forces=[MAX_JOINT_TORQUE] * n, positionGains=kp,
                                         velocityGains=kd)                     # trex_robot.py:404-411
            pb.stepSimulation()                                                # trex_env.py:150
        obs = self.observations()
        p = self.head_position()
        wd, we, wk = REWARD_WEIGHTS
        power = float(np.sum(np.fabs(obs[n:2 * n] * obs[2 * n:3 * n])))        # trex_robot.py:367-375
        reward = -wd * (2.5 - p[2]) ** 2 - wk * (p[0] ** 2 + p[1] ** 2) - we * power  # trex_env.py:186-192
        return obs, reward, p


def record(pb, reference_dir, out_prefix, derived_urdf=False, inertia_from_file=False, n_steps=1000, compare=False):
    assets = os.path.join(reference_dir, "assets")
    robot_urdf = os.path.join(assets, "trex.urdf")
    if derived_urdf:
        os.environ.setdefault("TREX_GYM_REFERENCE", reference_dir)
        from trex_gym_b200.model_compiler import emit_derived_urdf, load_builtin

        robot_urdf = emit_derived_urdf(robot_urdf, out_prefix + "_derived.urdf", model=load_builtin())
    loop = ReferenceLoop(pb, os.path.join(assets, "floor.urdf"), robot_urdf, inertia_from_file=inertia_from_file)
    actions = golden_actions(n_steps)
    reset_obs = loop.observations()
    reset_base = loop.base()
    obs, rew, head, base = [], [], [], []
    for t in range(n_steps):
        o, r, p = loop.step(actions[t])
        obs.append(o)
        rew.append(r)
        head.append(p)
        base.append(loop.base())
    np.savez_compressed(out_prefix + ".npz", actions=actions, reset_obs=reset_obs, reset_base=reset_base, obs=np.asarray(obs),
                        reward=np.asarray(rew), head=np.asarray(head), base=np.asarray(base),
                        joint_names=np.asarray(loop.revolute_names), revolute_indices=np.asarray(loop.revolute))

    def plain(x):
        if isinstance(x, bytes):
            return x.decode(errors="replace")
        if isinstance(x, (tuple, list)):
            return [plain(v) for v in x]
        if isinstance(x, dict):
            return {str(k): plain(v) for k, v in x.items()}
        return x

    meta = {
        "pybullet_api_version": getattr(pb, "getAPIVersion", lambda: None)(),
        "pybullet_module": getattr(pb, "__file__", None),
        "robot_urdf": robot_urdf, "derived_urdf": bool(derived_urdf), "inertia_from_file": bool(inertia_from_file),
        "engine_parameters": plain(loop.engine_parameters),
        "total_mass_links": loop.total_mass,
        "base_dynamics_info": plain(pb.getDynamicsInfo(loop.body, -1)),
        "dynamics_info": [plain(pb.getDynamicsInfo(loop.body, i)) for i in range(loop.n_joints)],
        "joint_info": [plain(info) for info in loop.joint_info],
        "floor_dynamics_info": plain(pb.getDynamicsInfo(loop.floor, -1)),
    }
    with open(out_prefix + ".json", "w") as f:
        json.dump(meta, f, indent=1)
    if compare:
        compare_with_oracle(out_prefix + ".npz", contacts=derived_urdf)
    return out_prefix + ".npz"


def compare_with_oracle(npz_path, contacts=True):
    """Step the CPU oracle on the recorded actions and print the per-step deltas against pybullet (free running)."""
    from oracle.oracle import Oracle
    from trex_gym_b200.model_compiler import load_builtin

    g = np.load(npz_path)
    o = Oracle(load_builtin().blob(), reward_weights=REWARD_WEIGHTS, contacts=contacts)
    ob0 = o.reset()
    print("reset: max |obs - pybullet| = %.3e   base %.3e" % (np.abs(ob0 - g["reset_obs"]).max(),
                                                             np.abs(o.get_state()[:13] - g["reset_base"]).max()))
    for t in range(len(g["actions"])):
        ob, r = o.step(g["actions"][t])
        if t < 20 or t % 100 == 99:
            d = np.abs(ob - g["obs"][t])
            print("step %4d: q %.3e  qd %.3e  tau %.3e  reward %.3e  head %.3e  base %.3e" % (
                t, d[:25].max(), d[25:50].max(), d[50:].max(), abs(r - g["reward"][t]),
                np.abs(o.head_position() - g["head"][t]).max(), np.abs(o.get_state()[:13] - g["base"][t]).max()))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", default=os.environ.get("TREX_GYM_REFERENCE", "/root/reference"))
    ap.add_argument("--out", default=os.path.join(HERE, "c1_pybullet"))
    ap.add_argument("--derived-urdf", action="store_true")
    ap.add_argument("--inertia-from-file", action="store_true")
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--compare", action="store_true")
    args = ap.parse_args(argv)
    try:
        import pybullet as pb
    except ImportError as e:  # the state of this image and of the GPU box
        print("pybullet is not importable here (%s): nothing recorded; parity stays unpinned" % e)
        return 2
    path = record(pb, args.reference, args.out, derived_urdf=args.derived_urdf, inertia_from_file=args.inertia_from_file,
                  n_steps=args.steps, compare=args.compare)
    print("wrote", path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
