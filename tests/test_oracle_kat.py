"""Analytic known-answer tests that pin the CPU oracle independently of Bullet.

PARITY UNPINNED with respect to pybullet itself: these tests prove the oracle is a correct
articulated-body simulator with the restated pybullet semantics, not that pybullet agrees."""
import numpy as np
import pytest

from trex_gym_b200.model_compiler import with_params


def _oracle(model, **kw):
    from oracle.oracle import Oracle

    return Oracle(model.blob(), **kw)


G, DT, MASS = 9.81, 0.002, 5180.275860952213


def test_reset_free_fall_step(model):
    """First reset step from rest: v_z = -g dt, z = 3 - g dt^2, no joint acceleration, tau = 0."""
    o = _oracle(model)
    obs = o.reset()
    s = o.get_state()
    assert abs(s[12] + G * DT) < 1e-15
    assert abs(s[2] - (3.0 - G * DT * DT)) < 1e-15
    assert np.abs(s[38:63]).max() < 1e-12 and np.abs(s[7:10]).max() < 1e-12
    assert np.all(obs[50:75] == 0.0)
    # observation layout: q | qd | tau in name-sorted order; crouch on femur/tibia/tarsometatarsus
    names = model.meta["obs_joint_names"]
    q = dict(zip(names, obs[:25]))
    assert abs(q["joint_femur_left"] + 0.6) < 1e-12 and abs(q["joint_tibia_right"] - 0.4) < 1e-12
    assert abs(q["joint_tarsometatarsus_left"] + 1.2) < 1e-12 and abs(q["joint_cranium"]) < 1e-12
    assert np.allclose(o.head_position(), [0.0, 3.100, 3.208 - G * DT * DT], atol=1e-3)


def test_total_mass_and_com(model):
    o = _oracle(model)
    o.reset()
    m = o.momentum()
    assert abs(m["mass"] - MASS) < 1e-9
    assert np.allclose(m["com"], [0.00177687, 1.27755908, 2.42532109 - G * DT * DT], atol=1e-6)
    assert abs(m["P"][2] + MASS * G * DT) < 1e-9  # momentum after one step = m g dt


def test_momentum_rate_without_contacts(model):
    """Contact-free, damping off, motors off: linear momentum changes by exactly m g dt per step up to the
    O(dt^2) error of the semi-implicit integrator; angular momentum about the COM is conserved."""
    m = with_params(model, linear_damping=0.0, angular_damping=0.0, max_coordinate_velocity=1e9)
    m.sections["full_damping"] = m.sections["full_damping"] * 0
    m.sections["full_lower"] = m.sections["full_lower"] * 0 - 100
    m.sections["full_upper"] = m.sections["full_upper"] * 0 + 100
    for ts, tol in ((0.01, 2.0), (0.001, 0.03)):
        mm = with_params(m, time_step=ts)
        o = _oracle(mm, contacts=False)
        o.reset()
        rng = np.random.default_rng(1)
        s = o.get_state()
        s[3:7] = [0.1, 0.2, 0.3, 0.9]
        s[3:7] /= np.linalg.norm(s[3:7])
        s[7:13] = rng.uniform(-1, 1, 6)
        s[38:63] = rng.uniform(-2, 2, 25)
        o.set_state(s)
        m0 = o.momentum()
        L0 = m0["L"] - np.cross(m0["com"], m0["P"])
        n = 10
        for _ in range(n):
            o.substep(np.zeros(25), 0.0)
        m1 = o.momentum()
        L1 = m1["L"] - np.cross(m1["com"], m1["P"])
        dP = m1["P"] - m0["P"] - np.array([0, 0, -MASS * G * ts / 5 * n])
        assert np.abs(dP).max() < tol * n, (ts, dP)
        assert np.abs(L1 - L0).max() < tol * n * 5, (ts, L1 - L0)
    # first-order convergence of the drift per unit time is at least linear in dt (ratio >= ~10 for dt/10)


def test_internal_impulses_carry_no_momentum(model):
    """Motor / limit / any joint-space impulse is internal: M^-1 e_j changes neither P nor L."""
    o = _oracle(model, contacts=False)
    o.reset()
    s = o.get_state()
    rng = np.random.default_rng(2)
    s[13:38] += rng.uniform(-0.3, 0.3, 25)
    s[3:7] = [0.3, -0.1, 0.2, 0.9]
    s[3:7] /= np.linalg.norm(s[3:7])
    o.set_state(s)
    M = np.stack([o.minv_column(d) for d in range(31)], 1)
    assert np.abs(M - M.T).max() < 1e-12
    assert np.linalg.eigvalsh((M + M.T) / 2).min() > 0
    for d in range(6, 31):
        s2 = s.copy()
        s2[7:13] = M[:6, d]
        s2[38:63] = M[6:, d]
        o.set_state(s2)
        m = o.momentum()
        assert np.abs(m["P"]).max() < 1e-10 and np.abs(m["L"] - np.cross(m["com"], m["P"])).max() < 1e-9
    # unit force on the base along world x gives unit momentum
    s2 = s.copy()
    s2[7:13] = M[:6, 3]
    s2[38:63] = M[6:, 3]
    o.set_state(s2)
    assert np.allclose(o.momentum()["P"], [1, 0, 0], atol=1e-10)


def test_energy_without_motors_decreases_only_by_damping(model):
    m = with_params(model, linear_damping=0.0, angular_damping=0.0, gravity=0.0, time_step=0.001)
    m.sections["full_damping"] = m.sections["full_damping"] * 0
    o = _oracle(m, contacts=False)
    o.reset()
    s = o.get_state()
    rng = np.random.default_rng(3)
    s[38:63] = rng.uniform(-1, 1, 25)
    s[7:13] = rng.uniform(-0.5, 0.5, 6)
    o.set_state(s)
    ke0 = o.momentum()["ke"]
    for _ in range(100):
        o.substep(np.zeros(25), 0.0)
    ke1 = o.momentum()["ke"]
    assert abs(ke1 - ke0) / ke0 < 2e-3  # semi-implicit Euler drift over 20 ms of simulated time


def test_motor_row_semantics(model):
    """btMultiBodyJointMotor: an unsaturated motor on a light distal joint reaches its velocity target
    kp*(theta-q)/dt + (1-kd)*qd within the step; appliedJointMotorTorque = impulse/dt; bound +-600."""
    o = _oracle(model, contacts=False)
    o.reset()
    tgt = o.get_state()[13:38].copy()
    d = model.meta["body_joint_names"].index("joint_toe_04_d_left") - 1
    tgt[d] += 0.2
    o.substep(tgt, 3e5 * DT)
    s = o.get_state()
    # 60 PGS iterations on 25 coupled rows: close to, not exactly, the isolated target 2.5 * 0.2
    assert abs(s[38 + d] - 2.5 * 0.2) < 0.05
    assert np.abs(s[63:88]).max() <= 3e5 + 1e-6
    # saturation: a huge error on the femur saturates at the torque bound
    tgt2 = s[13:38].copy()
    f = model.meta["body_joint_names"].index("joint_femur_left") - 1
    tgt2[f] += 1.5
    o.substep(tgt2, 3e5 * DT)
    assert abs(abs(o.get_state()[63 + f]) - 3e5) < 1e-6


def test_joint_limit_row_only_when_violated(model):
    """btMultiBodyJointLimitConstraint [RECALL]: a row exists only while the limit is violated; a SHALLOW violation
    (-0.04 < pen <= 0, the split-impulse threshold) gets the Baumgarte push-back erp * |pen| / dt on top of the
    velocity term, a DEEP one the velocity term only (its positional part lands in m_rhsPenetration, which the
    multibody solver never reads)."""
    o = _oracle(model, contacts=False)
    o.reset()
    assert o.last_num_limit_rows == 0
    d = model.meta["body_joint_names"].index("joint_toe_04_d_left") - 1  # light distal joint: the 100 N m s cap stays inactive
    upper = model["mb_upper"][d + 1]
    lower = model["mb_lower"][d + 1]
    # shallow: 0.01 rad beyond the upper limit -> target velocity 0.2 * 0.01 / 0.002 = 1 rad/s back towards the range
    s = o.get_state()
    s[7:13] = 0.0
    s[38:63] = 0.0
    s[13 + d] = upper + 0.01
    o.set_state(s)
    o.substep(s[13:38], 0.0)
    assert o.last_num_limit_rows == 1
    assert abs(o.get_state()[38 + d] - (-1.0)) < 1e-9
    # ... and beyond the lower limit, the other way
    s[13 + d] = lower - 0.01
    o.set_state(s)
    o.substep(s[13:38], 0.0)
    assert o.last_num_limit_rows == 1
    assert abs(o.get_state()[38 + d] - 1.0) < 1e-9
    # deep: 0.1 rad beyond -> velocity-only row: the joint is stopped from moving further out, not pushed back
    s[13 + d] = upper + 0.1
    o.set_state(s)
    o.substep(s[13:38], 0.0)
    assert o.last_num_limit_rows == 1
    assert abs(o.get_state()[38 + d]) < 1e-9
    s[38 + d] = 0.5  # moving further out: stopped
    o.set_state(s)
    o.substep(s[13:38], 0.0)
    assert abs(o.get_state()[38 + d]) < 1e-9
    s[38 + d] = -0.5  # moving back in: left alone (joint damping only)
    o.set_state(s)
    o.substep(s[13:38], 0.0)
    assert o.get_state()[38 + d] < -0.3


def test_contact_holds_the_standing_trex(model):
    """Holding the reset pose, the T-rex drops 0.25 m onto its feet and stands (floor z = 0.0005)."""
    o = _oracle(model)
    obs = o.reset()
    hold = obs[:25].copy()
    for _ in range(150):
        obs, rew = o.step(hold)
    s = o.get_state()
    assert o.last_num_contacts >= 6
    assert 1.9 < s[2] < 2.9 and np.abs(s[10:13]).max() < 0.5
    zs = [o.candidate_position(k)[2] for k in range(o.num_candidates)]
    assert min(zs) > -0.03  # no deep penetration
    assert abs(o.momentum()["P"][2]) < 0.2 * MASS  # not accelerating


def test_reward_formula(model):
    o = _oracle(model, reward_weights=(200.0, 1e-6, 1.0))
    o.reset()
    a = np.zeros(25)
    obs, rew = o.step(a)
    p = o.head_position()
    power = np.sum(np.abs(obs[25:50] * obs[50:75]))
    expect = -200.0 * (2.5 - p[2]) ** 2 - 1.0 * (p[0] ** 2 + p[1] ** 2) - 1e-6 * power
    assert abs(rew - expect) < 1e-9 * max(1, abs(expect))
    assert np.allclose(o.reward_terms(), [200.0 * (2.5 - p[2]) ** 2, p[0] ** 2 + p[1] ** 2, 1e-6 * power])


def test_substep_count_generalisation(model):
    for n in (1, 3, 8):
        o = _oracle(model, num_substeps=n)
        o.reset()
        z = o.get_state()[2]
        assert abs(z - (3.0 - G * (0.01 / n) ** 2)) < 1e-14
        o.step(np.zeros(25))
        assert o.total_substeps == 1 + n


# ---------------------------------------------------------------------------------------------------------------
# Independent KATs (SURVEY.md section 8c pins (1)): nothing below shares code or formulation with the oracle's ABA.
# ---------------------------------------------------------------------------------------------------------------
def _world_kinematics(model, quat_xyzw, pos, q_dof):
    """World pose of every URDF link of the FULL model by plain numpy / scipy rotations (independent of the oracle):
    returns COM (n+1,3), link->world rotation (n+1,3,3), and per revolute dof (world axis, a point on the axis, link)."""
    from scipy.spatial.transform import Rotation

    S = model.sections
    n = int(S["full_n_links"][0])
    parent, jtype, dof = S["full_parent"], S["full_jtype"], S["full_dof"]
    rot0 = S["full_rot0"].reshape(n, 3, 3)   # parent coordinates -> link coordinates at q = 0
    axis = S["full_axis"].reshape(n, 3)      # joint axis, link coordinates
    dvec = S["full_d"].reshape(n, 3)         # joint -> link COM, link coordinates
    evec = S["full_e"].reshape(n, 3)         # parent COM -> joint, parent coordinates
    Rlw = [Rotation.from_quat(quat_xyzw).as_matrix()]  # link -> world
    com = [np.asarray(pos, float)]
    joints = {}
    for i in range(n):
        p = parent[i] + 1
        R_pl = rot0[i]  # parent -> link
        if jtype[i] == 1:
            R_pl = Rotation.from_rotvec(-q_dof[dof[i]] * axis[i]).as_matrix() @ rot0[i]
        R = Rlw[p] @ R_pl.T
        joint_point = com[p] + Rlw[p] @ evec[i]
        Rlw.append(R)
        com.append(joint_point + R @ dvec[i])
        if jtype[i] == 1:
            joints[int(dof[i])] = (R @ axis[i], joint_point, i + 1)
    return np.asarray(com), np.asarray(Rlw), joints, parent


def _mass_matrix_by_jacobians(model, quat_xyzw, pos, q_dof):
    """H = sum_links m Jv^T Jv + Jw^T I_world Jw in the coordinates (base omega world, base v world, joint rates):
    the textbook kinetic-energy definition, no recursion, no spatial algebra."""
    S = model.sections
    com, Rlw, joints, parent = _world_kinematics(model, quat_xyzw, pos, q_dof)
    n = len(com) - 1
    mass, Idiag = S["full_mass"], S["full_inertia"].reshape(n + 1, 3)
    nd = len(joints)
    H = np.zeros((6 + nd, 6 + nd))
    for b in range(n + 1):
        Jw, Jv = np.zeros((3, 6 + nd)), np.zeros((3, 6 + nd))
        Jw[:, 0:3] = np.eye(3)
        Jv[:, 3:6] = np.eye(3)
        r = com[b] - com[0]
        Jv[:, 0:3] = -np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])  # omega x r
        a = b
        while a > 0:  # ancestors-or-self joints move link b
            for k, (ax, pt, link) in joints.items():
                if link == a:
                    Jw[:, 6 + k] = ax
                    Jv[:, 6 + k] = np.cross(ax, com[b] - pt)
            a = parent[a - 1] + 1
        Iw = Rlw[b] @ np.diag(Idiag[b]) @ Rlw[b].T
        H += mass[b] * Jv.T @ Jv + Jw.T @ Iw @ Jw
    return H


# diagonal of the joint-space mass matrix at the reset pose (inertia of the joint's subtree about its axis)
_APPENDIX_B_MKK = {
    "atlas_axis": 778.042530, "cranium": 307.322553, "femur_left": 1369.277762, "femur_right": 1367.687173,
    "tibia_left": 339.905413, "tibia_right": 339.268167, "tarsometatarsus_left": 39.844658, "tarsometatarsus_right": 39.852452,
    "toe_02_a_left": 0.678912, "toe_02_a_right": 0.674992, "toe_02_b_left": 0.169529, "toe_02_b_right": 0.167145,
    "toe_03_a_left": 1.398201, "toe_03_a_right": 1.396973, "toe_03_c_left": 0.078078, "toe_03_c_right": 0.077754,
    "toe_04_a_left": 0.969473, "toe_04_a_right": 0.966318, "toe_04_d_left": 0.040968, "toe_04_d_right": 0.040831,
    "vertebra_caudal_02": 958.947119, "vertebra_caudal_10": 190.447149, "vertebra_caudal_24": 3.935202,
    "vertebra_cervical_03": 973.609657, "vertebra_cervical_09": 1550.624604,
}


def test_unit_impulse_inverse_against_jacobian_mass_matrix(model):
    """trex_oracle_minv_column (Bullet's calcAccelerationDeltasMultiDof restated) must be the inverse of the mass matrix
    built independently from link Jacobians -- at the reset pose, where its joint diagonal is SURVEY Appendix B's M_kk
    table, and at a random pose with a tilted base."""
    o = _oracle(model, contacts=False)
    o.reset()
    s = o.get_state()
    names = model.meta["body_joint_names"][1:]  # dof order
    rng = np.random.default_rng(7)
    qlo, qhi = model["mb_lower"][1:], model["mb_upper"][1:]
    poses = [(s[3:7].copy(), s[0:3].copy(), s[13:38].copy())]
    quat = rng.normal(size=4)
    poses.append((quat / np.linalg.norm(quat), np.array([0.3, -0.2, 1.7]), rng.uniform(qlo, qhi)))
    for idx, (quat, pos, q) in enumerate(poses):
        s2 = s.copy()
        s2[0:3], s2[3:7], s2[13:38] = pos, quat, q
        s2[7:13] = 0.0
        s2[38:63] = 0.0
        o.set_state(s2)
        H = _mass_matrix_by_jacobians(model, quat, pos, q)
        assert np.allclose(H, H.T, rtol=0, atol=1e-9 * np.abs(H).max())
        Minv = np.stack([o.minv_column(d) for d in range(31)], 1)
        err = np.abs(H @ Minv - np.eye(31)).max()
        assert err < 1e-8, (idx, err)
        # column-wise relative agreement with the directly inverted matrix
        Hinv = np.linalg.inv(H)
        assert np.abs(Minv - Hinv).max() <= 1e-8 * np.abs(Hinv).max(), idx
        if idx == 0:
            # the reset pose is one physics step after the reference pose (q unchanged: free fall, zero joint rates)
            for k, name in enumerate(names):
                want = _APPENDIX_B_MKK[name.replace("joint_", "", 1)]
                assert abs(H[6 + k, 6 + k] - want) < 5e-7 * max(1.0, want), (name, H[6 + k, 6 + k], want)
            assert abs(H[3, 3] - MASS) < 1e-9 * MASS and abs(H[4, 4] - MASS) < 1e-9 * MASS


def _pendulum_blob(model, m, l, I_axis_com, gravity=9.81, dt=0.002):
    """A one-link model blob: fixed base, one revolute joint about x through the base COM, link COM at distance l below the
    joint (link -z), principal inertia I_axis_com about the joint-parallel axis through the COM; all damping off."""
    from collections import OrderedDict

    from trex_gym_b200 import model_blob

    names = model.meta["param_names"]
    p = model.sections["param_values"].copy()
    for k, v in (("time_step", dt), ("num_substeps", 1), ("gravity", gravity), ("linear_damping", 0.0), ("angular_damping", 0.0)):
        p[names.index(k)] = v
    S = OrderedDict()
    S["param_values"] = p
    S["full_n_links"] = np.asarray([1], np.int32)
    S["full_parent"] = np.asarray([-1], np.int32)
    S["full_jtype"] = np.asarray([1], np.int32)
    S["full_dof"] = np.asarray([0], np.int32)
    S["full_mass"] = np.asarray([1.0, m])
    S["full_inertia"] = np.asarray([1.0, 1.0, 1.0, I_axis_com, 0.3 * I_axis_com + 0.01, 0.7 * I_axis_com + 0.02])
    S["full_rot0"] = np.eye(3).reshape(-1)
    S["full_axis"] = np.asarray([1.0, 0.0, 0.0])
    S["full_d"] = np.asarray([0.0, 0.0, -l])
    S["full_e"] = np.zeros(3)
    S["full_lower"] = np.asarray([-10.0])
    S["full_upper"] = np.asarray([10.0])
    S["full_damping"] = np.zeros(1)
    S["full_start_q"] = np.zeros(1)
    S["full_head_link"] = np.asarray([0], np.int32)
    S["obs_dof"] = np.zeros(25, np.int32)
    S["full_cand_link"] = np.zeros(0, np.int32)
    S["full_cand_local"] = np.zeros(0)
    S["noncontact_order"] = np.asarray([1, 0], np.int32)  # the motor (id ndof + 0), then the limit constraint (id 0)
    return model_blob.pack(S)


def test_fixed_base_pendulum_closed_form(model):
    """A compound pendulum on a fixed base: q_dd = -(m g l / (I_com + m l^2)) sin q.  (a) the first step from rest at a
    large angle is exactly dt * q_dd (semi-implicit Euler); (b) small oscillations follow the exact solution of the
    symplectic-Euler recurrence q_{n+1} - 2 cos(theta) q_n + q_{n-1} = 0, cos(theta) = 1 - (w dt)^2 / 2, to round-off,
    and the continuous solution q0 cos(w t) to O(dt)."""
    from oracle.oracle import Oracle

    m, l, Ic, g, dt = 3.7, 0.45, 0.21, 9.81, 0.002
    o = Oracle(_pendulum_blob(model, m, l, Ic, g, dt), num_substeps=1, contacts=False)
    o.set_fixed_base(True)
    w2 = m * g * l / (Ic + m * l * l)
    tgt = np.zeros(25)

    def state(q, qd=0.0):
        s = np.zeros(o.state_dim)
        s[6] = 1.0
        s[13], s[38] = q, qd
        return s

    for q0 in (0.9, -2.2):  # (a) nonlinear, one step
        o.set_state(state(q0))
        o.substep(tgt, 0.0)
        s = o.get_state()
        assert abs(s[38] - (-dt * w2 * np.sin(q0))) < 1e-14
        assert abs(s[13] - (q0 + dt * s[38])) < 1e-15
        assert np.abs(s[7:13]).max() == 0.0 and np.abs(s[0:3]).max() == 0.0  # the base stays put
    q0, nstep = 1e-6, 2000  # (b) linear regime: sin q = q to 1e-13 relative
    o.set_state(state(q0))
    traj = []
    for _ in range(nstep):
        o.substep(tgt, 0.0)
        traj.append(o.get_state()[13])
    traj = np.asarray(traj)
    n = np.arange(1, nstep + 1)
    th = np.arccos(1.0 - w2 * dt * dt / 2.0)
    # q_0 = q0, q_1 = q0 (1 - (w dt)^2)  =>  q_n = A cos(n th) + B sin(n th)
    A = q0
    B = (q0 * (1.0 - w2 * dt * dt) - A * np.cos(th)) / np.sin(th)
    exact_discrete = A * np.cos(n * th) + B * np.sin(n * th)
    assert np.abs(traj - exact_discrete).max() < 1e-9 * q0
    assert np.abs(traj - q0 * np.cos(np.sqrt(w2) * n * dt)).max() < 2.0 * np.sqrt(w2) * dt * q0
    assert traj.min() < -0.99 * q0  # several full swings were covered
