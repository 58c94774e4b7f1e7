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
