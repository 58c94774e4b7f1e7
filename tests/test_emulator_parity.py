"""The PRODUCT kernel source (trex_gym_b200/csrc/trex_core.h) executed lane-for-lane on the CPU through
tests/emu (32-wide host structs instead of warp registers) and compared with the CPU oracle.
This is how the kernel arithmetic is covered in the GPU-less container; the real GPU runs are in
test_gpu_parity.py (-m gpu)."""
import numpy as np
import pytest

from conftest import STATE_BLOCKS, rel_err


@pytest.fixture(scope="module")
def emu_cls():
    from emu import EmuEnv

    return EmuEnv


def _oracle(model, **kw):
    from oracle.oracle import Oracle

    return Oracle(model.blob(), **kw)


def test_reset_matches_oracle(model, emu_cls):
    e, o = emu_cls(model.blob()), _oracle(model)
    eo, oo = e.reset(), o.reset()
    assert np.abs(eo - oo).max() < 1e-6
    assert np.abs(e.get_state(o.num_candidates) - o.get_state()).max() < 1e-6
    assert e.rec[153] == 1.0 and e.rec[152] == 0.0  # one episode started, zero steps


def test_per_step_parity_contact_free(model, emu_cls, action_limits):
    lo, hi = action_limits
    e, o = emu_cls(model.blob(), contacts=False), _oracle(model, contacts=False)
    e.reset()
    rng = np.random.default_rng(0)
    worst = 0.0
    for t in range(25):
        a = rng.uniform(lo, hi)
        o.set_state(e.get_state(o.num_candidates))
        oobs, orew = o.step(a)
        eobs, erew, done = e.step(a)
        so, se = o.get_state(), e.get_state(o.num_candidates)
        for k, sl in STATE_BLOCKS.items():
            worst = max(worst, rel_err(so[sl], se[sl]))
        assert abs(orew - erew) < 2e-5 * max(1.0, abs(orew))
        assert np.abs(oobs - eobs).max() < 2e-5 * max(1.0, np.abs(oobs).max())
        assert not done
        assert e.aux[6] == 300  # 5 substeps x 60 iterations (the solver never reaches 1e-7 here)
    assert worst < 2e-5, worst


@pytest.mark.parametrize("mode", ["row_space", "one_env"])
def test_per_substep_parity_with_contacts(model, emu_cls, action_limits, mode):
    """Single substeps in contact against the oracle.  "row_space": substeps with 1..4 contacts go through solve4<4>
    (four environments per warp, contact rows tracked in row space); "one_env": all of them through the
    one-environment sweep.  Both are the same Gauss-Seidel iteration in the same row order."""
    from trex_gym_b200.model_compiler import with_params

    lo, hi = action_limits
    sub = with_params(model, time_step=0.002, solver_iterations=60)
    e = emu_cls(sub.blob(), num_substeps=1, deferred=1 if mode == "row_space" else 5)
    o = _oracle(sub, num_substeps=1)
    e.reset()
    rng = np.random.default_rng(5)
    errs, n_contact = [], 0
    a = rng.uniform(lo, hi)
    for t in range(220):
        if t % 5 == 0:
            a = rng.uniform(lo, hi)
        o.set_state(e.get_state(o.num_candidates))
        o.step(a)
        e.step(a)
        so, se = o.get_state(), e.get_state(o.num_candidates)
        errs.append(max(rel_err(so[sl], se[sl]) for sl in STATE_BLOCKS.values()))
        n_contact += o.last_num_contacts > 0
        assert int(e.aux[7]) == o.last_num_contacts
    errs = np.asarray(errs)
    inline, free, rowspace = e.solve_counts()
    assert n_contact > 60
    assert (rowspace > 50 and inline < rowspace) if mode == "row_space" else (rowspace == 0 and inline > 60)
    assert np.percentile(errs, 50) < 2e-5 and np.percentile(errs, 99) < 2e-3, (np.percentile(errs, 50), errs.max())


def test_contact_capacity_keeps_the_deepest_points(model, emu_cls, action_limits):
    """More candidates inside the breaking distance than contact slots (model parameter max_contacts): kernel and
    oracle keep the same, deepest points (definition N2), so single substeps still agree."""
    from trex_gym_b200.model_compiler import with_params

    lo, hi = action_limits
    params = dict(zip(model.meta["param_names"], model["param_values"]))
    cap = int(params["max_contacts"])
    sub = with_params(model, time_step=0.002, solver_iterations=60)
    e, o = emu_cls(sub.blob(), num_substeps=1), _oracle(sub, num_substeps=1)
    nc = o.num_candidates
    e.reset()
    rng = np.random.default_rng(23)
    n_over, errs = 0, []
    for trial in range(12):
        e.reset()
        s = e.get_state(nc)
        # lay the robot on its side just above the floor: most candidate points come into range
        ang = rng.uniform(-0.2, 0.2)
        s[3:7] = [np.sin((np.pi / 2 + ang) / 2), 0.0, 0.0, np.cos((np.pi / 2 + ang) / 2)]
        s[13:38] = np.clip(s[13:38] + rng.uniform(-0.2, 0.2, 25), model["mb_lower"][1:], model["mb_upper"][1:])
        o.set_state(s)
        z = np.array([o.candidate_position(k)[2] for k in range(nc)])
        s[2] += params["floor_height"] + 0.005 - np.sort(z)[min(cap + 3, nc - 1)]  # cap+4 points at or below 5 mm
        s[88:152] = 0.0
        e.set_state(s)
        for t in range(4):
            a = rng.uniform(lo, hi)
            pre = e.get_state(nc)
            o.set_state(pre)
            z = np.array([o.candidate_position(k)[2] for k in range(nc)])
            in_range = int(((z - params["floor_height"]) < params["contact_breaking_threshold"]).sum())
            o.step(a)
            e.step(a)
            over = int(e.aux[7]) // 1000
            assert over == max(0, in_range - cap), (trial, t, over, in_range)
            assert int(e.aux[7]) % 1000 == o.last_num_contacts == min(in_range, cap)
            so, se = o.get_state(), e.get_state(nc)
            if over:
                n_over += 1
                errs.append(max(rel_err(so[sl], se[sl]) for sl in STATE_BLOCKS.values()))
    assert n_over >= 10
    assert e.heavy_solves() >= n_over  # 16 contacts: the row-space one-environment solver (solve_heavy)
    errs = np.asarray(errs)
    assert np.percentile(errs, 50) < 5e-5 and errs.max() < 5e-3, (np.percentile(errs, 50), errs.max())


def test_standing_drop_matches_oracle_qualitatively(model, emu_cls):
    """Free-running (no re-seeding) hold-pose drop: both land on the feet and stand at the same height."""
    e, o = emu_cls(model.blob()), _oracle(model)
    hold = o.reset()[:25].copy()
    e.reset()
    for _ in range(120):
        o.step(hold)
        e.step(hold)
    so, se = o.get_state(), e.get_state(o.num_candidates)
    assert abs(so[2] - se[2]) < 0.02 and abs(se[2] - 2.19) < 0.1
    assert int(e.aux[7]) >= 6


def test_standing_on_both_feet_per_substep(model, emu_cls):
    """A T-rex standing on both feet has 12-16 active contact points: those substeps are finished by solve_heavy (one
    environment per warp, contact rows in row space).  Per-substep parity against the oracle, and agreement with the
    one-environment sweep inside the front phase, which runs the same Gauss-Seidel iteration in coordinate space."""
    from trex_gym_b200.model_compiler import with_params

    sub = with_params(model, time_step=0.002, solver_iterations=60)
    e = emu_cls(sub.blob(), num_substeps=1)                 # row-space solvers
    f = emu_cls(sub.blob(), num_substeps=1, deferred=1 | 16)  # more than 8 contacts: front phase
    o = _oracle(sub, num_substeps=1)
    nc = o.num_candidates
    hold = o.reset()[:25].copy()
    e.reset()
    f.reset()
    errs, errs_f, errs_fo, ks = [], [], [], []
    for t in range(230):
        pre = e.get_state(nc)
        o.set_state(pre)
        f.set_state(pre)
        o.step(hold)
        e.step(hold)
        f.step(hold)
        so, se, sf = o.get_state(), e.get_state(nc), f.get_state(nc)
        if o.last_num_contacts > 8:
            errs.append(max(rel_err(so[sl], se[sl]) for sl in STATE_BLOCKS.values()))
            errs_fo.append(max(rel_err(so[sl], sf[sl]) for sl in STATE_BLOCKS.values()))
            errs_f.append(max(rel_err(sf[sl], se[sl]) for sl in STATE_BLOCKS.values()))
            ks.append(o.last_num_contacts)
            assert int(e.aux[7]) % 1000 == o.last_num_contacts
    errs, errs_f, errs_fo = np.asarray(errs), np.asarray(errs_f), np.asarray(errs_fo)
    assert len(errs) > 60 and max(ks) >= 14
    assert e.heavy_solves() == len(errs) and f.heavy_solves() == 0
    # 14-16 stacked contact points make an ill-conditioned FP32 problem: 1.5e-4 median per substep for either solver
    assert np.percentile(errs, 50) < 5e-4 and np.percentile(errs, 99) < 1e-2, (np.percentile(errs, 50), errs.max())
    assert np.percentile(errs, 50) < 1.5 * np.percentile(errs_fo, 50) and errs.max() < 1.5 * errs_fo.max()
    assert np.percentile(errs_f, 50) < 1e-4 and errs_f.max() < 1e-3, (np.percentile(errs_f, 50), errs_f.max())


def test_reward_bit_exact_from_outputs(model, emu_cls, action_limits):
    lo, hi = action_limits
    e = emu_cls(model.blob(), reward_weights=(200.0, 1e-6, 1.0))
    e.reset()
    rng = np.random.default_rng(1)
    f = np.float32
    for t in range(10):
        obs, rew, _ = e.step(rng.uniform(lo, hi))
        power = f(0)
        for k in range(25):
            power = f(power + f(abs(f(obs[25 + k] * obs[50 + k]))))
        x, y, z = e.aux[0], e.aux[1], e.aux[2]
        dz = f(f(2.5) - z)
        lifting = f(f(200.0) * f(dz * dz))
        station = f(f(1.0) * f(f(x * x) + f(y * y)))
        energy = f(f(1e-6) * power)
        assert f(f(-lifting - station) - energy) == f(rew)
        assert e.aux[3] == lifting and e.aux[4] == station and e.aux[5] == energy


def test_horizon_auto_reset(model, emu_cls, action_limits):
    lo, hi = action_limits
    e = emu_cls(model.blob(), max_episode_steps=3)
    reset_obs = e.reset()
    rng = np.random.default_rng(2)
    dones = []
    for t in range(7):
        obs, rew, done = e.step(rng.uniform(lo, hi))
        dones.append(done)
        if done:
            assert np.array_equal(obs, reset_obs)
    assert dones == [False, False, True, False, False, True, False]
    assert e.rec[153] == 3.0 and e.rec[154] == 0.0


def test_nan_guard_resets(model, emu_cls):
    e = emu_cls(model.blob())
    reset_obs = e.reset()
    e.rec[40] = np.nan
    obs, rew, done = e.step(np.zeros(25))
    assert done and np.array_equal(obs, reset_obs) and e.rec[154] == 1.0


def test_substep_sweep(model, emu_cls, action_limits):
    lo, hi = action_limits
    rng = np.random.default_rng(4)
    for n in (1, 2, 8):
        e, o = emu_cls(model.blob(), num_substeps=n), _oracle(model, num_substeps=n)
        e.reset()
        for t in range(3):
            a = rng.uniform(lo, hi)
            o.set_state(e.get_state(o.num_candidates))
            o.step(a)
            e.step(a)
            so, se = o.get_state(), e.get_state(o.num_candidates)
            assert max(rel_err(so[sl], se[sl]) for sl in STATE_BLOCKS.values()) < 2e-3
        assert e.aux[6] == n * int(300 / n)


def test_shared_slab_size(emu_cls):
    from emu import lib

    assert lib().emu_shared_bytes() <= 28 * 1024  # >= 8 resident warps per SM (227 KB)
    assert lib().emu_state_stride() == 160


def _philox4x32_10(counter, key):
    """Reference Philox4x32-10 (Salmon et al., SC'11) in Python integers."""
    c = [int(x) & 0xFFFFFFFF for x in counter]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return [np.float32(x >> 8) * np.float32(1.0 / 16777216.0) for x in c]


def test_fallen_start_sampler(model, emu_cls):
    """reset_mode 1 (BASELINE configs[4]): the sampled pose is reproducible from (seed, env id, episode) and the
    one physics step of the reset agrees with the oracle started from that pose."""
    seed, env_id = 7, 12345
    # contacts off: a random pose may start deep in the floor, where one solver step is chaotic
    e = emu_cls(model.blob(), reset_mode=1, seed=seed, env_id=env_id, max_episode_steps=2, contacts=False)
    o = _oracle(model, contacts=False)
    nc = o.num_candidates
    lower, upper = model["mb_lower"][1:].astype(np.float32), model["mb_upper"][1:].astype(np.float32)
    states = []
    for episode in range(3):
        if episode == 0:
            e.reset()
        else:
            for _ in range(2):
                _, _, done = e.step(np.zeros(25))
            assert done
        got = e.get_state(nc)
        # expected sample: block L>>2, word L&3 for joint L; block 8 = base height + orientation
        ep = episode  # episode counter before the increment
        q = np.zeros(25, np.float32)
        for L in range(25):
            u = _philox4x32_10((env_id, 0, ep, L >> 2), (seed, 0xFA11))[L & 3]
            q[L] = lower[L] + (upper[L] - lower[L]) * u
        ub = _philox4x32_10((env_id, 0, ep, 8), (seed, 0xFA11))
        z = np.float32(0.3) + np.float32(3.0 - 0.3) * ub[0]
        a, b = np.sqrt(np.float32(1) - ub[1]), np.sqrt(ub[1])
        t2, t3 = np.float32(6.283185307179586) * ub[2], np.float32(6.283185307179586) * ub[3]
        quat = np.array([a * np.sin(t2), a * np.cos(t2), b * np.sin(t3), b * np.cos(t3)], np.float64)
        s0 = np.zeros(o.state_dim)
        s0[2] = z
        s0[3:7] = quat
        s0[13:38] = q
        o.set_state(s0)
        o.substep(np.zeros(25), 0.0)  # the one zero-force physics step of reset
        so = o.get_state()
        for k, sl in STATE_BLOCKS.items():
            assert np.abs(so[sl] - got[sl]).max() < 2e-5 * max(1.0, np.abs(so[sl]).max()), (episode, k)
        assert abs(np.linalg.norm(got[3:7]) - 1.0) < 1e-5
        states.append(got)
    assert np.abs(states[0][13:38] - states[1][13:38]).max() > 0.05  # new episode, new pose
    e2 = emu_cls(model.blob(), reset_mode=1, seed=seed, env_id=env_id, contacts=False)
    e2.reset()
    assert np.array_equal(e2.get_state(nc), states[0])  # deterministic


def test_four_environments_per_warp(model, action_limits):
    """The kernel's work unit: one warp steps four consecutive environments; contact-free substeps defer their
    solve to solve4 (8 lanes per environment), substeps with contacts solve in the one-environment path.
    Every environment is checked against the oracle, and against the same run with the deferral disabled."""
    from emu import EmuWarp4

    lo, hi = action_limits
    qlo, qhi = model["mb_lower"][1:], model["mb_upper"][1:]
    w4 = EmuWarp4(model.blob(), n=4)
    w1 = EmuWarp4(model.blob(), n=4, deferred=False)
    wr = EmuWarp4(model.blob(), n=4, deferred=3)  # deferred environments packed into the lane groups in reverse
    wp = EmuWarp4(model.blob(), n=4, deferred=1 | 8)  # front phase as the library's 4-warp CTA: inward pass of the four
    o = _oracle(model)                                # environments by one warp (inward_packed), warps = host threads
    nc = o.num_candidates
    w4.reset()
    w1.reset()
    wr.reset()
    wp.reset()
    rng = np.random.default_rng(11)
    for e in range(4):
        s = w4.get_state(e, nc)
        s[13:38] = np.clip(s[13:38] + rng.uniform(-0.3, 0.3, 25), qlo - 0.02, qhi + 0.02)  # some limits violated
        s[2] -= 0.12 * e  # environments 2 and 3 start close to / on the floor
        w4.set_state(e, s)
        w1.set_state(e, s)
    saw_contact = saw_free = saw_limit = False
    for t in range(30):
        acts = rng.uniform(lo, hi, size=(4, 25))
        pre = [w4.get_state(e, nc) for e in range(4)]
        obs, rew, done = w4.step(acts)
        for e in range(4):
            o.set_state(pre[e])
            oobs, orew = o.step(acts[e])
            so, se = o.get_state(), w4.get_state(e, nc)
            err = max(rel_err(so[sl], se[sl]) for sl in STATE_BLOCKS.values())
            contact = o.last_num_contacts > 0
            saw_contact |= contact
            saw_free |= not contact
            saw_limit |= o.last_num_limit_rows > 0
            assert err < (5e-3 if contact else 1e-4), (t, e, err, contact)  # wild perturbed states: up to 4e-5 contact-free
            assert abs(orew - rew[e]) < (5e-3 if contact else 1e-4) * max(1.0, abs(orew))
            assert int(w4.aux[e, 7]) % 1000 == o.last_num_contacts  # + 1000 x points over capacity
            assert w4.aux[e, 6] == 300
        # same inputs through the non-deferred path: agreement to rounding (different summation order)
        for e in range(4):
            w1.set_state(e, pre[e])
            wr.set_state(e, pre[e])
            wp.set_state(e, pre[e])
        w1.step(acts)
        wr.step(acts)
        wp.step(acts)
        assert np.array_equal(wr.rec[:, :152], w4.rec[:, :152])  # lane-group assignment does not change a single bit
        assert np.array_equal(wp.rec[:, :152], w4.rec[:, :152])  # nor does the four-environments-per-warp inward pass
        for e in range(4):
            a, b = w4.get_state(e, nc), w1.get_state(e, nc)
            assert max(rel_err(a[sl], b[sl]) for sl in STATE_BLOCKS.values()) < 5e-3
    assert saw_contact and saw_free and saw_limit
    assert wp.env.packed_rounds() == 30 * 5 and w4.env.packed_rounds() == 0


def test_two_heavy_environments_per_warp(model):
    """solve2: environments with more than 8 contacts are solved two per warp, sixteen lanes each.  Three standing
    environments (different poses: 12-16 contacts) step as one work unit -- two share a warp, the third runs alone in
    the lower or (reverse packing) the upper lane group -- and every record equals, bit for bit, the record of the same
    environment stepped on its own: the pairing never changes a result."""
    from emu import EmuEnv, EmuWarp4

    o = _oracle(model)
    nc = o.num_candidates
    hold = o.reset()[:25].copy()
    w3 = EmuWarp4(model.blob(), n=3)
    wr = EmuWarp4(model.blob(), n=3, deferred=3)
    singles = [EmuEnv(model.blob()) for _ in range(3)]
    w3.reset()
    wr.reset()
    rng = np.random.default_rng(5)
    acts = np.stack([hold + rng.uniform(-0.02, 0.02, 25) * (e > 0) for e in range(3)]).astype(np.float32)
    for e, s1 in enumerate(singles):
        s1.reset()
    heavy_steps = 0
    for t in range(45):
        pre = w3.rec.copy()
        wr.rec[:] = pre
        for e, s1 in enumerate(singles):
            s1.rec[:] = pre[e]
        w3.step(acts)
        wr.step(acts)
        for e, s1 in enumerate(singles):
            s1.step(acts[e])
            assert np.array_equal(s1.rec[:152], w3.rec[e, :152]), (t, e)
        assert np.array_equal(wr.rec[:, :152], w3.rec[:, :152]), t
        ks = [int(w3.aux[e, 7]) % 1000 for e in range(3)]
        heavy_steps += all(k > 8 for k in ks)
        if t == 44:
            assert len(set(ks)) >= 1 and max(ks) >= 12
    assert heavy_steps >= 10 and w3.env.heavy_solves() >= 3 * 5 * 10
    # and the pair agrees with the oracle like a lone environment does
    o.set_state(np.concatenate([pre[1, :88], pre[1, 88:88 + nc]]).astype(np.float64))
    o.step(acts[1].astype(np.float64))
    so, se = o.get_state(), w3.get_state(1, nc)
    assert max(rel_err(so[sl], se[sl]) for sl in STATE_BLOCKS.values()) < 2e-2  # (one env step with 12+ contacts: p99 is 3e-3..5e-3)


def test_contact_primitives_model_per_substep(model, action_limits):
    """the contact model built from spheres / capsules fitted to the meshes (sphere candidates
    with a radius: contact point = centre - r * normal).  The kernel source and the oracle read the same candidate table:
    per-substep parity while the T-rex lands and stands on its toe capsules, and with random actions."""
    from emu import EmuEnv

    from trex_gym_b200.model_compiler import load_builtin, with_params

    prim = with_params(load_builtin("primitives"), time_step=0.002, solver_iterations=60)
    o = _oracle(prim, num_substeps=1)
    e = EmuEnv(prim.blob(), num_substeps=1)
    nc = o.num_candidates
    assert nc == len(prim["mb_cand_r"]) > 48
    lo, hi = action_limits
    rng = np.random.default_rng(4)
    hold = o.reset()[:25].copy()
    e.reset()
    errs, ks = [], []
    for t in range(260):
        a = hold if t < 180 else rng.uniform(lo, hi)
        pre = e.get_state(nc)
        o.set_state(pre)
        o.step(a)
        e.step(a)
        so, se = o.get_state(), e.get_state(nc)
        assert int(e.aux[7]) % 1000 == o.last_num_contacts
        if o.last_num_contacts:
            # (free fall with the pose held needs ~1e-3 N m of motor torque: that block is compared against a 1 N m floor)
            errs.append(max(rel_err(so[sl], se[sl]) if k != "tau" else float(np.abs(so[sl] - se[sl]).max() / max(np.abs(so[sl]).max(), 1.0))
                            for k, sl in STATE_BLOCKS.items()))
            ks.append(o.last_num_contacts)
            assert int(e.rec[158]) == o.signature or errs[-1] < 5e-3
    errs = np.asarray(errs)
    assert len(errs) > 100 and max(ks) >= 8
    # sphere candidates rest on the floor at centre - r: the standing height is that of the capsules, not of the mesh vertices
    zs = [o.candidate_position(k)[2] for k in range(nc)]
    assert -0.03 < min(zs) < 0.03
    assert np.percentile(errs, 50) < 5e-4 and np.percentile(errs, 99) < 2e-2, (np.percentile(errs, 50), errs.max())


def test_literal_reference_model(action_limits):
    """What the reference's own pybullet calls simulate with the checked-in URDF [RECALL, SURVEY.md H6 / section 0.4]:
    default-flag inertia (recomputed from the absent collision shapes: point masses) and no contact geometry at all.
    Shipped as ``load_builtin("literal")`` / ``TrexBulletEnv(literal_reference=True)``; the animal free-falls, and kernel
    source and oracle agree to the contact-free tolerance on it."""
    from emu import EmuEnv

    from trex_gym_b200.model_compiler import load_builtin

    lit, std = load_builtin("literal"), load_builtin()
    assert lit.meta["inertia_source"] == "bullet_default" and len(lit["mb_cand_body"]) == 0
    I = lit["full_inertia"].reshape(-1, 3)
    assert np.allclose(I[:, 0], lit["full_mass"] / 12.0 * 2 * 0.002 ** 2) and I.max() < 1e-2  # margin-sized boxes
    assert np.array_equal(lit["full_mass"], std["full_mass"]) and np.array_equal(lit["mb_parent"], std["mb_parent"])
    o = _oracle(lit, contacts=False)
    e = EmuEnv(lit.blob(), contacts=False)
    assert np.abs(o.reset() - e.reset()).max() < 1e-6
    lo, hi = action_limits
    rng = np.random.default_rng(0)
    errs = []
    for t in range(40):
        a = rng.uniform(lo, hi)
        pre = e.get_state(0)
        o.set_state(pre)
        o.step(a)
        e.step(a)
        so, se = o.get_state(), e.get_state(0)
        errs.append(max(rel_err(so[sl], se[sl]) for sl in STATE_BLOCKS.values()))
    assert max(errs) < 2e-5, max(errs)
    assert o.get_state()[2] < 2.3  # free fall from z = 3: nothing holds the literal model up


def test_solve2_with_the_delassus_matrix_in_tensor_memory(model):
    """solve2<true> keeps the Delassus matrices of its two environments in tensor memory (device: tcgen05.st / tcgen05.ld,
    a per-lane scratchpad; here an emulated 32 x 512 array) instead of shared memory.  The arithmetic is untouched: records
    are bit-identical with the shared-memory instance, so the two kernel instances can share one list of environments."""
    from emu import EmuWarp4

    o = _oracle(model)
    hold = o.reset()[:25].copy()
    a = EmuWarp4(model.blob(), n=3)
    b = EmuWarp4(model.blob(), n=3, deferred=1 | 32)
    a.reset()
    b.reset()
    rng = np.random.default_rng(6)
    acts = np.stack([hold + rng.uniform(-0.02, 0.02, 25) * (e > 0) for e in range(3)]).astype(np.float32)
    for t in range(40):
        a.step(acts)
        b.step(acts)
        assert np.array_equal(a.rec[:, :160], b.rec[:, :160]), t
    assert b.env.heavy_solves() >= 3 * 5 * 8 and int(b.aux[0, 7]) % 1000 >= 12


def test_solve4_with_delassus_blocks_and_sweep_responses_in_tensor_memory(model, action_limits):
    """solve4<8, TM = true> (the tensor-memory instance of the contact solver, 1-8 contacts) keeps the Delassus blocks A4 and the
    responses of every lane's own contact along the sweep in tensor memory (device: tcgen05.st / tcgen05.ld 32x32b, lane-private
    columns; here an emulated 32 x 512 array) instead of shared memory -- that kernel is bound by the shared-memory data
    pipe.  The arithmetic is untouched: records are bit-identical with the shared-memory instance."""
    from emu import EmuWarp4

    lo, hi = action_limits
    a = EmuWarp4(model.blob(), n=4)
    b = EmuWarp4(model.blob(), n=4, deferred=1 | 64)
    a.reset()
    b.reset()
    rng = np.random.default_rng(16)
    seen = set()
    for t in range(60):
        acts = rng.uniform(lo, hi, size=(4, 25)).astype(np.float32)
        a.step(acts)
        b.step(acts)
        assert np.array_equal(a.rec[:, :160], b.rec[:, :160]), t
        seen.update(int(k) % 1000 for k in b.aux[:, 7])
    assert len([k for k in seen if 1 <= k <= 8]) >= 3, seen  # several contact counts went through the tensor-memory instance
