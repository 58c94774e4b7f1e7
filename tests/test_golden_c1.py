"""BASELINE.json configs[0] (single env, random actions, 1,000 steps): regression pin of the ORACLE
trajectory (tests/golden/c1_oracle_trajectory.npz, made by tests/golden/make_c1_golden.py).
Not a pybullet golden vector -- pybullet cannot be run here; parity with pybullet stays unpinned."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_c1_fixture(model):
    from oracle.oracle import Oracle

    g = np.load(os.path.join(HERE, "golden", "c1_oracle_trajectory.npz"))
    o = Oracle(model.blob(), reward_weights=(200.0, 1e-6, 1.0))
    assert np.allclose(o.reset(), g["reset_obs"], rtol=0, atol=1e-12)
    want = {int(s): i for i, s in enumerate(g["steps"])}
    for t in range(200):
        ob, r = o.step(g["actions"][t])
        if t in want:
            i = want[t]
            # libm differences may perturb the last bits; the trajectory is chaotic once in contact
            tol = 1e-9 if t < 20 else 1e-5
            assert np.abs(ob - g["obs"][i]).max() <= tol * max(1.0, np.abs(g["obs"][i]).max()), t
            assert abs(r - g["reward"][i]) <= tol * max(1.0, abs(g["reward"][i])), t
            assert np.abs(o.head_position() - g["head"][i]).max() <= tol * 10, t
    assert g["contacts"].max() > 0 and abs(float(g["mean_iterations"]) - 60.0) < 1.0


def test_emulated_kernel_follows_c1_fixture(model):
    """Free-running FP32 kernel source (host emulation) against the fixture for the first 20 steps."""
    from emu import EmuEnv

    g = np.load(os.path.join(HERE, "golden", "c1_oracle_trajectory.npz"))
    e = EmuEnv(model.blob(), reward_weights=(200.0, 1e-6, 1.0))
    assert np.abs(e.reset() - g["reset_obs"]).max() < 1e-6
    for t in range(20):
        ob, r, done = e.step(g["actions"][t])
        assert not done and np.isfinite(ob).all() and np.isfinite(r)
        # Free-running FP32 vs FP64 in contact is chaotic.  Step 11 of this trajectory (touch-down of the second foot, 6
        # contacts) takes a different active set in FP32 and in FP64 -- with any ordering of the solver's roundings -- and
        # the two trajectories separate from there on (1e-2 .. 2e-1 of the joint velocities within a few steps; the re-seeded
        # per-step comparisons of tests/test_active_set_parity.py bucket and bound such steps).  Up to the flip the
        # free-running kernel follows the fixture: per block, 2e-3 of the block's magnitude.
        if t >= 11:
            continue
        for blk in (slice(0, 25), slice(25, 50), slice(50, 75)):
            assert np.abs(ob[blk] - g["obs"][t][blk]).max() < 2e-3 * max(1.0, np.abs(g["obs"][t][blk]).max()), t
        assert abs(r - g["reward"][t]) < 1e-3 * max(1.0, abs(g["reward"][t])), t
