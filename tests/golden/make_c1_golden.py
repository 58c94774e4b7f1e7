#!/usr/bin/env python
"""Generates tests/golden/c1_oracle_trajectory.npz: BASELINE.json configs[0] as far as it can be run here --
one environment, random actions, 1,000 steps -- stepped by the CPU ORACLE (pybullet is not installable in
this image, so this is a regression pin of the oracle, NOT a pybullet golden vector; parity stays unpinned).

    python tests/golden/make_c1_golden.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.oracle import Oracle  # noqa: E402
from trex_gym_b200.model_compiler import load_builtin  # noqa: E402


def main():
    model = load_builtin()
    lo = model["mb_lower"][1:][model["obs_dof"]]
    hi = model["mb_upper"][1:][model["obs_dof"]]
    rng = np.random.Generator(np.random.Philox(key=20261018))
    actions = rng.uniform(lo, hi, size=(1000, 25))
    o = Oracle(model.blob(), reward_weights=(200.0, 1e-6, 1.0))  # trex_train.py:66
    obs0 = o.reset()
    t0 = time.perf_counter()
    obs, rew, head, base, ncon = [], [], [], [], []
    for t in range(1000):
        ob, r = o.step(actions[t])
        if t < 50 or t % 10 == 9:
            obs.append(ob)
            rew.append(r)
            head.append(o.head_position())
            base.append(o.get_state()[:13])
            ncon.append(o.last_num_contacts)
    wall = time.perf_counter() - t0
    steps = np.array([t for t in range(1000) if t < 50 or t % 10 == 9])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_oracle_trajectory.npz")
    np.savez_compressed(out, actions=actions.astype(np.float64), steps=steps, reset_obs=obs0, obs=np.array(obs),
                        reward=np.array(rew), head=np.array(head), base=np.array(base), contacts=np.array(ncon),
                        final_state=o.get_state(), mean_iterations=o.total_iterations / o.total_substeps)
    print("wrote %s (%d samples); oracle: %.1f env-steps/s on one core" % (out, len(steps), 1000 / wall))


if __name__ == "__main__":
    main()
