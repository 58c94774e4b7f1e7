"""GPU parity: the CUDA path (through the C ABI) against the double-precision CPU oracle.

PARITY UNPINNED: the oracle restates pybullet from recall (oracle/trex_oracle.h); these tests
prove kernel == oracle, the oracle itself is pinned only by the analytic KATs.

Stated FP32 tolerances (relative to the largest magnitude of the state block, per env step of
5 substeps, oracle re-seeded with the kernel's pre-step state):
  contact-free steps : 2e-5     steps with floor contact : 2e-3
"""
import numpy as np
import pytest

from conftest import STATE_BLOCKS, rel_err

pytestmark = pytest.mark.gpu

TOL_FREE = 2e-5
TOL_CONTACT = 2e-3


def _oracle(model, **kw):
    from oracle.oracle import Oracle

    return Oracle(model.blob(), **kw)


def _sim(model, n, **kw):
    from trex_gym_b200.sim import TrexBatchSim

    return TrexBatchSim(n, device=0, model=model, **kw)


def _core(state_row, ncand):
    s = np.asarray(state_row, np.float64)
    return np.concatenate([s[:88], s[88:88 + ncand]])


def test_library_loaded_and_version():
    from trex_gym_b200 import _native

    assert b"sm_100a" in _native.lib().trex_version()


def test_reset_matches_oracle(model):
    sim = _sim(model, 8)
    o = _oracle(model)
    oobs = o.reset()
    obs = sim.reset().cpu().numpy()
    st = sim.get_state().cpu().numpy()
    for e in range(8):
        assert np.abs(obs[e] - oobs).max() < 1e-6
        assert np.abs(_core(st[e], o.num_candidates) - o.get_state()).max() < 1e-6
    # analytic KAT: one free-fall substep from rest
    assert abs(st[0, 12] - (-9.81 * 0.002)) < 1e-7
    assert abs(st[0, 2] - (3.0 - 9.81 * 0.002 ** 2)) < 1e-6


def _parity_campaign(model, contacts, n, n_check, steps, n_sub=5, model_for_sub=None):
    """Step a batch with random actions; for a subset of envs re-seed the oracle with the kernel's
    pre-step state every step and collect the per-step relative error of every state block."""
    import torch

    mdl = model_for_sub if model_for_sub is not None else model
    sim = _sim(mdl, n, contacts=contacts, num_substeps=n_sub)
    # per-env perturbation of the reset state: q0 += U(-0.05, 0.05) clipped to limits, seed = env id
    st = sim.get_state().cpu().numpy()
    qlo, qhi = model["mb_lower"][1:], model["mb_upper"][1:]
    for e in range(n):
        r = np.random.default_rng(e)
        st[e, 13:38] = np.clip(st[e, 13:38] + r.uniform(-0.05, 0.05, 25), qlo, qhi)
    sim.set_state(torch.from_numpy(st).cuda())
    o = _oracle(mdl, contacts=contacts, num_substeps=n_sub)
    nc = o.num_candidates
    errs = {k: [] for k in STATE_BLOCKS}
    rew_err, n_contact_steps = [], 0
    for t in range(steps):
        pre = sim.get_state().cpu().numpy()
        act = sim.random_actions(step=t, seed=0)
        obs, rew, done = sim.step(act)
        post = sim.get_state().cpu().numpy()
        a = act.cpu().numpy()
        rew = rew.cpu().numpy()
        assert not done.any().item()
        for e in range(0, n, n // n_check):
            o.set_state(_core(pre[e], nc))
            oobs, orew = o.step(a[e].astype(np.float64))
            so = o.get_state()
            n_contact_steps += o.last_num_contacts > 0
            for k, sl in STATE_BLOCKS.items():
                errs[k].append(rel_err(so[sl], post[e, sl]))
            rew_err.append(abs(orew - rew[e]) / max(1.0, abs(orew)))
    return {k: np.asarray(v) for k, v in errs.items()}, np.asarray(rew_err), n_contact_steps


def _report(tag, errs, rew_err):
    allv = np.max(np.stack(list(errs.values())), axis=0)
    print("%s: n=%d  p50 %.2e  p90 %.2e  p99 %.2e  max %.2e | per block max %s | reward max %.2e" % (
        tag, allv.size, np.percentile(allv, 50), np.percentile(allv, 90), np.percentile(allv, 99), allv.max(),
        {k: "%.1e" % v.max() for k, v in errs.items()}, rew_err.max()))
    return allv


def test_per_step_parity_contact_free(model):
    """BASELINE.json configs[1] on the literal reference model (no collision shapes => free fall):
    4,096 envs, per-step state delta vs the oracle on 64 of them, every step.  Tolerance 2e-5."""
    errs, rew_err, _ = _parity_campaign(model, False, 4096, 64, 40)
    allv = _report("contact-free per env-step", errs, rew_err)
    assert allv.max() < TOL_FREE
    assert rew_err.max() < 10 * TOL_FREE


def test_per_substep_parity_with_contacts(model):
    """With floor contact, per PHYSICS step (one stepSimulation of 2 ms, 60 PGS iterations): the
    cleanest measure of arithmetic agreement.  The solver is far from converged under random
    actions (residual >> 1), so FP32 round-off is amplified; the bound is on percentiles."""
    from trex_gym_b200.model_compiler import with_params

    sub = with_params(model, time_step=0.002, solver_iterations=60)
    errs, rew_err, n_contact = _parity_campaign(model, True, 4096, 64, 100, n_sub=1, model_for_sub=sub)
    allv = _report("contact per physics substep", errs, rew_err)
    assert n_contact > 500
    assert np.percentile(allv, 50) < 2e-5
    assert np.percentile(allv, 99) < 2e-3
    assert allv.max() < 0.1


def test_per_step_parity_with_contacts(model):
    """BASELINE.json configs[1] with the derived contact geometry: per env-step (5 substeps)."""
    errs, rew_err, n_contact = _parity_campaign(model, True, 4096, 64, 40)
    allv = _report("contact per env-step", errs, rew_err)
    assert n_contact > 200
    assert np.percentile(allv, 50) < 1e-4
    assert np.percentile(allv, 99) < 2e-2
    assert np.isfinite(allv).all()


def test_rollout_divergence_100_steps(model, action_limits):
    """Free-running 100-step divergence kernel vs oracle (reported; loose bound: chaotic under contact)."""
    n = 16
    sim = _sim(model, n, contacts=False)
    orcs = [_oracle(model, contacts=False) for _ in range(n)]
    for o in orcs:
        o.reset()
    div = []
    for t in range(100):
        act = sim.random_actions(step=t, seed=1)
        sim.step(act)
        a = act.cpu().numpy().astype(np.float64)
        st = sim.get_state().cpu().numpy()
        d = 0.0
        for e, o in enumerate(orcs):
            o.step(a[e])
            so = o.get_state()
            d = max(d, rel_err(so[13:38], st[e, 13:38]), rel_err(so[0:3], st[e, 0:3]))
        div.append(d)
    print("contact-free 100-step divergence: step10 %.2e step50 %.2e step100 %.2e" % (div[9], div[49], div[99]))
    assert div[9] < 1e-4
    assert np.isfinite(div[99])


def test_reward_bit_exact_from_outputs(model):
    """Reward recomputed in FP32 from the kernel's own outputs (same op order as trex_env.py:186-196)."""
    sim = _sim(model, 256, distance_weight=200.0, energy_weight=1e-6, drift_weight=1.0)  # trex_train.py:66
    f = np.float32
    for t in range(20):
        obs, rew, done = sim.step(sim.random_actions(step=t, seed=3))
        obs, rew, aux = obs.cpu().numpy(), rew.cpu().numpy(), sim.aux().cpu().numpy()
        for e in range(256):
            power = f(0)
            for k in range(25):
                power = f(power + f(abs(f(obs[e, 25 + k] * obs[e, 50 + k]))))
            x, y, z = aux[e, 0], aux[e, 1], aux[e, 2]
            dz = f(f(2.5) - z)
            lifting = f(f(200.0) * f(dz * dz))
            station = f(f(1.0) * f(f(x * x) + f(y * y)))
            energy = f(f(1e-6) * power)
            expect = f(f(-lifting - station) - energy)
            assert expect == rew[e], (t, e, expect, rew[e])
            assert aux[e, 3] == lifting and aux[e, 4] == station and aux[e, 5] == energy
        assert not done.any()


def test_done_flags_and_auto_reset(model):
    """Reference never terminates (trex_env.py:183-184); with a horizon every env is done exactly at
    the horizon, auto-reset, and returns the reset observation (VecEnv semantics)."""
    sim = _sim(model, 64, max_episode_steps=5)
    reset_obs = sim.reset().clone()
    for t in range(12):
        obs, rew, done = sim.step(sim.random_actions(step=t))
        d = done.cpu().numpy()
        if (t + 1) % 5 == 0:
            assert d.all()
            assert (obs == reset_obs).all().item()
        else:
            assert not d.any()
    s = sim.stats()
    assert s["episodes"] == 64 * 4  # create + explicit reset + 2 horizons
    assert s["nan_resets"] == 0
    sim2 = _sim(model, 64)
    for t in range(12):
        _, _, done = sim2.step(sim2.random_actions(step=t))
        assert not done.any().item()


def test_state_roundtrip_and_env_independence(model):
    import torch

    sim = _sim(model, 128)
    for t in range(5):
        sim.step(sim.random_actions(step=t))
    st = sim.get_state().clone()
    # environment 0's state copied everywhere, identical actions -> bitwise identical results
    st2 = st[0:1].repeat(128, 1).contiguous()
    sim.set_state(st2)
    assert (sim.get_state() == st2).all().item()
    a = sim.random_actions(step=99)[0:1].repeat(128, 1).contiguous()
    obs, rew, _ = sim.step(a)
    assert (obs == obs[0:1]).all().item() and (rew == rew[0]).all().item()
    assert (sim.get_state() == sim.get_state()[0:1]).all().item()


def test_random_actions_keyed_by_global_env(model, action_limits):
    lo, hi = action_limits
    a = _sim(model, 64).random_actions(step=7, seed=5, env_offset=0).cpu().numpy()
    b = _sim(model, 32).random_actions(step=7, seed=5, env_offset=32).cpu().numpy()
    assert (a[32:] == b).all()
    assert (a >= lo.astype(np.float32) - 1e-6).all() and (a <= hi.astype(np.float32) + 1e-6).all()
    assert np.abs(a.mean(0) - (lo + hi) / 2).max() < 0.6


@pytest.mark.parametrize("n", [96, 9000])
def test_step_host_equals_device_step(model, n):
    """trex_step_host == trex_step bit for bit: with one group of environments (one copy in, the step, one copy out) and with
    two (from 8,192 environments: every group's copies ride on its own stream and the second group starts behind the first
    group's first dynamics kernel, so the first group's copy out runs under the second group's last solve)."""
    s1, s2 = _sim(model, n), _sim(model, n)
    for t in range(4):
        act = s1.random_actions(step=t)
        obs, rew, done = s1.step(act)
        hobs, hrew, hdone = s2.step_host(act.cpu().numpy())
        assert (obs.cpu().numpy() == hobs).all() and (rew.cpu().numpy() == hrew).all() and (done.cpu().numpy() == hdone).all()
    import torch

    assert torch.equal(s1.get_state(), s2.get_state())


def test_gym_surface(model):
    from trex_gym_b200 import TrexBulletEnv, TrexVecEnv

    env = TrexBulletEnv(urdf_path=None, model=model)
    assert env.action_space.shape == (25,) and env.observation_space.shape == (75,)
    o = _oracle(model)
    oobs = o.reset()
    obs = env.reset()
    assert len(obs) == 75 and np.abs(np.asarray(obs) - oobs).max() < 1e-6
    a = env.action_space.sample()
    obs, rew, done, info = env.step(a)
    oobs2, orew = o.step(a.astype(np.float64))
    assert done is False and info == {}
    assert np.abs(np.asarray(obs) - oobs2).max() < 1e-3 * max(1.0, np.abs(oobs2).max())
    assert abs(rew - orew) < 1e-3 * max(1.0, abs(orew))
    assert len(env.model._revolute_joint_indices) == 25 and abs(env.model._total_mass - 4834.866376) < 1e-3
    assert np.allclose(env.model.get_head_position(), o.head_position(), atol=1e-4)
    with pytest.raises(ValueError):
        env.step(np.zeros(24))
    venv = TrexVecEnv(32, model=model)
    vobs = venv.reset()
    assert tuple(vobs.shape) == (32, 75)
    vobs, vrew, vdone, _ = venv.step(np.zeros((32, 25), np.float32))
    assert tuple(vrew.shape) == (32,) and not vdone.any().item()


def test_substep_sweep_matches_oracle(model):
    """BASELINE.json configs[4]: num_substeps 1..8, dt = 0.01/n, iterations = int(300/n) (trex_env.py:71-72)."""
    for n_sub in (1, 2, 3, 8):
        sim = _sim(model, 4, num_substeps=n_sub)
        o = _oracle(model, num_substeps=n_sub)
        o.reset()
        sim.reset()
        nc = o.num_candidates
        for t in range(3):
            pre = sim.get_state().cpu().numpy()
            act = sim.random_actions(step=t, seed=n_sub)
            sim.step(act)
            post = sim.get_state().cpu().numpy()
            o.set_state(_core(pre[0], nc))
            o.step(act[0].cpu().numpy().astype(np.float64))
            so = o.get_state()
            for k, sl in STATE_BLOCKS.items():
                assert rel_err(so[sl], post[0, sl]) < TOL_CONTACT, (n_sub, k)


def test_contact_stress_fallen_starts(model):
    """BASELINE.json configs[4]: fallen starts (reset_mode 1), every episode ends in auto-reset (horizon 16 here),
    substep counts 1..8.  Checks: finite states, exact done cadence, reproducible sampler keyed by global env id,
    and the post-reset state equals the oracle's one zero-force step from the sampled pose."""
    import torch

    for n_sub in (1, 5, 8):
        sim = _sim(model, 512, num_substeps=n_sub, max_episode_steps=16, reset_mode=1, seed=11, env_offset=1000)
        for t in range(33):
            obs, rew, done = sim.step(sim.random_actions(step=t, seed=2, env_offset=1000))
            assert bool(done.all().item()) == ((t + 1) % 16 == 0)
            assert torch.isfinite(obs).all().item() and torch.isfinite(rew).all().item()
        st = sim.stats()
        assert st["episodes"] == 512 * 3 and st["nan_resets"] == 0
        assert st["mean_contacts"] > 0.5
    a = _sim(model, 64, reset_mode=1, seed=5, env_offset=0).get_state().cpu().numpy()
    b = _sim(model, 32, reset_mode=1, seed=5, env_offset=32).get_state().cpu().numpy()
    assert (a[32:] == b).all()
    assert np.abs(a[0, 13:38] - a[1, 13:38]).max() > 0.05
    # orientation is a unit quaternion, height within the sampler's range minus one free-fall step
    assert np.abs(np.linalg.norm(a[:, 3:7], axis=1) - 1).max() < 1e-5
    assert (a[:, 2] > 0.29).all() and (a[:, 2] < 3.01).all()


def test_rollout_buffers_and_gae(model):
    """BASELINE.json configs[3] plumbing: in-place rollout collection with auto-reset and the
    GAE kernel, bit-exact against the baselines recursion evaluated in float32 numpy."""
    import torch

    from trex_gym_b200.rollout import RolloutBuffer, RunningMeanStd, normalize

    T, N = 24, 256
    sim = _sim(model, N, max_episode_steps=10, distance_weight=200.0, energy_weight=1e-6, drift_weight=1.0)
    sim.reset()
    buf = RolloutBuffer(sim, T)
    buf.collect(policy=None, seed=4)
    d = buf.dones.cpu().numpy()
    assert d[10].all() and d[20].all() and not d[1:10].any() and not d[11:20].any()
    assert torch.isfinite(buf.obs).all().item() and torch.isfinite(buf.rewards).all().item()
    assert (buf.obs[10] == buf.obs[20]).all().item()  # both are the (deterministic) reset observation
    g = torch.Generator(device="cpu").manual_seed(0)
    buf.values.copy_(torch.randn(T, N, generator=g).cuda())
    last_v = torch.randn(N, generator=g).cuda()
    adv, ret = buf.compute_gae(last_v, gamma=0.99, lam=0.95)
    f = np.float32
    r, v, dn = buf.rewards.cpu().numpy(), buf.values.cpu().numpy(), d
    exp = np.zeros((T, N), np.float32)
    a = np.zeros(N, np.float32)
    for t in reversed(range(T)):
        nonterm = (f(1.0) - dn[t + 1].astype(np.float32)).astype(np.float32)
        nv = last_v.cpu().numpy() if t == T - 1 else v[t + 1]
        delta = ((r[t] + (f(0.99) * nv) * nonterm).astype(np.float32) - v[t]).astype(np.float32)
        a = (delta + ((f(0.99) * f(0.95)) * nonterm).astype(np.float32) * a).astype(np.float32)
        exp[t] = a
    assert np.array_equal(adv.cpu().numpy(), exp)
    assert np.array_equal(ret.cpu().numpy(), (exp + v).astype(np.float32))
    rms = RunningMeanStd(75, sim.device)
    rms.update(buf.obs[:T])
    z = normalize(buf.obs[:T], rms)
    ref = torch.clamp((buf.obs[:T] - rms.mean.float()) / torch.sqrt(rms.var.float() + 1e-8), -10, 10)
    assert torch.allclose(z, ref, atol=1e-5) and z.abs().max().item() <= 10.0


def test_fused_policy_forward(model):
    """the trainer's policy / value networks as one kernel, against the same networks in plain
    PyTorch FP32 (tolerance 2e-5 on means and values), the diagonal-Gaussian neglogp recomputed from the outputs, the
    Philox noise (keyed by global environment and step: independent of the sharding) and the in-place rollout path."""
    import torch

    from trex_gym_b200.rollout import MlpPolicy, RolloutBuffer, RunningMeanStd

    torch.backends.cuda.matmul.allow_tf32 = False
    N = 1000  # not a multiple of the 128-row tile
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(3)
    obs = (torch.randn(N, 75, generator=g) * torch.logspace(-1, 1.5, 75)).to(dev)
    rms = RunningMeanStd(75, dev)
    rms.update(obs)
    pol = MlpPolicy(dev, seed=1, rms=rms)
    # non-trivial heads and log-std so every term of the outputs is exercised
    pol.view("pi_wo").mul_(50.0)
    pol.view("pi_bo").copy_(torch.linspace(-0.3, 0.3, 25))
    pol.view("vf_bo").fill_(0.7)
    pol.view("logstd").copy_(torch.linspace(-1.0, 0.2, 25))
    mean = torch.empty(N, 25, device=dev)
    a, v, nlp = pol(obs, step=5, seed=9)
    pol.act_into(obs, torch.empty_like(a), None, None, mean=mean, step=5, seed=9)
    ref_mean, ref_v = pol.reference_forward(obs)
    assert (mean - ref_mean).abs().max().item() < 2e-5 * max(1.0, ref_mean.abs().max().item())
    assert (v - ref_v).abs().max().item() < 2e-5 * max(1.0, ref_v.abs().max().item())
    # sampling: eps = (a - mean) / std is N(0,1); neglogp is the DiagGaussianPd formula on those outputs
    ls = pol.view("logstd")
    eps = ((a - mean) / torch.exp(ls)).double()
    assert abs(eps.mean().item()) < 0.02 and abs(eps.var().item() - 1.0) < 0.03 and eps.abs().max().item() < 6.5
    ref_nlp = 0.5 * (eps * eps).sum(1) + ls.double().sum() + 0.5 * np.log(2.0 * np.pi) * 25
    assert (nlp.double() - ref_nlp).abs().max().item() < 1e-3
    # deterministic mode returns the mean; the same (seed, step, global env) gives the same noise on any shard
    a_det, _, nlp_det = pol(obs, step=5, seed=9, deterministic=True)
    assert torch.equal(a_det, mean)
    a_shard = torch.empty(N - 300, 25, device=dev)
    pol.act_into(obs[300:].contiguous(), a_shard, None, None, step=5, seed=9, env_offset=300)
    assert torch.equal(a_shard, a[300:])
    a_other, _, _ = pol(obs, step=6, seed=9)
    assert not torch.equal(a_other, a)
    # rollout: the kernel reads obs[t] in place and fills actions / values / neglogp of the buffers
    sim = _sim(model, 256)
    sim.reset()
    buf = RolloutBuffer(sim, 6)
    buf.collect(policy=MlpPolicy(dev, seed=2), seed=4)
    assert torch.isfinite(buf.actions).all().item() and torch.isfinite(buf.neglogp).all().item()
    assert buf.values.abs().max().item() > 0 and (buf.actions[0] != buf.actions[1]).any().item()


def test_packed_inward_pass_is_bit_identical(model, action_limits):
    """4-warp CTAs run the inward pass of their four environments on one warp (inward_packed, eight lanes per
    environment, through the shared slabs): same operations in the same order as the one-environment pass, so the
    states agree bit for bit with the default 2-warp configuration -- also when the last CTA is only half full."""
    import torch

    n = 1026
    sims = [_sim(model, n, warps_per_block=w) for w in (2, 4)]
    for s in sims:
        s.reset()
    for t in range(40):
        a = sims[0].random_actions(step=t, seed=3)
        for s in sims:
            s.step(a)
    st = [s.get_state() for s in sims]
    assert torch.equal(st[0][:, :152], st[1][:, :152])
    assert sims[1].stats()["mean_contacts"] > 0.0


def test_heavy_contact_kernel_on_standing_batch(model):
    """Every environment standing on both feet (12-16 contacts each): trex_heavy_kernel forced for the whole batch against
    the default placement (a large share of such environments is solved inside the front kernel), and the kernel's
    per-env-step parity with the oracle on the same states.  Both solvers run the same Gauss-Seidel iteration."""
    import torch

    from trex_gym_b200.model_compiler import load_builtin

    n = 96
    names = list(model.meta["obs_joint_names"])
    hold = np.zeros(25, np.float32)
    for k, v in model.meta["starting_configuration"].items():
        hold[names.index(k)] = v
    a = torch.tensor(hold, device="cuda").repeat(n, 1).contiguous()
    forced = _sim(model, n)
    default = _sim(model, n, heavy_solver=False)
    forced.reset()
    default.reset()
    o = _oracle(model)
    nc = o.num_candidates
    errs, diffs, ks = [], [], []
    for t in range(60):
        pre = forced.get_state()
        default.set_state(pre)
        forced.step(a)
        default.step(a)
        sf, sd = forced.get_state().cpu().numpy().astype(np.float64), default.get_state().cpu().numpy().astype(np.float64)
        k = int(forced.aux()[0, 7].item()) % 1000
        if k > 8:
            ks.append(k)
            p = pre[0].cpu().numpy().astype(np.float64)
            o.set_state(np.concatenate([p[:88], p[88:88 + nc]]))
            o.step(hold)
            so = o.get_state()
            se = np.concatenate([sf[0, :88], sf[0, 88:88 + nc]])
            errs.append(max(rel_err(so[sl], se[sl]) for sl in STATE_BLOCKS.values()))
            diffs.append(max(rel_err(sd[:, sl], sf[:, sl]) for sl in STATE_BLOCKS.values()))
    errs, diffs = np.asarray(errs), np.asarray(diffs)
    print("standing batch, per env-step: heavy kernel vs oracle p50 %.2e max %.2e | vs front-kernel sweep p50 %.2e max %.2e | contacts %d..%d"
          % (np.percentile(errs, 50), errs.max(), np.percentile(diffs, 50), diffs.max(), min(ks), max(ks)))
    assert len(ks) >= 25 and max(ks) >= 14
    # measured: 4.5e-4 / 5.2e-3 against the oracle (14-16 stacked points: ill-conditioned in FP32), 5.4e-5 / 2.1e-3 between the solvers
    assert np.percentile(errs, 50) < 2e-3 and errs.max() < 2e-2, (np.percentile(errs, 50), errs.max())
    assert np.percentile(diffs, 50) < 3e-4 and diffs.max() < 1e-2, (np.percentile(diffs, 50), diffs.max())
    assert torch.isfinite(forced.get_state()).all().item()


def test_contact_primitives_model_on_gpu(model):
    """the contact model of spheres / capsules fitted to the meshes (62 sphere candidates with a
    radius) through the CUDA path: a batch that lands and stands on its toe capsules, then flails with random actions,
    per env step against the oracle on the same candidate table."""
    import torch

    from trex_gym_b200 import TrexVecEnv
    from trex_gym_b200.model_compiler import load_builtin

    prim = load_builtin("primitives")
    n = 256
    sim = _sim(prim, n)
    o = _oracle(prim)
    nc = o.num_candidates
    assert nc == 62
    names = list(prim.meta["obs_joint_names"])
    hold = np.zeros(25, np.float32)
    for k, v in prim.meta["starting_configuration"].items():
        hold[names.index(k)] = v
    a_hold = torch.tensor(hold, device="cuda").repeat(n, 1).contiguous()
    errs, ks = [], []
    for t in range(90):
        act = a_hold if t < 60 else sim.random_actions(step=t, seed=6)
        pre = sim.get_state().cpu().numpy()
        sim.step(act)
        post = sim.get_state().cpu().numpy()
        a = act.cpu().numpy().astype(np.float64)
        for e in (0, 100, 255):
            o.set_state(_core(pre[e], nc))
            o.step(a[e])
            if o.last_num_contacts:
                so = o.get_state()
                errs.append(max(rel_err(so[sl], post[e, sl]) if k != "tau" else float(np.abs(so[sl] - post[e, sl]).max() / max(np.abs(so[sl]).max(), 1.0))
                                for k, sl in STATE_BLOCKS.items()))
                ks.append(o.last_num_contacts)
    errs = np.asarray(errs)
    print("contact primitives, per env step: n=%d p50 %.2e p99 %.2e max %.2e contacts up to %d" % (
        len(errs), np.percentile(errs, 50), np.percentile(errs, 99), errs.max(), max(ks)))
    assert len(errs) > 100 and max(ks) >= 8
    assert np.percentile(errs, 50) < 5e-4 and np.percentile(errs, 95) < 2e-2
    st = sim.stats()
    assert st["nan_resets"] == 0 and st["contact_overflow"] == 0
    # the gym surface takes the contact model by name
    venv = TrexVecEnv(4, contact_model="primitives")
    assert venv.sim.model.meta["contact_model"] == "primitives"
    venv.close()


def test_solve2_tensor_memory_and_shared_memory_instances_agree(model):
    """The many-contact solver runs as two kernel instances on one task list: Delassus matrices in TENSOR MEMORY
    (tcgen05.st / tcgen05.ld as a per-lane scratchpad) and in shared memory.  Alone or together, with one or two groups of
    environments, they give bit-identical records on a batch where every environment stands on both feet (12-16 contacts)
    next to environments flailing with random actions."""
    import torch

    from trex_gym_b200 import _native

    n = 4096 + 6
    names = list(model.meta["obs_joint_names"])
    hold = torch.zeros(25, device="cuda")
    for k, v in model.meta["starting_configuration"].items():
        hold[names.index(k)] = v
    sims = [_sim(model, n, heavy_memory=m, pipelines=p) for m, p in ((_native.HEAVY_SHARED, 1), (_native.HEAVY_TENSOR, 1),
                                                                     (_native.HEAVY_BOTH, 1), (_native.HEAVY_BOTH, 2))]
    for t in range(45):
        a = sims[0].random_actions(step=t, seed=7)
        a[: n - 500] = hold
        for s in sims:
            s.step(a)
    st = [s.get_state() for s in sims]
    for x in st[1:]:
        assert torch.equal(st[0], x)
    stats = sims[0].stats()
    assert stats["mean_contacts"] > 10 and stats["nan_resets"] == 0 and stats["contact_overflow"] >= 0
    k = (sims[0].aux()[:, 7] % 1000)
    assert (k > 8).float().mean().item() > 0.8  # the batch really is in the many-contact class


def test_contact_solver_tensor_memory_and_shared_memory_instances_agree(model):
    """trex_config.contact_memory: the contact solver (1-8 contacts, four environments per warp) with its Delassus blocks and
    sweep responses in TENSOR MEMORY (persistent 4-warp CTAs pulling tasks from a counter; default) and in shared memory
    (one-warp CTAs, static assignment) give bit-identical records on a flailing batch, with one or two groups in flight."""
    import torch

    from trex_gym_b200 import _native

    n = 8192 + 10
    sims = [_sim(model, n, contact_memory=m, pipelines=p, seed=3) for m, p in ((_native.CONTACT_SHARED, 1), (_native.CONTACT_TENSOR, 1),
                                                                               (_native.CONTACT_TENSOR, 2))]
    for t in range(120):
        a = sims[0].random_actions(step=t, seed=5)
        for s in sims:
            s.step(a)
    st = [s.get_state() for s in sims]
    for x in st[1:]:
        assert torch.equal(st[0], x)
    k = (sims[0].aux()[:, 7] % 1000)
    assert ((k >= 1) & (k <= 8)).float().mean().item() > 0.2 and sims[0].stats()["nan_resets"] == 0  # the contact classes are populated
