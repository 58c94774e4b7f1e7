"""On-device rollout plumbing for the caller of the path (BASELINE.json configs[3]).

The reference trains with baselines PPO2 (trex_train.py:35-63): ``DummyVecEnv`` -> ``VecNormalize`` ->
``ppo2.Runner`` collects ``nsteps`` transitions, then GAE(lambda=0.95, gamma=0.99).  Here the simulator writes
observations / rewards / dones straight into preallocated time-major device buffers (no host round trip per
step) and the advantage scan runs as one CUDA kernel (``trex_gae``).  The RL algorithm itself stays out of scope.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native
from .sim import TrexBatchSim


class RunningMeanStd:
    """baselines ``RunningMeanStd`` (VecNormalize) on the device: parallel-variance update per batch."""

    def __init__(self, dim: int, device, epsilon: float = 1e-4):
        self.mean = torch.zeros(dim, device=device, dtype=torch.float64)
        self.var = torch.ones(dim, device=device, dtype=torch.float64)
        self.count = float(epsilon)

    def update(self, x: torch.Tensor) -> None:
        x = x.reshape(-1, self.mean.shape[0]).double()
        b_mean, b_var, b_count = x.mean(0), x.var(0, unbiased=False), x.shape[0]
        delta = b_mean - self.mean
        tot = self.count + b_count
        self.mean = self.mean + delta * (b_count / tot)
        m2 = self.var * self.count + b_var * b_count + delta * delta * (self.count * b_count / tot)
        self.var = m2 / tot
        self.count = tot


class RolloutBuffer:
    """Time-major device buffers ``[T, N, ...]`` filled in place by :class:`TrexBatchSim`."""

    def __init__(self, sim: TrexBatchSim, horizon: int):
        self.sim, self.T, self.N = sim, int(horizon), sim.num_envs
        dev = sim.device
        f32, u8 = torch.float32, torch.uint8
        self.obs = torch.zeros(self.T + 1, self.N, _native.OBS_DIM, device=dev, dtype=f32)  # obs[t] = state step t acts in
        self.actions = torch.zeros(self.T, self.N, _native.NUM_JOINTS, device=dev, dtype=f32)
        self.rewards = torch.zeros(self.T, self.N, device=dev, dtype=f32)
        self.dones = torch.zeros(self.T + 1, self.N, device=dev, dtype=u8)  # dones[t] = obs[t] starts an episode
        self.values = torch.zeros(self.T, self.N, device=dev, dtype=f32)
        self.neglogp = torch.zeros(self.T, self.N, device=dev, dtype=f32)
        self.advantages = torch.zeros(self.T, self.N, device=dev, dtype=f32)
        self.returns = torch.zeros(self.T, self.N, device=dev, dtype=f32)
        self.obs[0].copy_(sim.obs)

    def collect(self, policy=None, seed: int = 0, step_offset: int = 0):
        """Run ``T`` env steps.  ``policy(obs[N,75]) -> (action[N,25], value[N], neglogp[N])`` on the device, or
        ``None`` for uniform random actions (the benchmark's synthetic input)."""
        sim = self.sim
        for t in range(self.T):
            if policy is None:
                sim.random_actions(step=step_offset + t, seed=seed, out=self.actions[t])
            else:
                a, v, nlp = policy(self.obs[t])
                self.actions[t].copy_(a)
                self.values[t].copy_(v)
                self.neglogp[t].copy_(nlp)
            # obs[t+1], rewards[t], dones[t+1] written by the kernels, no copies
            sim.step_into(self.actions[t], self.obs[t + 1], self.rewards[t], self.dones[t + 1])
        return self

    def compute_gae(self, last_value: torch.Tensor, gamma: float = 0.99, lam: float = 0.95):
        """trex_train.py:52-53 (lam=0.95, gamma=0.99); one kernel, one thread per environment."""
        last_value = last_value.to(self.sim.device, torch.float32).contiguous()
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.sim.device).cuda_stream)
        p = lambda x: ctypes.c_void_p(x.data_ptr())  # noqa: E731
        idx = self.sim.device.index if self.sim.device.index is not None else torch.cuda.current_device()
        _native.check(_native.lib().trex_gae(idx, p(self.rewards), p(self.values), p(self.dones[: self.T]), p(last_value),
                                             p(self.dones[self.T]), float(gamma), float(lam), p(self.advantages), p(self.returns),
                                             self.T, self.N, stream), "trex_gae")
        return self.advantages, self.returns

    def roll(self):
        """Start the next rollout from the last observation."""
        self.obs[0].copy_(self.obs[self.T])
        self.dones[0].copy_(self.dones[self.T])


def normalize(x: torch.Tensor, rms: RunningMeanStd, clip: float = 10.0, eps: float = 1e-8) -> torch.Tensor:
    """VecNormalize observation filter (clip 10) as one CUDA kernel."""
    x = x.contiguous()
    out = torch.empty_like(x)
    dim = x.shape[-1]
    stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    mean, var = rms.mean.float().contiguous(), rms.var.float().contiguous()
    idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
    _native.check(_native.lib().trex_normalize(idx, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(mean.data_ptr()),
                                               ctypes.c_void_p(var.data_ptr()), float(eps), float(clip),
                                               ctypes.c_void_p(out.data_ptr()), x.numel() // dim, dim, stream), "trex_normalize")
    return out
