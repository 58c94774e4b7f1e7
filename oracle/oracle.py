"""ctypes binding of the CPU oracle (``libtrex_oracle.so``).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see ``trex_oracle.h``).  Importable only from ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs;
nothing under ``trex_gym_b200/`` imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtrex_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "trex_oracle.c")
    hdr = os.path.join(_HERE, "trex_oracle.h")
    stale = (
        force
        or not os.path.isfile(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtrex_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        L.trex_oracle_create.restype = ctypes.c_void_p
        L.trex_oracle_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.trex_oracle_destroy.argtypes = [ctypes.c_void_p]
        L.trex_oracle_last_error.restype = ctypes.c_char_p
        for name in ("state_dim", "num_candidates", "last_iterations", "last_num_contacts", "last_num_limit_rows"):
            f = getattr(L, "trex_oracle_" + name)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_void_p]
        for name in ("total_iterations", "total_substeps"):
            f = getattr(L, "trex_oracle_" + name)
            f.restype = ctypes.c_long
            f.argtypes = [ctypes.c_void_p]
        L.trex_oracle_get_state.argtypes = [ctypes.c_void_p, dp]
        L.trex_oracle_set_state.argtypes = [ctypes.c_void_p, dp]
        L.trex_oracle_set_substeps.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.trex_oracle_set_reward_weights.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 3
        L.trex_oracle_enable_contacts.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.trex_oracle_set_fixed_base.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.trex_oracle_signature.restype = ctypes.c_uint
        L.trex_oracle_signature.argtypes = [ctypes.c_void_p]
        L.trex_oracle_reset_signature.argtypes = [ctypes.c_void_p]
        L.trex_oracle_signature_words.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint)]
        L.trex_oracle_reset.argtypes = [ctypes.c_void_p, dp]
        L.trex_oracle_step.argtypes = [ctypes.c_void_p, dp, dp, dp]
        L.trex_oracle_run.argtypes = [ctypes.c_void_p, dp, ctypes.c_int, dp, dp]
        L.trex_oracle_substep.argtypes = [ctypes.c_void_p, dp, ctypes.c_double]
        L.trex_oracle_head_position.argtypes = [ctypes.c_void_p, dp]
        L.trex_oracle_reward_terms.argtypes = [ctypes.c_void_p, dp]
        L.trex_oracle_momentum.argtypes = [ctypes.c_void_p, dp]
        L.trex_oracle_minv_column.argtypes = [ctypes.c_void_p, ctypes.c_int, dp]
        L.trex_oracle_residual_history.argtypes = [ctypes.c_void_p, dp, ctypes.c_int]
        L.trex_oracle_candidate_position.argtypes = [ctypes.c_void_p, ctypes.c_int, dp]
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class Oracle:
    """One double-precision T-rex environment (the reference is one env per process)."""

    def __init__(self, blob: bytes, num_substeps: int | None = None, contacts: bool = True,
                 reward_weights=(1.0, 0.005, 0.002)):
        self._L = lib()
        self._h = self._L.trex_oracle_create(blob, len(blob))
        if not self._h:
            raise RuntimeError("trex_oracle_create: " + self._L.trex_oracle_last_error().decode())
        if num_substeps is not None:
            self._L.trex_oracle_set_substeps(self._h, int(num_substeps))
        self._L.trex_oracle_enable_contacts(self._h, int(bool(contacts)))
        d, e, k = reward_weights  # distance, energy, drift (trex_env.py:42-44)
        self._L.trex_oracle_set_reward_weights(self._h, float(d), float(e), float(k))
        self.state_dim = self._L.trex_oracle_state_dim(self._h)
        self.num_candidates = self._L.trex_oracle_num_candidates(self._h)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.trex_oracle_destroy(h)

    def set_fixed_base(self, on: bool) -> None:
        self._L.trex_oracle_set_fixed_base(self._h, int(bool(on)))

    @property
    def signature(self) -> int:
        """Active-set signature of the last step (== record slot 158 of the kernels when both took the same branches)."""
        return int(self._L.trex_oracle_signature(self._h))

    def signature_words(self):
        w = (ctypes.c_uint * 6)()
        self._L.trex_oracle_signature_words(self._h, w)
        return [int(x) for x in w]

    def get_state(self) -> np.ndarray:
        s = np.zeros(self.state_dim)
        self._L.trex_oracle_get_state(self._h, _dp(s))
        return s

    def set_state(self, s) -> None:
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert s.shape == (self.state_dim,)
        self._L.trex_oracle_set_state(self._h, _dp(s))

    def reset(self) -> np.ndarray:
        obs = np.zeros(75)
        self._L.trex_oracle_reset(self._h, _dp(obs))
        return obs

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        assert a.shape == (25,)
        obs = np.zeros(75)
        r = ctypes.c_double(0.0)
        self._L.trex_oracle_step(self._h, _dp(a), _dp(obs), ctypes.byref(r))
        return obs, r.value

    def run(self, actions, want_obs=True):
        a = np.ascontiguousarray(actions, dtype=np.float64)
        T = a.shape[0]
        obs = np.zeros((T, 75)) if want_obs else None
        rew = np.zeros(T)
        self._L.trex_oracle_run(self._h, _dp(a), T, _dp(obs) if want_obs else None, _dp(rew))
        return obs, rew

    def substep(self, target_dof_order, max_impulse: float) -> None:
        t = np.ascontiguousarray(target_dof_order, dtype=np.float64)
        self._L.trex_oracle_substep(self._h, _dp(t), float(max_impulse))

    def head_position(self) -> np.ndarray:
        p = np.zeros(3)
        self._L.trex_oracle_head_position(self._h, _dp(p))
        return p

    def reward_terms(self) -> np.ndarray:
        p = np.zeros(3)
        self._L.trex_oracle_reward_terms(self._h, _dp(p))
        return p

    def momentum(self) -> dict:
        m = np.zeros(11)
        self._L.trex_oracle_momentum(self._h, _dp(m))
        return {"P": m[0:3], "L": m[3:6], "ke": m[6], "mass": m[7], "com": m[8:11]}

    def minv_column(self, dof: int) -> np.ndarray:
        out = np.zeros(31)
        self._L.trex_oracle_minv_column(self._h, int(dof), _dp(out))
        return out

    def residual_history(self) -> np.ndarray:
        out = np.zeros(512)
        self._L.trex_oracle_residual_history(self._h, _dp(out), 512)
        return out[: self.last_iterations]

    def candidate_position(self, k: int) -> np.ndarray:
        p = np.zeros(3)
        self._L.trex_oracle_candidate_position(self._h, int(k), _dp(p))
        return p

    @property
    def last_iterations(self) -> int:
        return self._L.trex_oracle_last_iterations(self._h)

    @property
    def last_num_contacts(self) -> int:
        return self._L.trex_oracle_last_num_contacts(self._h)

    @property
    def last_num_limit_rows(self) -> int:
        return self._L.trex_oracle_last_num_limit_rows(self._h)

    @property
    def total_iterations(self) -> int:
        return self._L.trex_oracle_total_iterations(self._h)

    @property
    def total_substeps(self) -> int:
        return self._L.trex_oracle_total_substeps(self._h)
