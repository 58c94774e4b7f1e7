"""ctypes binding of the host warp emulator (TEST INFRASTRUCTURE ONLY).

``libtrex_emu.so`` is the product kernel source ``trex_gym_b200/csrc/trex_core.h`` compiled
against ``lane_emu.h``: the same statements the GPU executes per lane, run as 32-wide loops
on the CPU, so that the CPU test-suite exercises the product arithmetic without a GPU.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.normpath(os.path.join(_HERE, "..", "..", "trex_gym_b200", "csrc"))
_LIB = os.path.join(_HERE, "libtrex_emu.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, "emu_main.cpp"), os.path.join(_HERE, "lane_emu.h")] + [
        os.path.join(_CSRC, f) for f in ("trex_core.h", "trex_model.h", "trex_topology.h")
    ]
    stale = force or not os.path.isfile(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(s) for s in srcs)
    if stale:
        subprocess.check_call(
            ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-I" + _CSRC, "-o", _LIB,
             os.path.join(_HERE, "emu_main.cpp")]
        )
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        L.emu_create.restype = ctypes.c_void_p
        L.emu_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                 ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint]
        L.emu_destroy.argtypes = [ctypes.c_void_p]
        L.emu_last_error.restype = ctypes.c_char_p
        L.emu_state_stride.restype = ctypes.c_int
        L.emu_shared_bytes.restype = ctypes.c_int
        fp = ctypes.POINTER(ctypes.c_float)
        L.emu_step.argtypes = [ctypes.c_void_p, fp, fp, fp, fp, ctypes.POINTER(ctypes.c_uint8), fp, ctypes.c_int, ctypes.c_longlong]
        L.emu_step4.argtypes = [ctypes.c_void_p, ctypes.c_int, fp, fp, fp, fp, ctypes.POINTER(ctypes.c_uint8), fp, ctypes.c_int,
                                ctypes.c_longlong]
        L.emu_set_deferred.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


class EmuEnv:
    """One environment stepped by the emulated warp.  State record = float32[stride]."""

    def __init__(self, blob: bytes, num_substeps=5, reward_weights=(1.0, 0.005, 0.002), max_episode_steps=0,
                 contacts=True, reset_mode=0, seed=0, env_id=0, deferred=True):
        self._L = lib()
        d, e, k = reward_weights
        self._h = self._L.emu_create(blob, len(blob), num_substeps, d, e, k, max_episode_steps, int(contacts), int(reset_mode), int(seed))
        self.env_id = int(env_id)
        self._L.emu_set_deferred(self._h, int(deferred))  # bit 0: deferred solve, bit 1: pack lane groups in reverse
        if not self._h:
            raise RuntimeError(self._L.emu_last_error().decode())
        self.stride = self._L.emu_state_stride()
        self.rec = np.zeros(self.stride, np.float32)
        self.rec[6] = 1.0
        self.aux = np.zeros(8, np.float32)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.emu_destroy(h)

    def solve_counts(self):
        """Substeps finished by (the one-environment path, solve4<0>, solve4<KC>)."""
        out = (ctypes.c_longlong * 5)()
        self._L.emu_solve_counts(self._h, out)
        return tuple(int(x) for x in out)[:3]

    def heavy_solves(self):
        """Substeps finished by solve_heavy (more than 8 contacts, one environment per warp, row space)."""
        out = (ctypes.c_longlong * 5)()
        self._L.emu_solve_counts(self._h, out)
        return int(out[4])

    def packed_rounds(self):
        """Substep rounds whose front phase ran as a 4-warp CTA (inward pass of four environments by one warp)."""
        out = (ctypes.c_longlong * 5)()
        self._L.emu_solve_counts(self._h, out)
        return int(out[3])

    def reset(self):
        obs = np.zeros(75, np.float32)
        r = np.zeros(1, np.float32)
        d = np.zeros(1, np.uint8)
        self._L.emu_step(self._h, _fp(self.rec), None, _fp(obs), _fp(r), d.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                         _fp(self.aux), 1, self.env_id)
        return obs

    def step(self, action):
        a = np.ascontiguousarray(action, np.float32)
        obs = np.zeros(75, np.float32)
        r = np.zeros(1, np.float32)
        d = np.zeros(1, np.uint8)
        self._L.emu_step(self._h, _fp(self.rec), _fp(a), _fp(obs), _fp(r), d.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                         _fp(self.aux), 0, self.env_id)
        return obs, float(r[0]), bool(d[0])

    # state in the oracle's layout (63 core + 25 tau + n_cand lambda)
    def get_state(self, n_cand):
        return np.concatenate([self.rec[:88], self.rec[88 : 88 + n_cand]]).astype(np.float64)

    def set_state(self, s):
        s = np.asarray(s, np.float64)
        self.rec[:88] = s[:88]
        self.rec[88 : 88 + (len(s) - 88)] = s[88:]


class EmuWarp4:
    """Up to four consecutive environments stepped by ONE emulated warp (the kernel's work unit)."""

    def __init__(self, blob: bytes, n=4, deferred=True, **kw):
        self.env = EmuEnv(blob, deferred=deferred, **kw)
        self.n = n
        self.stride = self.env.stride
        self.rec = np.zeros((n, self.stride), np.float32)
        self.rec[:, 6] = 1.0
        self.aux = np.zeros((n, 8), np.float32)

    def _call(self, action, force_reset):
        L = self.env._L
        obs = np.zeros((self.n, 75), np.float32)
        r = np.zeros(self.n, np.float32)
        d = np.zeros(self.n, np.uint8)
        a = None if action is None else _fp(np.ascontiguousarray(action, np.float32))
        L.emu_step4(self.env._h, self.n, _fp(self.rec), a, _fp(obs), _fp(r), d.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                    _fp(self.aux), int(force_reset), self.env.env_id)
        return obs, r, d

    def reset(self):
        return self._call(None, True)[0]

    def step(self, actions):
        return self._call(actions, False)

    def get_state(self, e, n_cand):
        return np.concatenate([self.rec[e, :88], self.rec[e, 88:88 + n_cand]]).astype(np.float64)

    def set_state(self, e, s):
        s = np.asarray(s, np.float64)
        self.rec[e, :88] = s[:88]
        self.rec[e, 88:88 + (len(s) - 88)] = s[88:]
