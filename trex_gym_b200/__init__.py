"""trex_gym_b200: B200-native batched simulator for the trex-gym hot path.

Public surface (mirrors /root/reference/trex_gym): ``TrexBulletEnv`` (alias ``TrexEnv``),
``TrexRobot``, plus the batched ``TrexVecEnv`` / ``TrexBatchSim``.  Importing the package does not
need a GPU; constructing a simulator does (there is no CPU fallback).
"""
__all__ = ["TrexBulletEnv", "TrexEnv", "TrexVecEnv", "TrexRobot", "TrexBatchSim"]


def __getattr__(name):
    if name in ("TrexBulletEnv", "TrexEnv", "TrexVecEnv"):
        from . import trex_env

        return getattr(trex_env, name)
    if name == "TrexRobot":
        from .trex_robot import TrexRobot

        return TrexRobot
    if name == "TrexBatchSim":
        from .sim import TrexBatchSim

        return TrexBatchSim
    raise AttributeError(name)
