"""agarcl_b200 — B200-native batched AgarCL simulator (hot path of machado-research/AgarCL).

Public surface:
  Batch                      owner of a C-ABI batch handle (N lockstep instances on one GPU)
  GridEnvironment            drop-in for agarcl.GridEnvironment (environment/bindings.cpp:99-135), N = 1
  BatchedGridEnvironment     the same interface over N instances, device tensors out
  make_cfg / Cfg / Layout    configuration records of include/agarcl_b200.h
"""
from ._abi import Cfg, Layout, StateView, make_cfg, OBS_I16, OBS_I32, RNG_MT19937, RNG_PHILOX, RNG_REPLAY  # noqa: F401


def __getattr__(name):
    if name == "Batch":
        from .batch import Batch
        return Batch
    if name in ("GridEnvironment", "BatchedGridEnvironment"):
        from . import env
        return getattr(env, name)
    raise AttributeError(name)
