"""Host-side mirror of the reference's `agarcl.GridEnvironment` (environment/bindings.cpp:99-135).

`GridEnvironment` keeps the reference's constructor arguments, method names, argument meaning and
error behaviour, and runs on a size-1 batch of the CUDA library; `BatchedGridEnvironment` is the same
interface over N lockstep instances with device tensors out.  Nothing here computes game logic.

Behavioural notes versus the shipped reference ("quirk decisions"):
  * Q11: with `strict_reference=False` (default) the observation is the frame after the last tick of the
    step (what the reference's comments intend); `strict_reference=True` reproduces the shipped
    frame-index arithmetic (all-zero observation for ticks_per_step != num_frames).
  * Q15: `step()` returns rewards in the reference's own order (iteration order of the player map over
    the non-bot players); `BatchedGridEnvironment` returns them indexed by agent.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._abi import OBS_I16, OBS_I32, RNG_MT19937, RNG_PHILOX, make_cfg
from .batch import Batch


class _Base:
    def __init__(self, n_instances, num_agents, ticks_per_step, arena_size, pellet_regen, num_pellets, num_viruses,
                 num_bots, reward_type, c_death, mode_number, rng_mode, strict_reference, device, obs_dtype, instance_base=0):
        self._ctor = dict(n_instances=n_instances, num_agents=num_agents, ticks_per_step=ticks_per_step,
                          arena_size=arena_size, pellet_regen=pellet_regen, num_pellets=num_pellets, num_viruses=num_viruses,
                          num_bots=num_bots, reward_type=reward_type, c_death=c_death, mode_number=mode_number,
                          rng_mode=rng_mode, strict_reference=strict_reference, device=device, obs_dtype=obs_dtype,
                          instance_base=instance_base)
        self._obs_cfg = dict(num_frames=1, grid_size=128, observe_cells=True, observe_others=True, observe_viruses=True,
                             observe_pellets=True)
        self._seeds = None
        self._batch = None
        # validate early, like the reference constructor does (Engine::set_mode throws on a bad mode)
        _lib.make_layout(make_cfg(**self._ctor, **self._obs_cfg))

    def _ensure(self):
        if self._batch is None:
            self._batch = Batch(make_cfg(**self._ctor, **self._obs_cfg))
            if self._seeds is not None:
                self._batch.seed(self._seeds)
            self._batch.reset()  # the reference constructor ends with reset() (BaseEnvironment.hpp:66)
        return self._batch

    def configure_observation(self, config):
        """bindings.cpp:104-114 — dict with num_frames, grid_size, observe_cells/others/viruses/pellets"""
        for k in self._obs_cfg:
            if k in config:
                self._obs_cfg[k] = type(self._obs_cfg[k])(config[k])
        if self._batch is not None:
            self._batch.close()
            self._batch = None

    def observation_shape(self):
        L = _lib.make_layout(make_cfg(**self._ctor, **self._obs_cfg))
        return (self._obs_cfg["num_frames"] * L.obs_channels, self._obs_cfg["grid_size"], self._obs_cfg["grid_size"])

    def close(self):
        if self._batch is not None:
            self._batch.close()
            self._batch = None

    def render(self):
        raise RuntimeError("OpenGL rendering is out of scope of agarcl_b200 (SURVEY.md section 2, rows 11-13)")

    get_frame = render


class GridEnvironment(_Base):
    """Drop-in for agarcl.GridEnvironment(num_agents, ticks_per_step, arena_size, pellet_regen, num_pellets,
    num_viruses, num_bots, reward_type, c_death, mode_number) — bindings.cpp:102."""

    def __init__(self, num_agents, ticks_per_step, arena_size, pellet_regen, num_pellets, num_viruses, num_bots,
                 reward_type=0, c_death=0, mode_number=0, *, rng_mode=RNG_MT19937, strict_reference=False, device=0):
        super().__init__(1, num_agents, ticks_per_step, arena_size, pellet_regen, num_pellets, num_viruses, num_bots,
                         reward_type, c_death, mode_number, rng_mode, strict_reference, device, OBS_I32)
        self.num_agents = num_agents
        self._order_agents = None

    def seed(self, s):
        self._seeds = np.array([s], dtype=np.uint64)
        if self._batch is not None:
            self._batch.seed(self._seeds)

    def reset(self):
        self._ensure().reset()

    def take_actions(self, actions):
        """list of (dx, dy, action) per agent; wrong length raises like EnvironmentException (BaseEnvironment.hpp:142-144)"""
        if len(actions) != self.num_agents:
            raise RuntimeError(f"Number of actions ({len(actions)}) does not match number of agents ({self.num_agents})")
        dxdy = np.array([[a[0], a[1]] for a in actions], dtype=np.float32)
        act = np.array([int(a[2]) for a in actions], dtype=np.int32)
        self._ensure().set_actions(dxdy, act)

    def step(self):
        b = self._ensure()
        rew = np.zeros(self.num_agents, np.float64)
        b.step()
        import torch
        rew = b.rewards_tensor().cpu().numpy()
        if self._order_agents is None:
            L = b.layout
            self._order_agents = [p for p in list(L.order)[:L.P] if p < L.A]
        return [float(rew[p]) for p in self._order_agents]  # quirk Q15: player-map order

    def dones(self):
        return [bool(x) for x in self._ensure().dones_tensor().cpu().numpy()]

    def get_state(self):
        """list of int32 arrays (C, G, G), one fresh copy per agent (bindings.cpp:67-91)"""
        obs = self._ensure().obs_tensor().cpu().numpy()
        return [obs[a].copy() for a in range(self.num_agents)]

    def save_env_state(self, path):
        raise RuntimeError("JSON snapshots are a later row of the scope table (SURVEY.md 8f rank 2)")


class BatchedGridEnvironment(_Base):
    """N lockstep GridEnvironments on one GPU.  step() takes arrays/tensors and returns device tensors:
    obs [N*A, C*frames, G, G], rewards f64 [N*A], dones u8 [N*A]."""

    def __init__(self, n_instances, num_agents=1, ticks_per_step=4, arena_size=1000, pellet_regen=True, num_pellets=1000,
                 num_viruses=25, num_bots=25, reward_type=1, c_death=0, mode_number=0, *, rng_mode=RNG_PHILOX,
                 strict_reference=False, device=0, obs_dtype=OBS_I32, instance_base=0):
        super().__init__(n_instances, num_agents, ticks_per_step, arena_size, pellet_regen, num_pellets, num_viruses,
                         num_bots, reward_type, c_death, mode_number, rng_mode, strict_reference, device, obs_dtype,
                         instance_base)
        self.n_instances = n_instances
        self.num_agents = num_agents

    def seed(self, seeds):
        s = np.asarray(seeds, dtype=np.uint64)
        self._seeds = s if s.ndim else (np.arange(self.n_instances, dtype=np.uint64) + s)
        if self._batch is not None:
            self._batch.seed(self._seeds)

    def reset(self, mask=None):
        self._ensure().reset(mask)
        return self._batch.obs_tensor()

    def step(self, dxdy, act, stream=0):
        """dxdy: [N*A, 2] float32, act: [N*A] int32 — numpy (copied) or CUDA torch tensors (zero-copy)"""
        b = self._ensure()
        if hasattr(dxdy, "data_ptr"):
            assert dxdy.is_cuda and act.is_cuda and dxdy.is_contiguous() and act.is_contiguous()
            assert dxdy.numel() == self.n_instances * self.num_agents * 2, "Number of actions does not match number of agents"
            b.set_actions_device(dxdy.data_ptr(), act.data_ptr(), stream)
        else:
            b.set_actions(dxdy, act, stream)
        b.step(stream)
        return b.obs_tensor(), b.rewards_tensor(), b.dones_tensor()

    @property
    def batch(self):
        return self._ensure()
