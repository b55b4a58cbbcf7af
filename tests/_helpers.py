"""Shared test plumbing: loads the checkers (oracle port, compiled reference) and the product
library through ctypes and offers small wrappers.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from agarcl_b200._abi import (Cfg, Layout, RAM_RECORD, StateView, compare_states, make_cfg)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libagarcl_ref.so")
PRODUCT_SO = os.path.join(ROOT, "agarcl_b200", "libagarcl_b200.so")

_vp = C.c_void_p


def fptr(a):
    return a.ctypes.data_as(_vp)


def _build():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "oracle.c")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/agario"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        _build()
        _oracle = C.CDLL(ORACLE_SO)
        _oracle.oracle_make_layout.restype = C.c_int
    return _oracle


def ref_lib():
    """The compiled reference engine, or None when it has not been built (no /root/reference)."""
    global _ref
    if _ref is None:
        _build()
        if not os.path.exists(REF_SO):
            return None
        _ref = C.CDLL(REF_SO)
        _ref.ref_create.restype = _vp
        _ref.ref_bench.restype = C.c_double
    return _ref


def oracle_layout(cfg):
    L = Layout()
    rc = oracle_lib().oracle_make_layout(C.byref(cfg), C.byref(L))
    assert rc == 0, "oracle_make_layout failed"
    return L


class Oracle:
    """One instance stepped by oracle/oracle.c."""

    def __init__(self, cfg, layout=None, replay=None):
        self.lib = oracle_lib()
        self.cfg = cfg
        self.L = layout or oracle_layout(cfg)
        self.state = StateView(self.L)
        self.replay = np.ascontiguousarray(replay if replay is not None else np.zeros(0, np.float32), dtype=np.float32)

    def set_replay(self, replay):
        self.replay = np.ascontiguousarray(replay, dtype=np.float32)

    def seed_mt(self, seed, n=1 << 16):
        d = np.zeros(n, np.float32)
        self.lib.oracle_mt19937_draws(C.c_uint64(seed), fptr(d), n)
        self.replay = d

    def reset(self):
        self.lib.oracle_reset(C.byref(self.cfg), C.byref(self.L), self.state.ptr, fptr(self.replay), self.replay.size)

    def reset_keep_stream(self):
        """a later episode: the draw stream goes on (Engine::reset does not reseed)"""
        self.lib.oracle_reset_keep_stream(C.byref(self.cfg), C.byref(self.L), self.state.ptr, fptr(self.replay), self.replay.size)

    def set_actions(self, dxdy, act):
        dxdy = np.ascontiguousarray(dxdy, np.float32)
        act = np.ascontiguousarray(act, np.int32)
        self.lib.oracle_set_actions(C.byref(self.cfg), C.byref(self.L), self.state.ptr, fptr(dxdy), fptr(act))

    def step(self, with_obs=False):
        A = self.L.A
        rew = np.zeros(A, np.float64)
        dones = np.zeros(A, np.uint8)
        obs = None
        if with_obs:
            obs = np.zeros((A, self.cfg.num_frames * self.L.obs_channels, self.cfg.grid_size, self.cfg.grid_size), np.int32)
        self.lib.oracle_step(C.byref(self.cfg), C.byref(self.L), self.state.ptr, fptr(self.replay), self.replay.size,
                             fptr(rew), fptr(dones), fptr(obs) if with_obs else None)
        return rew, dones, obs

    def obs(self, agent):
        out = np.zeros((self.L.obs_channels, self.cfg.grid_size, self.cfg.grid_size), np.int32)
        self.lib.oracle_obs(C.byref(self.cfg), C.byref(self.L), self.state.ptr, agent, 0, fptr(out))
        return out

    # ---- structured ("ram") observation: records persist between calls (players with nothing in view keep theirs)
    def ram_clear(self):
        self.ram = np.zeros((self.L.P, RAM_RECORD), np.float32)

    def ram_obs(self):
        """oracle_ram_obs on the CURRENT state"""
        if not hasattr(self, "ram"):
            self.ram_clear()
        self.lib.oracle_ram_obs(C.byref(self.cfg), C.byref(self.L), self.state.ptr, fptr(self.ram))
        return self.ram

    def step_with_ram(self, with_obs=False):
        """oracle_step with the records taken where the reference takes them (after the ticks, before respawn)"""
        if not hasattr(self, "ram"):
            self.ram_clear()
        self.lib.oracle_set_ram_out(fptr(self.ram))
        try:
            return self.step(with_obs=with_obs)
        finally:
            self.lib.oracle_set_ram_out(None)


class Reference:
    """One instance of the compiled reference GridEnvironment<int,false>."""

    def __init__(self, cfg, layout):
        self.lib = ref_lib()
        assert self.lib is not None
        self.cfg, self.L = cfg, layout
        self.h = _vp(self.lib.ref_create(C.byref(cfg)))

    def __del__(self):
        try:
            self.lib.ref_destroy(self.h)
        except Exception:
            pass

    def seed(self, s):
        self.lib.ref_seed(self.h, C.c_uint(s))

    def reset(self):
        self.lib.ref_reset(self.h)

    def reset_native(self):
        """BaseEnvironment::reset as the reference runs it (pids keep counting, quirk Q3)"""
        self.lib.ref_reset_native(self.h)

    def pid_base(self):
        return int(self.lib.ref_pid_base(self.h))

    def order(self):
        o = (C.c_int32 * 64)()
        n = self.lib.ref_player_order(self.h, o)
        return list(o)[:n]

    def peek_draws(self, n):
        d = np.zeros(n, np.float32)
        self.lib.ref_rng_peek(self.h, fptr(d), n)
        return d

    def dump(self):
        sv = StateView(self.L)
        miss = self.lib.ref_dump_state(self.h, C.byref(self.L), sv.ptr)
        return sv, miss

    def set_actions(self, dxdy, act):
        dxdy = np.ascontiguousarray(dxdy, np.float32)
        act = np.ascontiguousarray(act, np.int32)
        self.lib.ref_take_actions(self.h, fptr(dxdy), fptr(act))

    def step(self):
        """rewards re-indexed by agent (the reference returns them in player-map order, quirk Q15)"""
        A = self.L.A
        raw = np.zeros(A, np.float64)
        self.lib.ref_step(self.h, fptr(raw))
        agents_in_map_order = [p for p in self.order() if p < A]
        rew = np.zeros(A, np.float64)
        for j, p in enumerate(agents_in_map_order):
            rew[p] = raw[j]
        d = np.zeros(A, np.uint8)
        self.lib.ref_dones(self.h, fptr(d))
        return rew, d

    def obs(self, agent):
        out = np.zeros((self.L.obs_channels, self.cfg.grid_size, self.cfg.grid_size), np.int32)
        self.lib.ref_obs(self.h, agent, fptr(out))
        return out

    def ram_clear(self):
        self.ram = np.zeros((self.L.P, RAM_RECORD), np.float32)
        self.lib.ref_ram_clear(self.h)

    def ram_obs(self):
        """GoBiggerObservation::add_frame on the current reference state, flattened into agarcl records"""
        if not hasattr(self, "ram"):
            self.ram_clear()
        self.lib.ref_ram_obs(self.h, self.L.P, fptr(self.ram))
        return self.ram

    def agent_pids(self):
        o = (C.c_int32 * 64)()
        n = self.lib.ref_agent_pids(self.h, o)
        return list(o)[:n]

    def save_env_state(self, path):
        assert self.lib.ref_save_env_state(self.h, str(path).encode()) == 0

    def load_env_state(self, path):
        assert self.lib.ref_load_env_state(self.h, str(path).encode()) == 0

    def set_cell_mass(self, pid, cell, mass):
        self.lib.ref_set_cell_mass(self.h, pid, cell, C.c_uint(mass))

    def set_cell_pos(self, pid, cell, x, y):
        self.lib.ref_set_cell_pos(self.h, pid, cell, C.c_float(x), C.c_float(y))

    def set_virus(self, idx, x, y):
        self.lib.ref_set_virus(self.h, idx, C.c_float(x), C.c_float(y))


def random_actions(rng, A, p_feed=1 / 3, p_split=1 / 3):
    dxdy = rng.uniform(-1, 1, size=(A, 2)).astype(np.float32)
    u = rng.random(A)
    act = np.where(u < p_feed, 1, np.where(u < p_feed + p_split, 2, 0)).astype(np.int32)
    return dxdy, act
