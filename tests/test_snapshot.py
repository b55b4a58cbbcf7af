"""CPU: the JSON environment snapshot (agarcl_snapshot_write / agarcl_snapshot_read, what
agarcl_batch_save_env_state / load_env_state run) interchanged with the REFERENCE's own
BaseEnvironment::save_env_state / load_env_state (BaseEnvironment.hpp:213-343, Engine.hpp:247-348)."""
import ctypes as C
import json

import numpy as np
import pytest

from _helpers import Oracle, Reference, oracle_layout, oracle_lib, random_actions, ref_lib
from agarcl_b200 import _lib
from agarcl_b200._abi import StateView, compare_states, make_cfg

CFG = dict(num_agents=2, num_bots=5, arena_size=300, num_pellets=200, num_viruses=6, cap_foods=1024)


def played_oracle(seed=9, steps=80):
    oracle_lib().oracle_set_trig_mode(0)
    cfg = make_cfg(**CFG)
    L = oracle_layout(cfg)
    ora = Oracle(cfg, L)
    ora.seed_mt(seed, 1 << 16)
    ora.reset()
    for a in range(L.A):
        ora.state.cells[a][0]["mass"] = 500
    rng = np.random.default_rng(seed)
    for _ in range(steps):
        dxdy, act = random_actions(rng, L.A, 0.3, 0.3)
        ora.set_actions(dxdy, act)
        ora.step()
    ora.state.hdr["seed_lo"] = seed
    return cfg, L, ora


def as_loaded_by_reference(sv):
    """what Engine::load_env_state keeps of a state: no splitting velocity, timers restarted, ticks 0, default action"""
    out = sv.copy()
    out.hdr["tick"] = 0
    out.hdr["done_sticky"] = 0
    out.viruses["hits"][:] = 0
    out.players["action"][:] = 0
    out.players["min_mass_cell"][:] = 25
    for f in ("svx", "svy"):
        out.cells[f][:] = 0
    out.cells["recomb_tick"][:] = 0
    return out


def test_write_read_round_trip_lossless(tmp_path):
    cfg, L, ora = played_oracle()
    lib = _lib.lib()
    path = str(tmp_path / "snap.json").encode()
    assert lib.agarcl_snapshot_write(C.byref(cfg), C.byref(L), ora.state.ptr, path) == 0
    doc = json.load(open(path))  # valid JSON with the reference's keys
    assert {"players", "pellets", "viruses", "foods", "mode_number", "seed", "arena_size", "num_agents", "pellet_count"} <= set(doc)
    assert {"pid", "name", "is_bot", "cells", "split_cooldown", "virus_eaten_ticks", "highest_mass"} <= set(doc["players"][0])
    back = StateView(L, ora.state.blob.copy())
    back.players["n_cells"][:] = 0  # prove that everything comes from the file
    back.hdr["n_pellets"] = 0
    assert lib.agarcl_snapshot_read(C.byref(cfg), C.byref(L), back.ptr, path, 1) == 0
    assert not compare_states(ora.state, back)
    assert int(back.hdr["rng_cursor"]) == int(ora.state.hdr["rng_cursor"]) and int(back.hdr["next_cell_id"]) == int(ora.state.hdr["next_cell_id"])
    lossy = StateView(L, ora.state.blob.copy())
    assert lib.agarcl_snapshot_read(C.byref(cfg), C.byref(L), lossy.ptr, path, 0) == 0
    assert not compare_states(as_loaded_by_reference(ora.state), lossy)


def test_bad_snapshots_are_rejected(tmp_path):
    cfg, L, ora = played_oracle(steps=3)
    lib = _lib.lib()
    path = tmp_path / "snap.json"
    assert lib.agarcl_snapshot_write(C.byref(cfg), C.byref(L), ora.state.ptr, str(path).encode()) == 0
    doc = json.load(open(path))
    sv = StateView(L, ora.state.blob.copy())
    for mutate in (lambda d: d.update(mode_number=3), lambda d: d["players"].pop(), lambda d: d["players"][0].update(name="HungryBot")):
        d = json.loads(json.dumps(doc))
        mutate(d)
        bad = tmp_path / "bad.json"
        bad.write_text(json.dumps(d))
        assert lib.agarcl_snapshot_read(C.byref(cfg), C.byref(L), sv.ptr, str(bad).encode(), 0) == -1
        assert lib.agarcl_last_error()
    assert lib.agarcl_snapshot_read(C.byref(cfg), C.byref(L), sv.ptr, str(tmp_path / "missing.json").encode(), 0) == -1
    (tmp_path / "garbage.json").write_text("{ not json")
    assert lib.agarcl_snapshot_read(C.byref(cfg), C.byref(L), sv.ptr, str(tmp_path / "garbage.json").encode(), 0) == -1


@pytest.mark.skipif(ref_lib() is None, reason="compiled reference not available")
def test_reference_loads_our_snapshot_and_continues_identically(tmp_path):
    cfg, L, ora = played_oracle()
    lib = _lib.lib()
    path = tmp_path / "ours.json"
    assert lib.agarcl_snapshot_write(C.byref(cfg), C.byref(L), ora.state.ptr, str(path).encode()) == 0
    ref = Reference(cfg, L)
    ref.reset()
    ref.load_env_state(path)
    rs, miss = ref.dump()
    expect = as_loaded_by_reference(ora.state)
    assert miss == 0 and not compare_states(rs, expect)
    # ... and both continue identically from there (the draw stream restarts from the snapshot's seed on both sides)
    cont = Oracle(cfg, L)
    cont.state.blob[:] = expect.blob
    cont.seed_mt(int(ora.state.hdr["seed_lo"]), 1 << 16)
    cont.state.hdr["rng_cursor"] = 0
    rng = np.random.default_rng(1)
    pids = ref.agent_pids()  # the reference re-derives agent -> pid from its player map on load: reversed here
    assert sorted(pids) == list(range(L.A))
    for st in range(40):
        dxdy, act = random_actions(rng, L.A, 0.3, 0.3)
        ref.set_actions(dxdy[pids], act[pids])
        cont.set_actions(dxdy, act)
        rr, _ = ref.step()
        orr, _, _ = cont.step()
        rs, _ = ref.dump()
        d = compare_states(rs, cont.state)
        assert not d and np.array_equal(rr, orr), f"step {st} after the reload: {d[:4]}"


@pytest.mark.skipif(ref_lib() is None, reason="compiled reference not available")
def test_we_load_the_reference_snapshot(tmp_path):
    oracle_lib().oracle_set_trig_mode(0)
    cfg = make_cfg(**CFG)
    L = oracle_layout(cfg)
    ref = Reference(cfg, L)
    ref.seed(21)
    ref.reset()
    for a in range(L.A):
        ref.set_cell_mass(a, 0, 500)
    rng = np.random.default_rng(2)
    for _ in range(60):
        dxdy, act = random_actions(rng, L.A, 0.3, 0.3)
        ref.set_actions(dxdy, act)
        ref.step()
    path = tmp_path / "ref.json"
    ref.save_env_state(path)
    rs, miss = ref.dump()
    sv = StateView(L)
    sv.players["bot_type"][:] = list(L.bot_type)[:L.P]
    assert _lib.lib().agarcl_snapshot_read(C.byref(cfg), C.byref(L), sv.ptr, str(path).encode(), 0) == 0, _lib.lib().agarcl_last_error()
    assert miss == 0 and not compare_states(as_loaded_by_reference(rs), sv)


def test_64_bit_seed_survives_the_round_trip(tmp_path):
    """the reference's "seed" key is 32 bits (Engine.hpp:346); the upper half of a 64-bit Philox key travels as "seed_hi",
    which the reference ignores; a file without a seed leaves the instance's key alone"""
    cfg, L, ora = played_oracle(steps=3)
    lib = _lib.lib()
    ora.state.hdr["seed_lo"], ora.state.hdr["seed_hi"] = 0x89abcdef, 0x01234567
    path = str(tmp_path / "snap.json").encode()
    assert lib.agarcl_snapshot_write(C.byref(cfg), C.byref(L), ora.state.ptr, path) == 0
    doc = json.load(open(path))
    assert doc["seed"] == 0x89abcdef and doc["seed_hi"] == 0x01234567
    back = StateView(L, ora.state.blob.copy())
    back.hdr["seed_lo"], back.hdr["seed_hi"] = 1, 2
    assert lib.agarcl_snapshot_read(C.byref(cfg), C.byref(L), back.ptr, path, 1) == 0
    assert int(back.hdr["seed_lo"]) == 0x89abcdef and int(back.hdr["seed_hi"]) == 0x01234567
    del doc["seed"], doc["seed_hi"]
    json.dump(doc, open(path, "w"))
    back.hdr["seed_lo"], back.hdr["seed_hi"] = 11, 22
    assert lib.agarcl_snapshot_read(C.byref(cfg), C.byref(L), back.ptr, path, 1) == 0
    assert int(back.hdr["seed_lo"]) == 11 and int(back.hdr["seed_hi"]) == 22
