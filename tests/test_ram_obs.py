"""CPU: the structured ("ram" / GoBigger-style) observation of oracle/oracle.c against the reference's own
GoBiggerObservation::add_frame (environment/envs/GoBiggerEnvironment.hpp:515-548) run on the reference
engine's state, step by step, element by element (libm trigonometry on both sides)."""
import numpy as np
import pytest

from _helpers import Oracle, Reference, oracle_layout, oracle_lib, random_actions, ref_lib
from agarcl_b200._abi import (RAM_KC, RAM_KP, RAM_KS, RAM_KV, RAM_OFF_CLONE, RAM_OFF_FOOD, RAM_OFF_SPORE, RAM_OFF_VIRUS,
                              RAM_RECORD, make_cfg)

pytestmark = pytest.mark.skipif(ref_lib() is None, reason="compiled reference not available")


def same(a, b):
    a = a.copy()
    b = b.copy()
    a[:, 5:7] = 0  # player x / y: an extra of the record, the reference's PlayerState does not carry it
    b[:, 5:7] = 0
    return np.array_equal(a.view(np.uint32), b.view(np.uint32)) or np.array_equal(a, b)


CASES = {
    "default_bots": (dict(), dict(steps=60)),
    "split_eject": (dict(num_agents=3, num_bots=6, arena_size=400, num_pellets=400, num_viruses=8, cap_foods=2048),
                    dict(steps=120, p_feed=0.4, p_split=0.3, boost=600)),
    "no_respawn_mode4": (dict(num_agents=2, num_bots=0, arena_size=150, num_pellets=100, num_viruses=3, mode_number=4, cap_foods=1024),
                         dict(steps=120, p_feed=0.2, p_split=0.2, boost=300)),
    "dense_overflow": (dict(num_agents=1, num_bots=2, arena_size=200, num_pellets=1500, num_viruses=30, cap_viruses=128), dict(steps=30, boost=160)),
    "grid64": (dict(num_bots=4, grid_size=64, arena_size=300, num_pellets=300, num_viruses=5), dict(steps=50)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_ram_records_match_reference(name):
    ck, rk = CASES[name]
    rk = dict(rk)
    steps, boost = rk.pop("steps"), rk.pop("boost", None)
    oracle_lib().oracle_set_trig_mode(0)
    cfg = make_cfg(**ck)
    L = oracle_layout(cfg)
    seed = 77 + len(name)
    ref = Reference(cfg, L)
    ref.seed(seed)
    ora = Oracle(cfg, L)
    ora.seed_mt(seed, 1 << 16)
    ref.reset()
    ora.reset()
    ref.ram_clear()
    ora.ram_clear()
    if boost:
        for a in range(L.A):
            ref.set_cell_mass(a, 0, boost)
            ora.state.cells[a][0]["mass"] = boost
    rng = np.random.default_rng(seed)
    seen = np.zeros(4)
    for st in range(steps):
        dxdy, act = random_actions(rng, L.A, rk.get("p_feed", 1 / 3), rk.get("p_split", 1 / 3))
        ref.set_actions(dxdy, act)
        ora.set_actions(dxdy, act)
        ref.step()
        ora.step()
        r, o = ref.ram_obs(), ora.ram_obs()
        if not same(r, o):
            bad = np.argwhere(r != o)[:8].tolist()
            raise AssertionError(f"step {st}: records differ at (player, slot) {bad}: {[(float(r[p, k]), float(o[p, k])) for p, k in bad]}")
        seen = np.maximum(seen, o[:, :4].max(axis=0))
    assert seen[0] > 0 and seen[3] > 0, "nothing was ever in view"
    if name == "split_eject":
        assert seen[2] > 0 and seen[3] > 1, "no spores / split cells observed"
    if name == "dense_overflow":
        assert seen[0] > RAM_KP, "the pellet capacity was never exceeded"


def test_record_layout_constants():
    assert RAM_OFF_FOOD == 8 and RAM_OFF_VIRUS == 8 + 4 * RAM_KP and RAM_OFF_SPORE == RAM_OFF_VIRUS + 4 * RAM_KV
    assert RAM_OFF_CLONE == RAM_OFF_SPORE + 4 * RAM_KS and RAM_RECORD == RAM_OFF_CLONE + 8 * RAM_KC == 1224
