"""CUDA path vs oracle port, step by step, on identical initial state and identical draws.

Everything discrete (counts, masses, timers, cooldowns, statistics, rewards, dones, flags) and
every fp32 field must be IDENTICAL: the oracle runs with trig_mode 1 (the same portable
trigonometry as the device), all other arithmetic is IEEE-exact on both sides.
"""
import numpy as np

from _helpers import Oracle, oracle_lib, oracle_layout, random_actions
from agarcl_b200 import RNG_REPLAY, make_cfg
from agarcl_b200._abi import compare_states
from agarcl_b200.batch import Batch


def philox_uniform_np(seed, instance, k):
    """numpy restatement of device_math.cuh::philox_uniform for draw indices k (test-side only)."""
    k = np.asarray(k, dtype=np.uint64)
    c = [(k >> np.uint64(2)).astype(np.uint64) & np.uint64(0xffffffff), np.zeros_like(k), np.full_like(k, instance), np.zeros_like(k)]
    key = [np.uint64(seed & 0xffffffff), np.uint64((seed >> 32) & 0xffffffff)]
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xffffffff)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ key[0]) & mask, lo1, (hi0 ^ c[3] ^ key[1]) & mask, lo0]
        key = [(key[0] + np.uint64(0x9E3779B9)) & mask, (key[1] + np.uint64(0xBB67AE85)) & mask]
    w = np.choose((k & np.uint64(3)).astype(np.int64), c)
    return ((w >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def run_parity(cfg_kwargs, seeds, steps, p_feed=1 / 3, p_split=1 / 3, boost=None, obs_every=5, replay_len=1 << 16,
               with_obs_in_step=True, ram=False):
    """Runs len(seeds) instances on the GPU in one batch and each one through the oracle."""
    oracle_lib().oracle_set_trig_mode(1)
    n = len(seeds)
    cfg = make_cfg(n_instances=n, rng_mode=RNG_REPLAY, cap_replay=replay_len, ram_obs=ram, **cfg_kwargs)
    b = Batch(cfg)
    L = b.layout
    Lo = oracle_layout(cfg)
    assert bytes(L) == bytes(Lo), "product and oracle layouts differ"
    oras = []
    for i, s in enumerate(seeds):
        o = Oracle(cfg, L)
        o.seed_mt(s, replay_len)
        b.set_replay(i, o.replay)
        oras.append(o)
    b.seed(np.asarray(seeds, dtype=np.uint64))
    b.reset()
    for o in oras:
        o.reset()
    if boost:
        for i, o in enumerate(oras):
            for a in range(L.A):
                o.state.cells[a][0]["mass"] = boost
            sv = b.download_state(i)
            for a in range(L.A):
                sv.cells[a][0]["mass"] = boost
            b.upload_state(i, sv)
    for i, o in enumerate(oras):
        d = compare_states(o.state, b.download_state(i))
        assert not d, f"reset mismatch inst {i}: {d[:5]}"
    import torch
    obs_t = b.obs_tensor()
    rew_t = b.rewards_tensor()
    done_t = b.dones_tensor()
    ram_t = b.ram_tensor() if ram else None
    if ram:
        for o in oras:
            o.ram_clear()
    rngs = [np.random.default_rng(s) for s in seeds]
    A = L.A
    for st in range(steps):
        dxdy = np.zeros((n, A, 2), np.float32)
        act = np.zeros((n, A), np.int32)
        for i in range(n):
            dxdy[i], act[i] = random_actions(rngs[i], A, p_feed, p_split)
        b.set_actions(dxdy, act)
        b.step()
        torch.cuda.synchronize()
        g_rew = rew_t.cpu().numpy().reshape(n, A)
        g_done = done_t.cpu().numpy().reshape(n, A)
        want_obs = (st % obs_every == 0)
        g_obs = obs_t.cpu().numpy().reshape(n, A, *b.obs_shape[1:]) if want_obs else None
        g_ram = ram_t.cpu().numpy() if ram else None
        for i, o in enumerate(oras):
            o.set_actions(dxdy[i], act[i])
            o_rew, o_done, o_obs = o.step_with_ram() if ram else o.step(with_obs=want_obs and with_obs_in_step)
            gs = b.download_state(i)
            d = compare_states(o.state, gs)
            assert not d, f"step {st} inst {i} (seed {seeds[i]}): {d[:6]} flags gpu={gs.flag_names()} oracle={o.state.flag_names()}"
            assert int(gs.hdr["flags"]) == int(o.state.hdr["flags"]), (st, i, gs.flag_names(), o.state.flag_names())
            assert int(gs.hdr["rng_cursor"]) == int(o.state.hdr["rng_cursor"]), (st, i)
            assert np.array_equal(g_rew[i], o_rew), f"step {st} inst {i}: rewards {g_rew[i]} vs {o_rew}"
            assert np.array_equal(g_done[i], o_done), f"step {st} inst {i}: dones {g_done[i]} vs {o_done}"
            if ram:
                same = (g_ram[i].view(np.uint32) == o.ram.view(np.uint32)) | (np.isnan(g_ram[i]) & np.isnan(o.ram))
                if not same.all():
                    bad = np.argwhere(~same)[:8].tolist()
                    raise AssertionError(f"step {st} inst {i}: ram records differ at (player, slot) {bad}: "
                                         f"{[(float(g_ram[i][p, k]), float(o.ram[p, k])) for p, k in bad]}")
            if want_obs:
                if o_obs is None:
                    o_obs = np.stack([o.obs(a) for a in range(A)])
                if not np.array_equal(g_obs[i], o_obs):
                    bad = [(a, c, int((g_obs[i][a][c] != o_obs[a][c]).sum())) for a in range(A) for c in range(o_obs.shape[1])
                           if not np.array_equal(g_obs[i][a][c], o_obs[a][c])]
                    raise AssertionError(f"step {st} inst {i}: obs differs (agent, channel, #cells): {bad[:8]}")
    stats = dict(cells_eaten=sum(int(o.state.players["cells_eaten"].sum()) for o in oras),
                 viruses_eaten=sum(int(o.state.players["viruses_eaten"].sum()) for o in oras),
                 food_eaten=sum(int(o.state.players["food_eaten"].sum()) for o in oras),
                 max_cells=max(int(o.state.players["n_cells"].max()) for o in oras),
                 foods=sum(int(o.state.hdr["n_foods"]) for o in oras),
                 flags=sorted({f for o in oras for f in o.state.flag_names()}))
    b.close()
    return stats
