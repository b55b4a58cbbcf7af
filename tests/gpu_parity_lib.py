"""CUDA path vs oracle port, step by step, on identical initial state and identical draws.

Everything discrete (counts, masses, timers, cooldowns, statistics, rewards, dones, flags) and
every fp32 field must be IDENTICAL: the oracle runs with trig_mode 1 (the same restatement of
glibc's atanf / sinf / cosf as the device, itself equal to libm: tests/test_trig_exact.py), all
other arithmetic is IEEE-exact on both sides.
"""
import numpy as np

from _helpers import Oracle, oracle_lib, oracle_layout, random_actions
from agarcl_b200 import RNG_REPLAY, make_cfg
from agarcl_b200._abi import compare_states
from agarcl_b200.batch import Batch


def philox4x32_10_np(counter, key):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123) on arrays of
    counters: counter = 4 uint32 words (arrays or scalars), key = 2 words.  Pinned against Random123's known-answer vectors
    in tests/test_philox_kat.py; the device's philox4x32_10 (device_math.cuh) is then pinned against THIS by the RNG_PHILOX
    parity runs."""
    c = [np.asarray(x, dtype=np.uint64) for x in counter]
    key = [np.uint64(key[0]), np.uint64(key[1])]
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xffffffff)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ key[0]) & mask, lo1, (hi0 ^ c[3] ^ key[1]) & mask, lo0]
        key = [(key[0] + np.uint64(0x9E3779B9)) & mask, (key[1] + np.uint64(0xBB67AE85)) & mask]
    return c


def philox_uniform_np(seed, instance, k):
    """numpy restatement of device_math.cuh::philox_uniform for draw indices k (test-side only): draw k of global instance
    g under 64-bit seed s = word (k & 3) of philox(counter = (k >> 2, 0, g, 0), key = (s_lo, s_hi)), top 24 bits -> [0, 1)."""
    k = np.asarray(k, dtype=np.uint64)
    c = philox4x32_10_np([(k >> np.uint64(2)) & np.uint64(0xffffffff), np.zeros_like(k), np.full_like(k, instance), np.zeros_like(k)],
                         [int(seed) & 0xffffffff, (int(seed) >> 32) & 0xffffffff])
    w = np.choose((k & np.uint64(3)).astype(np.int64), c)
    return ((w >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def run_parity(cfg_kwargs, seeds, steps, p_feed=1 / 3, p_split=1 / 3, boost=None, obs_every=5, replay_len=1 << 16,
               with_obs_in_step=True, ram=False, philox=False, state_every=1, instance_base=0):
    """Runs len(seeds) instances on the GPU in one batch and each one through the oracle.
    philox: the GPU draws from its own counter-based Philox stream (AGARCL_RNG_PHILOX, what bench.py runs) and the oracle is
    fed the same stream computed by philox_uniform_np; otherwise both replay the mt19937_64 stream of the seed.
    state_every: the full state is compared every that many steps (rewards and dones every step)."""
    oracle_lib().oracle_set_trig_mode(1)
    n = len(seeds)
    if philox:
        from agarcl_b200 import RNG_PHILOX
        cfg = make_cfg(n_instances=n, rng_mode=RNG_PHILOX, ram_obs=ram, instance_base=instance_base, **cfg_kwargs)
        ocfg = make_cfg(n_instances=n, rng_mode=RNG_REPLAY, cap_replay=replay_len, ram_obs=ram, instance_base=instance_base, **cfg_kwargs)
    else:
        cfg = ocfg = make_cfg(n_instances=n, rng_mode=RNG_REPLAY, cap_replay=replay_len, ram_obs=ram, **cfg_kwargs)
    b = Batch(cfg)
    L = b.layout
    Lo = oracle_layout(ocfg)
    if not philox:
        assert bytes(L) == bytes(Lo), "product and oracle layouts differ"
    oras = []
    for i, s in enumerate(seeds):
        o = Oracle(ocfg, Lo)
        if philox:
            o.set_replay(philox_uniform_np(s, instance_base + i, np.arange(replay_len)))
        else:
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
            o_rew, o_done, o_obs = (o.step_with_ram(with_obs=want_obs and with_obs_in_step) if ram
                                    else o.step(with_obs=want_obs and with_obs_in_step))
            if st % state_every == 0 or st == steps - 1:
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
