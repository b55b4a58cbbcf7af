"""-m gpu: the reference-shaped Python surface over the C ABI (GridEnvironment / BatchedGridEnvironment),
size-independent properties at the full BASELINE size, the int16 observation dtype, and strict_reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_grid_environment_drop_in_against_oracle():
    """agarcl.GridEnvironment semantics: seed -> reset -> take_actions -> step -> get_state/dones, with the
    mt19937_64 spawn stream the reference would consume for the same seed."""
    from _helpers import Oracle, oracle_lib
    from agarcl_b200 import make_cfg, RNG_REPLAY
    from agarcl_b200.env import GridEnvironment
    oracle_lib().oracle_set_trig_mode(1)
    env = GridEnvironment(2, 4, 500, True, 300, 5, 6, 1, 0, 0)
    env.seed(77)
    env.reset()
    cfg = make_cfg(num_agents=2, ticks_per_step=4, arena_size=500, num_pellets=300, num_viruses=5, num_bots=6, rng_mode=RNG_REPLAY,
                   cap_replay=16384)
    ora = Oracle(cfg)
    ora.seed_mt(77, 16384)
    ora.reset()
    L = ora.L
    agents_in_map_order = [p for p in list(L.order)[:L.P] if p < L.A]
    rng = np.random.default_rng(0)
    for st in range(40):
        acts = [(float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), int(rng.integers(0, 3))) for _ in range(2)]
        env.take_actions(acts)
        rew = env.step()
        ora.set_actions(np.array([[a[0], a[1]] for a in acts], np.float32), np.array([a[2] for a in acts], np.int32))
        orew, odone, oobs = ora.step(with_obs=True)
        assert rew == [float(orew[p]) for p in agents_in_map_order]  # quirk Q15 ordering
        assert env.dones() == [bool(x) for x in odone]
        state = env.get_state()
        assert len(state) == 2 and state[0].shape == (8, 128, 128) and state[0].dtype == np.int32
        for a in range(2):
            assert np.array_equal(state[a], oobs[a])
    env.close()


def test_full_size_batch_properties():
    """BASELINE.json configs[1] at full size (4096 instances): conservation-style invariants that do not need
    the oracle — counts within capacity, masses >= 25, cells inside the arena, observation consistent with the
    state (channel sums), rewards == mass deltas, no overflow flags, determinism across two identical runs."""
    import torch
    from agarcl_b200.env import BatchedGridEnvironment
    N = 4096
    finals = []
    for rep in range(2):
        env = BatchedGridEnvironment(N)
        env.seed(1000)
        env.reset()
        b = env.batch
        g = torch.Generator(device="cuda")
        g.manual_seed(5)
        mass_before = None
        for st in range(30):
            dxdy = (torch.rand((N, 2), device="cuda", generator=g) * 2 - 1).float()
            act = torch.randint(0, 3, (N,), device="cuda", generator=g, dtype=torch.int32)
            obs, rew, done = env.step(dxdy, act)
        torch.cuda.synchronize()
        assert obs.shape == (N, 8, 128, 128) and obs.dtype == torch.int32
        # observation structure: ch0 in {0,-1}; pellet presence <= count; own-mass channel sums to the agent mass when in view
        assert set(torch.unique(obs[:, 0]).tolist()) <= {0, -1}
        assert bool((obs[:, 1] <= obs[:, 2]).all()) and bool((obs[:, 1] >= 0).all())
        assert bool((obs[:, 6] <= obs[:, 7]).all())
        for i in (0, 1, N // 2, N - 1):
            sv = b.download_state(i)
            assert int(sv.hdr["flags"]) == 0, sv.flag_names()
            assert 0 < int(sv.hdr["n_pellets"]) <= 1000 and int(sv.hdr["tick"]) == 120
            for p in range(26):
                n = int(sv.players["n_cells"][p])
                assert 1 <= n <= 32  # mode 0 respawns every dead player at the end of the step
                c = sv.cells[p][:n]
                assert (c["mass"] >= 25).all()
                assert (c["x"] >= 0).all() and (c["x"] <= 1000).all() and (c["y"] >= 0).all() and (c["y"] <= 1000).all()
            own_mass = int(sv.cells[0][:int(sv.players["n_cells"][0])]["mass"].sum())
            assert int(obs[i, 5].sum()) == own_mass  # every own cell is inside its own view
        finals.append((obs.sum(dtype=torch.int64).item(), rew.sum().item(), b.download_state(N - 1).blob.copy()))
        env.close()
    assert finals[0][0] == finals[1][0] and finals[0][1] == finals[1][1]
    assert np.array_equal(finals[0][2], finals[1][2]), "two identical runs diverged (non-determinism)"


def test_results_independent_of_sharding():
    """An instance keyed by its GLOBAL index evolves identically whether it is instance 5 of one batch or
    instance 1 of a shard starting at 4 (what the N-GPU sharding relies on)."""
    import torch
    from agarcl_b200.env import BatchedGridEnvironment
    a = BatchedGridEnvironment(8, num_bots=5, arena_size=300, num_pellets=200, num_viruses=3)
    b = BatchedGridEnvironment(4, num_bots=5, arena_size=300, num_pellets=200, num_viruses=3, instance_base=4)
    a.seed(50)
    b.seed(np.arange(4, 8, dtype=np.uint64) + np.uint64(50))
    a.reset()
    b.reset()
    rng = np.random.default_rng(1)
    for st in range(25):
        dxdy = rng.uniform(-1, 1, size=(8, 2)).astype(np.float32)
        act = rng.integers(0, 3, size=8).astype(np.int32)
        a.step(dxdy, act)
        b.step(dxdy[4:], act[4:])
    torch.cuda.synchronize()
    sa, sb = a.batch.download_state(5), b.batch.download_state(1)
    from agarcl_b200._abi import compare_states
    assert not compare_states(sa, sb)
    assert np.array_equal(a.batch.obs_tensor()[5].cpu().numpy(), b.batch.obs_tensor()[1].cpu().numpy())


def test_results_independent_of_schedule(monkeypatch):
    """More instances than one round of the persistent grid holds (148 SMs x 16 warps): the aligned, cost-sorted, pooled
    schedule (default) and free-running warps on a ticket counter produce identical states, observations, rewards and
    dones for every instance -- scheduling only moves instances between warps and rounds."""
    import torch
    from agarcl_b200.env import BatchedGridEnvironment
    from agarcl_b200._abi import compare_states
    N = 2600
    kw = dict(num_bots=12, arena_size=400, num_pellets=300, num_viruses=12)
    outs = []
    for env in ({}, {"AGARCL_TICK_BARRIER": "0"}):
        for k in ("AGARCL_TICK_BARRIER",):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        e = BatchedGridEnvironment(N, **kw)
        e.configure_observation({"grid_size": 32})
        e.seed(123)
        e.reset()
        rng = np.random.default_rng(5)
        rew_sum = np.zeros(N, np.float64)
        for st in range(60):
            dxdy = rng.uniform(-1, 1, size=(N, 2)).astype(np.float32)
            act = rng.integers(0, 3, size=N).astype(np.int32)
            obs, rew, done = e.step(dxdy, act)
            rew_sum += rew.cpu().numpy()
        torch.cuda.synchronize()
        outs.append((obs.cpu().numpy().copy(), rew_sum, done.cpu().numpy().copy(), [e.batch.download_state(i) for i in (0, 777, 1500, 2368, 2599)]))
        e.close()
    a, b = outs
    assert np.array_equal(a[0], b[0]), "observations differ between schedules"
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for sa, sb in zip(a[3], b[3]):
        assert not compare_states(sa, sb)


def test_int16_observation_matches_int32():
    import torch
    from agarcl_b200 import OBS_I16
    from agarcl_b200.env import BatchedGridEnvironment
    e32 = BatchedGridEnvironment(16, num_agents=2, num_bots=8, arena_size=300, num_pellets=300, num_viruses=5)
    e16 = BatchedGridEnvironment(16, num_agents=2, num_bots=8, arena_size=300, num_pellets=300, num_viruses=5, obs_dtype=OBS_I16)
    for e in (e32, e16):
        e.seed(9)
        e.reset()
    rng = np.random.default_rng(2)
    for st in range(60):
        dxdy = rng.uniform(-1, 1, size=(32, 2)).astype(np.float32)
        act = rng.integers(0, 3, size=32).astype(np.int32)
        o32, _, _ = e32.step(dxdy, act)
        o16, _, _ = e16.step(dxdy, act)
        assert o16.dtype == torch.int16
        assert torch.equal(o32.clamp(-32768, 32767).to(torch.int16), o16), st
    torch.cuda.synchronize()
    # the int16 frame comes out of the fused step kernel like the int32 one (no separate observation kernel)
    assert e16.batch.launches_per_step() == e32.batch.launches_per_step() == 2  # k_step + k_order


def test_strict_reference_frame_quirk_q11():
    """With the shipped frame-index arithmetic (tps 4, num_frames 1) the reference's observation is all zeros."""
    import torch
    from agarcl_b200.env import BatchedGridEnvironment
    e = BatchedGridEnvironment(4, num_bots=3, arena_size=200, num_pellets=100, num_viruses=2, strict_reference=True)
    e.seed(1)
    e.reset()
    obs, _, _ = e.step(np.zeros((4, 2), np.float32), np.zeros(4, np.int32))
    torch.cuda.synchronize()
    assert int(obs.abs().sum()) == 0
    e2 = BatchedGridEnvironment(4, num_bots=3, arena_size=200, num_pellets=100, num_viruses=2, ticks_per_step=1, strict_reference=True)
    e2.seed(1)
    e2.reset()
    obs2, _, _ = e2.step(np.zeros((4, 2), np.float32), np.zeros(4, np.int32))
    torch.cuda.synchronize()
    assert int(obs2.abs().sum()) > 0  # tps == num_frames: frame 0 is filled


def test_gym_wrapper_mirrors_reference_interface():
    """gym_agario.AgarioEnv surface: ids, action format, 5-tuple step, HWC observation, episodic truncation."""
    from agarcl_b200.gym_env import make
    env = make("agario-grid-v0", num_bots=3, arena_size=200, num_pellets=100, num_viruses=2, number_steps=5)
    env.seed(3)
    obs, info = env.reset()
    assert obs.shape == (128, 128, 8) and obs.dtype == np.int32 and info == {}
    assert env.observation_space.shape == (128, 128, 8)
    done = False
    for t in range(6):
        obs, rew, done, trunc, info = env.step((np.array([0.5, -0.5], np.float32), 0))
        assert obs.shape == (128, 128, 8) and isinstance(rew, float) and trunc is False and info["steps"] == t + 1
    assert done is True  # number_steps reached (AgarioEnv.py:111-112)
    with pytest.raises(ValueError):
        env.step((np.array([2.0, 0.0], np.float32), 0))  # outside the action space
    with pytest.raises(ValueError):
        env.step([(np.zeros(2, np.float32), 0), (np.zeros(2, np.float32), 0)])  # wrong number of actions
    env.close()
    multi = make("agario-grid-v0", num_agents=2, num_bots=0, arena_size=100, num_pellets=50)
    obs, _ = multi.reset()
    assert isinstance(obs, list) and len(obs) == 2
    obs, rew, done, trunc, _ = multi.step([(np.zeros(2, np.float32), 0), (np.zeros(2, np.float32), 1)])
    assert len(rew) == 2 and len(done) == 2
    multi.close()


def test_gobigger_environment_get_state_against_oracle():
    """agarcl.GoBiggerEnvironment.get_state(): PlayerState objects rebuilt from the device records == oracle records."""
    from _helpers import Oracle, oracle_lib
    from agarcl_b200 import make_cfg, RNG_REPLAY
    from agarcl_b200.env import GoBiggerEnvironment
    oracle_lib().oracle_set_trig_mode(1)
    env = GoBiggerEnvironment(512, 512, 1000, 1, 4, 300, True, 200, 4, 5, 1)
    env.seed(5)
    env.reset()
    cfg = make_cfg(num_agents=1, arena_size=300, num_pellets=200, num_viruses=4, num_bots=5, rng_mode=RNG_REPLAY, cap_replay=16384)
    ora = Oracle(cfg)
    ora.seed_mt(5, 16384)
    ora.reset()
    ora.ram_clear()
    rng = np.random.default_rng(4)
    for st in range(25):
        a = (float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), int(rng.integers(0, 3)))
        env.take_actions([a])
        env.step()
        ora.set_actions(np.array([[a[0], a[1]]], np.float32), np.array([a[2]], np.int32))
        ora.step_with_ram()
    st = env.get_state()
    assert len(st) == 1 and set(st[0]) == {"global_state", "player_states"}
    assert st[0]["global_state"].get_map_width() == 512 and st[0]["global_state"].get_team_num() == 1
    ps = st[0]["player_states"].get_all_player_states()
    assert set(ps) == {p for p in range(6) if ora.ram[p, :4].sum() > 0}
    for p, s in ps.items():
        rec = ora.ram[p]
        assert [len(s.get_food_infos()), len(s.get_virus_infos()), len(s.get_spore_infos()), len(s.get_clone_infos())] == \
            [int(min(rec[0], 192)), int(min(rec[1], 16)), int(min(rec[2], 32)), int(min(rec[3], 32))]
        assert s.get_score() == float(rec[4])
        c0 = s.get_clone_infos()[0]
        assert np.float32(c0.position.x) == rec[968] and np.float32(c0.radius) == rec[970] and c0.owner == p
    assert env.observation_shape() == (25, 512, 512)
    env.close()


def test_batched_vector_env_auto_reset():
    import torch
    from agarcl_b200.gym_env import BatchedAgarioEnv
    env = BatchedAgarioEnv(64, obs_type="grid", num_bots=2, arena_size=150, num_pellets=80, num_viruses=2, number_steps=7)
    env.seed(11)
    obs, _ = env.reset()
    assert obs.shape == (64, 8, 128, 128) and obs.is_cuda
    n_done = 0
    for t in range(20):
        dxdy = torch.rand((64, 2), device="cuda") * 2 - 1
        act = torch.zeros(64, dtype=torch.int32, device="cuda")
        obs, rew, done, trunc, info = env.step(dxdy, act)
        n_done += int(done.sum())
        assert rew.dtype == torch.float32 and done.dtype == torch.bool
        assert int(info["steps"].max()) <= 8
    assert n_done == 64 * 2  # every instance hit number_steps twice in 20 steps (steps 8 and 16)
    ram = BatchedAgarioEnv(8, obs_type="ram", num_bots=3, arena_size=150, num_pellets=80, num_viruses=2)
    o, _ = ram.reset()
    o, *_ = ram.step(torch.zeros((8, 2), device="cuda"), torch.zeros(8, dtype=torch.int32, device="cuda"))
    assert o.shape == (8, 1, 1224) and float(o[:, 0, 3].min()) >= 1.0  # every agent sees its own cell
    env.close()
    ram.close()


def test_snapshot_save_load_through_the_device(tmp_path):
    """agarcl_batch_save_env_state / load_env_state: lossless round trip into another instance continues identically."""
    import torch
    from agarcl_b200._abi import compare_states
    from agarcl_b200.env import BatchedGridEnvironment
    e = BatchedGridEnvironment(4, num_agents=2, num_bots=5, arena_size=300, num_pellets=200, num_viruses=6)
    e.seed(np.array([5, 5, 6, 7], dtype=np.uint64))   # instances 0 and 1 share a seed (same Philox key) ...
    e.reset()
    b = e.batch
    rng = np.random.default_rng(3)

    def step_all(same01):
        dxdy = rng.uniform(-1, 1, size=(4, 2, 2)).astype(np.float32)
        act = rng.integers(0, 3, size=(4, 2)).astype(np.int32)
        if same01:
            dxdy[1], act[1] = dxdy[0], act[0]
        e.step(dxdy.reshape(8, 2), act.reshape(8))

    for _ in range(30):
        step_all(False)
    path = tmp_path / "inst0.json"
    b.save_env_state(0, path)
    b.load_env_state(1, path, lossless=True)          # ... but different instance ids: only equal until the next draw
    s0, s1 = b.download_state(0), b.download_state(1)
    assert not compare_states(s0, s1)
    for _ in range(6):                                 # fewer than the 120-tick regen period: no draws in between
        step_all(True)
    torch.cuda.synchronize()
    if int(b.download_state(0).hdr["rng_cursor"]) == int(s0.hdr["rng_cursor"]):
        assert not compare_states(b.download_state(0), b.download_state(1))
    with pytest.raises(RuntimeError):
        b.load_env_state(0, tmp_path / "missing.json")
    e.close()


def test_unseeded_default_env_spreads_its_pellets():
    """An unseeded reference environment draws from std::random_device (GameState.hpp:59): so does an unseeded drop-in
    (AGARCL_RNG_MT19937 is the default of GridEnvironment / make()); two unseeded environments differ."""
    from agarcl_b200.gym_env import make
    layouts = []
    for _ in range(2):
        env = make("agario-grid-v0", num_bots=3, arena_size=400, num_pellets=300)
        env.reset()
        sv = env._env._ensure().download_state(0)
        px, py = sv.pellets["x"][:300], sv.pellets["y"][:300]
        assert px.std() > 60 and py.std() > 60 and len(np.unique(px)) > 290, "pellets are not spread over the arena"
        assert len({(float(sv.cells[p][0]["x"]), float(sv.cells[p][0]["y"])) for p in range(4)}) == 4
        layouts.append(px.copy())
        env.close()
    assert not np.array_equal(layouts[0], layouts[1])


def test_mt19937_stream_never_runs_out_and_resets_continue_it():
    """AGARCL_RNG_MT19937: the host refills the draw ring ahead of the cursor, so a long game consumes far more draws than
    the ring holds and still equals the oracle fed the seed's whole mt19937_64 stream; a second reset() does NOT restart
    the stream (BaseEnvironment::reset does not reseed, BaseEnvironment.hpp:179-204) but goes on at the cursor."""
    from _helpers import Oracle, oracle_lib
    from agarcl_b200 import make_cfg, RNG_MT19937, RNG_REPLAY
    from agarcl_b200._abi import compare_states
    from agarcl_b200.batch import Batch
    oracle_lib().oracle_set_trig_mode(1)
    kw = dict(num_agents=1, num_bots=3, arena_size=60, num_pellets=150, num_viruses=2)
    cap = 1024  # one step can draw at most 2 * (150 + 34 + 4) = 376
    b = Batch(make_cfg(rng_mode=RNG_MT19937, cap_replay=cap, **kw))
    b.seed(123)
    b.reset()
    ocfg = make_cfg(rng_mode=RNG_REPLAY, cap_replay=1 << 16, **kw)
    ora = Oracle(ocfg)
    ora.seed_mt(123, 1 << 16)
    ora.reset()
    assert not compare_states(ora.state, b.download_state(0))
    rng = np.random.default_rng(0)
    for st in range(700):
        dxdy = rng.uniform(-1, 1, size=(1, 2)).astype(np.float32)
        act = np.zeros(1, np.int32)
        b.set_actions(dxdy, act)
        b.step()
        ora.set_actions(dxdy, act)
        ora.step()
        if st % 50 == 49:
            gs = b.download_state(0)
            assert not compare_states(ora.state, gs), st
            assert int(gs.hdr["rng_cursor"]) == int(ora.state.hdr["rng_cursor"])
    gs = b.download_state(0)
    cur = int(gs.hdr["rng_cursor"])
    assert cur > 2 * cap, f"only {cur} draws consumed: the ring was never wrapped"
    assert int(gs.hdr["flags"]) == int(ora.state.hdr["flags"]) and not (b.flags()[0] & 0x20)  # never AGARCL_FLAG_REPLAY_EXHAUSTED
    # second episode: the stream goes on where the first one stopped
    b.reset()
    ora2 = Oracle(ocfg)
    ora2.set_replay(ora.replay[cur:])
    ora2.reset()
    g2 = b.download_state(0)
    assert not compare_states(ora2.state, g2)
    assert int(g2.hdr["rng_cursor"]) == cur + int(ora2.state.hdr["rng_cursor"])
    # seed() restarts it
    b.seed(123)
    b.reset()
    ora3 = Oracle(ocfg)
    ora3.seed_mt(123, 1 << 16)
    ora3.reset()
    assert not compare_states(ora3.state, b.download_state(0))
    b.close()


def test_philox_episodes_differ_after_auto_reset():
    """counter-based stream: a reset goes on at the instance's cursor, so successive episodes of an instance differ"""
    from agarcl_b200.env import BatchedGridEnvironment
    e = BatchedGridEnvironment(3, num_bots=2, arena_size=200, num_pellets=100, num_viruses=2)
    e.seed(9)
    e.reset()
    first = e.batch.download_state(1).pellets["x"][:100].copy()
    e.reset(np.array([0, 1, 0], np.uint8))
    second = e.batch.download_state(1)
    assert not np.array_equal(first, second.pellets["x"][:100]) and int(second.hdr["rng_cursor"]) == 2 * (2 * (100 + 2 + 3))
    e.seed(9)
    e.reset()
    assert np.array_equal(first, e.batch.download_state(1).pellets["x"][:100])
    e.close()


def test_action_buffers_are_type_checked():
    import torch
    from agarcl_b200.env import BatchedGridEnvironment
    e = BatchedGridEnvironment(4, num_bots=1, arena_size=100, num_pellets=50, num_viruses=0)
    e.reset()
    dxdy = torch.zeros((4, 2), device="cuda")
    with pytest.raises(RuntimeError, match="int32"):
        e.step(dxdy, torch.zeros(4, dtype=torch.int64, device="cuda"))  # torch.randint's default dtype
    with pytest.raises(RuntimeError, match="float32"):
        e.step(dxdy.double(), torch.zeros(4, dtype=torch.int32, device="cuda"))
    with pytest.raises(RuntimeError, match="number of agents"):
        e.step(dxdy[:3].contiguous(), torch.zeros(4, dtype=torch.int32, device="cuda"))
    e.step(np.zeros((4, 2), np.float64), np.zeros(4, np.int64))  # host arrays are converted
    b = e.batch
    with pytest.raises(RuntimeError):
        b.step_mirror(np.zeros((4, 2), np.float32), np.zeros(4, np.int32), rewards_out=np.zeros(4, np.float32))
    e.close()


def test_flags_reduced_over_the_batch():
    """agarcl_batch_flags: OR and per-flag instance counts of hdr.flags, reduced on the device"""
    from agarcl_b200.env import BatchedGridEnvironment
    e = BatchedGridEnvironment(300, num_bots=1, arena_size=100, num_pellets=50, num_viruses=0)
    e.reset()
    assert e.flags() == (0, {})
    b = e.batch
    for i, f in ((7, 0x001), (150, 0x041), (299, 0x040)):
        sv = b.download_state(i)
        sv.hdr["flags"] = f
        b.upload_state(i, sv)
    assert e.flags() == (0x041, {"FOOD_OVERFLOW": 2, "PCD_TIE": 2})
    e.close()


def test_compiled_agarcl_module_drop_in_against_oracle(tmp_path):
    """`import agarcl` (the COMPILED pybind11 module, agarcl_b200/csrc/pybind_agarcl.cpp): GridEnvironment and GoBiggerEnvironment
    used exactly as gym_agario/AgarioEnv.py uses the reference's module, against the oracle with the seed's mt19937_64 stream."""
    import agarcl
    from _helpers import Oracle, oracle_lib
    from agarcl_b200 import make_cfg, RNG_REPLAY
    oracle_lib().oracle_set_trig_mode(1)
    env = agarcl.GridEnvironment(2, 4, 500, True, 300, 5, 6, 1, 0, 0)
    env.seed(77)
    env.reset()
    cfg = make_cfg(num_agents=2, ticks_per_step=4, arena_size=500, num_pellets=300, num_viruses=5, num_bots=6, rng_mode=RNG_REPLAY,
                   cap_replay=16384)
    ora = Oracle(cfg)
    ora.seed_mt(77, 16384)
    ora.reset()
    L = ora.L
    agents_in_map_order = [p for p in list(L.order)[:L.P] if p < L.A]
    first = env.get_state()
    assert [np.array_equal(first[a], ora.obs(a)) for a in range(2)] == [True, True] and env.dones() == [False, False]
    rng = np.random.default_rng(0)
    for st in range(40):
        acts = [(float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), int(rng.integers(0, 3))) for _ in range(2)]
        env.take_actions(acts)
        rew = env.step()
        ora.set_actions(np.array([[a[0], a[1]] for a in acts], np.float32), np.array([a[2] for a in acts], np.int32))
        orew, odone, oobs = ora.step(with_obs=True)
        assert rew == [float(orew[p]) for p in agents_in_map_order]  # quirk Q15 ordering
        assert env.dones() == [bool(x) for x in odone]
        state = env.get_state()
        assert len(state) == 2 and state[0].shape == (8, 128, 128) and state[0].dtype == np.int32 and state[0].flags.owndata
        for a in range(2):
            assert np.array_equal(state[a], oobs[a])
    with pytest.raises(RuntimeError, match="does not match number of agents"):
        env.take_actions([(0.0, 0.0, 0)])
    env.save_env_state(str(tmp_path / "s.json"))
    assert (tmp_path / "s.json").stat().st_size > 1000
    env.close()
    # GoBiggerEnvironment.get_state(): [{"global_state", "player_states"}] with the bound info classes
    g = agarcl.GoBiggerEnvironment(512, 512, 1000, 1, 4, 300, True, 200, 4, 5, 1)
    g.seed(5)
    g.reset()
    cfg = make_cfg(num_agents=1, arena_size=300, num_pellets=200, num_viruses=4, num_bots=5, rng_mode=RNG_REPLAY, cap_replay=16384)
    ora = Oracle(cfg)
    ora.seed_mt(5, 16384)
    ora.reset()
    ora.ram_clear()
    rng = np.random.default_rng(4)
    for st in range(25):
        a = (float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), int(rng.integers(0, 3)))
        g.take_actions([a])
        g.step()
        ora.set_actions(np.array([[a[0], a[1]]], np.float32), np.array([a[2]], np.int32))
        ora.step_with_ram()
    st = g.get_state()
    assert len(st) == 1 and set(st[0]) == {"global_state", "player_states"} and st[0]["global_state"].get_map_width() == 512
    ps = st[0]["player_states"].get_all_player_states()
    assert set(ps) == {p for p in range(6) if ora.ram[p, :4].sum() > 0}
    for p, s in ps.items():
        rec = ora.ram[p]
        assert [len(s.get_food_infos()), len(s.get_virus_infos()), len(s.get_spore_infos()), len(s.get_clone_infos())] == \
            [int(min(rec[0], 192)), int(min(rec[1], 16)), int(min(rec[2], 32)), int(min(rec[3], 32))]
        assert s.get_score() == float(rec[4]) and s.get_player_id() == p
        c0 = s.get_clone_infos()[0]
        assert np.float32(c0.position.x) == rec[968] and np.float32(c0.get_position_y()) == rec[969] and np.float32(c0.radius) == rec[970]
        assert c0.owner == p and c0.teamId == 0
    assert g.observation_shape() == (25, 512, 512)
    g.close()


def test_strict_reference_later_episodes_follow_pid_growth():
    """Quirk Q3 under strict_reference: the reference's player map and pid counter outlive a reset, so the k-th reset leaves the
    players in another iteration order than a fresh engine (26 players: from the first reset on) and CloneInfo::owner counts on.
    The batch replays that (agarcl_batch_reset, unmasked); the oracle is given the order of the episode
    (oracle_player_order_episode, pinned to the reference over four episodes in tests/test_oracle_vs_reference.py) and both are
    compared field by field over three episodes, the draw stream carried across the resets."""
    import ctypes
    import torch
    from _helpers import Oracle, oracle_lib, oracle_layout, random_actions
    from agarcl_b200 import make_cfg, RNG_REPLAY
    from agarcl_b200._abi import compare_states
    from agarcl_b200.batch import Batch
    oracle_lib().oracle_set_trig_mode(1)
    n, rl = 2, 1 << 15
    cfg = make_cfg(n_instances=n, strict_reference=True, ticks_per_step=1, rng_mode=RNG_REPLAY, cap_replay=rl, ram_obs=True,
                   arena_size=400, num_pellets=300, num_viruses=6)
    b = Batch(cfg)
    oras = []
    for i in range(n):
        o = Oracle(cfg, oracle_layout(cfg))
        o.seed_mt(900 + i, rl)
        b.set_replay(i, o.replay)
        oras.append(o)
    P, A = b.layout.P, b.layout.A
    assert P == 26
    rngs = [np.random.default_rng(40 + i) for i in range(n)]
    fresh_order = list(b.layout.order)[:P]
    orders = []
    try:
        for episode in range(3):
            b.reset()
            want = (ctypes.c_int * 64)()
            oracle_lib().oracle_player_order_episode(P, episode, want)
            assert list(b.layout.order)[:P] == list(want)[:P], episode
            orders.append(list(want)[:P])
            oracle_lib().oracle_set_pid_base(episode * P)
            for o in oras:
                for k in range(P):
                    o.L.order[k] = want[k]
                o.reset() if episode == 0 else o.reset_keep_stream()
                o.ram_clear()
            for i, o in enumerate(oras):
                d = compare_states(o.state, b.download_state(i))
                assert not d, f"episode {episode} reset, instance {i}: {d[:4]}"
            obs_t, rew_t, ram_t = b.obs_tensor(), b.rewards_tensor(), b.ram_tensor()
            for st in range(30):
                dxdy = np.zeros((n, A, 2), np.float32)
                act = np.zeros((n, A), np.int32)
                for i in range(n):
                    dxdy[i], act[i] = random_actions(rngs[i], A, 1 / 3, 1 / 3)
                b.set_actions(dxdy, act)
                b.step()
                torch.cuda.synchronize()
                g_rew = rew_t.cpu().numpy().reshape(n, A)
                g_obs = obs_t.cpu().numpy().reshape(n, A, *b.obs_shape[1:])
                g_ram = ram_t.cpu().numpy()
                for i, o in enumerate(oras):
                    o.set_actions(dxdy[i], act[i])
                    o_rew, o_done, o_obs = o.step_with_ram(with_obs=True)
                    d = compare_states(o.state, b.download_state(i))
                    assert not d, f"episode {episode} step {st} instance {i}: {d[:4]}"
                    assert np.array_equal(g_rew[i], o_rew)
                    assert np.array_equal(g_obs[i], o_obs)
                    same = (g_ram[i].view(np.uint32) == o.ram.view(np.uint32)) | (np.isnan(g_ram[i]) & np.isnan(o.ram))
                    assert same.all(), (episode, st, i, np.argwhere(~same)[:4].tolist())
            if episode > 0:  # CloneInfo::owner (record word 975 = 968 + 7) of the agent's own first clone counts on with the pids
                assert float(g_ram[0][0, 975]) >= episode * P
        assert orders[0] == fresh_order and orders[1] != fresh_order
    finally:
        oracle_lib().oracle_set_pid_base(0)
        b.close()
