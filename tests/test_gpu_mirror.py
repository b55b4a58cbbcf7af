"""-m gpu: the host-resident observation mirror (agarcl_batch_mirror / sync_mirror / step_mirror, mirror.cu) is
element-for-element identical to the device observation after every step, reset and render — for the int32 and
int16 dtypes, several frames, several agents, small grids, and when every image is forced onto the dense copy."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(env, steps, rng, reset_at=(), check_every=1):
    import torch
    b = env.batch
    NA = b.N * b.A
    env.reset()
    m = b.sync_mirror()
    assert m.shape == b.obs_shape and m.dtype == b.obs_dtype
    assert np.array_equal(m, b.obs_tensor().cpu().numpy()), "mirror differs after reset"
    rew = np.zeros(NA, np.float64)
    done = np.zeros(NA, np.uint8)
    dense_seen = 0
    for st in range(steps):
        dxdy = rng.uniform(-1, 1, size=(NA, 2)).astype(np.float32)
        act = rng.integers(0, 3, size=NA).astype(np.int32)
        m = b.step_mirror(dxdy, act, rew, done)
        dense_seen += b.mirror_stats()["dense_images"]
        if st % check_every == 0:
            dev = b.obs_tensor().cpu().numpy()
            if not np.array_equal(m, dev):
                bad = np.argwhere(m != dev)
                raise AssertionError(f"mirror differs from the device observation at step {st}: {len(bad)} elements, first {bad[:5].tolist()}, "
                                     f"mirror {m[tuple(bad[0])]} device {dev[tuple(bad[0])]}")
            assert np.array_equal(rew, b.rewards_tensor().cpu().numpy())
            assert np.array_equal(done, b.dones_tensor().cpu().numpy())
        if st in reset_at:
            mask = (rng.uniform(size=b.N) < 0.5).astype(np.uint8)
            b.reset(mask)
            m = b.sync_mirror()
            assert np.array_equal(m, b.obs_tensor().cpu().numpy()), "mirror differs after a masked reset"
    torch.cuda.synchronize()
    return dense_seen


def test_mirror_default_bots_int32():
    from agarcl_b200.env import BatchedGridEnvironment
    env = BatchedGridEnvironment(96)
    env.seed(21)
    dense = _run(env, 80, np.random.default_rng(0), reset_at=(20, 55))
    st = env.batch.mirror_stats()
    assert dense == 0, "the default workload must stay on the sparse path"
    assert 0 < st["entries"] < 96 * 1024
    assert st["d2h_bytes"] < 96 * 8 * 128 * 128 * 4 // 20  # less than 5 % of the dense copy
    env.close()


def test_mirror_small_arena_walls_and_deaths():
    """arena 150 < view: every frame has out-of-bounds rows and columns that move every step; agents die and respawn."""
    from agarcl_b200.env import BatchedGridEnvironment
    env = BatchedGridEnvironment(64, num_agents=2, num_bots=6, arena_size=150, num_pellets=120, num_viruses=3)
    env.seed(5)
    _run(env, 120, np.random.default_rng(1), reset_at=(40,))
    env.close()


def test_mirror_int16_frames_and_small_grid():
    from agarcl_b200 import OBS_I16, make_cfg
    from agarcl_b200.batch import Batch

    class E:  # minimal env shim over a Batch with a non-default observation configuration
        def __init__(self, **kw):
            self.batch = Batch(make_cfg(**kw))
            self.batch.seed(3)

        def reset(self):
            self.batch.reset()

    for kw in (dict(obs_dtype=OBS_I16), dict(num_frames=2, grid_size=64), dict(grid_size=40, num_frames=3, ticks_per_step=3),
               dict(observe_pellets=False, observe_others=False), dict(strict_reference=1, ticks_per_step=1)):
        e = E(n_instances=24, num_agents=2, num_bots=5, arena_size=260, num_pellets=250, num_viruses=4, **kw)
        _run(e, 40, np.random.default_rng(2), reset_at=(15,))
        e.batch.close()


def test_mirror_dense_fallback_is_exact():
    """Entry capacity 1 per image (64 per chunk of images on the fused path): the lists overflow, images take the dense
    copy and move between the sparse and the dense path from step to step; the mirror must not notice."""
    from agarcl_b200.env import BatchedGridEnvironment
    os.environ["AGARCL_MIRROR_CAP_IMG"] = "1"
    try:
        env = BatchedGridEnvironment(32, num_bots=4, arena_size=600, num_pellets=120, num_viruses=1)
        env.seed(8)
        dense = _run(env, 60, np.random.default_rng(3), reset_at=(10, 30))
        assert dense > 0
        env.close()
    finally:
        del os.environ["AGARCL_MIRROR_CAP_IMG"]


def test_mirror_full_size():
    """configs[1] at full size (4096 instances): the whole 2.1 GB mirror equals the device tensor after 20 steps."""
    import torch
    from agarcl_b200.env import BatchedGridEnvironment
    N = 4096
    env = BatchedGridEnvironment(N)
    env.seed(77)
    env.reset()
    b = env.batch
    rng = np.random.default_rng(4)
    for st in range(20):
        m = b.step_mirror(rng.uniform(-1, 1, size=(N, 2)).astype(np.float32), rng.integers(0, 3, size=N).astype(np.int32))
    dev = b.obs_tensor()
    for i in range(0, N, 512):  # element-for-element, 268 MB of the mirror at a time
        assert torch.equal(torch.from_numpy(m[i:i + 512].copy()).cuda(), dev[i:i + 512]), f"images {i}..{i + 511} differ"
    assert b.mirror_stats()["dense_images"] == 0
    env.close()


def test_observation_lists_decode_to_the_device_observation():
    """agarcl_batch_step_lists: the lists in pinned host memory -- per image the out-of-bounds row / column masks and the list of
    integer operations -- decode to exactly the device observation, by the C decoder (every image) and by a numpy decoder written
    from the layout documented in include/agarcl_b200.h (a sample); rewards and dones travel in the records."""
    from agarcl_b200 import OBS_I16, make_cfg
    from agarcl_b200.batch import Batch
    for kw in (dict(), dict(obs_dtype=OBS_I16, num_agents=2, arena_size=180, num_pellets=150, num_viruses=4, num_bots=6)):
        N = 80
        b = Batch(make_cfg(n_instances=N, **kw))
        b.seed(33)
        b.reset()
        NA = N * b.A
        rng = np.random.default_rng(5)
        rew, done = np.zeros(NA, np.float64), np.zeros(NA, np.uint8)
        for st in range(40):
            dxdy = rng.uniform(-1, 1, size=(NA, 2)).astype(np.float32)
            act = rng.integers(0, 3, size=NA).astype(np.int32)
            ol = b.step_lists(dxdy, act, rew, done)
            if st % 8 == 7:
                dev = b.obs_tensor().cpu().numpy()
                assert np.array_equal(rew, b.rewards_tensor().cpu().numpy()) and np.array_equal(done, b.dones_tensor().cpu().numpy())
                r2, d2 = ol.rewards_dones()
                assert np.array_equal(r2, rew) and np.array_equal(d2, done)
                for i in range(NA):
                    assert np.array_equal(ol.expand(i), dev[i]), f"step {st}: image {i} decodes differently"
                for i in (0, 1, NA // 2, NA - 1):
                    assert np.array_equal(ol.decode(i), dev[i]), f"step {st}: numpy decoder, image {i}"
                assert sorted(ol.slot_of.tolist()) == list(range(NA))
        b.close()
    with pytest.raises(RuntimeError, match="single fused step kernel"):
        b2 = Batch(make_cfg(n_instances=4, num_frames=2))
        b2.reset()
        b2.step_lists(np.zeros((4, 2), np.float32), np.zeros(4, np.int32))
