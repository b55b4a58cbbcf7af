"""Golden traces produced by the compiled reference (tests/golden/make_golden.py).
CPU: the oracle port replays every trace bit-exactly (states, rewards, dones, observations).
GPU (-m gpu): the CUDA path replays the same traces through the C ABI."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from _helpers import Oracle, oracle_layout, oracle_lib
from agarcl_b200._abi import Cfg, StateView, compare_states

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
assert GOLDEN, "no golden fixtures"


def load(path):
    z = np.load(path)
    cfg = Cfg.from_buffer_copy(z["cfg"].tobytes())
    return z, cfg


@pytest.mark.parametrize("trig_mode", [0, 1], ids=["libm", "restated"])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_replays_reference_trace(path, trig_mode):
    """trig_mode 0: libm, what the reference executes; 1: the restatement of glibc's atanf / sinf / cosf that the CUDA path
    runs (tests/test_trig_exact.py) -- both replay every trace bit-exactly, virus pops included."""
    oracle_lib().oracle_set_trig_mode(trig_mode)
    z, cfg = load(path)
    L = oracle_layout(cfg)
    assert list(L.order)[:L.P] == z["order"].tolist()
    ora = Oracle(cfg, L, replay=z["draws"])
    ora.reset()  # from the recorded draw stream
    if int(z["boost"]):
        for a in range(L.A):
            ora.state.cells[a][0]["mass"] = int(z["boost"])
    assert not compare_states(StateView(L, z["blob0"].copy()), ora.state)
    bi = oi = 0
    for st in range(z["dxdy"].shape[0]):
        ora.set_actions(z["dxdy"][st], z["act"][st])
        rew, done, _ = ora.step()
        assert np.array_equal(rew, z["rew"][st]) and np.array_equal(done, z["done"][st]), st
        if bi < len(z["blob_steps"]) and st == z["blob_steps"][bi]:
            d = compare_states(StateView(L, z["blobs"][bi].copy()), ora.state)
            assert not d, f"step {st}: {d[:5]}"
            bi += 1
        if oi < len(z["obs_steps"]) and st == z["obs_steps"][oi]:
            got = np.stack([ora.obs(a) for a in range(L.A)])
            assert np.array_equal(got, z["obs"][oi].astype(np.int32)), f"obs at step {st}"
            oi += 1
    assert bi == len(z["blob_steps"]) and oi == len(z["obs_steps"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_replays_reference_trace(path):
    """The reference trace is reproduced bit-exactly -- every discrete field, every fp32 field, rewards, dones and
    observations -- over its WHOLE length, virus contacts included: Engine::disrupt's atanf / cosf / sinf are glibc's
    own algorithms restated on the device (device_math.cuh g_atanf / g_sincosf, pinned by tests/test_trig_exact.py)."""
    import torch
    from agarcl_b200 import RNG_REPLAY
    from agarcl_b200.batch import Batch
    z, cfg = load(path)
    cfg.n_instances = 1
    cfg.rng_mode = RNG_REPLAY
    cfg.cap_replay = int(z["draws"].size)
    b = Batch(cfg)
    L = b.layout
    b.set_replay(0, z["draws"])
    b.reset()
    if int(z["boost"]):
        sv = b.download_state(0)
        for a in range(L.A):
            sv.cells[a][0]["mass"] = int(z["boost"])
        b.upload_state(0, sv)
    assert not compare_states(StateView(L, z["blob0"].copy()), b.download_state(0))
    hits = z["virus_hits"]
    bi = oi = 0
    for st in range(len(hits)):
        b.set_actions(z["dxdy"][st], z["act"][st])
        b.step()
        rew = b.rewards_tensor().cpu().numpy()
        done = b.dones_tensor().cpu().numpy()
        assert np.array_equal(rew, z["rew"][st]) and np.array_equal(done, z["done"][st]), st
        if bi < len(z["blob_steps"]) and st == z["blob_steps"][bi]:
            d = compare_states(StateView(L, z["blobs"][bi].copy()), b.download_state(0))
            assert not d, f"step {st}: {d[:5]}"
            bi += 1
        if oi < len(z["obs_steps"]) and st == z["obs_steps"][oi]:
            b.render()
            got = b.obs_tensor().cpu().numpy()
            assert np.array_equal(got, z["obs"][oi].astype(np.int32)), f"obs at step {st}"
            oi += 1
    assert bi == len(z["blob_steps"]) and oi == len(z["obs_steps"])
    print(os.path.basename(path), len(hits), "steps,", int(hits[-1]), "virus contacts,", bi, "state checkpoints,", oi, "observations")
    b.close()
