"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libagarcl_ref.so, built from
/root/reference by oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py

Each fixture holds, for one (config, seed): the config, the mt19937_64 draw stream the reference consumed
(ref_rng_peek after seeding), the action stream, the reference's state blob after reset and after every
`every` steps, rewards/dones of every step, and a few forced add_frame observations.  The oracle port and
the CUDA path are replayed from the reset blob + draws + actions and must reproduce all of it.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _helpers import Reference, oracle_layout, random_actions  # noqa: E402
from agarcl_b200._abi import make_cfg  # noqa: E402

CASES = {
    "c1_single_agent": (dict(num_bots=0, num_viruses=0), dict(seed=101, steps=120)),
    "c2_default_bots": (dict(), dict(seed=102, steps=120)),
    "c4_multi_agent": (dict(num_agents=4, num_bots=8, cap_foods=1024), dict(seed=103, steps=160, p_feed=0.3, p_split=0.3, boost=1000)),
    "dense_small_arena": (dict(num_agents=2, num_bots=25, arena_size=300, num_pellets=300, num_viruses=10, cap_foods=1024),
                          dict(seed=104, steps=160, boost=3000)),
    "dense_no_virus": (dict(num_agents=2, num_bots=20, arena_size=250, num_pellets=250, num_viruses=0, cap_foods=1024),
                       dict(seed=107, steps=160, boost=2000)),
    # many virus pops (Engine::disrupt with its libm trigonometry), fed viruses, 14-cell players
    "virus_heavy": (dict(num_agents=3, num_bots=10, arena_size=250, num_pellets=200, num_viruses=40, cap_foods=2048, cap_viruses=256),
                    dict(seed=108, steps=200, boost=400, p_feed=0.5, p_split=0.1)),
    "mode2_squares_decay": (dict(num_bots=0, num_viruses=0, arena_size=350, num_pellets=500, mode_number=2), dict(seed=105, steps=100)),
    "mode9_one_bot": (dict(num_bots=1, num_viruses=0, arena_size=100, num_pellets=50, mode_number=9), dict(seed=106, steps=120, boost=150)),
}


def make(name, cfg_kwargs, seed, steps, p_feed=1 / 3, p_split=1 / 3, boost=None, every=20):
    cfg = make_cfg(**cfg_kwargs)
    L = oracle_layout(cfg)
    ref = Reference(cfg, L)
    ref.seed(seed)
    draws = ref.peek_draws(1 << 15)
    ref.reset()
    if boost:
        for a in range(L.A):
            ref.set_cell_mass(a, 0, boost)
    s0, miss = ref.dump()
    assert miss == 0
    rng = np.random.default_rng(seed)
    A = L.A
    dxdy = np.zeros((steps, A, 2), np.float32)
    act = np.zeros((steps, A), np.int32)
    rew = np.zeros((steps, A), np.float64)
    done = np.zeros((steps, A), np.uint8)
    blobs, blob_steps, obs, obs_steps = [], [], [], []
    virus_hits = np.zeros(steps, np.int32)  # cumulative virus contacts (pop or eat) after each step
    for st in range(steps):
        dxdy[st], act[st] = random_actions(rng, A, p_feed, p_split)
        ref.set_actions(dxdy[st], act[st])
        rew[st], done[st] = ref.step()
        virus_hits[st] = int(ref.dump()[0].players["viruses_eaten"].sum())
        if (st + 1) % every == 0 or st == steps - 1:
            s, miss = ref.dump()
            assert miss == 0, (name, st, miss)
            blobs.append(s.blob.copy())
            blob_steps.append(st)
        if (st + 1) % (2 * every) == 0:
            obs.append(np.stack([ref.obs(a) for a in range(A)]))
            obs_steps.append(st)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), cfg=np.frombuffer(bytes(cfg), dtype=np.int32), seed=seed,
                        boost=boost or 0, draws=draws, dxdy=dxdy, act=act, rew=rew, done=done, blob0=s0.blob,
                        blobs=np.stack(blobs), blob_steps=np.array(blob_steps), obs=np.stack(obs).astype(np.int16),
                        obs_steps=np.array(obs_steps), order=np.array(ref.order()), virus_hits=virus_hits)
    print(name, "ok", os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    only = sys.argv[1:]  # fixture names; none = all
    for name, (ck, rk) in CASES.items():
        if not only or name in only:
            make(name, ck, **rk)
