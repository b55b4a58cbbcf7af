#!/usr/bin/env python
"""Work distribution of the self-collision pair solver at steady state (configs[1]), measured with the oracle
built with -DORACLE_STATS (gcc ... -DORACLE_STATS -o /tmp/liboracle_stats.so oracle/oracle.c).  Design input only."""
import ctypes as C
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(seed, settle=int(os.environ.get("SETTLE", 2000)), measure=int(os.environ.get("MEASURE", 50))):
    import _helpers as H
    from agarcl_b200._abi import make_cfg
    lib = C.CDLL("/tmp/liboracle_stats.so")
    H._oracle = lib
    lib.oracle_make_layout.restype = C.c_int
    cfg = make_cfg(n_instances=1, num_agents=1, ticks_per_step=4, arena_size=1000, pellet_regen=True, num_pellets=1000,
                   num_viruses=25, num_bots=25, reward_type=1, c_death=0, mode_number=0, num_frames=1, grid_size=128,
                   observe_cells=True, observe_others=True, observe_viruses=True, observe_pellets=True, rng_mode=1)
    o = H.Oracle(cfg)
    o.seed_mt(seed, 1 << 18)
    o.reset()
    rng = np.random.default_rng(seed)
    names = ["calls", "pairs", "passes", "static", "wave"]
    arr = {k: (C.c_longlong * 33).in_dll(lib, "oracle_stats_" + k) for k in names}
    hist = (C.c_longlong * (33 * 64)).in_dll(lib, "oracle_stats_pairhist")
    ncells_hist = np.zeros(33, np.int64)
    multi_per_inst = []
    for t in range(settle + measure):
        if t == settle:
            for k in names:
                for i in range(33):
                    arr[k][i] = 0
            for i in range(33 * 64):
                hist[i] = 0
        dxdy = rng.uniform(-1, 1, size=(1, 2)).astype(np.float32)
        act = rng.integers(0, 3, size=1).astype(np.int32)
        o.set_actions(dxdy, act)
        o.step()
        if t >= settle:
            n = np.array(o.state.players["n_cells"])
            for v in n:
                ncells_hist[min(int(v), 32)] += 1
            multi_per_inst.append(int((n >= 2).sum()))
    out = {k: np.array(list(arr[k]), np.int64) for k in names}
    out["hist"] = np.array(list(hist), np.int64).reshape(33, 64)
    out["ncells"] = ncells_hist
    out["multi"] = np.array(multi_per_inst)
    out["flags"] = int(o.state.hdr["flags"])
    return out


if __name__ == "__main__":
    nproc = int(os.environ.get("NPROC", 8))
    ninst = int(os.environ.get("NINST", 16))
    with Pool(nproc) as pool:
        res = pool.map(run, range(1, ninst + 1))
    tot = {k: sum(r[k] for r in res) for k in ("calls", "pairs", "passes", "static", "wave", "hist", "ncells")}
    measure = int(os.environ.get("MEASURE", 50))
    ticks = ninst * measure * 4
    print("flags", [hex(r["flags"]) for r in res])
    print("multi-cell players per instance (mean over steps):", np.mean([r["multi"].mean() for r in res]))
    print(f"per tick per instance: calls(n>=2) {tot['calls'][2:].sum() / ticks:.2f}  pairs {tot['pairs'].sum() / ticks:.1f}  "
          f"static {tot['static'].sum() / ticks:.1f}  passes {tot['passes'][2:].sum() / ticks:.2f}  wave {tot['wave'].sum() / ticks:.1f}")
    print(" n   players/step  calls/tick  pairs/call  passes/call  wave/call  static/call  share_of_pairs")
    for n in range(1, 33):
        if tot["calls"][n] == 0 and tot["ncells"][n] == 0:
            continue
        c = max(tot["calls"][n], 1)
        print(f"{n:2d}  {tot['ncells'][n] / (ninst * measure):10.3f}  {tot['calls'][n] / ticks:10.3f}  {tot['pairs'][n] / c:10.1f}  "
              f"{tot['passes'][n] / c:10.2f}  {tot['wave'][n] / c:10.1f} {tot['static'][n] / c:10.1f}   {tot['pairs'][n] / max(1, tot['pairs'].sum()):.3f}")
    print("pairs-per-call histogram (bins of 8) for n=14:", tot["hist"][14][:40].tolist())
