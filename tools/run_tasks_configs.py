#!/usr/bin/env python
"""Runs the reference's ten mini-game task configurations (bench/tasks_configs/mode_{1..10}.json: arena 350, 500 pellets, no
viruses, mode k, 0 bots (modes 1-6) or one bot of type k-7 (modes 7-10), episodes of 500 / 3000 / 10000 steps) through the
batched vector environment (agarcl_b200.gym_env.BatchedAgarioEnv: N lockstep instances, auto-reset) with the random-walk policy of
bench/go_bigger_example.py:100-103, and prints one JSON line per mode: env-steps/s, episodes finished, mean return, state flags.

    python tools/run_tasks_configs.py [--envs 4096] [--steps 600] [--modes 1 2 ...] [--dir /root/reference/bench/tasks_configs]

With --dir the mode_k.json files themselves are loaded (keys the batched grid environment does not have -- screen_len, video_path,
agent_view, render_mode, load_env_snapshot -- are ignored; obs_type "screen" is OpenGL and out of scope: the grid observation with
all channels is produced instead).  Without it the table below (the same numbers) is used, so the runner works on the GPU box."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# bench/tasks_configs/mode_k.json, the keys that reach the engine
TASKS = {k: dict(ticks_per_step=4, num_frames=1, arena_size=350, num_pellets=500, num_viruses=0, num_bots=(1 if k >= 7 else 0),
                 pellet_regen=True, grid_size=128, reward_type=1, c_death=0, mode=k,
                 number_steps=(500 if k <= 2 else 3000 if k <= 6 else 10000), env_type=0) for k in range(1, 11)}
KEEP = set(TASKS[1])


def load_dir(d):
    out = {}
    for k in range(1, 11):
        with open(os.path.join(d, f"mode_{k}.json")) as f:
            out[k] = {key: v for key, v in json.load(f).items() if key in KEEP}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--modes", type=int, nargs="*", default=list(range(1, 11)))
    ap.add_argument("--dir", default=None)
    args = ap.parse_args()
    import torch
    from agarcl_b200.gym_env import BatchedAgarioEnv
    tasks = load_dir(args.dir) if args.dir else TASKS
    N = args.envs
    for k in args.modes:
        kw = dict(tasks[k], observe_cells=True, observe_others=True, observe_viruses=True, observe_pellets=True)
        env = BatchedAgarioEnv(N, obs_type="grid", **kw)
        env.seed(1000 * k)
        env.reset()
        gen = torch.Generator(device="cuda")
        gen.manual_seed(k)
        ret = torch.zeros(N, device="cuda")
        finished, ret_sum = 0, 0.0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(args.steps):
            dxdy = (torch.rand((N, 2), device="cuda", generator=gen) * 2 - 1).float()
            act = torch.randint(0, 3, (N,), device="cuda", generator=gen, dtype=torch.int32)
            _, rew, done, _, _ = env.step(dxdy, act)
            ret += rew
            if bool(done.any()):
                finished += int(done.sum())
                ret_sum += float(ret[done].sum())
                ret[done] = 0
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        f, names = env.flags()
        print(json.dumps({"task": f"mode_{k}", "envs": N, "steps": args.steps, "env_steps_per_s": N * args.steps / sec,
                          "episode_steps": kw["number_steps"], "episodes_finished": finished,
                          "mean_return": (ret_sum / finished) if finished else None, "state_flags": names}))
        env.close()


if __name__ == "__main__":
    main()
