#!/usr/bin/env python
"""bench.py — grid-obs env-steps/sec of the batched AgarCL hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference C++ engine on the host cores

One "step" = one env-step of every instance of the batch: ticks_per_step engine ticks + bots +
regen/respawn + rewards/dones + one grid observation per agent.  Workload = BASELINE.json configs[1]:
4096 lockstep instances per GPU, 1 agent + 25 default bots, 25 viruses, 1000 pellets, arena 1000,
128x128x8 int32 grid observation.  Instances are independent, so N GPUs run N shards of 4096
instances with no collective on the step path ("weak" scaling); torch.distributed is used only for
the barrier and the max-over-ranks of the device time.

Game age: the cost of a step grows with the age of the games (players split, get popped by viruses into up to
16 cells, meet each other) and plateaus after about 1500 env-steps (tools/exp_age.py), so BOTH arms first settle
their instances for --settle env-steps (default 2000) and measure the steady state that a continuing task
(mode 0 never ends an episode) spends its life in; the young-game figure (age 100) is reported next to it as
`age_profile`.

`value` is timed with inputs (per-step action tensors) resident in HBM; `e2e` goes through the
reference-facing C-ABI call agarcl_batch_step_mirror with pinned HOST buffers: actions H2D, rewards / dones D2H and
the update of the dense host observation mirror are inside the timed region (the dense-copy call
agarcl_batch_step_host is timed beside it as e2e.dense_copy).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def host_cores():
    """cores this process may run on (affinity / cgroup aware, unlike os.cpu_count())"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1

_BASE = dict(num_agents=1, ticks_per_step=4, arena_size=1000, pellet_regen=True, num_pellets=1000, num_viruses=25,
             num_bots=25, reward_type=1, c_death=0, mode_number=0, num_frames=1, grid_size=128,
             observe_cells=True, observe_others=True, observe_viruses=True, observe_pellets=True)
# BASELINE.json configs[0..4] as concrete synthetic workloads.  c2 is the one the metric is quoted on and
# the default; the others are extra bench lines (`--config`).  p_feed / p_split < 0: a ~ U{0,1,2} (bench/go_bigger_example.py:100-103).
CONFIGS = {
    "c1": dict(workload=dict(_BASE, num_viruses=0, num_bots=0), instances=1, settle=500, obs="grid", p_feed=-1, p_split=-1, boost=0,
               name="agario-grid-v0 single instance, 1 agent, arena 1000, 1000 pellets, no viruses, no bots, tps 4, 8x128x128 int32 grid obs, "
                    "random-walk actions (BASELINE.json configs[0]; gym defaults gym_agario/AgarioEnv.py:314-322)"),
    "c2": dict(workload=dict(_BASE), instances=4096, settle=2000, obs="grid", p_feed=-1, p_split=-1, boost=0,
               name="agario-grid-v0 x 4096 lockstep instances per GPU, 1 agent + 25 default bots, 25 viruses, "
                    "1000 pellets, arena 1000, tps 4, 8x128x128 int32 grid obs (BASELINE.json configs[1])"),
    "c3": dict(workload=dict(_BASE, num_bots=8, num_viruses=10), instances=8192, settle=2000, obs="ram", p_feed=0.0, p_split=0.0, boost=0,
               name="agario-ram-v0 x 8192 lockstep instances per GPU (65 536 over 8 GPUs), 1 agent + 8 bots, 10 viruses, 1000 pellets, arena 1000, "
                    "tps 4, structured observation record per player, continuous random actions, a = 0 (BASELINE.json configs[2]; "
                    "bench/screen_obs_example.py:34-38 scaled to arena 1000)"),
    "c4": dict(workload=dict(_BASE, num_agents=4, num_bots=8, cap_foods=2048), instances=2048, settle=300, obs="grid", p_feed=0.3, p_split=0.3,
               boost=1000,
               name="agario-grid-v0 x 2048 lockstep instances per GPU, 4 agents + 8 bots, 25 viruses, arena 1000, tps 4, 4 x 8x128x128 int32 grid obs, "
                    "split w.p. 0.3 / feed w.p. 0.3, agents raised to mass 1000 after reset (BASELINE.json configs[3])"),
    "c5": dict(workload=dict(_BASE, arena_size=2000, num_pellets=4000, num_viruses=50, cap_viruses=128), instances=8192, settle=2000, obs="grid",
               p_feed=-1, p_split=-1, boost=0,
               name="agario-grid-v0 continuing task (mode 0) x 8192 lockstep instances per GPU, arena 2000, 4000 pellets, 50 viruses, 1 agent + 25 bots, "
                    "tps 4, 8x128x128 int32 grid obs (BASELINE.json configs[4]; the ten bench/tasks_configs/mode_k.json mini-games: tools/run_tasks_configs.py)"),
}
CONFIG = "c2"
WORKLOAD = CONFIGS[CONFIG]["workload"]
INSTANCES_PER_GPU = CONFIGS[CONFIG]["instances"]
WORKLOAD_NAME = CONFIGS[CONFIG]["name"]
METRIC = "grid-obs env-steps/sec"
UNIT = "env-steps/s"


def select_config(name):
    global CONFIG, WORKLOAD, INSTANCES_PER_GPU, WORKLOAD_NAME, METRIC
    CONFIG = name
    WORKLOAD = dict(CONFIGS[name]["workload"])
    INSTANCES_PER_GPU = CONFIGS[name]["instances"]
    WORKLOAD_NAME = CONFIGS[name]["name"]
    METRIC = "ram-obs env-steps/sec" if CONFIGS[name]["obs"] == "ram" else "grid-obs env-steps/sec"


def algorithmic_bytes(n_pel=1000, n_vir=25, n_food=0, n_cell=40, P=26, A=1, C_=8, G=128, s_obs=4, ram=False):
    """B = A*C*G^2*s_obs + 2*S_state + A*21 per env-step, split per kernel; with the structured ("ram")
    observation the frame is one record of 1224 float32 per player instead of the grid (include/agarcl_b200.h AGARCL_RAM_RECORD)."""
    s_state = 8 * n_pel + 24 * n_vir + 16 * n_food + 37 * n_cell + 64 * P + 16
    obs = P * 1224 * 4 if ram else A * C_ * G * G * s_obs
    return dict(obs_kernel=obs + s_state, sim_kernel=2 * s_state + A * 21, step=obs + 2 * s_state + A * 21, s_state=s_state)


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this same command (profiles/traffic.json); None when there is none."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if kernel is None or not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get(kernel)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class NvmlSampler:
    """SM clock / power / throttle reasons polled through NVML every few ms DURING the timed region (the
    region lasts tens of ms, too short for `nvidia-smi -lms`); same quantities as the B200_PROFILING.md recipe."""

    def __init__(self, index, period_s=0.003):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.period = period_s
        self.samples = []
        self._stop = threading.Event()
        self.thread = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        nv = self.nv
        self._stop.set()
        self.thread.join(timeout=1)
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted({k for _, rs in self.samples for k, bit in names.items() if rs & bit})
        sm = [x for x, _ in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml"}


def make_sampler(index):
    try:
        return NvmlSampler(index)
    except Exception:
        return ClockSampler(index)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the compiled reference engine (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------
def _ref_pool():
    from agarcl_b200._abi import make_cfg
    so = os.path.join(ROOT, "oracle", "_ref", "libagarcl_ref.so")
    if not os.path.exists(so):
        if os.path.isdir("/root/reference/agario"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
        else:
            return None, None
    lib = C.CDLL(so)
    lib.ref_pool_create.restype = C.c_void_p
    lib.ref_pool_run.restype = C.c_double
    lib.ref_pool_profile.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_uint]
    return lib, make_cfg(**WORKLOAD)


def _ref_make_pool(lib, cfg, inst):
    """M independent reference environments; the action mix / boost of the selected config; returns (pool, with_obs)"""
    cf = CONFIGS[CONFIG]
    pool = C.c_void_p(lib.ref_pool_create(C.byref(cfg), inst, 1234))
    if cf["p_feed"] >= 0 or cf["boost"]:
        lib.ref_pool_profile(pool, C.c_float(cf["p_feed"]), C.c_float(cf["p_split"]), C.c_uint(cf["boost"]))
    return pool, (2 if cf["obs"] == "ram" else 1)  # 1: forced GridObservation::add_frame per agent, 2: GoBiggerObservation per step


_REF_NOTE = ("the engines are independent and single-threaded (one reference environment per host thread at a time, working set ~100 KB), "
             "so the per-core rate does not depend on how many instances the sample holds")


def _cpu_reference_rate_inproc(seconds_target=12.0, threads=None, settle=2000):
    """Times the reference's own CPU implementation of the path on a bounded sample of the workload."""
    lib, cfg = _ref_pool()
    if lib is None:
        return None
    threads = threads or host_cores()
    inst = 2 * threads
    pool, wo = _ref_make_pool(lib, cfg, inst)
    if settle > 0:
        lib.ref_pool_run(pool, threads, settle, wo)  # untimed: the same game age as the GPU arm measures at
    sec = lib.ref_pool_run(pool, threads, 20, wo)  # calibration
    rate = inst * 20 / sec
    steps = max(5, int(seconds_target * rate / inst))
    sec = lib.ref_pool_run(pool, threads, steps, wo)
    lib.ref_pool_destroy(pool)
    return dict(value=inst * steps / sec, unit=UNIT, cores=threads, kind="reference",
                sample=f"{inst} instances settled for {settle} env-steps, then {steps} timed env-steps each (observation built every "
                       f"step), one reference engine per thread on {threads} host threads, {sec:.1f} s; {_REF_NOTE}")


def _reference_arm_inproc(steps, warmup, gpus, threads=None, settle=2000):
    lib, cfg = _ref_pool()
    if lib is None:
        return {"impl": "reference", "unavailable": "oracle/_ref/libagarcl_ref.so missing and /root/reference absent"}
    threads = threads or host_cores()
    inst = 2 * threads
    pool, wo = _ref_make_pool(lib, cfg, inst)
    if settle > 0:
        lib.ref_pool_run(pool, threads, settle, wo)  # untimed: the same game age as the GPU arm measures at
    sec = lib.ref_pool_run(pool, threads, 10, wo)
    rate = inst * 10 / sec
    per_step = max(2, int(0.5 * rate / inst))  # env-steps per instance in one bench "step" (about 0.5 s)
    per_step = min(per_step, max(2, int(150.0 * rate / inst / max(1, steps + warmup))))  # whole run within a few minutes
    for _ in range(warmup):
        lib.ref_pool_run(pool, threads, per_step, wo)
    t = 0.0
    for _ in range(steps):
        t += lib.ref_pool_run(pool, threads, per_step, wo)
    lib.ref_pool_destroy(pool)
    value = inst * per_step * steps / t
    sample = (f"{inst} instances settled for {settle} env-steps; each step = {inst} instances x {per_step} env-steps (observation built "
              f"every step), one reference engine per thread on {threads} host threads (plain std::thread workers over the unmodified engine: "
              f"the reference's own utils/thread-pool can hang); {_REF_NOTE}")
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * t / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "name": CONFIG, "settle_steps": settle, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def _ref_child(what, steps, warmup, gpus, threads, budget_s, settle=2000):
    """The reference engine is third-party code with process-wide globals (Ball::global_id, libc rand()): it runs in a
    CHILD process under a watchdog so that a hang or crash in it can never take the bench line down; one retry on
    half the threads."""
    last = "no attempt"
    for attempt in range(2):
        cmd = [sys.executable, os.path.abspath(__file__), "--_refchild", what, "--steps", str(steps), "--warmup", str(warmup),
               "--gpus", str(gpus), "--_threads", str(threads), "--settle", str(settle), "--config", CONFIG]
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=budget_s)
            lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
            if out.returncode == 0 and lines:
                return json.loads(lines[-1])
            last = f"child rc={out.returncode}: {out.stderr.strip()[-200:]}"
        except subprocess.TimeoutExpired:
            last = f"child exceeded {budget_s}s on {threads} threads"
        threads = max(1, threads // 2)
    return {"failed": last}


def cpu_reference_rate(settle):
    r = _ref_child("cpu_baseline", 0, 0, 1, host_cores(), 180, settle)
    if r is None or "failed" in r:
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {r and r['failed']}"}
    return r


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    budget = 240 + 2 * (args.steps + args.warmup + 3)  # settling + steps of at most about 0.5 s
    line = _ref_child("arm", args.steps, args.warmup, args.gpus, host_cores(), budget, args.settle)
    if "failed" in line:
        line = {"impl": "reference", "unavailable": line["failed"]}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from agarcl_b200 import make_cfg
    from agarcl_b200.batch import Batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: agarcl_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries the one JSON line and nothing else: NCCL's version banner (printed when the communicator is
        # created) goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    if world > 1:
        # the mirror's host threads: every rank gets its own contiguous slice of the box's cores and is BOUND to it before it
        # allocates anything, so that its patch threads do not migrate and its pinned mirror (first touch) sits on the slice's NUMA node
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            mine = cores[local_rank * per:(local_rank + 1) * per] or cores
            if os.environ.get("AGARCL_BIND_CORES", "1") != "0":
                os.sched_setaffinity(0, mine)
            os.environ.setdefault("AGARCL_HOST_THREADS", str(max(2, len(mine))))
        except (AttributeError, OSError):
            os.environ.setdefault("AGARCL_HOST_THREADS", str(max(2, host_cores() // world)))
    N = args.instances
    cf = CONFIGS[CONFIG]
    ram_mode = cf["obs"] == "ram"
    cfg = make_cfg(n_instances=N, device=local_rank, instance_base=rank * N, ram_obs=(2 if ram_mode else 0), **WORKLOAD)
    b = Batch(cfg)
    A = b.A
    b.seed(np.arange(N, dtype=np.uint64) + np.uint64(rank * N + 1))
    b.reset()
    if cf["boost"]:  # every agent's cell raised to this mass after reset (forces splits, merges, enemy eats)
        for i in range(N):
            sv = b.download_state(i)
            for a in range(A):
                sv.cells[a][0]["mass"] = cf["boost"]
            b.upload_state(i, sv)
    stream = torch.cuda.current_stream().cuda_stream
    K, W_ = args.steps, args.warmup
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    # per-step synthetic actions resident in HBM: dx,dy ~ U(-1,1), a ~ U{0,1,2}  (bench/go_bigger_example.py:100-103), or the config's mix
    n_act = 16
    dxdy = (torch.rand((n_act, N * A, 2), device="cuda", generator=gen) * 2 - 1).float().contiguous()
    if cf["p_feed"] < 0:
        act = torch.randint(0, 3, (n_act, N * A), device="cuda", generator=gen, dtype=torch.int32).contiguous()
    else:
        u = torch.rand((n_act, N * A), device="cuda", generator=gen)
        act = torch.where(u < cf["p_feed"], 1, torch.where(u < cf["p_feed"] + cf["p_split"], 2, 0)).to(torch.int32).contiguous()

    def one_step(i):
        b.set_actions_device(dxdy[i % n_act].data_ptr(), act[i % n_act].data_ptr(), stream)
        b.step(stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # let the games reach their steady state (see the module docstring); on the way, time the young games (age 100)
    age_profile = []
    young_at = 100
    for i in range(min(args.settle, young_at)):
        one_step(i)
    if args.settle > young_at:
        for i in range(W_):
            one_step(i)
        barrier()
        y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        y0.record()
        for i in range(K):
            one_step(i)
        y1.record()
        barrier()
        age_profile.append({"age_env_steps": young_at + W_, "ms_per_step": y0.elapsed_time(y1) / K})
        for i in range(max(0, args.settle - young_at - W_ - K)):
            one_step(i)
    for i in range(W_):
        one_step(i)
    barrier()
    sampler = make_sampler(local_rank)
    if rank == 0:
        sampler.start()
    b.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        one_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    sim_ms, obs_ms, tsteps = b.get_timing()
    b.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    launches = K * b.launches_per_step()

    # ---- e2e through the host-buffer C-ABI call (pinned host memory)
    Ke = max(3, min(K, args.e2e_steps))
    h_dxdy = torch.empty((N * A, 2), dtype=torch.float32, pin_memory=True)
    h_act = torch.empty((N * A,), dtype=torch.int32, pin_memory=True)
    h_rew = torch.empty((N * A,), dtype=torch.float64, pin_memory=True)
    h_done = torch.empty((N * A,), dtype=torch.uint8, pin_memory=True)
    rng = np.random.default_rng(7 + rank)
    h_dxdy.numpy()[:] = rng.uniform(-1, 1, size=(N * A, 2)).astype(np.float32)
    if cf["p_feed"] < 0:
        h_act.numpy()[:] = rng.integers(0, 3, size=N * A).astype(np.int32)
    else:
        uu = rng.random(N * A)
        h_act.numpy()[:] = np.where(uu < cf["p_feed"], 1, np.where(uu < cf["p_feed"] + cf["p_split"], 2, 0)).astype(np.int32)
    vp = C.c_void_p

    def time_host(fn, n):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    h2d = N * A * (2 * 4 + 4)
    dense_s, d2h_dense, mirror_ok, mstats, e2e_kernel_ms, e2e_ksteps, lists_s, lists_ok = None, None, None, None, 0.0, 0, None, None
    if ram_mode:
        # agario-ram-v0 hands every AGENT its record (gym_env.BatchedAgarioEnv._obs): pinned host actions in; the agents' records,
        # rewards and dones out, through the vector-env calls a user makes
        ram_t = b.ram_tensor()
        h_ram = torch.empty((N, A, ram_t.shape[2]), dtype=torch.float32, pin_memory=True)
        d_dxdy = torch.empty((N * A, 2), dtype=torch.float32, device="cuda")
        d_act = torch.empty((N * A,), dtype=torch.int32, device="cuda")
        rew_t, done_t = b.rewards_tensor(), b.dones_tensor()

        def host_step_ram():
            d_dxdy.copy_(h_dxdy, non_blocking=True)
            d_act.copy_(h_act, non_blocking=True)
            b.set_actions_device(d_dxdy.data_ptr(), d_act.data_ptr(), stream)
            b.step(stream)
            h_ram.copy_(ram_t[:, :A, :], non_blocking=True)
            h_rew.copy_(rew_t, non_blocking=True)
            h_done.copy_(done_t, non_blocking=True)
            torch.cuda.synchronize()

        Km = max(Ke, min(K, 50))
        e2e_s = time_host(host_step_ram, Km)
        d2h = h_ram.numel() * 4 + N * A * (8 + 1)
        Ke = Km
    else:
        h_obs = torch.empty(b.obs_shape, dtype=torch.int32, pin_memory=True)

        def host_step_dense():
            from agarcl_b200 import _lib
            _lib.check(_lib.lib().agarcl_batch_step_host(b._h, vp(h_dxdy.data_ptr()), vp(h_act.data_ptr()), vp(h_obs.data_ptr()),
                                                         vp(h_rew.data_ptr()), vp(h_done.data_ptr())))

        def host_step_mirror():
            from agarcl_b200 import _lib
            _lib.check(_lib.lib().agarcl_batch_step_mirror(b._h, vp(h_dxdy.data_ptr()), vp(h_act.data_ptr()),
                                                           vp(h_rew.data_ptr()), vp(h_done.data_ptr())))

        # (a) dense copy: the whole int32 observation crosses PCIe every step (kept for comparison)
        Kd = max(3, min(Ke, 5))
        dense_s = time_host(host_step_dense, Kd) / Kd
        d2h_dense = int(np.prod(b.obs_shape)) * 4 + N * A * (8 + 1)
        del h_obs
        # (b) host-resident mirror: same result in host memory (the full dense int32 tensor, kept identical to the device's by
        # moving the out-of-bounds masks + the non-zero lists and patching the mirror in place) -- this is `e2e`
        mirror = b.mirror()
        Km = max(Ke, min(K, 50))
        b.set_timing(True)
        e2e_s = time_host(host_step_mirror, Km)
        e2e_kernel_ms, _, e2e_ksteps = b.get_timing()
        b.set_timing(False)
        mstats = b.mirror_stats()
        d2h = mstats["d2h_bytes"] + N * A * (8 + 1)
        # (c) the observation as LISTS (agarcl_batch_step_lists): the same call without the host-side replay -- what the kernel wrote
        # into pinned host memory is handed to the caller as it is (include/agarcl_b200.h, agarcl_obs_lists)
        from agarcl_b200.batch import _ObsListsStruct
        from agarcl_b200 import _lib as _l
        ol = _ObsListsStruct()

        def host_step_lists():
            _l.check(_l.lib().agarcl_batch_step_lists(b._h, vp(h_dxdy.data_ptr()), vp(h_act.data_ptr()), vp(h_rew.data_ptr()),
                                                      vp(h_done.data_ptr()), C.byref(ol)))

        lists_s = time_host(host_step_lists, Km)
        lists_ok = True
        obs_dev = b.obs_tensor()
        dense_img = np.empty(b.obs_shape[1:], np.int32)
        for i in range(0, N * A, max(1, (N * A) // 64)):  # the C decoder on a spread of images against the device tensor
            _l.check(_l.lib().agarcl_batch_lists_expand(b._h, i, vp(dense_img.ctypes.data)))
            lists_ok = lists_ok and bool(torch.equal(torch.from_numpy(dense_img).cuda(), obs_dev[i]))
        host_step_mirror()  # (the dense mirror is compared below: bring it up to date with the state the lists steps left)
        # the WHOLE mirror against the device tensor, in slices (not a sample)
        obs_dev = b.obs_tensor()
        mirror_ok = True
        for i0 in range(0, N * A, 512):
            mirror_ok = mirror_ok and bool(torch.equal(torch.from_numpy(np.array(mirror[i0:i0 + 512])).cuda(), obs_dev[i0:i0 + 512]))
        Ke = Km
    # hdr.flags of ALL instances, reduced on the device (agarcl_batch_flags): a set bit = a fixed capacity was hit in that many
    # instances since their reset
    flags_seen, flag_counts = b.flags()

    # ---- the opt-in int16 observation (half the observation bytes; SURVEY 7 hard part 4): same workload, same game age, the
    # fused single-kernel step with 16-bit cells, timed on the device like `value`
    int16_profile = None
    if not ram_mode and not args.no_int16:
        from agarcl_b200 import OBS_I16
        cfg16 = make_cfg(n_instances=N, device=local_rank, instance_base=rank * N, obs_dtype=OBS_I16, **WORKLOAD)
        b16 = Batch(cfg16)
        b16.seed(np.arange(N, dtype=np.uint64) + np.uint64(rank * N + 1))
        b16.reset()
        if cf["boost"]:
            for i in range(N):
                sv16 = b16.download_state(i)
                for a in range(A):
                    sv16.cells[a][0]["mass"] = cf["boost"]
                b16.upload_state(i, sv16)

        def step16(i):
            b16.set_actions_device(dxdy[i % n_act].data_ptr(), act[i % n_act].data_ptr(), stream)
            b16.step(stream)

        young16 = None
        done16 = 0
        if args.settle > young_at:  # the young-game figure on the way, like the int32 arm's age_profile
            for i in range(young_at + W_):
                step16(i)
            barrier()
            j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            j0.record()
            for i in range(K):
                step16(i)
            j1.record()
            barrier()
            young16 = j0.elapsed_time(j1) / K
            done16 = young_at + W_ + K
        for i in range(max(0, args.settle - done16) + W_):
            step16(i)
        barrier()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        for i in range(K):
            step16(i)
        i1.record()
        barrier()
        ms16 = i0.elapsed_time(i1) / K
        l16 = b16.launches_per_step()
        b16.close()
        int16_profile = {"ms_per_step": ms16, "launches_per_step": l16, "young_ms_per_step": young16}

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, e2e_s, sim_ms, obs_ms, dense_s or 0.0, lists_s or 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, sim_ms, obs_ms, dense_max, lists_max = [float(x) for x in t.tolist()]
        dense_s = dense_max if dense_s is not None else None
        lists_s = lists_max if lists_s is not None else None
        if int16_profile is not None:
            t16 = torch.tensor([int16_profile["ms_per_step"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(t16, op=dist.ReduceOp.MAX)
            int16_profile["ms_per_step"] = float(t16.item())
        fl = torch.tensor([flags_seen], device="cuda", dtype=torch.int64)
        gathered = [torch.zeros_like(fl) for _ in range(world)]
        dist.all_gather(gathered, fl)
        for g in gathered:
            flags_seen |= int(g.item())
    if rank == 0:
        peak, peak_src = measured_peaks()
        sv = b.download_state(0)
        n_cell = int(sv.players["n_cells"].sum())
        ab = algorithmic_bytes(n_pel=int(sv.hdr["n_pellets"]), n_vir=int(sv.hdr["n_viruses"]), n_food=int(sv.hdr["n_foods"]),
                               n_cell=n_cell, P=b.layout.P, A=A, ram=ram_mode)
        sim_avg, obs_avg = sim_ms / max(tsteps, 1), obs_ms / max(tsteps, 1)
        fuse = int(os.environ.get("AGARCL_FUSE_CLEAR", "2"))
        obs_b = A * 8 * 128 * 128 * 4
        n_order = 1 if (int(os.environ.get("AGARCL_SORT_SCHEDULE", "1")) and int(os.environ.get("AGARCL_TICK_BARRIER", "2"))) else 0
        single = (launches == K * (1 + n_order))  # k_step (ticks + the whole observation) [+ k_order, the tiny schedule sort]
        if ram_mode:           # k_step (ticks only) + k_ram (one record per player)
            kern = {"k_step": (sim_avg, ab["sim_kernel"] * N), "k_ram": (obs_avg, ab["obs_kernel"] * N)}
            single = False
        elif single:
            kern = {"k_step": (sim_avg, ab["step"] * N)}
        elif fuse == 1:        # k_step also streams channels 1..7, k_obs writes channel 0 + scatter
            kern = {"k_step": (sim_avg, (ab["sim_kernel"] + obs_b * 7 // 8) * N),
                    "k_obs": (obs_avg, (obs_b // 8 + ab["s_state"]) * N)}
        else:
            kern = {"k_step": (sim_avg, ab["sim_kernel"] * N), "k_obs": (obs_avg, ab["obs_kernel"] * N)}
        dom = max(kern, key=lambda k: kern[k][0])
        traffic = measured_traffic(dom if single else None)
        def rf(k):
            t_ms, byt = kern[k]
            ach = byt / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
            return {"kernel": k, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "avg_launch_ms": t_ms, "algorithmic_bytes_per_launch": byt, "traffic": traffic if k == dom else None,
                    "peak_source": peak_src}
        value = world * N * K / (ms * 1e-3)
        for a in age_profile:
            a["value"] = world * N / (a["ms_per_step"] * 1e-3)
        age_profile.append({"age_env_steps": args.settle, "ms_per_step": ms / K, "value": value})
        whole = ab["step"] * value / 1e9 / world
        int16_line = None
        if int16_profile is not None:
            ab16 = algorithmic_bytes(n_pel=int(sv.hdr["n_pellets"]), n_vir=int(sv.hdr["n_viruses"]), n_food=int(sv.hdr["n_foods"]),
                                     n_cell=n_cell, P=b.layout.P, A=A, s_obs=2)
            v16 = world * N / (int16_profile["ms_per_step"] * 1e-3)
            int16_line = {"obs_dtype": "int16 (opt-in, saturating at 32767; the reference's is int32)", "age_env_steps": args.settle,
                          "ms_per_step": int16_profile["ms_per_step"], "value": v16, "unit": UNIT,
                          "launches_per_step": int16_profile["launches_per_step"],
                          "algorithmic_bytes_per_env_step": ab16["step"],
                          "roofline_frac": ab16["step"] * v16 / world / 1e9 / peak}
            if int16_profile["young_ms_per_step"]:
                vy = world * N / (int16_profile["young_ms_per_step"] * 1e-3)
                int16_line["young_games"] = {"age_env_steps": young_at + W_, "ms_per_step": int16_profile["young_ms_per_step"], "value": vy,
                                             "roofline_frac": ab16["step"] * vy / world / 1e9 / peak}
        if ram_mode:
            e2e_obj = {"value": world * N * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                       "note": "vector-env calls a user makes (BatchedGridEnvironment.step with ram_obs): pinned host actions copied in; every "
                               "agent's structured observation record [N, A, 1224] float32, rewards and dones copied out to pinned host memory"}
        else:
            e2e_obj = {"value": world * N * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "steps": Ke,
                       "note": "agarcl_batch_step_mirror: pinned host actions in; rewards + dones out; the dense int32 observation "
                               "[N*A,8,128,128] is left in the library-owned pinned HOST mirror, kept identical to the device tensor by "
                               "copying only the out-of-bounds masks and non-zero lists (listed by k_step itself while it scatters, fetched chunk by chunk "
                               f"while the kernel runs) and patching the mirror on {mstats['host_threads']} host threads; d2h_bytes_per_step is "
                               "what crossed PCIe in the last step",
                       "mirror_equals_device": mirror_ok, "mirror_images_compared": N * A, "mirror_entries_per_step": mstats["entries"],
                       "mirror_last_step_us": {"stage_and_launch": mstats["launch_us"], "device_wait": mstats["wait_us"],
                                               "collect_total": mstats["total_us"], "whole_call": mstats["call_us"]},
                       "k_step_ms_in_e2e": e2e_kernel_ms / max(e2e_ksteps, 1),
                       "mirror_dense_images": mstats["dense_images"],
                       "dense_copy": {"value": world * N / dense_s, "unit": UNIT, "d2h_bytes_per_step": d2h_dense,
                                      "note": "agarcl_batch_step_host: the whole int32 observation copied D2H every step (PCIe bound)"},
                       "lists": {"value": world * N * Ke / lists_s, "unit": UNIT, "decodes_to_device_observation": lists_ok,
                                 "note": "agarcl_batch_step_lists: pinned host actions in; rewards + dones out; the observation handed over as the "
                                         "lists the kernel wrote into pinned host memory (row / column masks + integer operations per image, "
                                         "include/agarcl_b200.h agarcl_obs_lists) instead of being replayed into a dense host tensor -- no host "
                                         "threads, so it scales with the GPU count where the dense mirror is bound by the box's cores"}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD_NAME, "name": CONFIG, "instances_per_gpu": N, "settle_steps": args.settle,
                           "cache": (f"working set per step (state {N * b.layout.stride / 1e6:.0f} MB + observation "
                                     f"{(N * b.layout.P * 4896 if ram_mode else N * obs_b) / 1e6:.0f} MB) "
                                     + ("larger than the 126 MB L2; no flush needed" if N * b.layout.stride > 126e6 or not ram_mode and N * obs_b > 126e6
                                        else "FITS the 126 MB L2: a latency-bound configuration, its roofline fraction is not a bandwidth statement")),
                           "game_age": f"instances settled for {args.settle} env-steps before the timed region (both arms)",
                           "rng": "philox4x32-10 per instance", "state_flags_seen": flags_seen,
                           "state_flag_instances": flag_counts, "instances_checked_for_flags": N},
                "age_profile": age_profile,
                "int16_profile": int16_line,
                "roofline": rf(dom),
                "roofline_all": {"kernels": [rf(k) for k in kern],
                                 "whole_step": {"achieved": whole, "peak": peak, "unit": "GB/s", "frac": whole / peak,
                                                "algorithmic_bytes_per_env_step": ab["step"]}},
                "e2e": e2e_obj,
                "gpu_launches": launches, "clocks": clocks}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference_rate(args.settle)  # child process under a watchdog: never takes the bench down
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    b.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configs[0..4]; c2 is the one the metric is quoted on")
    ap.add_argument("--instances", type=int, default=None, help="instances per GPU (default: the config's)")
    ap.add_argument("--settle", type=int, default=None, help="untimed env-steps before the timed region: game age (both arms; default: the config's)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-int16", action="store_true", help="skip the second measurement with the opt-in int16 observation")
    ap.add_argument("--tps", type=int, default=None, help="diagnostic only: ticks per env-step (the workload's is 4)")
    ap.add_argument("--_refchild", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--_threads", type=int, default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    select_config(args.config)
    if args.instances is None:
        args.instances = INSTANCES_PER_GPU
    if args.settle is None:
        args.settle = CONFIGS[CONFIG]["settle"]
    if args._refchild == "cpu_baseline":
        print(json.dumps(_cpu_reference_rate_inproc(threads=args._threads, settle=args.settle)))
        return 0
    if args._refchild == "arm":
        print(json.dumps(_reference_arm_inproc(args.steps, args.warmup, args.gpus, threads=args._threads, settle=args.settle)))
        return 0
    if args.warmup < 3:
        args.warmup = 3
    if args.tps is not None:
        WORKLOAD["ticks_per_step"] = args.tps
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
