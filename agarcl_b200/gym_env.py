"""Host-side mirror of the reference's gym wrapper (gym_agario/AgarioEnv.py, gym_agario/__init__.py).

`AgarioEnv` keeps the reference's constructor keywords, action format, return tuples and ids
(`agario-grid-v0`, `agario-gobigger-v0`; `agario-ram-v0` is the flat structured observation the reference's
disabled ram test expects, tests/ram_env_test.py:11,67-92) over a size-1 batch; `BatchedAgarioEnv` is the
vector form over N lockstep instances with device tensors and auto-reset.  gymnasium is used when it is
installed and replaced by minimal stand-ins when it is not (it is absent from the build image).

Deliberate differences from the shipped wrapper, which cannot run as written:
  * `_make_environment` references an undefined `args` and passes 11 positional arguments to a 10-argument
    binding (AgarioEnv.py:213-227): here the 10 documented arguments are passed;
  * `kwargs | grid_defaults` lets the defaults override the caller (AgarioEnv.py:226): here the caller wins;
  * the action noise is computed and then discarded (AgarioEnv.py:282-296): no noise is applied here either.
"""
import numpy as np

from .env import BatchedGridEnvironment, GoBiggerEnvironment, GridEnvironment

try:  # pragma: no cover - gymnasium is optional
    import gymnasium as gym
    from gymnasium import spaces
    _EnvBase = gym.Env
except Exception:  # minimal stand-ins with the attributes the wrapper and its users touch
    gym = None

    class _Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape) if shape is not None else (), np.dtype(dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

        def sample(self):
            return np.random.uniform(self.low, self.high, size=self.shape).astype(self.dtype)

    class _Discrete:
        def __init__(self, n):
            self.n = n

        def contains(self, x):
            return int(x) == x and 0 <= int(x) < self.n

        def sample(self):
            return int(np.random.randint(self.n))

    class _Tuple(tuple):
        def __new__(cls, items):
            return super().__new__(cls, items)

        def sample(self):
            return tuple(s.sample() for s in self)

    class spaces:  # noqa: N801
        Box, Discrete, Tuple = _Box, _Discrete, _Tuple

    class _EnvBase:
        pass


def _env_args(kwargs):
    """AgarioEnv._get_env_args (AgarioEnv.py:298-363): difficulty presets, overridable one by one"""
    difficulty = kwargs.get("difficulty", "normal").lower()
    if difficulty not in ("normal", "empty", "trivial"):
        raise ValueError(f"Unrecognized difficulty: {difficulty}")
    d = dict(num_agents=1, ticks_per_step=4, arena_size=1000, pellet_regen=True, num_pellets=1000, num_viruses=0, num_bots=0,
             reward_type=1, c_death=0, mode=0)
    if difficulty == "trivial":
        d.update(arena_size=50, num_pellets=200)
    for k in d:
        if k in kwargs:
            d[k] = kwargs[k]
    if type(d["ticks_per_step"]) is not int or d["ticks_per_step"] <= 0:
        raise ValueError("ticks_per_step must be a positive integer")
    return d


_OBS_KEYS = ("num_frames", "grid_size", "observe_cells", "observe_others", "observe_viruses", "observe_pellets")


class AgarioEnv(_EnvBase):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 60}

    def __init__(self, obs_type="grid", render_mode=None, **kwargs):
        if obs_type not in ("ram", "screen", "grid", "gobigger"):
            raise ValueError(obs_type)
        if obs_type == "screen":
            raise ValueError("agarcl_b200 does not include ScreenEnvironment (OpenGL, out of scope)")
        a = _env_args(kwargs)
        self.num_agents = a["num_agents"]
        self.multi_agent = bool(kwargs.get("multi_agent", False)) or self.num_agents > 1
        self.obs_type = obs_type
        self.render_mode = render_mode
        self.number_of_steps = kwargs.get("number_steps", 500)
        self.mode = a["mode"]
        self.env_type = kwargs.get("env_type", 0)  # 0 episodic, 1 continuing
        self.add_noise = kwargs.get("add_noise", True)
        self.steps = None
        self._seed = None
        pos = (a["num_agents"], a["ticks_per_step"], a["arena_size"], a["pellet_regen"], a["num_pellets"], a["num_viruses"],
               a["num_bots"], a["reward_type"], a["c_death"], a["mode"])
        extra = {k: kwargs[k] for k in ("device", "rng_mode") if k in kwargs}
        if obs_type == "grid":
            self._env = GridEnvironment(*pos, **extra)
            self._env.configure_observation({k: kwargs[k] for k in _OBS_KEYS if k in kwargs})
            c, w, h = self._env.observation_shape()
            self.observation_space = spaces.Box(-1, np.iinfo(np.int32).max, (w, h, c), dtype=np.int32)
        else:
            self._env = GoBiggerEnvironment(kwargs.get("map_width", 512), kwargs.get("map_height", 512),
                                            kwargs.get("frame_limit", 1000), *pos, **extra)
            self._env.configure_observation({k: kwargs[k] for k in _OBS_KEYS if k in kwargs})
            from ._abi import RAM_RECORD
            shape = (RAM_RECORD,) if obs_type == "ram" else self._env.observation_shape()
            self.observation_space = spaces.Box(-1e4 if obs_type == "ram" else 0, 1e6 if obs_type == "ram" else 255, shape,
                                                dtype=np.float32)
        self.action_space = spaces.Tuple((spaces.Box(low=-1, high=1, shape=(2,)), spaces.Discrete(3)))

    # ---- gym API (AgarioEnv.py:77-132)
    def step(self, actions):
        assert self.steps is not None, "Cannot call step() before calling reset()"
        actions = self._sanitize_actions(actions)
        self._env.take_actions(actions)
        rewards = self._env.step()
        assert len(rewards) == self.num_agents
        self.observations = self._make_observations()
        dones = self._env.dones()
        truncations = [False] * len(dones)
        if self.steps >= self.number_of_steps and self.env_type == 0:
            dones = [True] * len(dones)
        if not self.multi_agent:
            self.observations, rewards, dones, truncations = self.observations[0], rewards[0], dones[0], truncations[0]
        self.steps += 1
        return self.observations, rewards, dones, truncations, {"steps": self.steps, "untransformed_rewards": rewards}

    def reset(self, **kwargs):
        self.steps = 0
        self._env.reset()
        obs = self._make_observations()
        return (obs if self.multi_agent else obs[0]), {}

    def seed(self, seed=None):
        if seed is not None:
            self._seed = seed
            self._env.seed(seed)
            return [self._seed]

    def render(self):
        raise RuntimeError("OpenGL rendering is out of scope of agarcl_b200")

    def close(self):
        self._env.close()

    def save_env_state(self, filename):
        self._env.save_env_state(filename)

    def load_env_state(self, filename):
        self._env.load_env_state(filename)

    # ---- helpers
    def _make_observations(self):
        if self.obs_type == "grid":
            return [np.transpose(s, [1, 2, 0]) for s in self._env.get_state()]  # NCHW -> HWC view (AgarioEnv.py:192-194)
        if self.obs_type == "ram":
            ram = self._env.ram()
            return [ram[a].copy() for a in range(self.num_agents)]
        states = self._env.get_state()
        return states * self.num_agents if len(states) == 1 and self.num_agents > 1 else states

    def _sanitize_actions(self, actions):
        if not self.multi_agent and type(actions) is not list:
            actions = [actions]
        if type(actions) is not list:
            raise ValueError("Action list must be a list of two-element tuples")
        if len(actions) != self.num_agents:
            raise ValueError(f"Number of actions {len(actions)} does not match number of agents {self.num_agents}")
        out = []
        for tgt, a in actions:
            tgt = np.asarray(tgt, dtype=np.float32)
            if not (self.action_space[0].contains(tgt) and self.action_space[1].contains(a)):
                raise ValueError(f"action {(tgt, a)} not in action space")
            out.append((float(tgt[0]), float(tgt[1]), int(a)))
        return out


class BatchedAgarioEnv:
    """Vector environment: N lockstep instances on one GPU, device tensors in and out, auto-reset.

    step(dxdy [N*A, 2] float32, act [N*A] int32) -> obs, rewards (float32), dones, truncations (bool tensors), info.
    An instance whose agent 0 is done, or that reached `number_steps` (episodic), is reset before the next step and
    its observation is the first of the new episode (the final one is in info["final_obs_mask"] semantics of vector
    envs is not reproduced: rewards/dones of the finished step are returned as they were)."""

    def __init__(self, n_envs, obs_type="grid", auto_reset=True, **kwargs):
        if obs_type not in ("grid", "ram"):
            raise ValueError(obs_type)
        a = _env_args(kwargs)
        self.n_envs, self.num_agents, self.obs_type, self.auto_reset = n_envs, a["num_agents"], obs_type, auto_reset
        self.number_of_steps = kwargs.get("number_steps", 500)
        self.env_type = kwargs.get("env_type", 0)
        extra = {k: kwargs[k] for k in ("device", "rng_mode", "obs_dtype", "instance_base") if k in kwargs}
        self._env = BatchedGridEnvironment(n_envs, a["num_agents"], a["ticks_per_step"], a["arena_size"], a["pellet_regen"],
                                           a["num_pellets"], a["num_viruses"], a["num_bots"], a["reward_type"], a["c_death"],
                                           a["mode"], ram_obs=(2 if obs_type == "ram" else 0), **extra)
        self._env.configure_observation({k: kwargs[k] for k in _OBS_KEYS if k in kwargs})
        self._steps = None

    def seed(self, seed):
        self._env.seed(seed)

    def _obs(self):
        if self.obs_type == "grid":
            return self._env.batch.obs_tensor()
        ram = self._env.ram()
        return ram[:, :self.num_agents, :]

    def reset(self):
        import torch
        self._env.reset()
        self._steps = torch.zeros(self.n_envs, dtype=torch.int64, device=self._env.batch.obs_tensor().device)
        return self._obs(), {}

    def step(self, dxdy, act):
        import torch
        assert self._steps is not None, "Cannot call step() before calling reset()"
        _, rew, done = self._env.step(dxdy, act)
        A = self.num_agents
        done = done.view(self.n_envs, A).bool()
        trunc = torch.zeros_like(done)
        if self.env_type == 0:
            done = done | (self._steps >= self.number_of_steps).unsqueeze(1)
        self._steps += 1
        rewards = rew.float()
        if self.auto_reset:
            mask = done.any(dim=1)
            if bool(mask.any()):
                self._env.reset(mask.to(torch.uint8).cpu().numpy())
                self._steps[mask] = 0
        return self._obs(), rewards, done.view(-1), trunc.view(-1), {"steps": self._steps}

    def flags(self):
        """(OR over all instances, {AGARCL_FLAG name: instances}) -- fixed capacities hit since the instances' last reset"""
        return self._env.flags()

    def close(self):
        self._env.close()


_REGISTRY = {"agario-grid-v0": "grid", "agario-gobigger-v0": "gobigger", "agario-ram-v0": "ram"}


def make(env_id, **kwargs):
    """gym.make for the ids gym_agario registers (gym_agario/__init__.py:9-23)"""
    if env_id not in _REGISTRY:
        raise ValueError(f"unknown environment id {env_id}")
    return AgarioEnv(obs_type=_REGISTRY[env_id], **kwargs)


if gym is not None:  # pragma: no cover
    from gymnasium.envs.registration import register, registry
    for _id, _t in _REGISTRY.items():
        if _id not in registry:
            register(id=_id, entry_point="agarcl_b200.gym_env:AgarioEnv", kwargs={"obs_type": _t})
