"""Gymnasium-API environment shim.

The GPU image has no `gymnasium` (SURVEY.md §7 hard part 6).  The agents only need the small API
surface the reference touches (diamond/ppo.py:124-130, 163, 174-179, 291, 312): `spaces.Box/Discrete`,
an `Env` with reset/step, and a `SyncVectorEnv` with autoreset DISABLED whose `reset(options=
{"reset_mask": mask})` resets only the masked sub-envs.  If the real gymnasium is importable the
agents use it instead; these classes mirror its semantics.  CartPole-v1 and Pendulum-v1 follow the
published Gymnasium dynamics; LunarLander needs Box2D, so `SyntheticEnv(8, 4)` stands in for its
shapes (obs 8, 4 actions).
"""
from __future__ import annotations

import math
from typing import Any, Callable

import numpy as np


class Space:
    pass


class Box(Space):
    def __init__(self, low=-np.inf, high=np.inf, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()
        self.dtype = np.dtype(dtype)


class Discrete(Space):
    def __init__(self, n: int):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)


class Env:
    observation_space: Space
    action_space: Space

    def reset(self, *, seed: int | None = None, options: dict | None = None):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def close(self):
        pass


class CartPoleEnv(Env):
    """CartPole-v1 dynamics (Barto, Sutton & Anderson 1983; Gymnasium classic_control), 500-step limit."""
    gravity, masscart, masspole, length, force_mag, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    theta_limit, x_limit, max_steps = 12 * 2 * math.pi / 360, 2.4, 500

    def __init__(self):
        high = np.array([self.x_limit * 2, np.inf, self.theta_limit * 2, np.inf], dtype=np.float32)
        self.observation_space = Box(-high, high, (4,))
        self.action_space = Discrete(2)
        self.rng = np.random.default_rng()
        self.state = np.zeros(4)
        self.t = 0

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        self.state = self.rng.uniform(-0.05, 0.05, size=4)
        self.t = 0
        return self.state.astype(np.float32), {}

    def step(self, action):
        x, x_dot, th, th_dot = self.state
        force = self.force_mag if int(action) == 1 else -self.force_mag
        total_mass = self.masspole + self.masscart
        pml = self.masspole * self.length
        ct, st = math.cos(th), math.sin(th)
        temp = (force + pml * th_dot ** 2 * st) / total_mass
        th_acc = (self.gravity * st - ct * temp) / (self.length * (4.0 / 3.0 - self.masspole * ct ** 2 / total_mass))
        x_acc = temp - pml * th_acc * ct / total_mass
        self.state = np.array([x + self.tau * x_dot, x_dot + self.tau * x_acc, th + self.tau * th_dot, th_dot + self.tau * th_acc])
        self.t += 1
        terminated = bool(abs(self.state[0]) > self.x_limit or abs(self.state[2]) > self.theta_limit)
        truncated = bool(self.t >= self.max_steps and not terminated)
        return self.state.astype(np.float32), 1.0, terminated, truncated, {}


class PendulumEnv(Env):
    """Pendulum-v1 dynamics (Gymnasium classic_control), 200-step limit, torque in [-2, 2]."""
    max_speed, max_torque, dt, g, m, l, max_steps = 8.0, 2.0, 0.05, 10.0, 1.0, 1.0, 200

    def __init__(self):
        high = np.array([1.0, 1.0, self.max_speed], dtype=np.float32)
        self.observation_space = Box(-high, high, (3,))
        self.action_space = Box(-self.max_torque, self.max_torque, (1,))
        self.rng = np.random.default_rng()
        self.th, self.thdot, self.t = 0.0, 0.0, 0

    def _obs(self):
        return np.array([math.cos(self.th), math.sin(self.th), self.thdot], dtype=np.float32)

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        self.th, self.thdot = self.rng.uniform(-math.pi, math.pi), self.rng.uniform(-1.0, 1.0)
        self.t = 0
        return self._obs(), {}

    def step(self, action):
        u = float(np.clip(np.asarray(action).reshape(-1)[0], -self.max_torque, self.max_torque))
        th_n = ((self.th + math.pi) % (2 * math.pi)) - math.pi
        cost = th_n ** 2 + 0.1 * self.thdot ** 2 + 0.001 * u ** 2
        self.thdot = float(np.clip(self.thdot + (3 * self.g / (2 * self.l) * math.sin(self.th) + 3.0 / (self.m * self.l ** 2) * u) * self.dt,
                                   -self.max_speed, self.max_speed))
        self.th += self.thdot * self.dt
        self.t += 1
        return self._obs(), -cost, False, self.t >= self.max_steps, {}


class SyntheticEnv(Env):
    """Shape-only stand-in (e.g. LunarLander-v3: obs 8, 4 actions): i.i.d. normal observations,
    reward depends on (obs, action), random terminations/truncations."""

    def __init__(self, obs_dim=8, n_actions=4, continuous=False, p_term=0.01, p_trunc=0.005):
        self.observation_space = Box(-np.inf, np.inf, (obs_dim,))
        self.action_space = Box(-1.0, 1.0, (n_actions,)) if continuous else Discrete(n_actions)
        self.continuous, self.p_term, self.p_trunc = continuous, p_term, p_trunc
        self.rng = np.random.default_rng()
        self.obs = np.zeros(obs_dim, np.float32)

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        self.obs = self.rng.standard_normal(self.observation_space.shape).astype(np.float32)
        return self.obs, {}

    def step(self, action):
        a = float(np.sum(action)) if self.continuous else float(action)
        reward = float(self.obs[0]) * (1.0 if a > 0 else -1.0) * 0.1
        self.obs = self.rng.standard_normal(self.observation_space.shape).astype(np.float32)
        term = bool(self.rng.random() < self.p_term)
        trunc = bool((not term) and self.rng.random() < self.p_trunc)
        return self.obs, reward, term, trunc, {}


class SyncVectorEnv:
    """N sub-envs stepped serially; autoreset DISABLED (the caller resets done envs through
    reset(options={"reset_mask": mask}), so `step` returns the TRUE final observation, ppo.py:174-179)."""

    def __init__(self, env_fns: list[Callable[[], Env]], copy: bool = True, autoreset_mode: Any = "Disabled"):
        if str(autoreset_mode).lower().split(".")[-1] != "disabled":
            raise ValueError("this shim implements autoreset_mode='Disabled' only (what the agents use)")
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        self.single_observation_space = self.envs[0].observation_space
        self.single_action_space = self.envs[0].action_space
        shape = self.single_observation_space.shape
        self._obs = np.zeros((self.num_envs,) + tuple(shape), dtype=np.float32)

    def reset(self, *, seed: int | None = None, options: dict | None = None):
        mask = None if options is None else options.get("reset_mask")
        for i, env in enumerate(self.envs):
            if mask is None or mask[i]:
                self._obs[i], _ = env.reset(seed=None if seed is None else seed + i)
        return self._obs.copy(), {}

    def step(self, actions):
        n = self.num_envs
        rewards = np.zeros(n, dtype=np.float64)
        terms = np.zeros(n, dtype=bool)
        truncs = np.zeros(n, dtype=bool)
        for i, env in enumerate(self.envs):
            self._obs[i], rewards[i], terms[i], truncs[i], _ = env.step(actions[i])
        return self._obs.copy(), rewards, terms, truncs, {}

    def close(self):
        for e in self.envs:
            e.close()


class BatchedSyntheticVectorEnv:
    """Vectorised synthetic env for large N (config S: 4096 envs would take seconds per step through
    N Python sub-envs).  Same vector API and autoreset-disabled semantics as SyncVectorEnv."""

    def __init__(self, num_envs, obs_dim=64, n_actions=4, continuous=False, p_term=0.01, p_trunc=0.01, seed=0):
        self.num_envs = num_envs
        self.single_observation_space = Box(-np.inf, np.inf, (obs_dim,))
        self.single_action_space = Box(-1.0, 1.0, (n_actions,)) if continuous else Discrete(n_actions)
        self.continuous, self.p_term, self.p_trunc = continuous, p_term, p_trunc
        self.rng = np.random.default_rng(seed)
        self._obs = np.zeros((num_envs, obs_dim), np.float32)

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        mask = None if options is None else options.get("reset_mask")
        fresh = self.rng.standard_normal(self._obs.shape).astype(np.float32)
        self._obs = fresh if mask is None else np.where(np.asarray(mask)[:, None], fresh, self._obs)
        return self._obs.copy(), {}

    def step(self, actions):
        a = np.asarray(actions)
        a = a.reshape(self.num_envs, -1).sum(-1) if self.continuous else a
        rewards = (self._obs[:, 0] * np.where(a > 0, 1.0, -1.0) * 0.1).astype(np.float64)
        self._obs = self.rng.standard_normal(self._obs.shape).astype(np.float32)
        terms = self.rng.random(self.num_envs) < self.p_term
        truncs = (self.rng.random(self.num_envs) < self.p_trunc) & ~terms
        return self._obs.copy(), rewards, terms, truncs, {}

    def close(self):
        pass


class DeviceVectorEnv:
    """Vector environment that lives on the GPU (csrc/envs.cu: dppo_env_step / dppo_env_reset): CartPole-v1, Pendulum-v1
    (Gymnasium classic-control dynamics, float64 state) and the synthetic shape stand-ins.  Same vector API and
    autoreset-DISABLED semantics as SyncVectorEnv, so any agent or user code can drive it through `step` / `reset` with host
    arrays; the default-network agents use `step_into`, which writes the step straight into the device rollout buffer and
    resets finished environments inside the same kernel (diamond/ppo.py:160-182 without a PCIe crossing).

    Use as `env_fn = DeviceVectorEnv.factory("CartPole-v1")` (an `env_fn.vectorized` factory taking num_envs)."""

    device_resident = True
    _KINDS = {"CartPole-v1": (0, 4, 2, False), "Pendulum-v1": (1, 3, 1, True)}

    def __init__(self, env_id: str, num_envs: int, seed: int = 0, device=None, obs_dim: int = 64, n_actions: int = 4,
                 continuous: bool = False, p_term: float = 0.01, p_trunc: float = 0.01, env_offset: int = 0):
        import torch
        from . import _native as N
        self._torch, self._N = torch, N
        if env_id in self._KINDS:
            kind, D, A, cont = self._KINDS[env_id]
        elif env_id in ("Synthetic", "LunarLander-v3"):                   # Box2D unavailable: shape stand-in
            D, A, cont = (8, 4, False) if env_id == "LunarLander-v3" else (int(obs_dim), int(n_actions), bool(continuous))
            kind = 3 if cont else 2
        else:
            raise KeyError(f"unknown device env id {env_id!r}; available: {sorted(self._KINDS) + ['LunarLander-v3', 'Synthetic']}")
        self.env_id, self.num_envs, self.continuous = env_id, int(num_envs), cont
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ctx = N.get_context(self.device.index)
        if kind == 0:
            high = np.array([4.8, np.inf, 24 * math.pi / 360 * 2, np.inf], dtype=np.float32)
            self.single_observation_space, self.single_action_space = Box(-high, high, (4,)), Discrete(2)
        elif kind == 1:
            high = np.array([1.0, 1.0, 8.0], dtype=np.float32)
            self.single_observation_space, self.single_action_space = Box(-high, high, (3,)), Box(-2.0, 2.0, (1,))
        else:
            self.single_observation_space = Box(-np.inf, np.inf, (D,))
            self.single_action_space = Box(-1.0, 1.0, (A,)) if cont else Discrete(A)
        self.obs_dim, self.act_dim = D, A
        self.desc = N.EnvDesc(kind, self.num_envs, D, A, int(seed) & (2 ** 64 - 1), int(env_offset), float(p_term), float(p_trunc))
        n, dev = self.num_envs, self.device
        self.state = torch.zeros(n, 4, dtype=torch.float64, device=dev)
        self.steps = torch.zeros(n, dtype=torch.int32, device=dev)
        self.episode = torch.zeros(n, dtype=torch.int64, device=dev)
        self.ep_return = torch.zeros(n, dtype=torch.float64, device=dev)
        self.cur_obs = torch.zeros(n, D, dtype=torch.float32, device=dev)
        self._st = N.EnvState(self.state.data_ptr(), self.steps.data_ptr(), self.episode.data_ptr(), self.ep_return.data_ptr(),
                              self.cur_obs.data_ptr())
        f = dict(dtype=torch.float32, device=dev)
        self._row = dict(next_obs=torch.empty(1, n, D, **f), rew=torch.empty(1, n, **f), term=torch.empty(1, n, **f),
                         trunc=torch.empty(1, n, **f))
        self.ctx.env_reset(self.desc, self._st, None)

    @classmethod
    def factory(cls, env_id: str, **kwargs):
        def env_fn(num_envs):
            return cls(env_id, num_envs, **kwargs)
        env_fn.vectorized = True
        return env_fn

    def reset(self, *, seed: int | None = None, options: dict | None = None):
        torch = self._torch
        mask = None if options is None else options.get("reset_mask")
        if seed is not None:
            self.desc.seed = int(seed) & (2 ** 64 - 1)
            if mask is None:
                self.episode.zero_()
        m = None if mask is None else torch.as_tensor(np.asarray(mask, dtype=np.uint8) if not torch.is_tensor(mask) else mask).to(
            self.device, torch.uint8).contiguous()
        self.ctx.env_reset(self.desc, self._st, m)
        return self.cur_obs.cpu().numpy(), {}

    def _actions(self, actions):
        torch = self._torch
        a = actions if torch.is_tensor(actions) else torch.as_tensor(np.asarray(actions))
        if self.continuous:
            return a.to(self.device, torch.float32).reshape(self.num_envs, self.act_dim).contiguous()
        return a.to(self.device, torch.int64).reshape(self.num_envs).contiguous()

    def step(self, actions):
        """Gymnasium vector API with host arrays (autoreset disabled: the caller resets through reset(options=...))."""
        r = self._row
        self.ctx.env_step(self.desc, self._st, self._actions(actions), 0, False, None, r["next_obs"], None, r["rew"], r["term"],
                          r["trunc"])
        return (r["next_obs"][0].cpu().numpy(), r["rew"][0].double().cpu().numpy(), r["term"][0].bool().cpu().numpy(),
                r["trunc"][0].bool().cpu().numpy(), {})

    def step_into(self, buf, t: int, actions) -> None:
        """Fused rollout step: row t of the device rollout buffer (obs, next_obs, actions, rewards, terminations, truncations)
        is written by the kernel, finished environments are reset, `cur_obs` becomes the next step's observations."""
        self.ctx.env_step(self.desc, self._st, self._actions(actions), t, True, buf.obs, buf.next_obs, buf.actions, buf.rewards,
                          buf.terminations, buf.truncations)
        buf.filled = max(buf.filled, t + 1)

    def close(self):
        pass


_REGISTRY = {"CartPole-v1": CartPoleEnv, "Pendulum-v1": PendulumEnv,
             "LunarLander-v3": lambda: SyntheticEnv(8, 4)}       # Box2D unavailable: shape stand-in


def make(env_id: str, **kwargs) -> Env:
    """gym.make for the ids BASELINE.json's configs name."""
    if env_id not in _REGISTRY:
        raise KeyError(f"unknown env id {env_id!r}; available: {sorted(_REGISTRY)}")
    return _REGISTRY[env_id](**kwargs) if kwargs else _REGISTRY[env_id]()


class _Namespace:
    pass


# `import diamond.envs as gym` then gym.spaces.Box / gym.vector.SyncVectorEnv / gym.make work
spaces = _Namespace()
spaces.Space, spaces.Box, spaces.Discrete = Space, Box, Discrete
vector = _Namespace()
vector.SyncVectorEnv = SyncVectorEnv
