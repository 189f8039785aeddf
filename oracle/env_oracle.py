"""TEST INFRASTRUCTURE ONLY — CPU restatement of the environment dynamics the device environments (csrc/envs.cu) implement.

Not part of the product: only tests/ may import this module.

The dynamics are not in the reference tree: diamond-ppo steps third-party Gymnasium environments (`gymnasium>=1.0.0`,
pyproject.toml:38-44, unpinned; call sites diamond/ppo.py:124-130 `SyncVectorEnv(..., autoreset_mode=DISABLED)`, :163
`envs.step`, :174-179 masked reset).  gymnasium is not installed in this image, so the published algorithms are restated:

* CartPole-v1 — gymnasium/envs/classic_control/cartpole.py `CartPoleEnv.step` (Barto, Sutton & Anderson 1983): Euler
  integrator, gravity 9.8, masscart 1.0, masspole 0.1, half-length 0.5, force 10, tau 0.02; terminated when |x| > 2.4 or
  |theta| > 12 deg; reward 1 per step; TimeLimit 500 (registration of CartPole-v1) -> truncated; reset state ~ U(-0.05, 0.05)^4;
  float64 state, float32 observation.
* Pendulum-v1 — gymnasium/envs/classic_control/pendulum.py `PendulumEnv.step`: dt 0.05, g 10, m 1, l 1, torque clipped to
  [-2, 2], cost = angle_normalize(th)^2 + 0.1 thdot^2 + 0.001 u^2, thdot clipped to [-8, 8], obs (cos th, sin th, thdot);
  never terminates; TimeLimit 200; reset th ~ U(-pi, pi), thdot ~ U(-1, 1).

parity unpinned by golden vectors of gymnasium itself (absent); pinned against the independent per-env host implementation in
diamond/envs.py (tests/test_oracle.py) and against hand-computed single steps."""
import numpy as np


def cartpole_step(state: np.ndarray, actions: np.ndarray, steps: np.ndarray):
    """state [N,4] float64, actions [N] int, steps [N] int (taken so far) -> next_state, obs f32, reward, terminated, truncated."""
    x, x_dot, th, th_dot = (state[:, i].astype(np.float64) for i in range(4))
    gravity, masscart, masspole, length, force_mag, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    total_mass, pml = masspole + masscart, masspole * length
    force = np.where(np.asarray(actions) == 1, force_mag, -force_mag)
    ct, st = np.cos(th), np.sin(th)
    temp = (force + pml * th_dot ** 2 * st) / total_mass
    th_acc = (gravity * st - ct * temp) / (length * (4.0 / 3.0 - masspole * ct ** 2 / total_mass))
    x_acc = temp - pml * th_acc * ct / total_mass
    nxt = np.stack([x + tau * x_dot, x_dot + tau * x_acc, th + tau * th_dot, th_dot + tau * th_acc], axis=1)
    terminated = (np.abs(nxt[:, 0]) > 2.4) | (np.abs(nxt[:, 2]) > 12 * 2 * np.pi / 360)
    truncated = ((np.asarray(steps) + 1) >= 500) & ~terminated
    return nxt, nxt.astype(np.float32), np.ones(len(nxt)), terminated, truncated


def pendulum_step(state: np.ndarray, actions: np.ndarray, steps: np.ndarray):
    """state [N,2+] float64 (theta, theta_dot), actions [N,1] float -> next_state [N,2], obs f32 [N,3], reward, terminated, truncated."""
    th, thdot = state[:, 0].astype(np.float64), state[:, 1].astype(np.float64)
    u = np.clip(np.asarray(actions, dtype=np.float32).reshape(len(th), -1)[:, 0].astype(np.float64), -2.0, 2.0)
    thn = ((th + np.pi) % (2 * np.pi)) - np.pi
    cost = thn ** 2 + 0.1 * thdot ** 2 + 0.001 * u ** 2
    new_thdot = np.clip(thdot + (3 * 10.0 / (2 * 1.0) * np.sin(th) + 3.0 / (1.0 * 1.0 ** 2) * u) * 0.05, -8.0, 8.0)
    new_th = th + new_thdot * 0.05
    nxt = np.stack([new_th, new_thdot], axis=1)
    obs = np.stack([np.cos(new_th), np.sin(new_th), new_thdot], axis=1).astype(np.float32)
    return nxt, obs, -cost, np.zeros(len(th), bool), (np.asarray(steps) + 1) >= 200
