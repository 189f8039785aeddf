"""Host-side helpers the agents construct (reference: diamond/utils.py).  They are NOT on the data
path (out of scope); these are thin, plotly-free stand-ins that keep the agent
constructors and `ticker.tick(...)` calls working with the same signatures and payload formats."""
from __future__ import annotations

import time
from collections import deque
from contextlib import contextmanager
from pathlib import Path
from typing import Any

import numpy as np
import torch


class Ticker:
    """Episode statistics + console progress (reference: diamond/utils.py:20-215)."""

    def __init__(self, total_steps: int, num_envs: int, rollout_steps: int, *, window_size: int = 100,
                 print_every: int = 5, num_checkpoints: int = 20, verbose: bool = True) -> None:
        self.total_steps, self.num_envs, self.rollout_steps = total_steps, num_envs, rollout_steps
        self.window_size, self.print_every, self.verbose = window_size, print_every, verbose
        per_rollout = rollout_steps * num_envs
        iters = total_steps // per_rollout if per_rollout else 0
        self.checkpoints = (np.arange(1, num_checkpoints + 1) * iters // num_checkpoints) * per_rollout
        self.current_step = 0
        self.current_episode = 1
        self.current_returns = np.zeros(num_envs, dtype=np.float64)
        self.current_lengths = np.zeros(num_envs, dtype=np.int64)
        self.recent_returns: deque = deque(maxlen=window_size)
        self.recent_lengths: deque = deque(maxlen=window_size)
        self.custom_logs: dict[str, Any] = {}
        self.start_time = time.time()
        self._header = False

    def reset(self) -> None:
        self.__init__(self.total_steps, self.num_envs, self.rollout_steps, window_size=self.window_size,
                      print_every=self.print_every, num_checkpoints=len(self.checkpoints), verbose=self.verbose)

    def tick(self, rewards: np.ndarray, dones: np.ndarray, **custom_logs: Any) -> None:
        self.current_step += self.num_envs
        self.current_returns += rewards
        self.current_lengths += 1
        if dones.any():
            for r, l in zip(self.current_returns[dones], self.current_lengths[dones]):
                self.recent_returns.append(float(r))
                self.recent_lengths.append(int(l))
                self.current_episode += 1
            self.current_returns[dones] = 0.0
            self.current_lengths[dones] = 0
        self.custom_logs.update(custom_logs)
        if self.verbose:
            self.print_logs()

    def tick_rollout(self, num_steps: int, finished: int, returns, lengths, **custom_logs: Any) -> None:
        """`num_steps` vector steps at once, for rollouts whose bookkeeping ran on the device (dppo_episode_stats): `finished`
        episodes ended, `returns` / `lengths` are the last min(window, finished) of them in tick() order.  Leaves the same window,
        step and episode counters as `num_steps` tick() calls (the per-environment running sums stay on the device)."""
        self.current_step += self.num_envs * num_steps
        self.current_episode += int(finished)
        for r, l in zip(returns, lengths):
            self.recent_returns.append(float(r))
            self.recent_lengths.append(int(l))
        self.custom_logs.update(custom_logs)
        if self.verbose:
            self.print_logs(force=True)

    def print_logs(self, force: bool = False) -> None:
        at_checkpoint = self.current_step in self.checkpoints
        due = (force or self.current_step % (self.num_envs * self.print_every) == 0) and len(self.recent_returns) > 0
        if not (due or at_checkpoint) or not self.recent_returns:
            return
        if not self._header:
            print(f"{'Progress':>9} | {'Step':>10} | {'Episode':>8} | {'Mean Rew':>9} | {'Mean Len':>8} | {'FPS':>8} | {'Time':>8}")
            self._header = True
        elapsed = time.time() - self.start_time
        h, rem = divmod(int(elapsed), 3600)
        m, s = divmod(rem, 60)
        row = (f"{100.0 * self.current_step / max(self.total_steps, 1):8.1f}% | {self.current_step:10,d} | "
               f"{self.current_episode - 1:8,d} | {np.mean(self.recent_returns):9.2f} | {np.mean(self.recent_lengths):8.1f} | "
               f"{self.current_step / (elapsed + 1e-6):8,.0f} | {h:02d}:{m:02d}:{s:02d}")
        print("\r" + row, end="\n" if at_checkpoint else "")

    @property
    def logs(self) -> dict[str, Any]:
        flush = getattr(self, "_flush_pending", None)          # device-side statistics of the last rollout still in flight
        if flush is not None:
            flush()
        elapsed = time.time() - self.start_time
        return dict(total_steps=self.current_step, total_episodes=self.current_episode - 1,
                    episode_returns=list(self.recent_returns), episode_lengths=list(self.recent_lengths),
                    best_reward=max(self.recent_returns, default=None), total_duration=elapsed,
                    mean_fps=self.current_step / (elapsed + 1e-6), custom_logs=self.custom_logs.copy())


class Logger:
    """Named scalar series (reference: diamond/utils.py:270-458; plotting needs plotly and is omitted)."""

    def __init__(self):
        self.logs: dict[str, dict[str, list[Any]]] = {}

    def log(self, log_name: str, step: int, value: Any):
        series = self.logs.setdefault(log_name, {"steps": [], "values": []})
        series["steps"].append(step)
        series["values"].append(value)

    def plot(self, *a, **k):
        raise NotImplementedError("plotting is outside the hot-path scope (needs plotly); use logger.logs")


class Timer:
    """Wall-clock section timer with the reference's `with timer.time(name)` interface
    (diamond/utils.py:461-543).  `cuda=True` brackets the section with device synchronisation."""

    def __init__(self, cuda: bool = False):
        self.timings: dict[str, dict[str, float]] = {}
        self.cuda = cuda

    def reset(self) -> None:
        self.timings = {}

    @contextmanager
    def time(self, name: str):
        if self.cuda and torch.cuda.is_available():
            torch.cuda.synchronize()
        t0 = time.time()
        try:
            yield
        finally:
            if self.cuda and torch.cuda.is_available():
                torch.cuda.synchronize()
            dt = time.time() - t0
            rec = self.timings.setdefault(name, {"avg_time": 0.0, "count": 0})
            rec["count"] += 1
            rec["avg_time"] += (dt - rec["avg_time"]) / rec["count"]


class Checkpointer:
    """`{run_name}-step{step:06d}.pt` files holding {"step", "model_state", "opt_state"} — the payload
    keys of the reference (diamond/utils.py:584-612) so checkpoints are interchangeable."""

    def __init__(self, folder: str | Path = "models", run_name: str = "run", *, keep_last: int | None = None) -> None:
        self.folder, self.run_name, self.keep_last = Path(folder), run_name, keep_last

    def save(self, step: int, model: torch.nn.Module, optimizer: torch.optim.Optimizer | None = None, *,
             extra: dict | None = None) -> Path:
        self.folder.mkdir(parents=True, exist_ok=True)
        path = self.folder / f"{self.run_name}-step{step:06d}.pt"
        payload = {"step": step, "model_state": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}}
        if optimizer is not None:
            osd = optimizer.state_dict()                   # moments are views of the device buffers: detach them from the live run
            payload["opt_state"] = {"state": {i: {k: (v.detach().cpu().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                                              for i, st in osd["state"].items()}, "param_groups": osd["param_groups"]}
        if extra:
            payload.update(extra)                          # e.g. "train_state" (resume); readers of the reference's keys ignore it
        torch.save(payload, path)
        if self.keep_last is not None:
            for old in sorted(self.folder.glob(f"{self.run_name}-step*.pt"))[:-self.keep_last]:
                old.unlink(missing_ok=True)
        return path

    def load(self, path: str | Path, model: torch.nn.Module, optimizer: torch.optim.Optimizer | None = None) -> int:
        chk = torch.load(path, map_location="cpu", weights_only=False)
        model.load_state_dict(chk["model_state"])          # copies into the flat device buffers in place
        if optimizer is not None and "opt_state" in chk:
            optimizer.load_state_dict(chk["opt_state"])
        return int(chk.get("step", 0))
