"""Default actor-critic networks with the reference's module structure and custom-network
interface (discrete: diamond/ppo.py:40-108; continuous: diamond/continuous_ppo.py:50-121).

Module/parameter names equal the reference's, so `state_dict()` payloads are interchangeable.
When an agent adopts a default network (`engine` attribute set) the parameters become views into
the flat device buffers the CUDA kernels update, and `get_actions` runs the fused forward +
sampling kernels.  The torch methods below remain the public interface for user code
(evaluation, custom subclasses); the agents' fused update never goes through them.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def _obs_dim(space) -> int:
    assert hasattr(space, "shape") and not hasattr(space, "n"), "Only Box obs spaces are supported."
    return int(np.prod(space.shape))


def _mlp(inp: int, hidden: int, out: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(inp, hidden), nn.Tanh(), nn.Linear(hidden, out))


class ActorCriticNetwork(nn.Module):
    """Discrete policy: shared trunk (2 x Linear+Tanh), actor head -> logits, critic head -> value."""

    def __init__(self, observation_space, action_space, cfg) -> None:
        super().__init__()
        assert hasattr(action_space, "n"), "Only Discrete action spaces are supported."
        d, h = _obs_dim(observation_space), int(cfg.network_hidden_dim)
        self.base = nn.Sequential(nn.Linear(d, h), nn.Tanh(), nn.Linear(h, h), nn.Tanh())
        self.actor_head = _mlp(h, h, int(action_space.n))
        self.actor_out_layer = self.actor_head[-1]
        self.critic_head = _mlp(h, h, 1)
        self.engine = None            # set by the agent when the fused CUDA path adopts this module

    def get_actions(self, observations: np.ndarray, device: torch.device) -> np.ndarray:
        if self.engine is not None:
            return self.engine.sample_actions(observations)
        obs = torch.as_tensor(observations, dtype=torch.float32, device=device)
        with torch.inference_mode():
            logits = self.actor_head(self.base(obs))
        return torch.distributions.Categorical(logits=logits).sample().cpu().numpy()

    def get_values(self, observations: torch.Tensor) -> torch.Tensor:
        with torch.inference_mode():
            return self.critic_head(self.base(observations)).squeeze(-1)

    def get_logits_and_values(self, x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        x = self.base(x)
        return self.actor_head(x), self.critic_head(x).squeeze(-1)


class ContinuousActorCriticNetwork(nn.Module):
    """Gaussian policy: mean head + state-independent log-std parameter (1, A)."""

    def __init__(self, observation_space, action_space, cfg) -> None:
        super().__init__()
        assert hasattr(action_space, "shape") and not hasattr(action_space, "n"), "Only Box action spaces are supported."
        d, h, a = _obs_dim(observation_space), int(cfg.network_hidden_dim), int(np.prod(action_space.shape))
        self.base = nn.Sequential(nn.Linear(d, h), nn.Tanh(), nn.Linear(h, h), nn.Tanh())
        self.actor_mean_head = _mlp(h, h, a)
        self.actor_log_std = nn.Parameter(torch.zeros(1, a))
        self.critic_head = _mlp(h, h, 1)
        self.engine = None

    def get_actions(self, observations: np.ndarray, device: torch.device) -> np.ndarray:
        if self.engine is not None:
            return self.engine.sample_actions(observations)
        obs = torch.as_tensor(observations, dtype=torch.float32, device=device)
        with torch.inference_mode():
            mean = self.actor_mean_head(self.base(obs))
        std = torch.broadcast_to(self.actor_log_std, mean.shape).exp()
        return torch.normal(mean, std).cpu().numpy()

    def get_values(self, observations: torch.Tensor) -> torch.Tensor:
        with torch.inference_mode():
            return self.critic_head(self.base(observations)).squeeze(-1)

    def get_means_log_stds_and_values(self, observations: torch.Tensor):
        x = self.base(observations)
        mean = self.actor_mean_head(x)
        return mean, torch.broadcast_to(self.actor_log_std, mean.shape), self.critic_head(x).squeeze(-1)


def network_parameter_init_(network: nn.Module, gain: float = 1.0, small_actor_out: bool = True) -> None:
    """Orthogonal weights / zero biases for every nn.Linear; `actor_out_layer` (if present and
    `small_actor_out`) re-drawn with gain 0.01 (diamond/ppo.py:99-108).  The continuous variant has
    no small-gain output layer (continuous_ppo.py:114-121)."""
    with torch.no_grad():
        for m in network.modules():
            if isinstance(m, nn.Linear):
                nn.init.orthogonal_(m.weight, gain=gain)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        if small_actor_out and hasattr(network, "actor_out_layer"):
            nn.init.orthogonal_(network.actor_out_layer.weight, gain=0.01)
