"""Flat parameter layout of the default actor-critic MLP (include/dppo.h dppo_mlp_layout).

The kernels read parameters, write gradients and keep Adam moments in four flat fp32 buffers of
identical layout.  The torch module's nn.Parameters are views into the parameter buffer, so
`state_dict()`, checkpoints and user code keep working while the kernels update the same memory.
actor_head.0 and critic_head.0 are adjacent ([2H, H]) so both first head layers run as one product.
"""
from __future__ import annotations

import torch

from . import _native as N


def param_slices(desc: N.MlpDesc, lay: N.MlpLayout):
    """name (reference module naming, diamond/ppo.py:53-71 / continuous_ppo.py:64-82) -> (offset, shape)."""
    D, H, A = desc.obs_dim, desc.hidden, desc.act_dim
    head = "actor_mean_head" if desc.continuous else "actor_head"
    s = {
        "base.0.weight": (lay.w1, (H, D)), "base.0.bias": (lay.b1, (H,)),
        "base.2.weight": (lay.w2, (H, H)), "base.2.bias": (lay.b2, (H,)),
        f"{head}.0.weight": (lay.w3, (H, H)), f"{head}.0.bias": (lay.b3, (H,)),
        f"{head}.2.weight": (lay.wa, (A, H)), f"{head}.2.bias": (lay.ba, (A,)),
        "critic_head.0.weight": (lay.w3 + H * H, (H, H)), "critic_head.0.bias": (lay.b3 + H, (H,)),
        "critic_head.2.weight": (lay.wc, (1, H)), "critic_head.2.bias": (lay.bc, (1,)),
    }
    if desc.continuous:
        s["actor_log_std"] = (lay.log_std, (1, A))
    return s


class FlatMlp:
    """Descriptor + layout + named views over a flat buffer."""

    def __init__(self, obs_dim: int, hidden: int, act_dim: int, continuous: bool):
        self.desc = N.MlpDesc(int(obs_dim), int(hidden), int(act_dim), int(bool(continuous)))
        self.layout = N.mlp_layout(self.desc)
        self.slices = param_slices(self.desc, self.layout)
        self.total = int(self.layout.total)

    def views(self, flat: torch.Tensor):
        out = {}
        for name, (off, shape) in self.slices.items():
            n = 1
            for d in shape:
                n *= d
            out[name] = flat[off:off + n].view(*shape)
        return out

    def pack(self, named: dict, device=None, dtype=torch.float32) -> torch.Tensor:
        flat = torch.zeros(self.total, dtype=dtype, device=device)
        v = self.views(flat)
        for name in self.slices:
            v[name].copy_(torch.as_tensor(named[name]).to(dtype).reshape(v[name].shape))
        return flat
