"""TEST INFRASTRUCTURE ONLY — CPU restatement of the Diamond PPO hot path.

A from-scratch restatement (explicit forward, explicit hand-derived backward, explicit
clip and Adam arithmetic — no autograd, no torch.optim, no torch.distributions) of what
the reference computes, written so that every CUDA kernel has a line-by-line CPU
counterpart.  Each function cites the reference lines it follows (paths relative to
/root/reference, and torch 2.11.0 library files where the arithmetic lives there).

Pinned by tests/test_oracle.py against tests/golden/*.npz, which were produced by
executing the UNMODIFIED reference (tests/golden/make_golden.py).  Never imported by the
product package; used only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import math

import numpy as np
import torch

F32 = torch.float32

# Parameter names, in the reference's module registration order (diamond/ppo.py:53-71).
DISCRETE_PARAM_NAMES = [
    "base.0.weight", "base.0.bias", "base.2.weight", "base.2.bias",
    "actor_head.0.weight", "actor_head.0.bias", "actor_head.2.weight", "actor_head.2.bias",
    "critic_head.0.weight", "critic_head.0.bias", "critic_head.2.weight", "critic_head.2.bias",
]
# diamond/continuous_ppo.py:64-82 (actor_log_std is registered after actor_mean_head)
CONTINUOUS_PARAM_NAMES = [
    "base.0.weight", "base.0.bias", "base.2.weight", "base.2.bias",
    "actor_mean_head.0.weight", "actor_mean_head.0.bias", "actor_mean_head.2.weight", "actor_mean_head.2.bias",
    "actor_log_std",
    "critic_head.0.weight", "critic_head.0.bias", "critic_head.2.weight", "critic_head.2.bias",
]


# --------------------------------------------------------------------------------------
# GAE  (diamond/ppo.py:188-222; identical copies continuous_ppo.py:200-234, recurrent_ppo.py:265-299)
# --------------------------------------------------------------------------------------
def gae(rewards, terminations, truncations, values, next_values, gamma=0.99, gae_lambda=0.95):
    """numpy float32, same operation order as the reference: (gamma*nv)*nt ; ((gamma*lambda)*nt)*ntr)*adv."""
    r, te, tr, v, nv = (np.asarray(x, dtype=np.float32) for x in (rewards, terminations, truncations, values, next_values))
    T = r.shape[0]
    out = np.zeros_like(r)
    adv = np.zeros_like(r[0])
    g = np.float32(gamma)
    gl = np.float32(gamma * gae_lambda)          # python double product first (ppo.py:215-216), then cast by torch
    one = np.float32(1.0)
    for t in reversed(range(T)):
        nt = one - te[t]
        ntr = one - tr[t]
        delta = r[t] + g * nv[t] * nt - v[t]                     # ppo.py:206-210
        adv = delta + gl * nt * ntr * adv                        # ppo.py:213-220
        out[t] = adv
    return out


def returns_and_normalise(values, advantages, advantage_norm=True):
    """ppo.py:241-243: returns from UN-normalised advantages; unbiased std; eps added to the std."""
    a = torch.as_tensor(np.asarray(advantages, dtype=np.float32))
    v = torch.as_tensor(np.asarray(values, dtype=np.float32))
    ret = v + a
    if advantage_norm:
        a64 = a.double()
        mean = a64.mean()
        std = torch.sqrt(((a64 - mean) ** 2).sum() / (a.numel() - 1))
        a = ((a - mean.float()) / (std.float() + 1e-6))
    return ret.numpy(), a.numpy()


# --------------------------------------------------------------------------------------
# legacy np.random.permutation  (ppo.py:254; numpy/random/mtrand.pyx shuffle + _mt19937.c, numpy 2.3.5)
# --------------------------------------------------------------------------------------
class MT19937:
    """Pure-Python MT19937 with numpy's legacy integer seeding (init_genrand)."""

    def __init__(self, seed):
        self.key = [0] * 624
        self.key[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.key[i] = (1812433253 * (self.key[i - 1] ^ (self.key[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.pos = 624

    def _gen(self):
        k = self.key
        for i in range(624):
            y = (k[i] & 0x80000000) | (k[(i + 1) % 624] & 0x7FFFFFFF)
            k[i] = k[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.pos = 0

    def next_u32(self):
        if self.pos == 624:
            self._gen()
        y = self.key[self.pos]
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y


def legacy_permutation(mt: MT19937, n: int) -> np.ndarray:
    """np.random.permutation(n): arange then Fisher-Yates from the end, j = random_interval(i)
    (masked rejection on 32-bit draws; numpy/random/src/distributions/distributions.c random_interval)."""
    arr = list(range(n))
    for i in range(n - 1, 0, -1):
        mask = i
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16
        while True:
            j = mt.next_u32() & mask
            if j <= i:
                break
        arr[i], arr[j] = arr[j], arr[i]
    return np.asarray(arr, dtype=np.int64)


# --------------------------------------------------------------------------------------
# Actor-critic MLP  (ppo.py:53-71, 91-96; continuous_ppo.py:64-82, 102-111)
# --------------------------------------------------------------------------------------
def _names(continuous):
    head = "actor_mean_head" if continuous else "actor_head"
    return head


def mlp_forward(p, x, continuous=False, keep=False):
    """x [M,D] f32 -> (head_out [M,A], values [M]) (+ activations if keep)."""
    head = _names(continuous)
    h1 = torch.tanh(x @ p["base.0.weight"].T + p["base.0.bias"])
    h2 = torch.tanh(h1 @ p["base.2.weight"].T + p["base.2.bias"])
    ha = torch.tanh(h2 @ p[f"{head}.0.weight"].T + p[f"{head}.0.bias"])
    hc = torch.tanh(h2 @ p["critic_head.0.weight"].T + p["critic_head.0.bias"])
    out = ha @ p[f"{head}.2.weight"].T + p[f"{head}.2.bias"]
    v = (hc @ p["critic_head.2.weight"].T + p["critic_head.2.bias"]).squeeze(-1)
    if keep:
        return out, v, (h1, h2, ha, hc)
    return out, v


def categorical_log_prob(logits, actions):
    """torch/distributions/categorical.py:78 (logits - logsumexp) and :151-158 (gather)."""
    lsm = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    return lsm.gather(-1, actions.long().unsqueeze(-1)).squeeze(-1), lsm


def normal_log_prob(mean, log_std, actions):
    """torch/distributions/normal.py:87-102 summed over action dims (continuous_ppo.py:40-43)."""
    scale = log_std.exp()
    var = scale ** 2
    lp = -((actions - mean) ** 2) / (2 * var) - scale.log() - math.log(math.sqrt(2 * math.pi))
    return lp.sum(-1)


def prepass(p, obs, next_obs, actions, continuous=False):
    """ppo.py:235-238: old log-probs, values, next_values with the pre-update parameters."""
    out, values = mlp_forward(p, obs, continuous)
    if continuous:
        log_std = p["actor_log_std"].expand_as(out)
        logp = normal_log_prob(out, log_std, actions)
    else:
        logp, _ = categorical_log_prob(out, actions)
    _, next_values = mlp_forward(p, next_obs, continuous)
    return logp, values, next_values


def loss_and_grads(p, x, actions, old_logp, adv, ret, cfg, continuous=False):
    """Forward (ppo.py:261), loss (ppo.py:264-280) and the hand-derived backward (ppo.py:283).

    Returns (losses dict, grads dict).  Gradients:
      policy : dL/dratio = (1/M) * (-A) where the unclipped branch is active (torch.max sends the
               gradient to the larger argument; inside the clip range both are equal and clamp's
               gradient is 1, so the sum is again -A), 0 where the clipped branch wins;
               dL/dlogp_new = dL/dratio * ratio.
      discrete: dlogits_k = dlogp*(1[k=a]-p_k) + (beta/M) * p_k*(lsm_k + H)
      gaussian: dmu_d = dlogp*(a-mu)/var ; dlogstd_d = sum_rows dlogp*((a-mu)^2/var - 1) - beta
      value  : dv = w_v * (v - ret)/M
    """
    head = _names(continuous)
    M = x.shape[0]
    out, v, (h1, h2, ha, hc) = mlp_forward(p, x, continuous, keep=True)
    eps = cfg["ppo_clip"]
    beta, wv = cfg["entropy_beta"], cfg["value_loss_weight"]
    if continuous:
        log_std = p["actor_log_std"].expand_as(out)
        scale = log_std.exp()
        var = scale ** 2
        new_logp = normal_log_prob(out, log_std, actions)
        ent_rows = (0.5 + 0.5 * math.log(2 * math.pi) + scale.log()).sum(-1)    # normal.py:114-115
    else:
        new_logp, lsm = categorical_log_prob(out, actions)
        probs = torch.softmax(lsm, dim=-1)
        ent_rows = -(lsm.clamp(min=torch.finfo(F32).min) * probs).sum(-1)      # categorical.py:160-163
    ratio = (new_logp - old_logp).exp()
    s1 = -adv * ratio
    s2 = -adv * torch.clamp(ratio, 1.0 - eps, 1.0 + eps)
    l_policy = torch.max(s1, s2).mean()
    l_value = 0.5 * ((v - ret) ** 2).mean()
    entropy = ent_rows.mean()
    total = l_policy + wv * l_value + -beta * entropy
    losses = dict(policy=float(l_policy), value=float(l_value), entropy=float(entropy), total=float(total))

    in_range = (ratio >= 1.0 - eps) & (ratio <= 1.0 + eps)
    w1 = torch.where(in_range, torch.ones_like(ratio),
                     torch.where(s1 > s2, torch.ones_like(ratio),
                                 torch.where(s1 == s2, torch.full_like(ratio, 0.5), torch.zeros_like(ratio))))
    dlogp = (-adv * w1 / M) * ratio
    g = {}
    if continuous:
        diff = actions - out
        dout = dlogp.unsqueeze(-1) * diff / var
        g["actor_log_std"] = ((dlogp.unsqueeze(-1) * (diff ** 2 / var - 1.0)).sum(0, keepdim=True)
                              - beta * torch.ones_like(p["actor_log_std"]))
    else:
        onehot = torch.zeros_like(out).scatter_(-1, actions.long().unsqueeze(-1), 1.0)
        dout = dlogp.unsqueeze(-1) * (onehot - probs) + (beta / M) * probs * (lsm + ent_rows.unsqueeze(-1))
    dv = wv * (v - ret) / M

    g[f"{head}.2.weight"] = dout.T @ ha
    g[f"{head}.2.bias"] = dout.sum(0)
    g["critic_head.2.weight"] = dv.unsqueeze(0) @ hc
    g["critic_head.2.bias"] = dv.sum().reshape(1)
    dha = (dout @ p[f"{head}.2.weight"]) * (1 - ha * ha)
    dhc = (dv.unsqueeze(-1) * p["critic_head.2.weight"]) * (1 - hc * hc)
    g[f"{head}.0.weight"] = dha.T @ h2
    g[f"{head}.0.bias"] = dha.sum(0)
    g["critic_head.0.weight"] = dhc.T @ h2
    g["critic_head.0.bias"] = dhc.sum(0)
    dh2 = (dha @ p[f"{head}.0.weight"] + dhc @ p["critic_head.0.weight"]) * (1 - h2 * h2)
    g["base.2.weight"] = dh2.T @ h1
    g["base.2.bias"] = dh2.sum(0)
    dh1 = (dh2 @ p["base.2.weight"]) * (1 - h1 * h1)
    g["base.0.weight"] = dh1.T @ x
    g["base.0.bias"] = dh1.sum(0)
    return losses, g


def clip_grad_norm_(grads, names, max_norm):
    """torch/nn/utils/clip_grad.py:96-106,165-182: norm of per-tensor norms; coef = max_norm/(total+1e-6)
    clamped to <= 1; gradients are ALWAYS multiplied by the clamped coefficient."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(grads[n]) for n in names]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for n in names:
        grads[n] = grads[n] * coef
    return float(total)


def adam_step_(p, grads, state, names, lr, eps, beta1=0.9, beta2=0.999):
    """torch/optim/adam.py single-tensor path (:457 lerp_, :476 mul_/addcmul_, :531-547 bias corrections
    in python double; denom = sqrt(v)/sqrt(bc2) + eps; p -= (lr/bc1) * m/denom)."""
    state["step"] += 1
    t = state["step"]
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    step_size = lr / bc1
    bc2_sqrt = bc2 ** 0.5
    for n in names:
        m, v, gr = state["exp_avg"][n], state["exp_avg_sq"][n], grads[n]
        m.lerp_(gr, 1 - beta1)
        v.mul_(beta2).addcmul_(gr, gr, value=1 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p[n].addcdiv_(m, denom, value=-step_size)


def linear_lr_factor(k, total_iters, start=1.0, end=1.0):
    """Closed form of torch LinearLR after k scheduler steps (optim/lr_scheduler.py:939-984)."""
    k = min(k, total_iters) if total_iters > 0 else 0
    if total_iters <= 0:
        return end
    return start + (end - start) * k / total_iters


def new_adam_state(p, names):
    return dict(step=0, exp_avg={n: torch.zeros_like(p[n]) for n in names},
                exp_avg_sq={n: torch.zeros_like(p[n]) for n in names})


def learn(p, state, obs, next_obs, actions, rewards, terminations, truncations, cfg, perms, continuous=False, lr=None):
    """One PPO.learn() (ppo.py:224-287 / continuous_ppo.py:236-299) on [T,N,...] numpy inputs.

    p: dict name -> torch f32 tensor (updated in place); state: Adam state (new_adam_state);
    perms: int array [E, B] (np.random.permutation outputs, ppo.py:254).  Returns list of loss dicts
    (one per minibatch) and the intermediate (advantages, returns, normalised advantages)."""
    names = CONTINUOUS_PARAM_NAMES if continuous else DISCRETE_PARAM_NAMES
    T, N = rewards.shape[:2]
    obs_t = torch.as_tensor(np.asarray(obs, dtype=np.float32))
    nobs_t = torch.as_tensor(np.asarray(next_obs, dtype=np.float32))
    act_t = torch.as_tensor(np.asarray(actions, dtype=np.float32 if continuous else np.int64))
    r, te, tr = (np.asarray(x).astype(np.float32) for x in (rewards, terminations, truncations))    # ppo.py:230-232
    B = T * N
    flat_obs = obs_t.reshape(B, -1)
    logp, values, next_values = prepass(p, flat_obs, nobs_t.reshape(B, -1),
                                        act_t.reshape(B, -1) if continuous else act_t.reshape(B), continuous)
    adv = gae(r, te, tr, values.reshape(T, N).numpy(), next_values.reshape(T, N).numpy(), cfg["gamma"], cfg["gae_lambda"])
    ret, adv_n = returns_and_normalise(values.reshape(T, N).numpy(), adv, cfg["advantage_norm"])
    adv_f, ret_f = torch.as_tensor(adv_n).reshape(B), torch.as_tensor(ret).reshape(B)
    act_f = act_t.reshape(B, -1) if continuous else act_t.reshape(B)
    MB = cfg["num_minibatches"]
    M = B // MB
    lr = cfg["lr"] if lr is None else lr
    all_losses = []
    for perm in np.asarray(perms):
        for mb in perm.reshape(MB, M):
            idx = torch.as_tensor(mb.astype(np.int64))
            losses, grads = loss_and_grads(p, flat_obs[idx], act_f[idx], logp[idx], adv_f[idx], ret_f[idx], cfg, continuous)
            losses["grad_norm"] = clip_grad_norm_(grads, names, cfg["grad_norm_clip"])
            adam_step_(p, grads, state, names, lr, cfg["adam_eps"])
            all_losses.append(losses)
    return all_losses, dict(advantages=adv, returns=ret, adv_norm=adv_n, values=values.reshape(T, N).numpy(),
                            next_values=next_values.reshape(T, N).numpy(), log_probs=logp.reshape(T, N).numpy())


def default_cfg(**kw):
    """PPOConfig defaults (ppo.py:17-37) as a plain dict."""
    d = dict(lr=3e-4, adam_eps=1e-5, gamma=0.99, gae_lambda=0.95, num_epochs=4, num_minibatches=8, ppo_clip=0.2,
             value_loss_weight=1.0, entropy_beta=0.01, advantage_norm=True, grad_norm_clip=0.5)
    d.update(kw)
    return d


# --------------------------------------------------------------------------------------
# Recurrent core (recurrent_ppo.py:46-91, 127-149): GRU cell with done-masked resets
# --------------------------------------------------------------------------------------
def gru_forward(p, x, hx, dones):
    """x [T,B,I], hx [B,Hg], dones [T,B] bool -> outputs [T,B,Hg], final hx.  Gate order r,z,n
    (torch nn.GRU): n = tanh(W_in x + b_in + r*(W_hn h + b_hn)); h' = (1-z)*n + z*h."""
    Wi, Wh, bi, bh = p["gru.weight_ih_l0"], p["gru.weight_hh_l0"], p["gru.bias_ih_l0"], p["gru.bias_hh_l0"]
    Hg = Wh.shape[1]
    outs = []
    h = hx
    for t in range(x.shape[0]):
        h = torch.where(dones[t].unsqueeze(-1), torch.zeros_like(h), h)        # recurrent_ppo.py:84
        gi = x[t] @ Wi.T + bi
        gh = h @ Wh.T + bh
        r = torch.sigmoid(gi[:, :Hg] + gh[:, :Hg])
        z = torch.sigmoid(gi[:, Hg:2 * Hg] + gh[:, Hg:2 * Hg])
        n = torch.tanh(gi[:, 2 * Hg:] + r * gh[:, 2 * Hg:])
        h = (1 - z) * n + z * h
        outs.append(h)
    return torch.stack(outs), h
