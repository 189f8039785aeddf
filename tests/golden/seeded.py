"""TEST INFRASTRUCTURE ONLY — inputs of the large learn() fixtures, regenerated from a seed.

The synthetic-scale configurations (4096 envs x 128 steps, obs 64, hidden 256) are
too large to commit as arrays (2 x 134 MB of observations), so `make_golden.py` and the GPU tests both call
these functions: same seed -> same numpy PCG64 stream -> bit-identical experience and initial parameters on
the build container (where the UNMODIFIED reference produced the committed losses / updated parameters) and
on the GPU box.  The fixture stores a checksum of what was generated; the tests verify it before comparing.
"""
import numpy as np

DISCRETE_PARAM_SHAPES = lambda D, H, A: {  # noqa: E731  (names/shapes of diamond/ppo.py:53-71)
    "base.0.weight": (H, D), "base.0.bias": (H,), "base.2.weight": (H, H), "base.2.bias": (H,),
    "actor_head.0.weight": (H, H), "actor_head.0.bias": (H,), "actor_head.2.weight": (A, H), "actor_head.2.bias": (A,),
    "critic_head.0.weight": (H, H), "critic_head.0.bias": (H,), "critic_head.2.weight": (1, H), "critic_head.2.bias": (1,)}


def seeded_experience(seed, T, N, D, A, p_term=0.01, p_trunc=0.01):
    """Arrays in the layout PPO.rollout() produces (ppo.py:165-172), stacked over T:
    obs, next_obs f32 [T,N,D]; actions i64 [T,N]; rewards f64 [T,N]; terminations, truncations bool [T,N]."""
    rng = np.random.default_rng(seed)
    obs = rng.standard_normal((T, N, D), dtype=np.float32)
    nobs = rng.standard_normal((T, N, D), dtype=np.float32)
    act = rng.integers(0, A, size=(T, N)).astype(np.int64)
    rew = rng.standard_normal((T, N)).astype(np.float64)
    term = rng.random((T, N)) < p_term
    trunc = (rng.random((T, N)) < p_trunc) & ~term
    return obs, nobs, act, rew, term, trunc


def seeded_params(seed, D, H, A):
    """Initial parameters with the scale of the reference's init (orthogonal gain sqrt(2) ~ N(0, 2/fan_in);
    actor output layer gain 0.01, ppo.py:99-108) and small non-zero biases so that every bias gradient path matters."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in DISCRETE_PARAM_SHAPES(D, H, A).items():
        if name.endswith("weight"):
            gain = 0.01 if name == "actor_head.2.weight" else np.sqrt(2.0)
            out[name] = (rng.standard_normal(shape) * gain / np.sqrt(shape[1])).astype(np.float32)
        else:
            out[name] = (rng.standard_normal(shape) * 0.01).astype(np.float32)
    return out


def state_dict_of(params):
    """The 12 tensors + the `actor_out_layer` alias the reference's module registers (ppo.py:65)."""
    sd = dict(params)
    sd["actor_out_layer.weight"], sd["actor_out_layer.bias"] = params["actor_head.2.weight"], params["actor_head.2.bias"]
    return sd


def checksum(arrays):
    """Order-sensitive fp64 checksum of a sequence of arrays (cheap guard that both sides generated the same bits)."""
    s = 0.0
    for i, a in enumerate(arrays):
        a = np.asarray(a, dtype=np.float64).ravel()
        s += float((a[::7] * (1.0 + (i % 5))).sum()) + float(a[-1]) * 3.0
    return s
