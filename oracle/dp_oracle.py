"""TEST INFRASTRUCTURE ONLY — host restatement of the env-sharded split of the reference's minibatches.

The reference draws ONE permutation of the whole batch per epoch and cuts it into num_minibatches consecutive minibatches
(diamond/ppo.py:252-255); flat sample index i = t*N + env (ppo.py:246-249).  Under env-sharded data parallelism rank r owns
envs [lo, lo + n_local) of the N_global environments, so its share of global minibatch k is the subsequence of that minibatch
whose env falls in its range, re-indexed to its local buffer.  `shard_filter` states exactly that with numpy; the product does
it on the device (dppo_perm_shard_filter) and tests compare the two.
"""
import numpy as np


def shard_filter(perm, n_global, lo, n_local, num_minibatches, m_pad):
    """perm: int array [B_global] (one epoch).  Returns (idx [MB, m_pad] int32 padded with -1, counts [MB], overflow flag)."""
    perm = np.asarray(perm, dtype=np.int64)
    M = perm.size // num_minibatches
    idx = np.full((num_minibatches, m_pad), -1, dtype=np.int32)
    counts = np.zeros(num_minibatches, dtype=np.int32)
    overflow = 0
    for k in range(num_minibatches):
        seg = perm[k * M:(k + 1) * M]
        t, env = np.divmod(seg, n_global)
        mine = (env >= lo) & (env < lo + n_local)
        local = (t[mine] * n_local + (env[mine] - lo)).astype(np.int32)
        if local.size > m_pad:
            overflow = 1
            local = local[:m_pad]
        idx[k, :local.size] = local
        counts[k] = local.size
    return idx, counts, overflow
