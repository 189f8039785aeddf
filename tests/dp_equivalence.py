"""Multi-GPU check, launched by torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_equivalence.py

Env-sharded data-parallel learn() over G ranks (global permutation, gradient all-reduce) must equal the
single-GPU learn() on the concatenated buffer up to fp32 summation order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "diamond-ppo_b200")):
    sys.path.insert(0, p)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from diamond import PPO, PPOConfig, envs
    D, A, H, T, NL = 16, 4, 64, 32, 24
    NG = NL * world
    rng = np.random.default_rng(0)
    full = [[rng.standard_normal((NG, D)).astype(np.float32), rng.standard_normal((NG, D)).astype(np.float32),
             rng.integers(0, A, NG), rng.standard_normal(NG), rng.random(NG) < 0.05, rng.random(NG) < 0.05] for _ in range(T)]
    lo, hi = rank * NL, (rank + 1) * NL
    local_exp = [[x[lo:hi] for x in step] for step in full]

    def env_fn(n):
        return envs.BatchedSyntheticVectorEnv(n, D, A)
    env_fn.vectorized = True
    cfg = PPOConfig(num_envs=NL, rollout_steps=T, network_hidden_dim=H, num_epochs=2, num_minibatches=4, verbose=False)
    agent = PPO(env_fn, cfg, dp=True)
    init = {k: v.detach().clone() for k, v in agent.network.state_dict().items()}
    np.random.seed(123)
    agent.learn(local_exp)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        cfg1 = PPOConfig(num_envs=NG, rollout_steps=T, network_hidden_dim=H, num_epochs=2, num_minibatches=4, verbose=False)
        single = PPO(env_fn, cfg1, dp=False)
        single.network.load_state_dict(init)
        np.random.seed(123)
        single.learn(full)
        torch.cuda.synchronize()
        worst = 0.0
        for (n, p), (_, q) in zip(agent.network.named_parameters(), single.network.named_parameters()):
            err = float((p - q).abs().max() / q.abs().max().clamp_min(1e-12))
            worst = max(worst, err)
        lerr = float((agent.last_losses - single.last_losses).abs().max())
        ok = worst <= 2e-5 and lerr <= 1e-5
        print(f"dp{world} vs single: max param err {worst:.2e}, max loss err {lerr:.2e} -> {'OK' if ok else 'FAIL'}", flush=True)
    # every rank holds identical parameters after the update
    flat = agent.engine.P.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(flat, ref))
    t = torch.tensor([int(ok and same)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(t) != 1:
        print(f"rank {rank}: replicas identical={same}", flush=True)
        sys.exit(1)


if __name__ == "__main__":
    main()
