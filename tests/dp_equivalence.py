"""Multi-GPU check, launched by torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_equivalence.py

Env-sharded data-parallel learn() over G ranks with the GLOBAL permutation must equal the single-GPU learn() on the
concatenated buffer up to fp32 summation order -- for both exchange paths: the fused NVLink peer-memory
kernel (dppo_dp_allreduce_clip_adam) and NCCL all_reduce + dppo_clip_adam_step.  The default rank-local permutation mode
is checked for replica consistency (bit-identical parameters on every rank) and finite losses."""
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
    # hidden 256 and >= 1024 rows per rank and step: every GEMM of the update runs on the 3xTF32 tcgen05 kernels
    D, A, H, T, NL = 64, 4, 256, 32, 320
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
    ok, same = True, True
    single_result = None
    for exchange in ("fused", "nccl"):
        agent = PPO(env_fn, cfg, dp=True, dp_permutation="global", dp_exchange=exchange)
        if exchange == "fused":
            assert agent.engine.dpx is not None, "fused exchange was not set up"
        init = {k: v.detach().clone() for k, v in agent.network.state_dict().items()}
        np.random.seed(123)
        agent.learn(local_exp)
        torch.cuda.synchronize()
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
            good = worst <= 2e-5 and lerr <= 2e-5
            ok = ok and good
            print(f"dp{world} [{exchange}] vs single: max param err {worst:.2e}, max loss err {lerr:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
        # every rank holds identical parameters after the update
        flat = agent.engine.P.clone()
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        same = same and bool(torch.equal(flat, ref))
    # default mode: rank-local permutations, fused exchange
    agent = PPO(env_fn, cfg, dp=True)
    np.random.seed(123)
    agent.learn(local_exp)
    agent.learn(local_exp)
    torch.cuda.synchronize()
    flat = agent.engine.P.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    local_ok = bool(torch.equal(flat, ref)) and bool(torch.isfinite(agent.last_losses).all()) and bool(torch.isfinite(flat).all())
    if rank == 0:
        print(f"dp{world} [local permutation, fused]: replicas identical and finite -> {'OK' if local_ok else 'FAIL'}", flush=True)
    same = same and local_ok
    # graph replay under DP (device-resident exchange sequence numbers) equals the eager DP path bit for bit -- rank-local and global
    # permutation (padded fixed-shape steps from the device shard filter)
    from diamond.agents import RolloutBuffer
    for mode in ("local", "global"):
        finals = []
        for use_graphs in (False, True):
            ag = PPO(env_fn, cfg, dp=True, dp_permutation=mode)
            ag.engine.use_graphs = use_graphs
            buf = RolloutBuffer.from_lists(ag.ctx, local_exp, False, ag.device)
            np.random.seed(77)
            for _ in range(3):
                ag.learn(buf)
            torch.cuda.synchronize()
            ag.engine.check_health()
            finals.append(ag.engine.P.clone())
            if use_graphs:
                assert any(b_.get("graph") for b_ in ag.engine._bufs.values()), "graph path was not taken under DP"
        graph_ok = bool(torch.equal(finals[0], finals[1]))
        if rank == 0:
            print(f"dp{world} [{mode} permutation, graph replay vs eager]: bit-identical -> {'OK' if graph_ok else 'FAIL'}", flush=True)
        same = same and graph_ok
    # device permutation generator (fast mode): the sharded update equals the single-GPU update that draws the same keyed permutation
    ag = PPO(env_fn, PPOConfig(num_envs=NL, rollout_steps=T, network_hidden_dim=H, num_epochs=2, num_minibatches=4, verbose=False, seed=9),
             dp=True, dp_permutation="global", minibatch_permutation="device")
    init = {k: v.detach().clone() for k, v in ag.network.state_dict().items()}
    ag.learn(local_exp)
    torch.cuda.synchronize()
    ag.engine.check_health()
    flat = ag.engine.P.clone(); ref = flat.clone(); dist.broadcast(ref, src=0)
    dev_ok = bool(torch.equal(flat, ref))
    if rank == 0:
        single = PPO(env_fn, PPOConfig(num_envs=NG, rollout_steps=T, network_hidden_dim=H, num_epochs=2, num_minibatches=4, verbose=False,
                                       seed=9), dp=False, minibatch_permutation="device")
        single.network.load_state_dict(init)
        single.learn(full)
        torch.cuda.synchronize()
        worst = max(float((p_ - q_).abs().max() / q_.abs().max().clamp_min(1e-12))
                    for (_, p_), (_, q_) in zip(ag.network.named_parameters(), single.network.named_parameters()))
        dev_ok = dev_ok and worst <= 2e-5
        print(f"dp{world} [device permutation] vs single: max param err {worst:.2e} -> {'OK' if dev_ok else 'FAIL'}", flush=True)
    same = same and dev_ok
    # a user network_cls under DP (autograd module, NCCL gradient sum): replicas identical, and -- with every rank holding the SAME data
    # and the same permutation -- equal to the single-GPU update on one copy (sums of G identical shard gradients / G x the rows)
    import torch.nn as nn

    class TinyNet(nn.Module):
        def __init__(self, observation_space, action_space, cfg):
            super().__init__()
            d = int(np.prod(observation_space.shape))
            self.actor = nn.Sequential(nn.Linear(d, 32), nn.ReLU(), nn.Linear(32, action_space.n))
            self.critic = nn.Sequential(nn.Linear(d, 32), nn.ReLU(), nn.Linear(32, 1))
            self.actor_out_layer = self.actor[-1]

        def get_actions(self, observations, device):
            with torch.inference_mode():
                return torch.distributions.Categorical(logits=self.actor(torch.as_tensor(observations, dtype=torch.float32, device=device))).sample().cpu().numpy()

        def get_values(self, observations):
            with torch.inference_mode():
                return self.critic(observations).squeeze(-1)

        def get_logits_and_values(self, x):
            return self.actor(x), self.critic(x).squeeze(-1)

    cfg_c = PPOConfig(num_envs=NL, rollout_steps=T, num_epochs=2, num_minibatches=4, verbose=False, seed=3)
    shard0 = [[x[0:NL] for x in step] for step in full]                 # every rank learns from shard 0
    ag = PPO(env_fn, cfg_c, network_cls=TinyNet, dp=True)
    init = {k: v.detach().clone() for k, v in ag.network.state_dict().items()}
    np.random.seed(31)
    ag.learn(shard0)
    torch.cuda.synchronize()
    flat = ag.engine.P.clone(); ref = flat.clone(); dist.broadcast(ref, src=0)
    cust_ok = bool(torch.equal(flat, ref)) and bool(torch.isfinite(flat).all())
    if rank == 0:
        single = PPO(env_fn, cfg_c, network_cls=TinyNet, dp=False)
        single.network.load_state_dict(init)
        np.random.seed(31)
        single.learn(shard0)
        torch.cuda.synchronize()
        worst = max(float((p_ - q_).abs().max() / q_.abs().max().clamp_min(1e-12))
                    for (_, p_), (_, q_) in zip(ag.network.named_parameters(), single.network.named_parameters()))
        cust_ok = cust_ok and worst <= 2e-5
        print(f"dp{world} [custom network_cls, NCCL gradient sum] vs single on replicated data: max param err {worst:.2e} -> {'OK' if cust_ok else 'FAIL'}", flush=True)
    same = same and cust_ok
    # a rank whose numpy stream drifted is detected (asynchronously) instead of silently mis-assigning minibatch members
    ag = PPO(env_fn, cfg, dp=True, dp_permutation="global")
    if rank == world - 1:
        np.random.random()
    ag.learn(local_exp)
    torch.cuda.synchronize()
    try:
        ag.engine.check_health()
        caught = False
    except RuntimeError:
        caught = True
    if rank == 0:
        print(f"dp{world} [numpy stream drift]: detected -> {'OK' if caught else 'FAIL'}", flush=True)
    same = same and caught
    np.random.seed(1)
    # end to end under DP with device-resident environments: every rank owns envs [rank*N, (rank+1)*N) of the global run (reset and
    # sampling draws keyed by global env id), rollout() on the device with recorded values, learn() with the fused exchange
    from diamond.envs import DeviceVectorEnv
    n_loc, t_roll = 64, 64
    cfg_e = PPOConfig(num_envs=n_loc, rollout_steps=t_roll, verbose=False, seed=5, total_steps=n_loc * t_roll * 6)
    ag = PPO(DeviceVectorEnv.factory("CartPole-v1", seed=5, env_offset=rank * n_loc), cfg_e, dp=True)
    ag.train()
    torch.cuda.synchronize()
    flat = ag.engine.P.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    st0 = ag.envs.state.clone()
    gathered = [torch.empty_like(st0) for _ in range(world)]
    dist.all_gather(gathered, st0)
    distinct = all(not torch.equal(gathered[0], g_) for g_ in gathered[1:])        # shards simulate different environments
    env_ok = bool(torch.equal(flat, ref)) and bool(torch.isfinite(flat).all()) and bool(torch.isfinite(ag.last_losses).all()) and distinct
    if rank == 0:
        print(f"dp{world} [device envs, train()]: replicas identical, shards distinct, finite -> {'OK' if env_ok else 'FAIL'}", flush=True)
    same = same and env_ok
    # RecurrentPPO under env-sharded DP (rank-local permutations, NCCL gradient sum; the GRU scans shard by environment): with every
    # rank holding the SAME rollout and permutation the update equals the single-GPU update on one copy, replicas stay identical;
    # train() end to end with distinct shards keeps the replicas identical
    from diamond import RecurrentPPO, RecurrentPPOConfig
    n_r, t_r = 64, 32
    # (advantage_norm off: duplicating the data changes the unbiased (n - 1) standard deviation by 1 / (4 B), which is all that would
    # distinguish the two runs)
    cfg_r = RecurrentPPOConfig(num_envs=n_r, rollout_steps=t_r, num_epochs=2, num_minibatches=4, verbose=False, seed=11,
                               total_steps=n_r * t_r * 1000, advantage_norm=False)

    def recurrent_agent(dp):
        a = RecurrentPPO(DeviceVectorEnv.factory("CartPole-v1", seed=11), cfg_r, dp=dp)
        a.envs.desc.env_offset = 0                                     # replicated data: every rank simulates the same environments
        a.engine.env_offset = 0
        a.current_observations, _ = a.envs.reset(seed=11)
        a.prev_dones = np.zeros(n_r, dtype=bool)
        a.current_hx = torch.zeros(1, n_r, cfg_r.gru_hidden_dim, device=a.device)
        return a

    ag = recurrent_agent(True)
    init = {k: v.detach().clone() for k, v in ag.network.state_dict().items()}
    ro = ag.rollout()
    np.random.seed(5)
    ag.learn(ro)
    torch.cuda.synchronize()
    flat = ag.engine.P.clone(); ref = flat.clone(); dist.broadcast(ref, src=0)
    rec_ok = bool(torch.equal(flat, ref)) and bool(torch.isfinite(flat).all())
    if rank == 0:
        single = recurrent_agent(False)
        single.network.load_state_dict(init)
        ro1 = single.rollout()
        assert torch.equal(ro1.obs, ro.obs) and torch.equal(ro1.actions, ro.actions), "replicated rollouts differ"
        np.random.seed(5)
        single.learn(ro1)
        torch.cuda.synchronize()
        worst = max(float((p_ - q_).abs().max() / q_.abs().max().clamp_min(1e-12))
                    for (_, p_), (_, q_) in zip(ag.network.named_parameters(), single.network.named_parameters()))
        lerr = float((ag.last_losses - single.last_losses).abs().max())
        rec_ok = rec_ok and worst <= 2e-5 and lerr <= 2e-5
        print(f"dp{world} [RecurrentPPO, NCCL gradient sum] vs single on replicated data: max param err {worst:.2e}, max loss err {lerr:.2e} "
              f"-> {'OK' if rec_ok else 'FAIL'}", flush=True)
    same = same and rec_ok
    cfg_rt = RecurrentPPOConfig(num_envs=n_r, rollout_steps=t_r, verbose=False, seed=13, total_steps=n_r * t_r * 4)
    ag = RecurrentPPO(DeviceVectorEnv.factory("CartPole-v1", seed=13), cfg_rt, dp=True)
    ag.train()
    torch.cuda.synchronize()
    flat = ag.engine.P.clone(); ref = flat.clone(); dist.broadcast(ref, src=0)
    st0 = ag.envs.state.clone()
    gathered = [torch.empty_like(st0) for _ in range(world)]
    dist.all_gather(gathered, st0)
    rec_train_ok = (bool(torch.equal(flat, ref)) and bool(torch.isfinite(flat).all()) and bool(torch.isfinite(ag.last_losses).all())
                    and all(not torch.equal(gathered[0], g_) for g_ in gathered[1:]))
    if rank == 0:
        print(f"dp{world} [RecurrentPPO, device envs, train()]: replicas identical, shards distinct, finite -> {'OK' if rec_train_ok else 'FAIL'}", flush=True)
    same = same and rec_train_ok
    t = torch.tensor([int(ok and same)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(t) != 1:
        print(f"rank {rank}: replicas identical={same}", flush=True)
        sys.exit(1)


if __name__ == "__main__":
    main()
