"""A/B of the head training kernel (tc_debug bit 256 = generic shared-memory-accumulator kernel) inside whole learn() calls."""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
import bench
from diamond import PPO, PPOConfig, envs, _native
from diamond.agents import RolloutBuffer
T, N_ENVS, D, H, A, E, MB = bench.T, bench.N_ENVS, bench.D, bench.H, bench.A, bench.E, bench.MB
ctx = _native.get_context(0)
def env_fn(n): return envs.BatchedSyntheticVectorEnv(n, D, A)
env_fn.vectorized = True
host = bench.synth_host_rollout(1)
res = {}
BIT = int(sys.argv[1]) if len(sys.argv) > 1 else 256      # 256: generic head kernel; 512: no programmatic dependent launch
for dbg in (BIT, 0, BIT, 0, BIT, 0):
    ctx.set_option("tc_debug", dbg)
    torch.manual_seed(0)
    cfg = PPOConfig(num_envs=N_ENVS, rollout_steps=T, network_hidden_dim=H, num_epochs=E, num_minibatches=MB, verbose=False, total_steps=T*N_ENVS*1000)
    agent = PPO(env_fn, cfg)
    buf = RolloutBuffer(ctx, T, N_ENVS, D, 1, False, agent.device)
    buf.load_host(*host)
    np.random.seed(123)
    for _ in range(3): agent.learn(buf)
    torch.cuda.synchronize()
    uss = []
    for _ in range(5):
        ev = {}
        agent.learn(buf, events=ev)
        torch.cuda.synchronize()
        uss.append(ev["gae_end"].elapsed_time(ev["update_end"]) * 1e3 / (E * MB))
    us = min(uss)
    print("   repeats:", " ".join(f"{u:.1f}" for u in uss))
    p = torch.cat([q.detach().flatten() for q in agent.network.parameters()]).double().cpu()
    print(f"tc_debug {dbg}: {us:.1f} us/step", flush=True)
    if dbg in res:
        pass
    res.setdefault(dbg, p)
d = (res[0] - res[BIT]).abs().max().item() / res[BIT].abs().max().item()
print(f"params after 8 learn() calls, tc_debug 0 vs {BIT}: max rel diff {d:.2e}")
