"""A/B inside one process: eager optimiser steps timed with CUDA events (8 steps), a few repeats."""
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
cfg = PPOConfig(num_envs=N_ENVS, rollout_steps=T, network_hidden_dim=H, num_epochs=E, num_minibatches=MB, verbose=False, total_steps=T*N_ENVS*1000)
agent = PPO(env_fn, cfg)
buf = RolloutBuffer(ctx, T, N_ENVS, D, 1, False, agent.device)
buf.load_host(*bench.synth_host_rollout(1))
np.random.seed(123)
for _ in range(3): agent.learn(buf)
torch.cuda.synchronize()
for rep in range(3):
    ev = {}
    agent.learn(buf, events=ev)
    torch.cuda.synchronize()
    print("learn: prepass %.3f gae %.3f update %.3f ms -> %.1f us/step" % (ev["start"].elapsed_time(ev["prepass_end"]), ev["prepass_end"].elapsed_time(ev["gae_end"]),
          ev["gae_end"].elapsed_time(ev["update_end"]), ev["gae_end"].elapsed_time(ev["update_end"]) * 1e3 / (E * MB)), flush=True)
