"""rollout() throughput: host vector env (numpy, 2 PCIe crossings per step) vs device-resident env (csrc/envs.cu)."""
import sys, os, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
from diamond import PPO, PPOConfig, envs
from diamond.envs import DeviceVectorEnv

def run(name, env_fn, N, T, H, reps=3):
    cfg = PPOConfig(num_envs=N, rollout_steps=T, network_hidden_dim=H, verbose=False, seed=1, total_steps=N * T * 1000)
    agent = PPO(env_fn, cfg)
    agent.ticker = None
    agent.current_observations, _ = agent.envs.reset(seed=1)
    for _ in range(4): agent.rollout()          # eager, eager, capture, replay
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): agent.rollout()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    for _ in range(2):
        buf = agent.rollout(); agent.learn(buf)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(reps):
        buf = agent.rollout(); agent.learn(buf)
    torch.cuda.synchronize()
    both = (time.perf_counter() - t1) * 1e3 / reps
    print(f"{name:44s} rollout {ms:9.2f} ms = {N * T / ms * 1e3 / 1e6:8.3f} M env-steps/s   rollout+learn {both:8.2f} ms", flush=True)

def host_synth(n): return envs.BatchedSyntheticVectorEnv(n, 64, 4)
host_synth.vectorized = True
run("host numpy synthetic  N=4096 T=128 D=64 H=256", host_synth, 4096, 128, 256)
run("device synthetic      N=4096 T=128 D=64 H=256", DeviceVectorEnv.factory("Synthetic", obs_dim=64, n_actions=4), 4096, 128, 256)
run("host CartPole (SyncVectorEnv) N=64 T=128 H=64", lambda: envs.make("CartPole-v1"), 64, 128, 64)
run("device CartPole       N=64 T=128 H=64", DeviceVectorEnv.factory("CartPole-v1"), 64, 128, 64)
run("device CartPole       N=4096 T=128 H=64", DeviceVectorEnv.factory("CartPole-v1"), 4096, 128, 64)
