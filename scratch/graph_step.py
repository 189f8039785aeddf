import sys, os, time, numpy as np, torch
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
host = bench.synth_host_rollout(1)
buf = RolloutBuffer(ctx, T, N_ENVS, D, 1, False, agent.device)
buf.load_host(*host)
np.random.seed(123)
for _ in range(2): agent.learn(buf)
torch.cuda.synchronize()
eng = agent.engine
b = eng._alloc(T, N_ENVS, E, MB)
B = T * N_ENVS; M = B // MB
hyper = eng._hyper(cfg, M, B); hyper.step = 100
obs_flat = buf.obs.view(B, D); actions = buf.actions.view(B)
old_logp, adv, ret = b["old_logp"].view(B), b["adv"].view(B), b["ret"].view(B)
losses = b["losses"]
def step(k):
    ctx.mlp_grad_minibatch(eng.fm.desc, eng.P, eng.G, obs_flat, actions, old_logp, adv, ret, b["stats"], b["idx"][0][k*M:(k+1)*M], M, hyper, losses[k], b["train_ws"])
    ctx.clip_adam_step(eng.P, eng.G, eng.M, eng.V, hyper, eng.adam_ws, eng.grad_norm)
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def eager():
    for k in range(MB): step(k)
print("eager 8 steps: %.3f ms -> %.1f us/step" % (timeit(eager), timeit(eager) * 1e3 / MB))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for k in range(MB): step(k)
print("graph 8 steps: %.3f ms -> %.1f us/step" % (timeit(g.replay), timeit(g.replay) * 1e3 / MB))
