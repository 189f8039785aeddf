"""learn() latency of the small (launch-latency-bound) configurations of BASELINE.json next to the CPU port of the reference
algorithm (oracle/ppo_oracle.py learn(), 1 host thread: more threads are slower at these sizes).  lives under tests/ because it runs the oracle (checker / CPU baseline only)."""
import sys, os, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
from diamond import PPO, PPOConfig, ContinuousPPO, ContinuousPPOConfig, envs
from oracle import ppo_oracle as O

def run(name, D, A, H, N, T, cont):
    rng = np.random.default_rng(0)
    Agent, Cfg = (ContinuousPPO, ContinuousPPOConfig) if cont else (PPO, PPOConfig)
    cfg = Cfg(num_envs=N, rollout_steps=T, network_hidden_dim=H, verbose=False, seed=1, total_steps=N * T * 1000)
    agent = Agent(lambda: envs.SyntheticEnv(D, A, continuous=cont), cfg)
    exp = [[rng.standard_normal((N, D)).astype(np.float32), rng.standard_normal((N, D)).astype(np.float32),
            rng.standard_normal((N, A)).astype(np.float32) if cont else rng.integers(0, A, N), rng.standard_normal(N),
            rng.random(N) < 0.02, rng.random(N) < 0.02] for _ in range(T)]
    from diamond.agents import RolloutBuffer
    buf = RolloutBuffer.from_lists(agent.ctx, exp, cont, agent.device)
    np.random.seed(0)
    for _ in range(3): agent.learn(buf)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); reps = 10
    for _ in range(reps): agent.learn(buf)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    # CPU port of the reference algorithm
    torch.set_num_threads(1)
    names = O.CONTINUOUS_PARAM_NAMES if cont else O.DISCRETE_PARAM_NAMES
    p = {n: q.detach().cpu().clone() for n, q in agent.network.named_parameters()}
    p = {n: p[n] for n in names}
    state = O.new_adam_state(p, names)
    obs, nobs, act, rew, term, trunc = (np.asarray(x) for x in zip(*exp))
    ocfg = O.default_cfg()
    E, B = ocfg["num_epochs"], T * N
    perms = np.stack([np.random.permutation(B) for _ in range(E)])
    O.learn(p, state, obs, nobs, act, rew, term, trunc, ocfg, perms, cont)
    t0 = time.perf_counter()
    for _ in range(3): O.learn(p, state, obs, nobs, act, rew, term, trunc, ocfg, perms, cont)
    cpu_ms = (time.perf_counter() - t0) * 1e3 / 3
    print(f"{name:34s} learn() {ms:7.2f} ms ({E * B / ms * 1e3 / 1e6:6.2f} M sample-updates/s)   CPU port, 1 thread: {cpu_ms:8.1f} ms  ({cpu_ms / ms:5.1f}x)", flush=True)

run("C  CartPole  D=4 A=2 N=8 T=128", 4, 2, 64, 8, 128, False)
run("L  LunarLander D=8 A=4 N=8 T=128", 8, 4, 64, 8, 128, False)
run("Pn Pendulum  D=3 act=1 N=64 T=64", 3, 1, 64, 64, 64, True)
