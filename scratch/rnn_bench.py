"""RecurrentPPO.learn() latency: fused recurrent kernels (csrc/rnn.cu) vs the autograd engine (custom network_cls path)."""
import sys, os, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
from diamond import RecurrentPPO, RecurrentPPOConfig, envs
from diamond.recurrent import RecurrentActorCriticNetwork, RecurrentRollout

class Custom(RecurrentActorCriticNetwork):
    pass

def run(name, D, A, H, Hg, N, T, E, MB, cls):
    cfg = RecurrentPPOConfig(num_envs=N, rollout_steps=T, num_epochs=E, num_minibatches=MB, verbose=False, network_hidden_dim=H,
                             gru_hidden_dim=Hg, seed=1, total_steps=N * T * 1000)
    agent = RecurrentPPO(lambda: envs.SyntheticEnv(D, A), cfg, network_cls=cls)
    dev = agent.device
    ro = RecurrentRollout(T, N, D, Hg, dev)
    g = torch.Generator(device=dev).manual_seed(0)
    ro.obs.normal_(generator=g); ro.actions.random_(0, A, generator=g); ro.rewards.normal_(generator=g)
    ro.terminations.copy_((torch.rand(T, N, device=dev, generator=g) < 0.02).float()); ro.truncations.zero_()
    ro.prev_dones.copy_(torch.rand(T, N, device=dev, generator=g) < 0.02)
    ro.log_probs.fill_(-np.log(A)); ro.values.normal_(generator=g); ro.next_values.normal_(generator=g); ro.filled = T
    np.random.seed(0)
    for _ in range(2): agent.learn(ro)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); reps = 3
    for _ in range(reps): agent.learn(ro)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    print(f"{name:28s} {'fused' if cls is RecurrentActorCriticNetwork else 'autograd':8s} learn() {ms:9.2f} ms  "
          f"{E * MB * 1e3 / ms:8.1f} optimiser steps/s  {E * T * N / ms * 1e3 / 1e6:8.3f} M sample-updates/s", flush=True)

for cls in (RecurrentActorCriticNetwork, Custom):
    run("R (N=32,T=32,Hg=16,H=64)", 4, 2, 64, 16, 32, 32, 10, 1, cls)
    run("N=1024,T=128,Hg=32,H=128", 16, 4, 128, 32, 1024, 128, 4, 4, cls)
    run("N=4096,T=128,Hg=64,H=256", 64, 4, 256, 64, 4096, 128, 2, 4, cls)
