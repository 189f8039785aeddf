"""Small driver for ncu: one recurrent learn() (scan kernels) and one device rollout (environment + sampling kernels)."""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200"))
from diamond import PPO, PPOConfig, RecurrentPPO, RecurrentPPOConfig, envs
from diamond.envs import DeviceVectorEnv
from diamond.recurrent import RecurrentRollout
N, T, D, A, H, Hg = 4096, 128, 64, 4, 256, 64
cfg = RecurrentPPOConfig(num_envs=N, rollout_steps=T, num_epochs=1, num_minibatches=4, verbose=False, network_hidden_dim=H, gru_hidden_dim=Hg, seed=1)
agent = RecurrentPPO(lambda: envs.SyntheticEnv(D, A), cfg)
dev = agent.device
ro = RecurrentRollout(T, N, D, Hg, dev)
g = torch.Generator(device=dev).manual_seed(0)
ro.obs.normal_(generator=g); ro.actions.random_(0, A, generator=g); ro.rewards.normal_(generator=g)
ro.terminations.copy_((torch.rand(T, N, device=dev, generator=g) < 0.02).float()); ro.truncations.zero_()
ro.prev_dones.copy_(torch.rand(T, N, device=dev, generator=g) < 0.02)
ro.log_probs.fill_(-np.log(A)); ro.values.normal_(generator=g); ro.next_values.normal_(generator=g); ro.filled = T
np.random.seed(0)
agent.learn(ro); agent.learn(ro)
torch.cuda.synchronize()
cfg2 = PPOConfig(num_envs=N, rollout_steps=8, network_hidden_dim=H, verbose=False, seed=1)
for env_id, kw in (("CartPole-v1", {}), ("Synthetic", dict(obs_dim=D, n_actions=A))):
    a2 = PPO(DeviceVectorEnv.factory(env_id, seed=1, **kw), cfg2)
    a2.ticker = None
    a2.current_observations, _ = a2.envs.reset(seed=1)
    a2.rollout(); a2.rollout()
torch.cuda.synchronize()
print("done")
