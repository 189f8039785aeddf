"""GPU: device-resident vector environments (csrc/envs.cu) against oracle/env_oracle.py, and the PCIe-free rollout path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env_id", ["CartPole-v1", "Pendulum-v1"])
def test_device_env_steps_match_oracle(env_id):
    """Every step of 600 vector steps: the kernel's next state / observation / reward / flags equal the oracle's step from the
    same float64 state and actions (flags bit-exact, floats to 1e-12); autoreset-disabled contract of the vector API."""
    from diamond.envs import DeviceVectorEnv
    from oracle import env_oracle as EO
    N_ = 96
    env = DeviceVectorEnv(env_id, N_, seed=5)
    step_fn = EO.cartpole_step if env_id == "CartPole-v1" else EO.pendulum_step
    rng = np.random.default_rng(1)
    obs0, _ = env.reset(seed=5)
    s = env.state.cpu().numpy()
    if env_id == "CartPole-v1":
        assert np.all(np.abs(s) <= 0.05) and np.array_equal(obs0, s.astype(np.float32))
    else:
        assert np.all(np.abs(s[:, 0]) <= np.pi) and np.all(np.abs(s[:, 1]) <= 1.0)
        np.testing.assert_allclose(obs0, np.stack([np.cos(s[:, 0]), np.sin(s[:, 0]), s[:, 1]], 1).astype(np.float32), atol=1e-7)
    n_done = 0
    for it in range(600):
        state = env.state.cpu().numpy().copy()
        steps = env.steps.cpu().numpy().copy()
        actions = rng.integers(0, 2, N_) if env_id == "CartPole-v1" else rng.uniform(-2.5, 2.5, (N_, 1)).astype(np.float32)
        nobs, rew, term, trunc, _ = env.step(actions)
        ref_next, ref_obs, ref_rew, ref_term, ref_trunc = step_fn(state, actions, steps)
        np.testing.assert_array_equal(term, ref_term)
        np.testing.assert_array_equal(trunc, ref_trunc)
        np.testing.assert_allclose(env.state.cpu().numpy()[:, :ref_next.shape[1]], ref_next, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(nobs, ref_obs, rtol=0, atol=1e-6)
        np.testing.assert_allclose(rew, ref_rew, rtol=1e-6, atol=1e-6)            # rewards are stored as f32 (ppo.py:230)
        dones = term | trunc
        if dones.any():
            n_done += int(dones.sum())
            before = env.state.cpu().numpy().copy()
            obs, _ = env.reset(options={"reset_mask": dones})
            after = env.state.cpu().numpy()
            assert np.array_equal(after[~dones], before[~dones])                     # only the masked environments are reset
            assert np.all(env.steps.cpu().numpy()[dones] == 0)
            assert np.array_equal(obs[~dones], nobs[~dones])
    assert n_done > 0


def test_device_env_reset_draws_are_keyed_by_seed_and_global_env_id():
    from diamond.envs import DeviceVectorEnv
    a = DeviceVectorEnv("CartPole-v1", 32, seed=9)
    b = DeviceVectorEnv("CartPole-v1", 16, seed=9, env_offset=16)          # the second shard of a 32-env run
    c = DeviceVectorEnv("CartPole-v1", 32, seed=10)
    assert torch.equal(a.state[16:], b.state)
    assert not torch.equal(a.state, c.state)


@pytest.mark.parametrize("env_id", ["CartPole-v1", "Pendulum-v1", "LunarLander-v3"])
def test_rollout_on_device_envs_keeps_rollout_semantics(env_id):
    """agent.rollout() with a DeviceVectorEnv: no host env stepping; the buffer obeys ppo.py:165-185 (obs[t+1] is next_obs[t]
    unless the env finished at t, in which case it is a freshly reset observation), and learn() consumes it."""
    from diamond import PPO, PPOConfig, ContinuousPPO, ContinuousPPOConfig
    from diamond.envs import DeviceVectorEnv
    cont = env_id == "Pendulum-v1"
    Agent, Cfg = (ContinuousPPO, ContinuousPPOConfig) if cont else (PPO, PPOConfig)
    T, N_ = 64 if cont else 128, 64
    extra = dict(p_term=0.03, p_trunc=0.02) if env_id == "LunarLander-v3" else {}
    cfg = Cfg(num_envs=N_, rollout_steps=T, verbose=False, seed=3, total_steps=T * N_ * 4)
    agent = Agent(DeviceVectorEnv.factory(env_id, seed=3, **extra), cfg)
    agent.current_observations, _ = agent.envs.reset(seed=3)
    first = agent.envs.cur_obs.clone()
    buf = agent.rollout()
    torch.cuda.synchronize()
    assert torch.equal(buf.obs[0], first)
    done = (buf.terminations + buf.truncations) > 0
    cont_rows = ~done[:-1]
    assert torch.equal(buf.obs[1:][cont_rows], buf.next_obs[:-1][cont_rows])
    if done[:-1].any():
        assert not torch.equal(buf.obs[1:][done[:-1]], buf.next_obs[:-1][done[:-1]])
    assert torch.equal(agent.envs.cur_obs[~done[-1]], buf.next_obs[-1][~done[-1]])
    if env_id == "CartPole-v1":
        assert done.any() and torch.all(buf.rewards == 1.0)
        assert torch.all((buf.next_obs[..., 0].abs() > 2.4) | (buf.next_obs[..., 2].abs() > 12 * 2 * np.pi / 360) == (buf.terminations > 0))
    before = {k: v.detach().clone() for k, v in agent.network.state_dict().items()}
    agent.learn(buf)
    assert torch.isfinite(agent.last_losses).all()
    assert any(not torch.equal(before[k], v) for k, v in agent.network.state_dict().items())


def test_cartpole_learns_on_device():
    """End-to-end train() on device-resident CartPole-v1: the mean episode length grows well beyond the random policy's ~22."""
    from diamond import PPO, PPOConfig
    from diamond.envs import DeviceVectorEnv
    cfg = PPOConfig(num_envs=64, rollout_steps=128, verbose=False, seed=1, total_steps=64 * 128 * 40)
    agent = PPO(DeviceVectorEnv.factory("CartPole-v1", seed=1), cfg)
    agent.train()
    lengths = list(agent.ticker.recent_lengths)
    assert len(lengths) > 10 and np.mean(lengths) > 100.0, np.mean(lengths)


@pytest.mark.parametrize("env_id", ["CartPole-v1", "Pendulum-v1", "Synthetic"])
def test_learn_reuses_rollout_values_and_matches_full_prepass(env_id):
    """with log_prob(action) / V(obs) recorded at sampling time and next_values[t] = values[t+1] where the env did
    not finish, learn() skips the pre-update pass; the result equals the full pass (ppo.py:235-238) on the same buffer."""
    from diamond import PPO, PPOConfig, ContinuousPPO, ContinuousPPOConfig
    from diamond.envs import DeviceVectorEnv
    cont = env_id == "Pendulum-v1"
    Agent, Cfg = (ContinuousPPO, ContinuousPPOConfig) if cont else (PPO, PPOConfig)
    N_, T, H = (2048, 32, 128) if env_id == "Synthetic" else (64, 64, 64)           # Synthetic: tensor-core forward path
    kw = dict(obs_dim=32, n_actions=4, p_term=0.03, p_trunc=0.02) if env_id == "Synthetic" else {}

    def make_agent():
        cfg = Cfg(num_envs=N_, rollout_steps=T, network_hidden_dim=H, verbose=False, seed=11, total_steps=T * N_ * 8)
        agent = Agent(DeviceVectorEnv.factory(env_id, seed=11, **kw), cfg)
        agent.current_observations, _ = agent.envs.reset(seed=11)
        return agent, cfg

    def run(reuse):
        agent, cfg = make_agent()
        agent.engine.reuse_rollout_values = reuse
        buf = agent.rollout()
        assert buf.policy_stamp == agent.engine.policy_stamp()
        np.random.seed(4)
        agent.learn(buf)
        torch.cuda.synchronize()
        return agent.last_losses.clone(), torch.cat([q.detach().flatten() for q in agent.network.parameters()]).clone()

    # identical parameters and seeds -> identical rollouts; one learn() with and without the recorded values
    (la, pa), (lf, pf) = run(True), run(False)
    assert torch.allclose(la, lf, rtol=1e-4, atol=5e-6)
    err = (pa - pf).abs().max().item() / pf.abs().max().item()
    assert err <= 1e-4, err

    # GAE inputs of the two pre-update passes on the same buffers, over several rollout / learn iterations
    agent, cfg = make_agent()
    eng = agent.engine
    for it in range(3):
        buf = agent.rollout()
        b = eng._alloc(T, N_, cfg.num_epochs, cfg.num_minibatches)
        eng.prepass(buf, b)
        full = {k: b[k].clone() for k in ("values", "next_values", "old_logp")}
        for k in full:
            b[k].fill_(float("nan"))
        eng.prepass_from_rollout(buf, b)
        for k, ref in full.items():
            assert (b[k] - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()), (it, k)
        agent.learn(buf)

    # a buffer whose sampling policy is no longer the current one falls back to the full pass
    from diamond.agents import FusedMlpEngine
    cfg = Cfg(num_envs=N_, rollout_steps=T, network_hidden_dim=H, verbose=False, seed=11, total_steps=T * N_ * 8)
    agent = Agent(DeviceVectorEnv.factory(env_id, seed=11, **kw), cfg)
    agent.current_observations, _ = agent.envs.reset(seed=11)
    buf = agent.rollout()
    with torch.no_grad():
        next(agent.network.parameters()).mul_(1.01)
    assert buf.policy_stamp != agent.engine.policy_stamp()


@pytest.mark.parametrize("env_id", ["CartPole-v1", "Pendulum-v1", "Synthetic"])
def test_rollout_graph_replay_is_bit_identical_to_eager(env_id):
    """From the third rollout on, the T x (forward, head, sampling, environment) launches of rollout() are replayed as one CUDA graph
    whose draw counters are relative to a device-resident base: buffers, recorded values / log-probs, environment state and
    the parameters after learn() must equal the eagerly launched run bit for bit."""
    from diamond import PPO, PPOConfig, ContinuousPPO, ContinuousPPOConfig
    from diamond.envs import DeviceVectorEnv
    cont = env_id == "Pendulum-v1"
    Agent, Cfg = (ContinuousPPO, ContinuousPPOConfig) if cont else (PPO, PPOConfig)
    N_, T, H = (2048, 16, 128) if env_id == "Synthetic" else (64, 32, 64)
    kw = dict(obs_dim=32, n_actions=4, p_term=0.03, p_trunc=0.02) if env_id == "Synthetic" else {}

    def run(use_graphs):
        cfg = Cfg(num_envs=N_, rollout_steps=T, network_hidden_dim=H, verbose=False, seed=13, total_steps=T * N_ * 16)
        agent = Agent(DeviceVectorEnv.factory(env_id, seed=13, **kw), cfg)
        agent.engine.use_graphs = use_graphs
        agent.current_observations, _ = agent.envs.reset(seed=13)
        out = []
        for it in range(5):
            buf = agent.rollout()
            out.append([x.clone() for x in (buf.obs, buf.next_obs, buf.actions, buf.rewards, buf.terminations, buf.truncations,
                                            buf.values, buf.logp, agent.envs.state, agent.envs.cur_obs)])
            np.random.seed(it)
            agent.learn(buf)
            out[-1].append(torch.cat([q.detach().flatten() for q in agent.network.parameters()]).clone())
        replayed = agent._rollout_graph is not None
        return out, replayed, agent.engine.draws

    (a, ra, da), (b, rb, db) = run(True), run(False)
    assert ra and not rb and da == db == 5 * T
    for it, (xa, xb) in enumerate(zip(a, b)):
        for k, (u, v) in enumerate(zip(xa, xb)):
            assert torch.equal(u, v), (it, k)


@pytest.mark.parametrize("env_id,N_,T,rollouts", [("CartPole-v1", 64, 128, 4), ("CartPole-v1", 2500, 32, 3), ("Pendulum-v1", 16, 250, 3),
                                                  ("Synthetic", 4096, 16, 2)])
def test_device_episode_statistics_equal_host_ticker(env_id, N_, T, rollouts):
    """the device-side episode statistics (dppo_episode_stats, consumed lazily once per rollout) leave the Ticker in
    the state the reference's per-step `ticker.tick(rewards, dones)` (utils.py:99-123, ppo.py:181-182) leaves it in: same step and
    episode counters, same window of the last 100 (return, length) pairs in the same order -- returns bit-equal (same fp64 sums)."""
    from diamond import PPO, PPOConfig, ContinuousPPO, ContinuousPPOConfig
    from diamond.envs import DeviceVectorEnv
    from diamond.utils import Ticker
    cont = env_id == "Pendulum-v1"
    Agent, Cfg = (ContinuousPPO, ContinuousPPOConfig) if cont else (PPO, PPOConfig)
    kw = dict(obs_dim=8, n_actions=4, p_term=0.002, p_trunc=0.001) if env_id == "Synthetic" else {}
    cfg = Cfg(num_envs=N_, rollout_steps=T, verbose=False, seed=3, total_steps=T * N_ * 100)
    agent = Agent(DeviceVectorEnv.factory(env_id, seed=3, **kw), cfg)
    agent.current_observations, _ = agent.envs.reset(seed=3)
    host = Ticker(cfg.total_steps, N_, T, verbose=False)
    for _ in range(rollouts):
        buf = agent.rollout()
        rew = buf.rewards.double().cpu().numpy()
        dones = ((buf.terminations + buf.truncations) > 0).cpu().numpy()
        for t in range(T):
            host.tick(rew[t], dones[t])                     # the reference's per-step bookkeeping, replayed on the host
    got, ref = agent.ticker.logs, host.logs
    assert got["total_steps"] == ref["total_steps"] and got["total_episodes"] == ref["total_episodes"]
    assert got["episode_lengths"] == ref["episode_lengths"]
    assert got["episode_returns"] == ref["episode_returns"]
    if env_id != "Pendulum-v1" or T * rollouts >= 200:
        assert ref["total_episodes"] > 0
    # running sums of unfinished episodes carried across rollouts on the device
    np.testing.assert_array_equal(agent._epstats["ep_len"].cpu().numpy(), host.current_lengths)
    np.testing.assert_array_equal(agent._epstats["ep_return"].cpu().numpy(), host.current_returns)
