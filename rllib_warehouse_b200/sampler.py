"""RolloutSampler — a ray-free stand-in for RLlib's sampler loop (`ray/rllib/evaluation/sampler.py`,
`_env_runner`: poll the BaseEnv, batch the per-agent observations, evaluate the policy, send the actions
back, reset finished envs, collect episode returns), used to exercise and to time `WarehouseVectorEnv`
where `ray` is not installed (`scripts/train.py:29-43` is the reference's entry to the real one).

Two modes:
  * "base_env": strictly through the BaseEnv protocol (`poll` / `send_actions` / `try_reset`): nested
    env_id -> agent_id dicts of host arrays — what RLlib's own sampler would drive. The cost is dominated
    by building and walking those Python dicts, exactly as it is for RLlib with any vector env;
  * "tensor":   the adapter's zero-copy API (`reset_tensors` / `step_tensors`): observations stay in HBM,
    one step kernel + one policy forward per iteration — what a vectorised torch sampler uses.
The policy is any callable mapping float32 observations [B, 9R+1] on the env's device to logits [B, 9].
"""
import time

import numpy as np
import torch


class RolloutSampler:
    def __init__(self, venv, policy):
        assert venv.flat_obs, "the sampler feeds the policy RLlib-flattened observations (flat_obs=True)"
        self.venv, self.policy = venv, policy
        self.device = venv.env.device
        self.episode_returns = []            # per finished episode: mean per-agent return (train.py:18-23)

    # ------------------------------------------------------------------------------------------
    def _act(self, batch):
        with torch.no_grad():
            return self.policy(batch).argmax(dim=-1)

    def run_base_env(self, iterations):
        """`iterations` sampler iterations through poll / send_actions / try_reset. Returns a dict with
        env_steps, agent_steps, seconds."""
        venv = self.venv
        returns = {}
        obs, _, _, _, _ = venv.poll()
        if not obs:                          # already initialised by an earlier run: start from fresh resets
            obs = {e: venv.try_reset(e) for e in range(venv.num_envs)}
        env_steps = agent_steps = 0
        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        for _ in range(iterations):
            keys = [(e, a) for e, agents in obs.items() for a in agents]
            batch = torch.from_numpy(np.stack([obs[e][a] for e, a in keys])).to(self.device, non_blocking=True)
            acts = self._act(batch).cpu().numpy()
            action_dict = {e: {} for e in obs}
            for (e, a), v in zip(keys, acts):
                action_dict[e][a] = int(v)
            venv.send_actions(action_dict)
            obs, rewards, dones, _, _ = venv.poll()
            env_steps += len(obs)
            agent_steps += len(keys)
            for e, agent_rewards in rewards.items():
                acc = returns.setdefault(e, {})
                for a, r in agent_rewards.items():
                    acc[a] = acc.get(a, 0.0) + float(r)
                if dones[e]["__all__"]:
                    self.episode_returns.append(sum(acc.values()) / len(acc))
                    returns[e] = {}
                    obs[e] = venv.try_reset(e)
        torch.cuda.synchronize(self.device)
        return dict(env_steps=env_steps, agent_steps=agent_steps, seconds=time.perf_counter() - t0, mode="base_env")

    def run_tensor(self, iterations):
        """`iterations` sampler iterations through reset_tensors / step_tensors (the env must have been built
        with auto_reset=True: finished envs restart in-kernel and show their reset observation)."""
        venv, env = self.venv, self.venv.env
        N, R, F = env.N, env.R, venv.F
        flat = venv.reset_tensors()
        ret = torch.zeros((N, R), device=self.device)
        alive = (torch.arange(R, device=self.device)[None, :] < env.state["num_agents"][:, None]).float()
        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        agent_rows = 0
        for _ in range(iterations):
            actions = self._act(flat.view(-1, F)).view(N, R).to(torch.int32)
            flat, rewards, dones = venv.step_tensors(actions)
            ret += rewards
            agent_rows += N * R
            d = dones.bool()
            if bool(d.any()):                # episode bookkeeping only when something finished (one host sync per step)
                mean_ret = (ret * alive).sum(dim=1) / alive.sum(dim=1)
                self.episode_returns.extend(mean_ret[d].tolist())
                ret[d] = 0
                alive = (torch.arange(R, device=self.device)[None, :] < env.state["num_agents"][:, None]).float()
        torch.cuda.synchronize(self.device)
        return dict(env_steps=N * iterations, agent_steps=agent_rows, seconds=time.perf_counter() - t0, mode="tensor")


def mlp_policy(obs_dim, hidden=256, device="cuda:0", dtype=torch.float32, seed=0):
    """The fcnet the PPO specs under scripts/experiments/*-ppo configure (2 x 256, ReLU, 9 logits)."""
    torch.manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(obs_dim, hidden), torch.nn.ReLU(), torch.nn.Linear(hidden, hidden),
                              torch.nn.ReLU(), torch.nn.Linear(hidden, 9)).to(device=device, dtype=dtype)
    if dtype == torch.float32:
        return net
    return lambda x: net(x.to(dtype))
