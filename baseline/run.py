#!/usr/bin/env python
"""Greedy-baseline driver with the reference's command line (baseline/run.py:79-99):

    python baseline/run.py {small,medium,large} NUM_AGENTS RANDOM_ACTION_PROB [-r] [--envs N]

Without --envs it runs the single-env MultiAgentEnv loop exactly like the reference (reset, then
solver.compute_action -> env.step until dones["__all__"], checking every observation against
observation_space) on the CUDA environment. With --envs N it runs the same episode for N
environments at once through the batched tensors and prints aggregate statistics.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from warehouse import WarehouseLarge, WarehouseMedium, WarehouseSmall  # noqa: E402
from solvers import WarehouseRandomGreedySolver  # noqa: E402

ENV_TYPES = {"small": WarehouseSmall, "medium": WarehouseMedium, "large": WarehouseLarge}


def check_observations(env, observations):
    for agent_id, ob in observations.items():
        assert env.observation_space.contains(ob), f"observation of agent {agent_id} is outside the space"


def run_single(env_size, num_agents, random_action_prob, render):
    env = ENV_TYPES[env_size](num_agents)
    solver = WarehouseRandomGreedySolver(env.num_agents, env.num_requests, random_action_prob, env.action_space)
    observations = env.reset()
    check_observations(env, observations)
    returns = {f"{i}": 0.0 for i in range(env.num_agents)}
    think_s = step_s = 0.0
    steps, done = 0, False
    while not done:
        if render:
            env.render(animate=True)
        t0 = time.perf_counter()
        actions = solver.compute_action(observations)
        t1 = time.perf_counter()
        observations, rewards, dones, _ = env.step(action_dict=actions)
        t2 = time.perf_counter()
        think_s, step_s = think_s + (t1 - t0), step_s + (t2 - t1)
        check_observations(env, observations)
        for k, r in rewards.items():
            returns[k] += float(r)
        done = dones["__all__"]
        steps += 1
    total = sum(returns.values())
    print(f"\n=== Done ({steps} steps) ===")
    print("Rewards:", *returns.values())
    print(f"Total: {total}, Per Agent: {total / len(returns)}")
    print(f"Step avg FPS: {steps / step_s:.1f}, think avg FPS: {steps / think_s:.1f}")
    return total


def run_batched(env_size, num_agents, random_action_prob, num_envs, rollout_kernel=False):
    import torch
    from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse
    env = BatchedWarehouse(VARIANTS[env_size], num_envs, num_agents=num_agents, seed=int(time.time()))
    env.reset()
    total = torch.zeros(num_envs, device=env.device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = 0
    if rollout_kernel:
        # the whole episode in one launch: state in registers, no per-step observations
        steps = VARIANTS[env_size].episode_duration
        total = env.greedy_rollout(steps, random_action_prob=random_action_prob, solver_seed=1,
                                   with_obs=False).sum(dim=1)
    else:
        while True:
            _, rewards, dones = env.greedy_step(random_action_prob=random_action_prob, solver_seed=1)
            total += rewards.sum(dim=1)
            steps += 1
            if bool(dones[0].item()):
                break
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"\n=== Done ({steps} steps x {num_envs} envs) ===")
    print(f"Return per env: mean {total.mean().item():.2f} std {total.std().item():.2f}; "
          f"per agent {total.mean().item() / num_agents:.2f}")
    what = "solver + step, one launch per episode" if rollout_kernel else "solver + step + observations"
    print(f"{num_envs * num_agents * steps / dt:.3e} agent-steps/s ({what})")
    print("Episode statistics:", env.stats_dict())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("env_size", type=str, choices=list(ENV_TYPES), help="environment size")
    ap.add_argument("num_agents", type=int, help="number of agents")
    ap.add_argument("random_action_prob", type=float, help="probability of a random action [0.0, 1.0]")
    ap.add_argument("-r", "--render", action="store_true", help="render the environment on each step")
    ap.add_argument("--envs", type=int, default=0, help="run N environments at once on the GPU")
    ap.add_argument("--rollout-kernel", action="store_true",
                    help="with --envs: the whole episode in one kernel launch (no per-step observations)")
    a = ap.parse_args()
    if a.envs:
        run_batched(a.env_size, a.num_agents, a.random_action_prob, a.envs, a.rollout_kernel)
    else:
        run_single(a.env_size, a.num_agents, a.random_action_prob, a.render)
