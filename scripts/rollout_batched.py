#!/usr/bin/env python
"""Batched evaluation on the GPU: thousands of episodes at once, observations never
leave HBM. The policy is either the greedy baseline (in-kernel) or a TorchScript module mapping
RLlib-flattened float32 observations [B, 9R+1] to action logits [B, 9].

    python scripts/rollout_batched.py large --envs 65536 --episodes 2 [--policy policy.pt] [--train-variant]
    torchrun --nproc-per-node 8 scripts/rollout_batched.py large --envs 262144     # sharded over GPUs

Prints the reference's episode metrics (avg_agent_reward_all / avg_agent_reward_{n}, train.py:18-23)
reduced over all GPUs with one NCCL all-reduce at the end of the rollout.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(a):
    import torch
    from rllib_warehouse_b200 import VARIANTS, BatchedWarehouse
    from rllib_warehouse_b200.parallel import allreduce_stats, init_from_env, shard_range, stats_to_metrics
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rank, world = init_from_env(device=dev)
    lo, hi = shard_range(a.envs, rank, world)
    cfg = VARIANTS[a.env_size].replace(random_num_agents=a.train_variant)
    env = BatchedWarehouse(cfg, hi - lo, num_agents=a.num_agents, device=dev, seed=a.seed, env_id0=lo,
                           auto_reset=True)
    policy = torch.jit.load(a.policy, map_location=dev).eval() if a.policy else None
    env.reset()
    steps = a.episodes * cfg.episode_duration
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        if policy is None:
            for _ in range(steps):
                env.greedy_step(random_action_prob=a.random_action_prob, solver_seed=a.seed + 1, want_actions=False)
        else:
            # one launch per step: the step kernel itself emits the RLlib-flattened observations the policy
            # reads (reset-flavour ones for envs that were just auto-reset); nothing is synchronised per step
            flat = env.build_obs_flat(1)
            for _ in range(steps):
                logits = policy(flat.view(-1, flat.shape[-1]))
                actions = logits.argmax(dim=-1).view(env.N, env.R).to(torch.int32)
                flat, _, _ = env.step_flat(actions)
    stats = allreduce_stats(env.stats)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        m = stats_to_metrics(stats, cfg.num_requests)
        print(f"{a.envs} envs x {steps} steps on {world} GPU(s): {dt:.2f}s")
        for k, v in m.items():
            print(f"  {k}: {v:.4f}" if isinstance(v, float) else f"  {k}: {v}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("env_size", choices=["small", "medium", "large"])
    ap.add_argument("--envs", type=int, default=65536, help="total environments (sharded over ranks)")
    ap.add_argument("--episodes", type=int, default=1)
    ap.add_argument("--num-agents", type=int, default=None)
    ap.add_argument("--train-variant", action="store_true", help="random agent count per episode (*Train)")
    ap.add_argument("--policy", type=str, default=None, help="TorchScript policy file (default: greedy solver)")
    ap.add_argument("--random-action-prob", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=0)
    main(ap.parse_args())
