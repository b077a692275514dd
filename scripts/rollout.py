#!/usr/bin/env python
"""Checkpoint rollout with the reference's command line (scripts/rollout.py:90-113):

    python scripts/rollout.py EXPERIMENT_DIR NUM_AGENTS [-c CHECKPOINT] [-n EPISODES] [--no-render]

Restores the RLlib trainer of a Tune trial (params.json + checkpoint_N/checkpoint-N, picking the
latest or the one closest to -c) and evaluates it on the fixed-size variant of the trained
environment with NUM_AGENTS agents. Needs ray[rllib]; without it the script says so and exits 2.
For policies that are plain torch modules use scripts/rollout_batched.py, which keeps observations
on the GPU and evaluates thousands of episodes at once.
"""
import argparse
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pick_checkpoint(experiment_dir, wanted):
    found = {}
    for name in os.listdir(experiment_dir):
        m = re.fullmatch(r"checkpoint_(\d+)", name)
        if m:
            found[int(m.group(1))] = os.path.join(experiment_dir, name, f"checkpoint-{m.group(1)}")
    if not found:
        raise FileNotFoundError(f"no checkpoint_N directories under {experiment_dir}")
    key = max(found) if wanted is None else min(found, key=lambda k: abs(k - wanted))
    return key, found[key]


def main(a):
    try:
        import ray
        from ray.rllib.agents.registry import get_agent_class
        from ray.tune.registry import register_env
    except ImportError as e:
        print(f"scripts/rollout.py needs ray[rllib] ({e}); it is not installed here.", file=sys.stderr)
        return 2
    import functools
    from warehouse import (WarehouseLarge, WarehouseLargeTrain, WarehouseMedium, WarehouseMediumTrain,
                           WarehouseSmall, WarehouseSmallTrain)
    with open(os.path.join(a.experiment_dir, "params.json")) as f:
        params = json.load(f)
    train_envs = {"WarehouseSmall-v0": (WarehouseSmallTrain, WarehouseSmall),
                  "WarehouseMedium-v0": (WarehouseMediumTrain, WarehouseMedium),
                  "WarehouseLarge-v0": (WarehouseLargeTrain, WarehouseLarge)}
    ray.init()
    for env_id, (train_cls, _) in train_envs.items():
        register_env(env_id, functools.partial(lambda cls, _cfg: cls(), train_cls))
    number, path = pick_checkpoint(a.experiment_dir, a.checkpoint)
    print(f"restoring checkpoint {number}: {path}")
    trainer = get_agent_class(a.run)(env=params["env"], config=params)   # the reference hard-codes SAC (rollout.py:46)
    trainer.restore(path)
    env = train_envs[params["env"]][1](a.num_agents)
    for ep in range(a.num_episodes):
        obs, done = env.reset(), False
        returns = {str(i): 0.0 for i in range(env.num_agents)}
        while not done:
            if not a.no_render:
                env.render()
            actions = {agent: trainer.compute_action(ob) for agent, ob in obs.items()}
            obs, rewards, dones, _ = env.step(actions)
            for k, r in rewards.items():
                returns[k] += float(r)
            done = dones["__all__"]
        total = sum(returns.values())
        print(f"episode {ep}: total {total}, per agent {total / len(returns)}")
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("experiment_dir", type=str, help="Tune trial directory (contains params.json)")
    ap.add_argument("num_agents", type=int)
    ap.add_argument("-c", "--checkpoint", type=int, default=None, help="checkpoint number (default: latest)")
    ap.add_argument("-n", "--num-episodes", type=int, default=1)
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--run", type=str, default="SAC", help="RLlib algorithm the trial was trained with")
    sys.exit(main(ap.parse_args()))
