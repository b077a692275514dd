#!/usr/bin/env python
"""RLlib training driver with the reference's command line (scripts/train.py:45-50):

    python scripts/train.py scripts/experiments/warehouse-small-ppo/warehouse-small-ppo.yaml

Registers "Warehouse{Small,Medium,Large}-v0" -> the *Train variants (random agent count per
episode, variants.py:65-98) backed by the CUDA environment, injects the reference's per-episode
metrics (avg_agent_reward_all / avg_agent_reward_{n}, train.py:18-23) and hands the Tune
experiment dict to `run_experiments`. The algorithm is whatever the YAML's `run:` names.

Needs `ray[rllib]` (0.8.x API, as the reference). ray is not installed in the build image; without
it this script explains that and exits with status 2 — use scripts/rollout_batched.py for a
ray-free batched rollout.
"""
import argparse
import functools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ENV_IDS = ("WarehouseSmall-v0", "WarehouseMedium-v0", "WarehouseLarge-v0")


def episode_metrics(info):
    """train.py:18-23: mean per-agent return of the finished episode, overall and bucketed by
    the number of agents that episode had."""
    episode = info["episode"]
    returns = list(episode.agent_rewards.values())
    mean_return = sum(returns) / len(returns)
    episode.custom_metrics["avg_agent_reward_all"] = [mean_return]
    episode.custom_metrics[f"avg_agent_reward_{len(returns)}"] = [mean_return]


def _make_env(cls, _env_config):
    return cls()


def main(config_path):
    try:
        import ray
        import yaml
        from ray.tune.registry import register_env
        from ray.tune.tune import run_experiments
    except ImportError as e:
        print(f"scripts/train.py needs ray[rllib] and pyyaml ({e}); they are not installed here.", file=sys.stderr)
        return 2
    from warehouse import WarehouseLargeTrain, WarehouseMediumTrain, WarehouseSmallTrain
    ray.init()
    for env_id, cls in zip(ENV_IDS, (WarehouseSmallTrain, WarehouseMediumTrain, WarehouseLargeTrain)):
        register_env(env_id, functools.partial(_make_env, cls))   # bind cls now, not at call time
    with open(config_path) as f:
        experiments = yaml.safe_load(f)
    for spec in experiments.values():
        spec.setdefault("config", {})["callbacks"] = {"on_episode_end": episode_metrics}
    run_experiments(experiments)
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("config_path", type=str, help="path to the experiment config file")
    sys.exit(main(ap.parse_args().config_path))
