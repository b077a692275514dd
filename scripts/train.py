#!/usr/bin/env python
"""RLlib training driver with the reference's command line (scripts/train.py:45-50):

    python scripts/train.py scripts/experiments/warehouse-small-ppo/warehouse-small-ppo.yaml

Registers "Warehouse{Small,Medium,Large}-v0" exactly like the reference (train.py:29-35) — but the env
creator hands RLlib ONE vectorised `WarehouseVectorEnv` (a `BaseEnv`: every `send_actions` is a single
kernel launch for all `num_envs` environments, observations arrive RLlib-flattened from the step kernel)
instead of a one-env `MultiAgentEnv` that RLlib would copy `num_envs_per_worker` times and step in a
Python loop. The environments are the *Train variants (random agent count per episode,
variants.py:65-98). Per-experiment knobs come from the spec's `env_config`:

    env_config: {num_envs: 256, flat_obs: true, device: "cuda:0", seed: 0, vector: true}

`vector: false` (or an RLlib without `ray.rllib.env.base_env.BaseEnv`) falls back to the single-env
`Warehouse*Train` classes, i.e. the reference's own wiring. The reference's per-episode metrics
(avg_agent_reward_all / avg_agent_reward_{n}, train.py:18-23) are injected as callbacks and the Tune
experiment dict is handed to `run_experiments`; the algorithm is whatever the YAML's `run:` names.

Needs `ray[rllib]` (0.8.x API, as the reference). ray is not installed in the build image; without it
this script explains that and exits with status 2 — `rllib_warehouse_b200.RolloutSampler` is the
ray-free stand-in for the sampler loop, scripts/rollout_batched.py the ray-free batched evaluation.
"""
import argparse
import functools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ENV_IDS = {"WarehouseSmall-v0": "small", "WarehouseMedium-v0": "medium", "WarehouseLarge-v0": "large"}


def episode_metrics(info):
    """train.py:18-23: mean per-agent return of the finished episode, overall and bucketed by
    the number of agents that episode had."""
    episode = info["episode"]
    returns = list(episode.agent_rewards.values())
    mean_return = sum(returns) / len(returns)
    episode.custom_metrics["avg_agent_reward_all"] = [mean_return]
    episode.custom_metrics[f"avg_agent_reward_{len(returns)}"] = [mean_return]


def have_base_env():
    try:
        from ray.rllib.env.base_env import BaseEnv  # noqa: F401
        return True
    except Exception:  # noqa: BLE001
        return False


def make_env(size, env_config=None, vector=None):
    """The env creator registered for every env id (train.py:34-35 passes `lambda _: val()`; here the
    EnvContext / dict RLlib passes is honoured). vector=None: vectorised iff RLlib has BaseEnv."""
    cfg = dict(env_config or {})
    if vector is None:
        vector = bool(cfg.get("vector", True)) and have_base_env()
    if not vector:
        from warehouse import WarehouseLargeTrain, WarehouseMediumTrain, WarehouseSmallTrain
        return {"small": WarehouseSmallTrain, "medium": WarehouseMediumTrain, "large": WarehouseLargeTrain}[size]()
    from rllib_warehouse_b200 import VARIANTS, WarehouseVectorEnv
    worker = int(getattr(env_config, "worker_index", 0) or 0)        # EnvContext: distinct env ids per rollout worker
    num_envs = int(cfg.get("num_envs", 64))
    return WarehouseVectorEnv(VARIANTS[size].replace(random_num_agents=True), num_envs,
                              device=cfg.get("device", "cuda:0"), seed=int(cfg.get("seed", 0)),
                              flat_obs=bool(cfg.get("flat_obs", True)), env_id0=worker * num_envs)


def main(config_path):
    try:
        import ray
        import yaml
        from ray.tune.registry import register_env
        from ray.tune.tune import run_experiments
    except ImportError as e:
        print(f"scripts/train.py needs ray[rllib] and pyyaml ({e}); they are not installed here.", file=sys.stderr)
        return 2
    ray.init()
    for env_id, size in ENV_IDS.items():
        register_env(env_id, functools.partial(make_env, size))   # bind the size now, not at call time
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
