#!/usr/bin/env python
"""Checkpoint rollout with the reference's command line (scripts/rollout.py:90-113):

    python scripts/rollout.py TRIAL_DIR NUM_AGENTS [-i ITERATION] [-r] [--run ALGO] [-n EPISODES]

Restores the RLlib trainer of a Tune trial (params.json + checkpoint_N/checkpoint-N: the latest one, or
the one closest to -i/--iteration, with the reference's messages, rollout.py:33-56) and evaluates it on
the fixed-size variant of the trained environment with NUM_AGENTS agents, one `trainer.compute_action`
per agent and step (rollout.py:71-73). Rendering is off unless -r/--render is given, as in the reference.
Extras beyond the reference: --run (the reference hard-codes SAC, rollout.py:46) and -n/--num-episodes.
Needs ray[rllib]; without it the script says so and exits 2. For policies that are plain torch modules
use scripts/rollout_batched.py, which keeps observations on the GPU and evaluates thousands of episodes
at once.
"""
import argparse
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pick_checkpoint(trial_dir, iteration):
    """rollout.py:33-56: the newest checkpoint for iteration == -1, else the one closest to `iteration`.
    Returns (chosen iteration, restore path, message)."""
    found = {}
    for name in os.listdir(trial_dir):
        m = re.fullmatch(r"checkpoint_(\d+)", name)
        if m:
            found[int(m.group(1))] = os.path.join(trial_dir, name, f"checkpoint-{m.group(1)}")
    if not found:
        raise FileNotFoundError(f"no checkpoint_N directories under {trial_dir}")
    if iteration == -1:
        choice = max(found)
        msg = f"Loading the lastest checkpoint at iteration {choice}."
    else:
        choice = min(sorted(found), key=lambda k: abs(k - iteration))
        msg = (f"Loading the selected checkpoint at iteration {choice}." if iteration in found else
               f"Checkpoint at iteration {iteration} doesn't exist, loading the closest one at {choice} instead.")
    return choice, found[choice], msg


def main(a):
    try:
        import ray
        from ray.tune.registry import register_env
    except ImportError as e:
        print(f"scripts/rollout.py needs ray[rllib] ({e}); it is not installed here.", file=sys.stderr)
        return 2
    import functools
    from warehouse import WarehouseLarge, WarehouseMedium, WarehouseSmall
    with open(os.path.join(a.trial_dir, "params.json")) as f:
        params = json.load(f)
    env_types = {"WarehouseSmall-v0": WarehouseSmall, "WarehouseMedium-v0": WarehouseMedium,
                 "WarehouseLarge-v0": WarehouseLarge}
    ray.init()
    for env_id, cls in env_types.items():                       # rollout.py:25-31 (class bound eagerly)
        register_env(env_id, functools.partial(lambda c, _cfg: c(1), cls))
    _, path, msg = pick_checkpoint(a.trial_dir, a.iteration)
    if a.run == "SAC":
        from ray.rllib.agents.sac.sac import SACTrainer as trainer_cls          # rollout.py:10,46
    else:
        from ray.rllib.agents.registry import get_agent_class
        trainer_cls = get_agent_class(a.run)
    trainer = trainer_cls(config=params)
    trainer.restore(path)
    print(msg)
    env = env_types[params["env"]](a.num_agents)
    for _ in range(a.num_episodes):
        observations = env.reset()
        acc_rewards = [0.0 for _ in range(len(observations))]
        done, step_count = False, 0
        while not done:
            if a.render:
                env.render(animate=True)
            action_dict = {f"{i}": trainer.compute_action(observations[f"{i}"]) for i in range(env.num_agents)}
            observations, rewards, dones, _ = env.step(action_dict=action_dict)
            acc_rewards = [acc_rewards[i] + rewards[f"{i}"] for i in range(env.num_agents)]
            done = dones["__all__"]
            if a.render:
                print(f"\n=== Step {step_count} ===")
                print("Rewards:", *acc_rewards)
                print(f"Total: {sum(acc_rewards)}, Per Agent: {sum(acc_rewards) / len(acc_rewards)}")
            step_count += 1
        print(f"\n=== Done ({step_count} steps) ===")
        print("Rewards:", *acc_rewards)
        print(f"Total: {sum(acc_rewards)}, Per Agent: {sum(acc_rewards) / len(acc_rewards)}")
    return 0


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("trial_dir", type=str, help="path to the folder of the saved training trial")
    ap.add_argument("num_agents", type=int, help="number of agents")
    ap.add_argument("-i", "--iteration", type=int, default=-1, help="the iteration of the checkpoint to be loaded")
    ap.add_argument("-r", "--render", help="render the environment on each step", action="store_true")
    ap.add_argument("--run", type=str, default="SAC", help="RLlib algorithm the trial was trained with (extra)")
    ap.add_argument("-n", "--num-episodes", type=int, default=1, help="episodes to roll out (extra)")
    return ap.parse_args(argv)


if __name__ == "__main__":
    sys.exit(main(parse()))
