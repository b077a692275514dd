"""GPU parity, part 1: the CUDA path (through the C ABI) replays the fixtures recorded from the
UNMODIFIED reference bit-for-bit — every state array, observation key, reward and done flag."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def make_env(kw, n, num_agents):
    from rllib_warehouse_b200 import BatchedWarehouse, WarehouseConfig
    cfg = WarehouseConfig(kw["num_requests"], kw["area_dimension"], tuple(kw["racks"]), kw["episode"], kw["wait"])
    return BatchedWarehouse(cfg, n, num_agents=num_agents)


def greedy_fn(kw, obs, num_agents, rand_prob, is_random, random_actions):
    env = make_env(kw, len(num_agents), None)
    env.load_state(num_agents=num_agents)
    return env.greedy_actions(obs, 0.0, 0, is_random, random_actions).cpu().numpy()


@pytest.mark.parametrize("size", gu.SIZES)
def test_episodes(size):
    d = gu.load(f"episodes_{size}.npz")
    steps = sum(gu.check_episode(make_env, d, p) for p in gu.episode_prefixes(d))
    assert steps == 4 * 210 + 3 * 60


@pytest.mark.parametrize("size", gu.SIZES)
def test_single_steps(size):
    assert gu.check_single_steps(make_env, gu.load(f"single_steps_{size}.npz")) == 400


def test_quirk_scenarios():
    d = gu.load("quirks_small.npz")
    names = [str(s) for s in d.pop("names")]
    assert gu.check_single_steps(make_env, d, names) == len(names)


@pytest.mark.parametrize("size", gu.SIZES)
def test_solver(size):
    assert gu.check_solver(greedy_fn, gu.load(f"solver_{size}.npz")) == 120


@pytest.mark.parametrize("name", ["small_random", "medium_greedy", "large_random", "small_train_greedy", "large_train_random"])
def test_full_size_reference_digests(name):
    """BASELINE configs[1] exactly as SURVEY §8d config 2 states it — 4 096 Small envs, 200 steps,
    env e = the unmodified reference seeded with BASE+e, PCG64 action tensor — plus the Medium
    greedy-solver and Large replay subsets: every state tensor, observation key, action, reward and
    done flag of the CUDA path has the reference's CRC at every step."""
    d = gu.load(f"batch_{name}.npz")
    n = gu.check_batch_digests(make_env, d, greedy_fn=lambda env: env.greedy_actions().clone())
    assert n == int(d["n"]) * 200
