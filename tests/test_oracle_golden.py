"""Pins the C oracle (oracle/wh_oracle.c) against fixtures produced by executing the UNMODIFIED
reference (oracle/make_golden.py). CPU only."""
import numpy as np
import pytest

import golden_util as gu
from oracle import wh_oracle as wo


def make_env(kw, n, num_agents):
    return wo.OracleEnv(wo.make_config(**kw), n, num_agents=num_agents)


def greedy_fn(kw, obs, num_agents, rand_prob, is_random, random_actions):
    env = wo.OracleEnv(wo.make_config(**kw), len(num_agents))
    env.state["num_agents"][:] = num_agents
    obs = {k: np.ascontiguousarray(v.reshape(env.obs[k].shape), env.obs[k].dtype) for k, v in obs.items()}
    return env.greedy(obs=obs, rand_prob=rand_prob, is_random=is_random, random_actions=random_actions).copy()


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert wo.philox(0, 0, 0, 0, 0) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert wo.philox(*([0xFFFFFFFF] * 4), 0xFFFFFFFFFFFFFFFF) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert wo.philox(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0x299F31D0A4093822) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


@pytest.mark.parametrize("size", gu.SIZES)
def test_episodes(size):
    d = gu.load(f"episodes_{size}.npz")
    prefixes = gu.episode_prefixes(d)
    assert len(prefixes) == 7
    steps = sum(gu.check_episode(make_env, d, p) for p in prefixes)
    assert steps == 4 * 210 + 3 * 60


@pytest.mark.parametrize("size", gu.SIZES)
def test_single_steps(size):
    assert gu.check_single_steps(make_env, gu.load(f"single_steps_{size}.npz")) == 400


def test_quirk_scenarios():
    d = gu.load("quirks_small.npz")
    names = [str(s) for s in d.pop("names")]
    assert gu.check_single_steps(make_env, d, names) == len(names) >= 28


@pytest.mark.parametrize("size", gu.SIZES)
def test_solver(size):
    d = gu.load(f"solver_{size}.npz")
    assert gu.check_solver(greedy_fn, d) == 120
    # p = 0: same obs, no random branch -> pure greedy; the non-random rows must equal the golden ones
    keep = d["is_random"] == 0
    d0 = dict(d)
    d0["is_random"] = np.zeros_like(d["is_random"])
    pure = greedy_fn(gu.cfg_kwargs(d), {k: d["obs_" + k] for k in gu.OBS_KEYS},
                     np.full(120, d["actions"].shape[1], np.int32), 0.0, None, None)
    assert np.array_equal(pure[:, : d["actions"].shape[1]][keep], d["actions"][keep])


@pytest.mark.parametrize("name", ["small_random", "medium_greedy", "large_random", "small_train_greedy", "large_train_random"])
def test_full_size_reference_digests(name):
    """BASELINE configs[1] at full size (4 096 Small envs x 200 steps) and the configs[2]/[3] replay
    subsets: the C oracle reproduces the reference's per-step CRCs of every output array."""
    d = gu.load(f"batch_{name}.npz")
    n = gu.check_batch_digests(make_env, d, greedy_fn=lambda env: env.greedy().copy())
    assert n == int(d["n"]) * 200
