"""Pins the numpy port (oracle/ref_port.py) against the fixtures recorded from the reference."""
import numpy as np
import pytest

import golden_util as gu
from oracle import ref_port as rp


def make_env(kw, n, num_agents):
    return rp.PortEnv(kw, n, num_agents)


@pytest.mark.parametrize("size", gu.SIZES)
def test_episodes(size):
    d = gu.load(f"episodes_{size}.npz")
    for p in gu.episode_prefixes(d)[:5]:
        gu.check_episode(make_env, d, p)


@pytest.mark.parametrize("size", gu.SIZES)
def test_single_steps(size):
    d = gu.load(f"single_steps_{size}.npz")
    d = {k: (v[:150] if getattr(v, "ndim", 0) and v.shape[0] == 400 else v) for k, v in d.items()}
    assert gu.check_single_steps(make_env, d) == 150


def test_quirk_scenarios():
    d = gu.load("quirks_small.npz")
    names = [str(s) for s in d.pop("names")]
    gu.check_single_steps(make_env, d, names)


@pytest.mark.parametrize("size", gu.SIZES)
def test_solver(size):
    def greedy_fn(kw, obs, num_agents, rand_prob, is_random, random_actions):
        n, R = len(num_agents), kw["num_requests"]
        out = np.full((n, R), -1, np.int32)
        for e in range(n):
            A = int(num_agents[e])
            per_agent = [{k: obs[k][e][i] for k in gu.OBS_KEYS} for i in range(A)]
            out[e, :A] = rp.greedy_actions(per_agent, A, R, is_random[e], random_actions[e])
        return out
    gu.check_solver(greedy_fn, gu.load(f"solver_{size}.npz"))
